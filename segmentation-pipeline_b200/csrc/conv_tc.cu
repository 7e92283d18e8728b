// Tensor-core convolution engine for sm_100a: implicit GEMM on tcgen05.mma with TMEM accumulators, operands
// staged by TMA, one CTA per output tile.
//
// Geometry (all three modes share it)
//   * GEMM M = 128 output positions of one z-plane: a 16 (y) x 8 (x) patch of voxels; row m = y*8 + x.
//   * GEMM N = output channels of up to 3-4 NEIGHBOURING z-planes at once ("dz fusion"): an input plane
//     contributes to several output planes through different dz taps, so one A operand (a shifted view of the
//     input halo) is multiplied by the concatenation [W(dz=2) | W(dz=1) | W(dz=0)] and accumulated into the
//     adjacent TMEM column ranges of those planes.  This raises N from Cout (40..80) to 3*Cout, which is what
//     lifts the MMA off the shared-memory operand-bandwidth bound measured in profiles/r01_probe_mma_rate.log
//     (cycles/MMA = max(N/2, (4096 + 32 N)/128)).
//   * GEMM K = 16 per MMA = two 8-channel chunks of one tap, or (for an odd chunk) the same chunk at two taps.
//   * A operand: the input halo of ONE z-plane, 18 (y) x 10 (x) voxels x 2 chunks, lands in shared memory by one
//     TMA box load as [chunk][y][x][8ch] (no swizzle).  Tap (dy,dx) is the same bytes read at byte offset
//     (dy*10+dx)*16 with SBO = 160 B (next y row) and LBO = 2880 B (next chunk): no im2col, no re-load per tap.
//     Zero padding at the PATCH border (components.py conv padding=1) is TMA out-of-bounds zero fill.
//   * B operand: weights pre-packed on the host (models/_plan.py::pack_tc_weight) in exactly the shared-memory
//     image [step][k-half][row][8], row = (plane block, cout), plus 16 zero rows so that N rounded up to a
//     multiple of 16 multiplies zeros ("spill" columns add 0 to the next plane's accumulator).
//   * accumulators: TZ planes x Cpad fp32 columns in TMEM; first touch of a plane uses accumulate=0.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
// (TMEM -> registers -> folded BN / bias, activation, residual, bf16 store or softmax -> fp32 NCDHW).
#include <cstring>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace b200seg {
using namespace ptx;

constexpr int kHX = 10, kHY = 18;                 // halo extent of a 8 x 16 tile
constexpr int kChunkBytes = kHX * kHY * 16;       // 2880: one 8-channel chunk of one halo plane
constexpr int kAStageBytes = 2 * kChunkBytes;     // 5760
constexpr int kNA = 4;                            // A ring depth
constexpr int kMaxBuf = 2;                        // B ring depth (1 or 2, chosen per launch)
constexpr int kThreadsTc = 192;
constexpr int kMaxZin = 28;                       // input planes a tile may walk (DOWN: 2*TZ+2, TZ <= 12)
constexpr int kMaxBlk = 4;                        // MMAs (column blocks) per step and plane

struct TcMaps {
    CUtensorMap m[8];  // K3/UP: [0] = 2-chunk box, [1] = 1-chunk box.  DOWN: [pp] 2-chunk, [4+pp] 1-chunk.
};

struct TcParams {
    int mode;
    int TZ;          // output planes per tile
    int Cpad;        // output channels rounded up to 8
    int NB;          // rows of one B k-half (blocks*Cpad + 16)
    int G;           // chunk groups (pairs); the last may be a single chunk
    int lone_last;   // 1 if the last group has one chunk
    int steps_full, steps_lone;
    int bimg_stride;  // bytes between consecutive B images in wpacked
    int maxp;         // planes per MMA
    int n_pass, n_bimg;  // passes and B images per pass
    int nbuf;         // B ring depth
    int chunk_base;   // in.c8_off
    int c8_total;     // in.c8_total
    int tiles_x, tiles_y, tiles_z;
    int out_z, out_y, out_x;   // extent the tile grid covers (K3/DOWN: output; UP: low-res input y/x, output z)
    int zin_count;    // input planes a tile walks (zi range)
    const uint8_t* wpacked;
    DEpilogue epi;
};

__host__ __device__ __forceinline__ uint32_t pad16(uint32_t n) { return (n + 15u) & ~15u; }

// byte offset / LBO of the A operand for step `st`
__device__ __forceinline__ void step_desc(int mode, bool lone, int pp, int st, uint32_t& off, uint32_t& lbo) {
    if (mode == B200SEG_TC_K3) {
        if (!lone) {
            off = ((st / 3) * kHX + (st % 3)) * 16;
            lbo = kChunkBytes;
        } else if (st < 3) {
            off = (st * kHX) * 16;  // taps (dy=st,dx=0) + (dy=st,dx=1)
            lbo = 16;
        } else if (st == 3) {
            off = 2 * 16;           // taps (0,2) + (1,2)
            lbo = kHX * 16;
        } else {
            off = (2 * kHX + 1) * 16;  // (2,1) with zero weights + (2,2)
            lbo = 16;
        }
    } else {
        const int py = pp >> 1, px = pp & 1;
        // first halo offset used by this parity: DOWN parity 1 -> 0, parity 0 -> 1 ; UP parity 0 -> 0, parity 1 -> 1
        const int sy0 = (mode == B200SEG_TC_DOWN) ? (1 - py) : py;
        const int sx0 = (mode == B200SEG_TC_DOWN) ? (1 - px) : px;
        if (!lone) {
            off = ((sy0 + st / 2) * kHX + (sx0 + st % 2)) * 16;
            lbo = kChunkBytes;
        } else {
            off = ((sy0 + st) * kHX + sx0) * 16;  // taps (sy, sx) + (sy, sx+1)
            lbo = 16;
        }
    }
}

// planes touched by input plane zi: [lo, hi], B row block of plane lo, first plane that is touched for the first time
__device__ __forceinline__ void plane_window(int mode, int TZ, int zi, int& lo, int& hi, int& jlo, int& ft) {
    if (mode == B200SEG_TC_K3) {
        lo = max(zi - 2, 0);
        hi = min(zi, TZ - 1);
        jlo = 2 - zi + lo;
        ft = zi < TZ ? zi : TZ;
    } else if (mode == B200SEG_TC_DOWN) {
        const int q = zi >> 1;
        lo = max(q - 1, 0);
        hi = min(q, TZ - 1);
        jlo = lo - (q - 1);
        ft = ((zi & 1) == 0 && q < TZ) ? q : TZ;
    } else {
        lo = max(2 * zi - 3, 0);
        hi = min(2 * zi, TZ - 1);
        jlo = lo - (2 * zi - 3);
        ft = max(2 * zi - 1, 0);
    }
}

// One MMA of a (plane, step): accumulator column, instruction descriptor, B row offset (16-byte units), accumulate flag
struct MmaBlk {
    uint32_t dcol, idesc, brow, acc;
};
struct PlaneTab {
    MmaBlk blk[2][kMaxBlk];  // [0] = normal step, [1] = the very first step of a pass (first-touch split)
    int nblk[2];
};

__device__ inline void build_plane_tab(const TcParams& p, int zi, PlaneTab& t) {
    int lo, hi, jlo, ft;
    plane_window(p.mode, p.TZ, zi, lo, hi, jlo, ft);
    for (int first = 0; first < 2; ++first) {
        const int split = first ? min(max(ft, lo), hi + 1) : hi + 1;
        int nb = 0;
        for (int q = lo; q <= hi;) {
            const bool overwrite = q >= split;
            const int lim = overwrite ? hi + 1 : split;
            const int np = min(p.maxp, lim - q);
            MmaBlk b;
            b.dcol = q * p.Cpad;
            b.idesc = make_idesc_bf16(128, pad16(np * p.Cpad));
            b.brow = (jlo + (q - lo)) * p.Cpad;
            b.acc = overwrite ? 0u : 1u;
            if (nb < kMaxBlk) t.blk[first][nb] = b;
            ++nb;
            q += np;
        }
        t.nblk[first] = nb;
    }
}

template <uint32_t kTmemCols>
__global__ void __launch_bounds__(kThreadsTc, 2)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + kNA * kAStageBytes;
    const int bbuf_bytes = (p.bimg_stride + 127) & ~127;
    uint8_t* after_b = sB + p.nbuf * bbuf_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(after_b);
    uint64_t* full_a = bars;
    uint64_t* empty_a = bars + kNA;
    uint64_t* full_b = bars + 2 * kNA;
    uint64_t* empty_b = full_b + kMaxBuf;
    uint64_t* acc_full = empty_b + kMaxBuf;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    // tables: A step descriptors [lone][pp][9], per-plane MMA blocks, epilogue parameters
    uint32_t* step_tab = tmem_slot + 2;                                   // 2 * 4 * 9 words
    PlaneTab* plane_tab = reinterpret_cast<PlaneTab*>(step_tab + 72);     // kMaxZin entries (8-byte aligned)
    float* s_scale = reinterpret_cast<float*>(plane_tab + kMaxZin);
    float* s_shift = s_scale + p.Cpad;
    float* s_slope = s_shift + p.Cpad;

    const int warp = threadIdx.x / 32;
    const int lane = threadIdx.x % 32;

    // ---- tile coordinates
    int tile = blockIdx.x;
    const int tx = tile % p.tiles_x;
    tile /= p.tiles_x;
    const int ty = tile % p.tiles_y;
    const int tz = tile / p.tiles_y;
    const int n = blockIdx.y;
    const int x0 = tx * 8, y0 = ty * 16, z0 = tz * p.TZ;
    const int chunk0 = n * p.c8_total + p.chunk_base;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kNA; ++i) {
            mbar_init(&full_a[i], 1);
            mbar_init(&empty_a[i], 1);
        }
        for (int i = 0; i < kMaxBuf; ++i) {
            mbar_init(&full_b[i], 1);
            mbar_init(&empty_b[i], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    // cooperative table build
    for (int i = threadIdx.x; i < 72; i += blockDim.x) {
        const int lone = i / 36, pp = (i / 9) % 4, st = i % 9;
        uint32_t off, lbo;
        step_desc(p.mode, lone != 0, pp, st, off, lbo);
        step_tab[i] = (off >> 4) | ((lbo >> 4) << 16);
    }
    for (int zi = threadIdx.x; zi < p.zin_count; zi += blockDim.x) build_plane_tab(p, zi, plane_tab[zi]);
    for (int c = threadIdx.x; c < p.Cpad; c += blockDim.x) {
        s_scale[c] = p.epi.scale[c];
        s_shift[c] = p.epi.shift[c];
        s_slope[c] = p.epi.slope[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // =============================================================== TMA producer
        if (elect_one()) {
            const int nmap = p.mode == B200SEG_TC_DOWN ? 8 : 2;
            for (int i = 0; i < nmap; ++i) prefetch_tmap(&maps.m[i]);
            uint32_t a_it = 0, b_it = 0;
            for (int pass = 0; pass < p.n_pass; ++pass) {
                for (int bi = 0; bi < p.n_bimg; ++bi) {
                    const int g = bi % p.G;
                    const bool lone = p.lone_last && g == p.G - 1;
                    const int nsteps = lone ? p.steps_lone : p.steps_full;
                    // ---- B image
                    {
                        const uint32_t s = b_it % p.nbuf, ph = (b_it / p.nbuf) & 1;
                        mbar_wait(&empty_b[s], ph ^ 1);
                        const uint32_t bytes = nsteps * 2 * p.NB * 16;
                        mbar_arrive_expect_tx(&full_b[s], bytes);
                        const uint8_t* src = p.wpacked + static_cast<size_t>(pass * p.n_bimg + bi) * p.bimg_stride;
                        bulk_load(sB + s * bbuf_bytes, src, bytes, &full_b[s]);
                        ++b_it;
                    }
                    // ---- A stages
                    int zi_start = 0, zi_step = 1, pp = 0;
                    if (p.mode == B200SEG_TC_DOWN) {
                        zi_start = bi / (4 * p.G);  // z parity
                        zi_step = 2;
                        pp = (bi / p.G) % 4;
                    } else if (p.mode == B200SEG_TC_UP) {
                        pp = pass;
                    }
                    (void)pp;
                    for (int zi = zi_start; zi < p.zin_count; zi += zi_step) {
                        const uint32_t s = a_it % kNA, ph = (a_it / kNA) & 1;
                        mbar_wait(&empty_a[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full_a[s], lone ? kChunkBytes : kAStageBytes);
                        uint8_t* dst = sA + s * kAStageBytes;
                        if (p.mode == B200SEG_TC_K3) {
                            tma_load_4d(dst, &maps.m[lone ? 1 : 0], &full_a[s], (x0 - 1) * 8, y0 - 1, z0 - 1 + zi,
                                        chunk0 + 2 * g);
                        } else if (p.mode == B200SEG_TC_UP) {
                            // tile origin is in low-res input coordinates; z0 counts OUTPUT planes (even)
                            tma_load_4d(dst, &maps.m[lone ? 1 : 0], &full_a[s], (x0 - 1) * 8, y0 - 1,
                                        z0 / 2 - 1 + zi, chunk0 + 2 * g);
                        } else {
                            tma_load_5d(dst, &maps.m[(lone ? 4 : 0) + pp], &full_a[s], 0, x0 - 1, y0 - 1,
                                        2 * z0 - 1 + zi, chunk0 + 2 * g);
                        }
                        ++a_it;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer
        if (elect_one()) {
            uint32_t a_it = 0, b_it = 0;
            const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
            // descriptor high words: SBO | version(1) at bit 46 ; A: SBO = 160 B, B: SBO = 128 B
            const uint64_t a_hi = (static_cast<uint64_t>((kHX * 16) >> 4) | (1ull << 14)) << 32;
            const uint64_t b_hi = (static_cast<uint64_t>(128 >> 4) | (1ull << 14)) << 32;
            const uint32_t b_lbo = static_cast<uint32_t>(p.NB) << 16;   // LBO = NB * 16 bytes
            const uint32_t b_step16 = 2 * p.NB;                         // one step of a B image, in 16-byte units
            for (int pass = 0; pass < p.n_pass; ++pass) {
                if (pass > 0) {
                    mbar_wait(acc_empty, (pass - 1) & 1);
                    tc_fence_after();
                }
                for (int bi = 0; bi < p.n_bimg; ++bi) {
                    const int g = bi % p.G;
                    const bool lone = p.lone_last && g == p.G - 1;
                    const int nsteps = lone ? p.steps_lone : p.steps_full;
                    const uint32_t bs = b_it % p.nbuf, bph = (b_it / p.nbuf) & 1;
                    mbar_wait(&full_b[bs], bph);
                    tc_fence_after();
                    const uint32_t bimg16 = sB16 + bs * (bbuf_bytes >> 4);
                    int zi_start = 0, zi_step = 1, pp = 0;
                    if (p.mode == B200SEG_TC_DOWN) {
                        zi_start = bi / (4 * p.G);
                        zi_step = 2;
                        pp = (bi / p.G) % 4;
                    } else if (p.mode == B200SEG_TC_UP) {
                        pp = pass;
                    }
                    const uint32_t* steps = step_tab + (lone ? 36 : 0) + pp * 9;
                    for (int zi = zi_start; zi < p.zin_count; zi += zi_step) {
                        const uint32_t s = a_it % kNA, ph = (a_it / kNA) & 1;
                        mbar_wait(&full_a[s], ph);
                        tc_fence_after();
                        const PlaneTab& pt = plane_tab[zi];
                        const uint32_t a16 = sA16 + s * (kAStageBytes >> 4);
                        int st = 0;
                        if (bi == 0) {
                            // first step of the pass: first-touch split
                            const uint64_t adesc = a_hi | static_cast<uint64_t>(steps[0] + a16);
                            const int nb = pt.nblk[1];
                            for (int b = 0; b < nb; ++b) {
                                const MmaBlk k = pt.blk[1][b];
                                umma_bf16(tmem + k.dcol, adesc, b_hi | static_cast<uint64_t>((bimg16 + k.brow) | b_lbo),
                                          k.idesc, k.acc);
                            }
                            st = 1;
                        }
                        const int nb = pt.nblk[0];
                        const MmaBlk k0 = pt.blk[0][0];
                        if (nb == 1) {
                            const uint32_t d0 = tmem + k0.dcol;
                            uint32_t blo = (bimg16 + st * b_step16 + k0.brow) | b_lbo;
#pragma unroll 1
                            for (; st < nsteps; ++st) {
                                umma_bf16(d0, a_hi | static_cast<uint64_t>(steps[st] + a16), b_hi | static_cast<uint64_t>(blo),
                                          k0.idesc, 1u);
                                blo += b_step16;
                            }
                        } else {
#pragma unroll 1
                            for (; st < nsteps; ++st) {
                                const uint64_t adesc = a_hi | static_cast<uint64_t>(steps[st] + a16);
                                const uint32_t bst = bimg16 + st * b_step16;
                                for (int b = 0; b < nb; ++b) {
                                    const MmaBlk k = pt.blk[0][b];
                                    umma_bf16(tmem + k.dcol, adesc, b_hi | static_cast<uint64_t>((bst + k.brow) | b_lbo),
                                              k.idesc, 1u);
                                }
                            }
                        }
                        umma_commit(&empty_a[s]);
                        ++a_it;
                    }
                    umma_commit(&empty_b[bs]);
                    ++b_it;
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // =============================================================== epilogue (warps 2..5)
        const int lg = warp & 3;            // TMEM lane quarter this warp may read
        const int m = lg * 32 + lane;       // accumulator row
        const int my = m >> 3, mx = m & 7;
        const DEpilogue& e = p.epi;
        const int c8 = p.Cpad / 8;
        for (int pass = 0; pass < p.n_pass; ++pass) {
            mbar_wait(acc_full, pass & 1);
            tc_fence_after();
            int oy, ox;
            bool valid;
            if (p.mode == B200SEG_TC_UP) {
                const int py = pass >> 1, px = pass & 1;
                oy = 2 * (y0 + my) + py;
                ox = 2 * (x0 + mx) + px;
                valid = (y0 + my) < p.out_y && (x0 + mx) < p.out_x;
            } else {
                oy = y0 + my;
                ox = x0 + mx;
                valid = oy < p.out_y && ox < p.out_x;
            }
            for (int q = 0; q < p.TZ; ++q) {
                const int oz = z0 + q;
                if (oz >= p.out_z) break;
                const uint32_t taddr = tmem + (static_cast<uint32_t>(lg * 32) << 16) + q * p.Cpad;
                if (e.out_ncdhw == nullptr) {
                    // 5 chunks (40 channels) per round: all TMEM loads, then all residual loads, then math + stores
                    for (int c0 = 0; c0 < c8; c0 += 5) {
                        uint32_t r[5][8];
#pragma unroll
                        for (int j = 0; j < 5; ++j)
                            if (c0 + j < c8) tmem_ld8(taddr + (c0 + j) * 8, r[j]);
                        uint4 res[5];
                        const bool has_res = e.residual.data != nullptr;
                        if (valid && has_res) {
#pragma unroll
                            for (int j = 0; j < 5; ++j)
                                if (c0 + j < c8 && c0 + j < e.split_c8)
                                    res[j] = __ldg(reinterpret_cast<const uint4*>(e.residual.data) +
                                                   vox_index(e.residual, n, c0 + j, oz, oy, ox));
                        }
                        tmem_ld_wait();
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                const int cc = c0 + j;
                                if (cc >= c8) continue;
                                float v[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const int c = cc * 8 + i;
                                    float t = fmaf(__uint_as_float(r[j][i]), s_scale[c], s_shift[c]);
                                    v[i] = t > 0.f ? t : t * s_slope[c];
                                }
                                const bool to0 = cc < e.split_c8;
                                if (to0 && has_res) {
                                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&res[j]);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        float2 f = __bfloat1622float2(h[i]);
                                        v[2 * i] += f.x;
                                        v[2 * i + 1] += f.y;
                                    }
                                }
                                uint4 o;
                                __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                                for (int i = 0; i < 4; ++i) oh[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                                if (to0)
                                    reinterpret_cast<uint4*>(e.dst0.data)[vox_index(e.dst0, n, cc, oz, oy, ox)] = o;
                                else
                                    reinterpret_cast<uint4*>(e.dst1.data)[vox_index(e.dst1, n, cc - e.split_c8, oz, oy, ox)] = o;
                            }
                        }
                    }
                } else {
                    // final layer: affine (+bias), optional channel softmax, fp32 NCDHW store (cout <= 16)
                    uint32_t r0[8], r1[8];
                    tmem_ld8(taddr, r0);
                    if (c8 > 1) tmem_ld8(taddr + 8, r1);
                    tmem_ld_wait();
                    if (valid) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float t = fmaf(__uint_as_float(r0[j]), s_scale[j], s_shift[j]);
                            v[j] = t > 0.f ? t : t * s_slope[j];
                        }
                        if (c8 > 1) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float t = fmaf(__uint_as_float(r1[j]), s_scale[8 + j], s_shift[8 + j]);
                                v[8 + j] = t > 0.f ? t : t * s_slope[8 + j];
                            }
                        }
                        const long long vox = 1LL * p.out_z * p.out_y * p.out_x;
                        float* dst = e.out_ncdhw + static_cast<long long>(n) * e.cout * vox +
                                     (static_cast<long long>(oz) * p.out_y + oy) * p.out_x + ox;
                        if (e.softmax) {
                            float mx_ = -INFINITY;
#pragma unroll
                            for (int c = 0; c < 16; ++c)
                                if (c < e.cout) mx_ = fmaxf(mx_, v[c]);
                            float s = 0.f;
#pragma unroll
                            for (int c = 0; c < 16; ++c)
                                if (c < e.cout) {
                                    v[c] = expf(v[c] - mx_);
                                    s += v[c];
                                }
#pragma unroll
                            for (int c = 0; c < 16; ++c)
                                if (c < e.cout) dst[c * vox] = v[c] / s;
                        } else {
#pragma unroll
                            for (int c = 0; c < 16; ++c)
                                if (c < e.cout) dst[c * vox] = v[c];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(acc_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

struct TcGeom {
    int Cpad, blocks, NB, steps_full, steps_lone, G, lone_last, n_pass, n_bimg, bimg_stride, TZmax;
};

static int tc_geometry(int mode, int cin_chunks, int cout, TcGeom* g) {
    B200SEG_CHECK_ARG(mode >= 0 && mode <= 2, "conv3d_tc: bad mode %d", mode);
    B200SEG_CHECK_ARG(cout >= 1 && cout <= 80, "conv3d_tc: cout %d not in [1,80] (split wider layers)", cout);
    B200SEG_CHECK_ARG(cin_chunks >= 1, "conv3d_tc: no input chunks");
    g->Cpad = (cout + 7) / 8 * 8;
    g->blocks = mode == B200SEG_TC_K3 ? 3 : (mode == B200SEG_TC_DOWN ? 2 : 4);
    g->NB = g->blocks * g->Cpad + 16;
    g->steps_full = mode == B200SEG_TC_K3 ? 9 : 4;
    g->steps_lone = mode == B200SEG_TC_K3 ? 5 : 2;
    g->G = (cin_chunks + 1) / 2;
    g->lone_last = cin_chunks & 1;
    g->n_pass = mode == B200SEG_TC_UP ? 4 : 1;
    g->n_bimg = mode == B200SEG_TC_DOWN ? 8 * g->G : g->G;
    g->bimg_stride = g->steps_full * 2 * g->NB * 16;
    g->TZmax = (256 - 16) / g->Cpad;   // <= 256 TMEM columns per CTA so that two CTAs share an SM
    if (g->TZmax > 12) g->TZmax = 12;
    return B200SEG_OK;
}

static int encode_map(CUtensorMap* tm, int rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    B200SEG_CHECK_ARG(enc != nullptr, "conv3d_tc: cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("conv3d_tc: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        return B200SEG_ERR_CUDA;
    }
    return B200SEG_OK;
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int64_t b200seg_conv3d_tc_wbytes(int32_t mode, int32_t cin_chunks, int32_t cout) {
    TcGeom g;
    if (tc_geometry(mode, cin_chunks, cout, &g)) return -1;
    return static_cast<int64_t>(g.n_pass) * g.n_bimg * g.bimg_stride;
}

extern "C" int b200seg_conv3d_tc(int32_t mode, b200seg_view in, const void* wpacked, int64_t wpacked_bytes,
                                 int32_t cout, const b200seg_epilogue* epi, void* stream) {
    int rc = validate_view(in, "conv3d_tc in");
    if (rc) return rc;
    B200SEG_CHECK_ARG(in.dtype == B200SEG_BF16, "conv3d_tc: activations must be bf16");
    B200SEG_CHECK_ARG(wpacked != nullptr && (reinterpret_cast<uintptr_t>(wpacked) & 15) == 0,
                      "conv3d_tc: wpacked must be a 16-byte aligned device pointer");
    B200SEG_CHECK_ARG(epi != nullptr, "conv3d_tc: null epilogue");
    const int cin_chunks = (in.c + 7) / 8;
    TcGeom g;
    rc = tc_geometry(mode, cin_chunks, cout, &g);
    if (rc) return rc;
    const int64_t need = static_cast<int64_t>(g.n_pass) * g.n_bimg * g.bimg_stride;
    B200SEG_CHECK_ARG(wpacked_bytes == need, "conv3d_tc: wpacked holds %lld bytes, geometry needs %lld",
                      static_cast<long long>(wpacked_bytes), static_cast<long long>(need));
    // ---- output extent
    int oz, oy, ox;
    if (mode == B200SEG_TC_K3) {
        oz = in.z; oy = in.y; ox = in.x;
    } else if (mode == B200SEG_TC_DOWN) {
        B200SEG_CHECK_ARG(in.z % 2 == 0 && in.y % 2 == 0 && in.x % 2 == 0, "conv3d_tc DOWN: input extent must be even");
        oz = in.z / 2; oy = in.y / 2; ox = in.x / 2;
    } else {
        oz = in.z * 2; oy = in.y * 2; ox = in.x * 2;
    }
    B200SEG_CHECK_ARG(epi->out_ncdhw == nullptr || mode == B200SEG_TC_K3, "conv3d_tc: out_ncdhw only in K3 mode");
    DEpilogue de;
    rc = make_depilogue(epi, cout, in.n, oz, oy, ox, B200SEG_BF16, &de);
    if (rc) return rc;

    TcParams p{};
    p.mode = mode;
    p.Cpad = g.Cpad;
    p.NB = g.NB;
    p.G = g.G;
    p.lone_last = g.lone_last;
    p.steps_full = g.steps_full;
    p.steps_lone = g.steps_lone;
    p.bimg_stride = g.bimg_stride;
    p.n_pass = g.n_pass;
    p.n_bimg = g.n_bimg;
    p.chunk_base = in.c8_off;
    p.c8_total = in.c8_total;
    p.wpacked = static_cast<const uint8_t*>(wpacked);
    p.epi = de;
    // planes per MMA: N = pad16(np * Cpad) <= 256; keep np*Cpad a multiple of 16 for non-final column blocks
    p.maxp = 256 / g.Cpad;
    if (p.maxp > 4) p.maxp = 4;
    if ((g.Cpad % 16) != 0 && p.maxp >= 2) p.maxp &= ~1;
    if (p.maxp < 1) p.maxp = 1;
    // ---- z tiling
    int tzmax = g.TZmax;
    if (mode == B200SEG_TC_UP) tzmax &= ~1;
    B200SEG_CHECK_ARG(tzmax >= 1, "conv3d_tc: cout too wide for TMEM");
    int ntz = (oz + tzmax - 1) / tzmax;
    int TZ = (oz + ntz - 1) / ntz;
    if (mode == B200SEG_TC_UP && (TZ & 1)) ++TZ;
    {
        // small grids: thinner z tiles until there are at least two waves of CTAs (or TZ bottoms out)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int plane_tiles = (mode == B200SEG_TC_UP ? ((in.x + 7) / 8) * ((in.y + 15) / 16)
                                                       : ((ox + 7) / 8) * ((oy + 15) / 16)) * in.n;
        const int tz_min = mode == B200SEG_TC_UP ? 2 : 1;
        while (TZ > tz_min && plane_tiles * ((oz + TZ - 1) / TZ) < 4 * sms) {
            TZ = (TZ + 1) / 2;
            if (mode == B200SEG_TC_UP && (TZ & 1)) ++TZ;
        }
    }
    p.TZ = TZ;
    p.tiles_z = (oz + TZ - 1) / TZ;
    if (mode == B200SEG_TC_UP) {
        p.tiles_x = (in.x + 7) / 8;
        p.tiles_y = (in.y + 15) / 16;
        p.out_y = in.y;  // validity is tested in low-res coordinates
        p.out_x = in.x;
        p.zin_count = TZ / 2 + 2;
    } else {
        p.tiles_x = (ox + 7) / 8;
        p.tiles_y = (oy + 15) / 16;
        p.out_y = oy;
        p.out_x = ox;
        p.zin_count = mode == B200SEG_TC_K3 ? TZ + 2 : 2 * TZ + 2;
    }
    p.out_z = oz;
    if (de.out_ncdhw != nullptr) {
        p.out_y = oy;
        p.out_x = ox;
    }
    // ---- tensor maps
    TcMaps maps;
    memset(&maps, 0, sizeof maps);
    const cuuint64_t X = in.x, Y = in.y, Z = in.z, NC = static_cast<cuuint64_t>(in.n) * in.c8_total;
    if (mode != B200SEG_TC_DOWN) {
        cuuint64_t dims[4] = {X * 8, Y, Z, NC};
        cuuint64_t strides[3] = {X * 16, Y * X * 16, Z * Y * X * 16};
        for (int one = 0; one < 2; ++one) {
            cuuint32_t box[4] = {kHX * 8, kHY, 1, static_cast<cuuint32_t>(one ? 1 : 2)};
            rc = encode_map(&maps.m[one], 4, in.data, dims, strides, box);
            if (rc) return rc;
        }
    } else {
        cuuint64_t dims[5] = {8, X / 2, Y / 2, Z, NC};
        cuuint64_t strides[4] = {32, 2 * X * 16, Y * X * 16, Z * Y * X * 16};
        for (int pp = 0; pp < 4; ++pp) {
            const int py = pp >> 1, px = pp & 1;
            uint8_t* base = static_cast<uint8_t*>(in.data) + (static_cast<size_t>(py) * in.x + px) * 16;
            for (int one = 0; one < 2; ++one) {
                cuuint32_t box[5] = {8, kHX, kHY, 1, static_cast<cuuint32_t>(one ? 1 : 2)};
                rc = encode_map(&maps.m[(one ? 4 : 0) + pp], 5, base, dims, strides, box);
                if (rc) return rc;
            }
        }
    }
    // ---- launch
    const int bbuf = (g.bimg_stride + 127) & ~127;
    const size_t fixed = static_cast<size_t>(kNA) * kAStageBytes + (2 * kNA + 2 * kMaxBuf + 2) * 8 + 16 + 72 * 4 +
                         sizeof(PlaneTab) * kMaxZin + 3 * static_cast<size_t>(g.Cpad) * 4 + 64;
    // double-buffer the weights when two CTAs per SM still fit (113 KB each)
    p.nbuf = (fixed + 2 * static_cast<size_t>(bbuf) <= 112 * 1024) ? 2 : 1;
    const size_t smem = fixed + static_cast<size_t>(p.nbuf) * bbuf;
    B200SEG_CHECK_ARG(smem <= 227 * 1024, "conv3d_tc: %zu bytes of shared memory needed", smem);
    B200SEG_CHECK_ARG(p.zin_count <= kMaxZin, "conv3d_tc: tile walks %d input planes (max %d)", p.zin_count, kMaxZin);
    const uint32_t cols_needed = TZ * g.Cpad + 16;
    dim3 grid(static_cast<unsigned>(p.tiles_x * p.tiles_y * p.tiles_z), static_cast<unsigned>(in.n));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_TC(COLS)                                                                                          \
    do {                                                                                                         \
        B200SEG_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                static_cast<int>(smem)));                                        \
        conv_tc_kernel<COLS><<<grid, kThreadsTc, smem, s>>>(maps, p);                                            \
    } while (0)
    if (cols_needed <= 64) LAUNCH_TC(64);
    else if (cols_needed <= 128) LAUNCH_TC(128);
    else if (cols_needed <= 256) LAUNCH_TC(256);
    else LAUNCH_TC(512);
#undef LAUNCH_TC
    return check_launch("conv3d_tc");
}
