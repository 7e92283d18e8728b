// Tensor-core convolution engine for sm_100a: implicit GEMM on tcgen05.mma with TMEM accumulators, operands
// staged by TMA, persistent CTAs.
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
//     Up to 120 output channels per launch (TZ = 2); the host keeps large layers at <= 80 (TZ = 3) and uses the wide
//     form only on the small deep levels, where a second launch costs more than the thinner z tiles.
//
//   * K3T (final layer, Cout <= 4): the nine in-plane taps are accumulator COLUMNS instead of MMA steps -- one
//     unshifted MMA per chunk group, the epilogue sums each voxel's nine neighbours through shared memory; tiles
//     own the 14 x 6 interior of the 16 x 8 voxels they compute (see run_image_t and the K3T epilogue).
//
// Execution: persistent CTAs loop over work units (tile x pass).
//   warp 0  A producer : ring of halo planes (TMA box loads), in consumption order
//   warp 2  B producer : weight images, resident (loaded once) or a ring of bulk copies; in dual mode it also
//                        feeds the second A ring and therefore only polls its barriers
//   warp 1  MMA issuer : per plane, the steps of the image are issued from an UNROLLED loop whose descriptor
//   (warp 3)             deltas are compile-time constants, so one MMA costs a couple of uniform adds.  Dual mode:
//                        warps 1 and 3 each own one accumulator set, half of the A ring and every other unit
//   epilogue warps     : TMEM -> registers -> folded BN / bias, activation, residual, bf16 store
//                        (or softmax -> fp32 NCDHW for the last layer); kEpi != 0 are compile-time
//                        specialisations for 10 / 5 chunks with / without residual
// kSets == 1: two CTAs per SM, one accumulator set each (the CTAs overlap each other's epilogue); test variant.
// kSets == 2: one CTA per SM, two accumulator sets (epilogue of unit u overlaps the MMAs of unit u+1); default.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace b200seg {
using namespace ptx;

constexpr int kHX = 10, kHY = 18;                 // halo extent of a 8 x 16 tile
constexpr int kChunkBytes = kHX * kHY * 16;       // 2880: one 8-channel chunk of one halo plane
constexpr int kAStageBytes = 2 * kChunkBytes;     // 5760
constexpr int kMaxA = 12;                         // A ring depth (upper bound)
constexpr int kMaxB = 16;                         // B image slots (ring of 1-2, or ALL images when they fit: resident)
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
    int bbuf_bytes;   // bytes of one B image slot in shared memory
    int maxp;         // planes per MMA
    int n_pass, n_bimg;  // passes and B images per pass
    int gp;              // chunk pairs ("images") handled per visit: 1, or 2 = "super groups" (one A stage holds two chunk
                         // pairs, one B slot two consecutive weight images; halves the per-visit issue overhead)
    int SG, n_simg;      // super groups per parity block (= ceil(G / gp)) and super images per pass (= n_bimg / G * SG)
    int na, nbuf;     // A ring depth, B ring depth
    int resident;     // 1: every weight image has its own slot and is loaded once per CTA (nbuf = n_pass * n_bimg)
    int dual;         // 1: two MMA issuers, one per accumulator set, each with half of the A ring
    int batch;
    int chunk_base;   // in.c8_off
    int c8_total;     // in.c8_total
    int tiles_x, tiles_y, tiles_z;
    int tile_sx, tile_sy;      // tile pitch in voxels: 8 x 16, or 6 x 14 in K3T mode (the tile computes a 1-voxel border it does not own)
    int stage_bytes;           // bytes of one A ring stage: 5760 (one chunk pair), K3T: all chunk groups of the plane
    int cin_chunks;            // input chunks of the layer
    int tcout;                 // K3T: real output channels (Cpad holds 9 * tcout columns per plane)
    int tpitch;                // K3T: floats per row of the epilogue's shared-memory exchange buffer
    int out_z, out_y, out_x;   // extent the tile grid covers (K3/DOWN: output; UP: low-res input y/x, output z)
    int zin_count;    // input planes a tile walks (zi range)
    const uint8_t* wpacked;
    DEpilogue epi;
};

__host__ __device__ __forceinline__ uint32_t pad16(uint32_t n) { return (n + 15u) & ~15u; }

// ------------------------------------------------------------------------------------------------ step tables
// A-descriptor low word of step `st`, relative to the stage base and the parity base:
//   (halo offset in 16-byte units) | (LBO in 16-byte units) << 16.        These are compile-time constants.
enum StepKind { kK3Full = 0, kK3Lone = 1, kS2Full = 2, kS2Lone = 3, kT = 4 };

template <int KIND> struct Steps;
template <> struct Steps<kK3Full> {
    static constexpr int n = 9;
    __device__ static constexpr uint32_t delta(int st) {
        return static_cast<uint32_t>((st / 3) * kHX + (st % 3)) | (static_cast<uint32_t>(kChunkBytes >> 4) << 16);
    }
};
template <> struct Steps<kK3Lone> {
    // taps (dy,0)+(dy,1) for dy = 0..2 ; (0,2)+(1,2) ; (2,1)[zero weights]+(2,2)
    static constexpr int n = 5;
    __device__ static constexpr uint32_t delta(int st) {
        return st < 3 ? (static_cast<uint32_t>(st * kHX) | (1u << 16))
             : st == 3 ? (2u | (static_cast<uint32_t>(kHX) << 16))
                       : (static_cast<uint32_t>(2 * kHX + 1) | (1u << 16));
    }
};
template <> struct Steps<kS2Full> {
    static constexpr int n = 4;
    __device__ static constexpr uint32_t delta(int st) {
        return static_cast<uint32_t>((st / 2) * kHX + (st % 2)) | (static_cast<uint32_t>(kChunkBytes >> 4) << 16);
    }
};
template <> struct Steps<kS2Lone> {
    static constexpr int n = 2;
    __device__ static constexpr uint32_t delta(int st) { return static_cast<uint32_t>(st * kHX) | (1u << 16); }
};
// Parity base (16-byte units) of the stride-2 modes: DOWN parity 1 -> halo offset 0, parity 0 -> 1; UP: parity.
__device__ __forceinline__ uint32_t parity_base(int mode, int pp) {
    if (mode == B200SEG_TC_K3 || mode == B200SEG_TC_K3T) return 0;
    const int py = pp >> 1, px = pp & 1;
    const int sy0 = (mode == B200SEG_TC_DOWN) ? (1 - py) : py;
    const int sx0 = (mode == B200SEG_TC_DOWN) ? (1 - px) : px;
    return static_cast<uint32_t>(sy0 * kHX + sx0);
}

// planes touched by input plane zi: [lo, hi], B row block of plane lo, first plane that is touched for the first time
__device__ __forceinline__ void plane_window(int mode, int TZ, int zi, int& lo, int& hi, int& jlo, int& ft) {
    if (mode == B200SEG_TC_K3 || mode == B200SEG_TC_K3T) {
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
struct alignas(16) PlaneTab {
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

constexpr uint64_t kADescHi = (static_cast<uint64_t>((kHX * 16) >> 4) | (1ull << 14)) << 32;  // SBO 160 B, version 1
constexpr uint64_t kBDescHi = (static_cast<uint64_t>(128 >> 4) | (1ull << 14)) << 32;         // SBO 128 B, version 1

// Issue steps [st0, n) of one plane: descriptor deltas are immediates, the B image advances by bstep16 per step.
// The normal-step MMA blocks of one plane, held in registers (prefetched one plane ahead by the issuer).
struct PlaneRegs {
    MmaBlk k0, k1, k2;
    int nb;
};
__device__ __forceinline__ PlaneRegs load_plane_regs(const PlaneTab* pt) {
    PlaneRegs r;
    const uint4 a = *reinterpret_cast<const uint4*>(&pt->blk[0][0]);
    const uint4 b = *reinterpret_cast<const uint4*>(&pt->blk[0][1]);
    const uint4 c = *reinterpret_cast<const uint4*>(&pt->blk[0][2]);
    r.k0 = MmaBlk{a.x, a.y, a.z, a.w};
    r.k1 = MmaBlk{b.x, b.y, b.z, b.w};
    r.k2 = MmaBlk{c.x, c.y, c.z, c.w};
    r.nb = pt->nblk[0];
    return r;
}

// Steps [ST0, n) of one plane.  The B descriptor low word is (b16 + brow + st*bstep16) | LBO<<16; the sum stays below
// 2^14, so the step offset can be added AFTER the OR -- one value per block, then uniform adds per MMA.
template <int KIND, int ST0>
__device__ __forceinline__ void issue_plane(uint32_t a16, uint32_t b16, uint32_t b_lbo, uint32_t bstep16,
                                            uint32_t tacc, const PlaneRegs& pr) {
    const int nb = pr.nb;
    const MmaBlk k0 = pr.k0;
    const uint32_t d0 = tacc + k0.dcol;
    const uint32_t bl0 = (b16 + k0.brow) | b_lbo;
    if (nb == 1) {
#pragma unroll
        for (int st = ST0; st < Steps<KIND>::n; ++st)
            umma_bf16(d0, kADescHi | static_cast<uint64_t>(a16 + Steps<KIND>::delta(st)),
                      kBDescHi | static_cast<uint64_t>(bl0 + st * bstep16), k0.idesc, 1u);
    } else {
        const MmaBlk k1 = pr.k1;
        const MmaBlk k2 = nb > 2 ? pr.k2 : pr.k1;
        const uint32_t d1 = tacc + k1.dcol, d2 = tacc + k2.dcol;
        const uint32_t bl1 = (b16 + k1.brow) | b_lbo, bl2 = (b16 + k2.brow) | b_lbo;
#pragma unroll
        for (int st = ST0; st < Steps<KIND>::n; ++st) {
            const uint64_t adesc = kADescHi | static_cast<uint64_t>(a16 + Steps<KIND>::delta(st));
            umma_bf16(d0, adesc, kBDescHi | static_cast<uint64_t>(bl0 + st * bstep16), k0.idesc, 1u);
            umma_bf16(d1, adesc, kBDescHi | static_cast<uint64_t>(bl1 + st * bstep16), k1.idesc, 1u);
            if (nb > 2) umma_bf16(d2, adesc, kBDescHi | static_cast<uint64_t>(bl2 + st * bstep16), k2.idesc, 1u);
        }
    }
}

// All planes of one weight image: wait for the plane's halo, issue its steps, release the stage.  The next plane's
// MMA blocks are fetched from shared memory before blocking on the current plane's data.
template <int KIND, int KIND2 = -1>
__device__ __forceinline__ void run_image(bool first_image, int n_planes, int zstep, const PlaneTab* pt,
                                          uint64_t* full_a, uint64_t* empty_a, uint32_t& a_s, uint32_t& a_ph, int na,
                                          uint32_t sA16, uint32_t pbase, uint32_t b16, uint32_t b_lbo,
                                          uint32_t bstep16, uint32_t tacc, uint32_t stage16, uint32_t bimg16) {
    PlaneRegs cur = load_plane_regs(pt);
    for (int j = 0; j < n_planes; ++j) {
        const PlaneTab* ptn = pt + zstep;
        PlaneRegs nxt = cur;
        if (j + 1 < n_planes) nxt = load_plane_regs(ptn);
        const uint32_t s = a_s;
        mbar_wait(&full_a[s], a_ph);
        tc_fence_after();
        if (++a_s == static_cast<uint32_t>(na)) {
            a_s = 0;
            a_ph ^= 1;
        }
        const uint32_t a16 = sA16 + s * stage16 + pbase;
        if (first_image) {
            // first step of the pass: first-touch split (planes seen for the first time overwrite)
            const uint64_t adesc = kADescHi | static_cast<uint64_t>(a16 + Steps<KIND>::delta(0));
            const int nb = pt->nblk[1];
            for (int b = 0; b < nb; ++b) {
                const MmaBlk k = pt->blk[1][b];
                umma_bf16(tacc + k.dcol, adesc, kBDescHi | static_cast<uint64_t>((b16 + k.brow) | b_lbo), k.idesc, k.acc);
            }
            issue_plane<KIND, 1>(a16, b16, b_lbo, bstep16, tacc, cur);
        } else {
            issue_plane<KIND, 0>(a16, b16, b_lbo, bstep16, tacc, cur);
        }
        // super group: the stage holds a second chunk pair, the slot a second weight image -- same plane, same blocks
        if constexpr (KIND2 >= 0)
            issue_plane<KIND2, 0>(a16 + (kAStageBytes >> 4), b16 + bimg16, b_lbo, bstep16, tacc, cur);
        umma_commit(&empty_a[s]);
        cur = nxt;
        pt = ptn;
    }
}

// K3T ("taps in N", tiny Cout): the plane's stage holds every chunk group; step g multiplies chunk group g, unshifted
// (row m of the MMA is halo voxel m), by the image's step g.  A lone last chunk pairs with its x-neighbour (LBO = 16 B)
// against zero weights.  One MMA per group and column block: the per-plane overhead is paid once for the whole plane.
__device__ __forceinline__ void run_image_t(int G, bool lone_last, int n_planes, const PlaneTab* pt, uint64_t* full_a,
                                            uint64_t* empty_a, uint32_t& a_s, uint32_t& a_ph, int na, uint32_t sA16,
                                            uint32_t b16, uint32_t b_lbo, uint32_t bstep16, uint32_t tacc,
                                            uint32_t stage16) {
    constexpr int kMaxG = 6;     // host limit: at most 12 input chunks
    for (int j = 0; j < n_planes; ++j, ++pt) {
        // MMA blocks of this plane, fetched before blocking on its data: [1] = step 0 (first-touch split), [0] = the rest
        const PlaneRegs norm = load_plane_regs(pt);
        const uint4 f0 = *reinterpret_cast<const uint4*>(&pt->blk[1][0]);
        const uint4 f1 = *reinterpret_cast<const uint4*>(&pt->blk[1][1]);
        const uint4 f2 = *reinterpret_cast<const uint4*>(&pt->blk[1][2]);
        const int nb1 = pt->nblk[1];
        const uint32_t s = a_s;
        mbar_wait(&full_a[s], a_ph);
        tc_fence_after();
        if (++a_s == static_cast<uint32_t>(na)) {
            a_s = 0;
            a_ph ^= 1;
        }
        const uint32_t a16 = sA16 + s * stage16;
        const uint32_t lbo_full = static_cast<uint32_t>(kChunkBytes >> 4) << 16;
        {
            const uint64_t adesc = kADescHi | static_cast<uint64_t>(a16 | ((lone_last && G == 1) ? (1u << 16) : lbo_full));
            umma_bf16(tacc + f0.x, adesc, kBDescHi | static_cast<uint64_t>((b16 + f0.z) | b_lbo), f0.y, f0.w);
            if (nb1 > 1) umma_bf16(tacc + f1.x, adesc, kBDescHi | static_cast<uint64_t>((b16 + f1.z) | b_lbo), f1.y, f1.w);
            if (nb1 > 2) umma_bf16(tacc + f2.x, adesc, kBDescHi | static_cast<uint64_t>((b16 + f2.z) | b_lbo), f2.y, f2.w);
        }
#pragma unroll
        for (int g = 1; g < kMaxG; ++g) {
            if (g < G) {
                const uint32_t lbo = (lone_last && g == G - 1) ? (1u << 16) : lbo_full;
                const uint64_t adesc = kADescHi | static_cast<uint64_t>((a16 + g * (kAStageBytes >> 4)) | lbo);
                const uint32_t bg = b16 + g * bstep16;
                umma_bf16(tacc + norm.k0.dcol, adesc, kBDescHi | static_cast<uint64_t>((bg + norm.k0.brow) | b_lbo),
                          norm.k0.idesc, 1u);
                if (norm.nb > 1)
                    umma_bf16(tacc + norm.k1.dcol, adesc, kBDescHi | static_cast<uint64_t>((bg + norm.k1.brow) | b_lbo),
                              norm.k1.idesc, 1u);
                if (norm.nb > 2)
                    umma_bf16(tacc + norm.k2.dcol, adesc, kBDescHi | static_cast<uint64_t>((bg + norm.k2.brow) | b_lbo),
                              norm.k2.idesc, 1u);
            }
        }
        umma_commit(&empty_a[s]);
    }
}

// ------------------------------------------------------------------------------------------------ epilogue
// One epilogue warp owns 32 accumulator rows (TMEM lanes) and, per output plane, NCH consecutive 8-channel chunks
// starting at chunk `cbeg`.  The chunk count is a template parameter so that the per-chunk code is straight-line:
// one tcgen05.ld, six 128-bit parameter loads at immediate offsets, 8 x (FFMA, FMUL, FMNMX), 4 packs, one 16-byte
// store.  A plane's chunks form one item (NCH <= 3) or two (4 -> 2+2, 5 -> 3+2); the TMEM / residual loads of the
// next item are in flight while the current one is converted and stored (two register buffers).
struct EpiArgs {
    uint32_t tbase;        // TMEM address of (lane quarter, accumulator set, column 0)
    int Cpad;              // TMEM columns per plane
    int nq;                // planes of this unit that exist in the output
    bool valid;            // this lane's (y, x) lies inside the output
    bool slope01;
    uint4* p0;             // dst0 at (chunk 0, plane 0, this lane's voxel)
    uint4* p1;             // dst1, pre-offset by -split_c8 chunks
    const uint4* pr;       // residual
    uint32_t cs;           // chunk stride in 16-byte vectors
    uint32_t plane;        // plane stride in 16-byte vectors
    int split_c8;          // chunks below go to dst0 (and receive the residual), the others to dst1
    const float* s_par;    // shared: per chunk 8 scales, 8 shifts, 8 slopes
};

template <int S, bool RES>
__device__ __forceinline__ void epi_load(const EpiArgs& a, int q, int cc0, uint32_t (&r)[3][8], uint4 (&res)[3]) {
    const uint32_t taddr = a.tbase + q * a.Cpad + cc0 * 8;
#pragma unroll
    for (int j = 0; j < S; ++j) tmem_ld8(taddr + j * 8, r[j]);
    if (RES) {
        if (a.valid) {
            const uint4* prq = a.pr + static_cast<size_t>(static_cast<uint32_t>(q) * a.plane);
#pragma unroll
            for (int j = 0; j < S; ++j)
                if (cc0 + j < a.split_c8) res[j] = __ldg(prq + static_cast<size_t>((cc0 + j) * a.cs));
        }
    }
}

template <int S, bool RES, bool ID>
__device__ __forceinline__ void epi_finish(const EpiArgs& a, int q, int cc0, const uint32_t (&r)[3][8],
                                           const uint4 (&res)[3]) {
    if (!a.valid) return;
    const size_t qoff = static_cast<size_t>(static_cast<uint32_t>(q) * a.plane);
    if constexpr (ID) {
        // identity epilogue (scale 1, shift 0, no activation, no residual, one destination): the blur convolutions
        // of the stride-2 / transposed layers.  tcgen05.ld -> 4 packs -> one 16-byte store per chunk.
#pragma unroll
        for (int j = 0; j < S; ++j) {
            uint4 o;
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                oh[k] = __floats2bfloat162_rn(__uint_as_float(r[j][2 * k]), __uint_as_float(r[j][2 * k + 1]));
            a.p0[qoff + static_cast<size_t>((cc0 + j) * a.cs)] = o;
        }
        return;
    }
    const float4* par = reinterpret_cast<const float4*>(a.s_par + cc0 * 24);
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const int cc = cc0 + j;
        const float4 s0 = par[6 * j], s1 = par[6 * j + 1], h0 = par[6 * j + 2], h1 = par[6 * j + 3],
                     l0 = par[6 * j + 4], l1 = par[6 * j + 5];
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float sl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        // affine in packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two exact fp32 operations per instruction)
        f32x2 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            t[k] = fma2(pack2(__uint_as_float(r[j][2 * k]), __uint_as_float(r[j][2 * k + 1])),
                        pack2(sc[2 * k], sc[2 * k + 1]), pack2(sh[2 * k], sh[2 * k + 1]));
        float v[8];
        if (a.slope01) {
            // 0 <= slope <= 1:  v > 0 ? v : v*slope  ==  max(v, v*slope)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const f32x2 u = mul2(t[k], pack2(sl[2 * k], sl[2 * k + 1]));
                float t0, t1, u0, u1;
                unpack2(t[k], t0, t1);
                unpack2(u, u0, u1);
                t[k] = pack2(fmaxf(t0, u0), fmaxf(t1, u1));
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t0, t1;
                unpack2(t[k], t0, t1);
                t[k] = pack2(t0 > 0.f ? t0 : t0 * sl[2 * k], t1 > 0.f ? t1 : t1 * sl[2 * k + 1]);
            }
        }
        const bool to0 = cc < a.split_c8;
        if (RES) {
            if (to0) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&res[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __bfloat1622float2(h[k]);
                    t[k] = add2(t[k], pack2(f.x, f.y));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack2(t[k], v[2 * k], v[2 * k + 1]);
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        (to0 ? a.p0 : a.p1)[qoff + static_cast<size_t>(cc * a.cs)] = o;
    }
}

// planes q0, q0 + qstep, ... of one accumulator set; chunks [cbeg, cbeg + NCH)
template <int NCH, bool RES, bool ID = false>
__device__ __forceinline__ void epi_planes(const EpiArgs& a, int cbeg, int q0, int qstep) {
    constexpr int S0 = NCH <= 3 ? NCH : (NCH + 1) / 2;
    constexpr int S1 = NCH - S0;
    uint32_t rA[3][8], rB[3][8];
    uint4 resA[3], resB[3];
    int q = q0;
    if (q >= a.nq) return;
    epi_load<S0, RES>(a, q, cbeg, rA, resA);
    if (S1 > 0) {
        constexpr int S1c = S1 > 0 ? S1 : 1;
        for (;;) {
            tmem_ld_wait();
            epi_load<S1c, RES>(a, q, cbeg + S0, rB, resB);
            epi_finish<S0, RES, ID>(a, q, cbeg, rA, resA);
            const int qn = q + qstep;
            tmem_ld_wait();
            if (qn < a.nq) epi_load<S0, RES>(a, qn, cbeg, rA, resA);
            epi_finish<S1c, RES, ID>(a, q, cbeg + S0, rB, resB);
            if (qn >= a.nq) break;
            q = qn;
        }
    } else {
        for (;;) {
            tmem_ld_wait();
            int qn = q + qstep;
            if (qn < a.nq) epi_load<S0, RES>(a, qn, cbeg, rB, resB);
            epi_finish<S0, RES, ID>(a, q, cbeg, rA, resA);
            if (qn >= a.nq) break;
            q = qn;
            qn = q + qstep;
            tmem_ld_wait();
            if (qn < a.nq) epi_load<S0, RES>(a, qn, cbeg, rA, resA);
            epi_finish<S0, RES, ID>(a, q, cbeg, rB, resB);
            if (qn >= a.nq) break;
            q = qn;
        }
    }
}

// kEpi selects the epilogue: 0 = any chunk count (run-time item cursor) and the plain final layer; 1 / 2 = 10 chunks
// without / with residual; 3 / 4 = 5 chunks without / with residual; 5 = K3T final layer.
template <int kSets, int kEpi>
__global__ void __launch_bounds__(kSets == 2 ? 384 : 224, kSets == 2 ? 1 : 2)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams p) {
    constexpr int kEpiWarp0 = kSets == 2 ? 4 : 3;
    constexpr int kEpiWarps = kSets == 2 ? 8 : 4;
    constexpr uint32_t kTmemCols = kSets == 2 ? 512 : 256;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                               // na x 5760
    uint8_t* sB = smem + p.na * p.stage_bytes;        // nbuf x bbuf_bytes
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + p.nbuf * p.bbuf_bytes);
    uint64_t* full_a = bars;                 // [kMaxA]
    uint64_t* empty_a = bars + kMaxA;        // [kMaxA]
    uint64_t* full_b = bars + 2 * kMaxA;     // [kMaxB]
    uint64_t* empty_b = full_b + kMaxB;      // [kMaxB]
    uint64_t* acc_full = empty_b + kMaxB;    // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    // 16-byte aligned tables; offsets are computed on the byte OFFSET (smem is 1024-byte aligned) so that the
    // pointers keep their shared-memory address space (LDS, not generic loads)
    const uint32_t bars_off = static_cast<uint32_t>(p.na * p.stage_bytes + p.nbuf * p.bbuf_bytes);
    const uint32_t tab_off = (bars_off + (2 * kMaxA + 2 * kMaxB + 4) * 8 + 8 + 15u) & ~15u;
    PlaneTab* plane_tab = reinterpret_cast<PlaneTab*>(smem + tab_off);                  // kMaxZin entries
    // epilogue parameters, per 8-channel chunk: 8 scales, 8 shifts, 8 slopes (float4 reads)
    float* s_par = reinterpret_cast<float*>(smem + tab_off + sizeof(PlaneTab) * kMaxZin);
    // K3T: per half of the epilogue warps, a 128-row exchange buffer for the tap products of one plane
    float* t_xbuf = s_par + 3 * p.Cpad;

    const int warp = threadIdx.x / 32;
    const int lane = threadIdx.x % 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kMaxA; ++i) {
            mbar_init(&full_a[i], 1);
            mbar_init(&empty_a[i], 1);
        }
        for (int i = 0; i < kMaxB; ++i) {
            mbar_init(&full_b[i], 1);
            mbar_init(&empty_b[i], (kSets == 2 && p.dual) ? 2 : 1);   // a streamed image is released by both issuers
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], kEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    for (int zi = threadIdx.x; zi < p.zin_count; zi += blockDim.x) build_plane_tab(p, zi, plane_tab[zi]);
    const int n_par = p.mode == B200SEG_TC_K3T ? 8 : p.Cpad;   // K3T: Cpad counts tap columns, not channels
    for (int c = threadIdx.x; c < n_par; c += blockDim.x) {
        float* par = s_par + (c >> 3) * 24 + (c & 7);
        par[0] = p.epi.scale[c];
        par[8] = p.epi.shift[c];
        par[16] = p.epi.slope[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_tiles = p.tiles_x * p.tiles_y * p.tiles_z * p.batch;
    const int n_img_planes = p.mode == B200SEG_TC_DOWN ? p.TZ + 1 : p.zin_count;

    // Unit walk.  A unit is (tile, pass); the CTA owns tiles bid, bid + G, ...  Units are numbered q = 0, 1, ...; unit q
    // uses accumulator set q & 1 with barrier phase (q >> 1) & 1.
    //   single issuer: q = tile_it * n_pass + pass.
    //   dual issuers : q = 2 * (k * n_pass + pass) + r  is (tile 2k + r, pass): issuer r (warp 1 / warp 3) owns set r, the
    //                  odd/even tiles and half r of the A ring (fed by warp 0 / warp 2); both issuers walk the same
    //                  (pass, image) sequence, so a streamed weight image is shared and released by both.
    // A single thread issuing MMAs spends ~100 instructions per plane and is the bottleneck of every layer whose MMAs are
    // short; two issuers on different accumulators double that rate and keep the results deterministic (each accumulator
    // still sees its MMAs in one fixed order).
    const int dual = kSets == 2 ? p.dual : 0;
    const int ring_n = dual ? p.na / 2 : p.na;
    const int bid = static_cast<int>(blockIdx.x), grid_n = static_cast<int>(gridDim.x);
    const int my_tiles = (n_tiles - bid + grid_n - 1) / grid_n;
    const int n_seq = dual ? ((my_tiles + 1) / 2) * 2 * p.n_pass : my_tiles * p.n_pass;
    auto unit_of = [&](int q, int& tile, int& pass) -> bool {
        int tile_it;
        if (dual) {
            const int group = q >> 1, k = group / p.n_pass;
            pass = group - k * p.n_pass;
            tile_it = 2 * k + (q & 1);
        } else {
            tile_it = q / p.n_pass;
            pass = q - tile_it * p.n_pass;
        }
        tile = bid + tile_it * grid_n;
        return tile_it < my_tiles;
    };
    const int n_img_round = p.n_pass * p.n_simg;
    // A-plane load request of super image si of a tile: everything that is constant over the planes of the image is
    // computed here, once, so that a stage costs its producer thread a handful of instructions (the thread that feeds
    // ring 1 also streams the weights and must stay cheap)
    struct ALoad {
        int x0, y0, zc, zstep;
        int cnt, cc0;                 // chunk pairs of the stage, first chunk index (incl. sample offset)
        uint32_t bytes;               // expected bytes of the stage
        const CUtensorMap* map_full;  // box of a full chunk pair
        const CUtensorMap* map_last;  // box of the last pair (a 1-chunk box when the pair is lone)
    };
    auto image_loads = [&](int tile, int si, ALoad& a) {
        int t = tile;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y; t /= p.tiles_y;
        const int tz = t % p.tiles_z;
        const int n = t / p.tiles_z;
        a.x0 = tx * p.tile_sx;
        a.y0 = ty * p.tile_sy;
        const int z0 = tz * p.TZ;
        const int outer = si / p.SG;
        const int sg = si - outer * p.SG;
        a.zstep = 1;
        if (p.mode == B200SEG_TC_K3T) {
            a.zc = z0 - 1;
            a.cnt = 1;
            a.cc0 = n * p.c8_total + p.chunk_base;
            a.bytes = static_cast<uint32_t>(p.cin_chunks) * kChunkBytes;   // one box with every chunk of the plane
            a.map_full = a.map_last = &maps.m[2];
            return;
        }
        const int g0 = p.gp * sg;
        a.cnt = min(p.gp, p.G - g0);
        const bool last_lone = p.lone_last && g0 + a.cnt == p.G;
        a.cc0 = n * p.c8_total + p.chunk_base + 2 * g0;
        a.bytes = static_cast<uint32_t>(a.cnt * kAStageBytes - (last_lone ? kChunkBytes : 0));
        if (p.mode == B200SEG_TC_K3) {
            a.zc = z0 - 1;
            a.map_full = &maps.m[0];
            a.map_last = &maps.m[last_lone ? 1 : 0];
        } else if (p.mode == B200SEG_TC_UP) {
            a.zc = z0 / 2 - 1;        // tile origin is in low-res input coordinates; z0 counts OUTPUT planes
            a.map_full = &maps.m[0];
            a.map_last = &maps.m[last_lone ? 1 : 0];
        } else {
            a.zc = 2 * z0 - 1 + outer / 4;
            a.zstep = 2;
            a.map_full = &maps.m[outer % 4];
            a.map_last = &maps.m[(last_lone ? 4 : 0) + outer % 4];
        }
    };
    // one stage = the chunk pairs of input plane zc: one TMA box per pair, one barrier
    auto issue_a = [&](uint8_t* dst, uint64_t* bar, const ALoad& a, int zc) {
        mbar_arrive_expect_tx(bar, a.bytes);
        if (p.mode == B200SEG_TC_DOWN) {
            if (a.cnt == 2) tma_load_5d(dst, a.map_full, bar, 0, a.x0 - 1, a.y0 - 1, zc, a.cc0);
            tma_load_5d(dst + (a.cnt - 1) * kAStageBytes, a.map_last, bar, 0, a.x0 - 1, a.y0 - 1, zc,
                        a.cc0 + 2 * (a.cnt - 1));
        } else {
            if (a.cnt == 2) tma_load_4d(dst, a.map_full, bar, (a.x0 - 1) * 8, a.y0 - 1, zc, a.cc0);
            tma_load_4d(dst + (a.cnt - 1) * kAStageBytes, a.map_last, bar, (a.x0 - 1) * 8, a.y0 - 1, zc,
                        a.cc0 + 2 * (a.cnt - 1));
        }
    };
    // super image idx of the round (pass, outer, sg) = the weight images [first, first + cnt) of wpacked, one barrier
    auto issue_b = [&](int idx, uint32_t slot) {
        const int pass = idx / p.n_simg, si = idx - pass * p.n_simg;
        const int outer = si / p.SG, sg = si - outer * p.SG;
        const int g0 = p.gp * sg;
        const int cnt = min(p.gp, p.G - g0);
        const bool last_lone = p.lone_last && g0 + cnt == p.G;
        const uint32_t full_bytes = static_cast<uint32_t>(p.steps_full) * 2 * p.NB * 16;
        const uint32_t lone_bytes = static_cast<uint32_t>(p.steps_lone) * 2 * p.NB * 16;
        const int first = pass * p.n_bimg + outer * p.G + g0;
        mbar_arrive_expect_tx(&full_b[slot], (cnt - 1) * full_bytes + (last_lone ? lone_bytes : full_bytes));
        for (int q = 0; q < cnt; ++q)
            bulk_load(sB + slot * p.bbuf_bytes + q * p.bimg_stride,
                      p.wpacked + static_cast<size_t>(first + q) * p.bimg_stride,
                      (last_lone && q == cnt - 1) ? lone_bytes : full_bytes, &full_b[slot]);
    };
    if (warp == 0) {
        // =============================================================== A producer (ring 0)
        if (elect_one()) {
            const int nmap = p.mode == B200SEG_TC_DOWN ? 8 : (p.mode == B200SEG_TC_K3T ? 3 : 2);
            for (int i = 0; i < nmap; ++i) prefetch_tmap(&maps.m[i]);
            uint32_t s = 0, ph = 0;
            for (int q = 0; q < n_seq; q += dual ? 2 : 1) {
                int tile, pass;
                unit_of(q, tile, pass);          // ring 0 units always exist
                for (int si = 0; si < p.n_simg; ++si) {
                    ALoad al;
                    image_loads(tile, si, al);
                    int zc = al.zc;
                    for (int j = 0; j < n_img_planes; ++j, zc += al.zstep) {
                        mbar_wait(&empty_a[s], ph ^ 1);
                        issue_a(sA + s * p.stage_bytes, &full_a[s], al, zc);
                        if (++s == static_cast<uint32_t>(ring_n)) {
                            s = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 2) {
        if (elect_one()) {
            if (!dual) {
                // =========================================================== B producer (blocking)
                // resident mode: every image has its own slot, loaded once; otherwise a ring re-streamed per unit
                const int total = p.resident ? n_img_round : n_seq * p.n_simg;
                uint32_t s = 0, ph = 0;
                int idx = 0;
                for (int it = 0; it < total; ++it) {
                    mbar_wait(&empty_b[s], ph ^ 1);
                    issue_b(idx, s);
                    if (++idx == n_img_round) idx = 0;
                    if (++s == static_cast<uint32_t>(p.nbuf)) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            } else {
                // =========================================================== B producer + A producer of ring 1
                // One thread serves two queues, so it must never block on either: both are polled (test_wait).  Blocking
                // on a full A ring while issuer 1 waits for a weight image that only this thread can load would deadlock.
                const int b_total = p.resident ? n_img_round : (n_seq / 2) * p.n_simg;
                int b_next = 0, b_idx = 0;
                uint32_t b_s = 0, b_ph = 0;
                uint64_t* const full_r = full_a + ring_n;
                uint64_t* const empty_r = empty_a + ring_n;
                uint8_t* const sA_r = sA + ring_n * p.stage_bytes;
                uint32_t s = 0, ph = 0;
                int q = 1, bi = 0, j = 0, tile = 0, pass = 0;
                int zc = 0;
                ALoad al{};
                bool a_done = true;
                // position the A cursor on the first existing unit of ring 1
                auto seek = [&]() {
                    a_done = true;
                    for (; q < n_seq; q += 2) {
                        if (unit_of(q, tile, pass)) {
                            a_done = false;
                            bi = 0;
                            j = 0;
                            image_loads(tile, 0, al);
                            zc = al.zc;
                            break;
                        }
                    }
                };
                seek();
                while (!a_done || b_next < b_total) {
                    bool progress = false;
                    if (b_next < b_total && (p.resident || mbar_test(&empty_b[b_s], b_ph ^ 1))) {
                        issue_b(b_idx, b_s);
                        ++b_next;
                        if (++b_idx == n_img_round) b_idx = 0;
                        if (++b_s == static_cast<uint32_t>(p.nbuf)) {
                            b_s = 0;
                            b_ph ^= 1;
                        }
                        progress = true;
                    }
                    if (!a_done && mbar_test(&empty_r[s], ph ^ 1)) {
                        issue_a(sA_r + s * p.stage_bytes, &full_r[s], al, zc);
                        if (++s == static_cast<uint32_t>(ring_n)) {
                            s = 0;
                            ph ^= 1;
                        }
                        zc += al.zstep;
                        if (++j == n_img_planes) {
                            j = 0;
                            if (++bi == p.n_simg) {
                                q += 2;
                                seek();
                            } else {
                                image_loads(tile, bi, al);
                                zc = al.zc;
                            }
                        }
                        progress = true;
                    }
                    if (!progress) __nanosleep(40);
                }
            }
        }
    } else if (warp == 1 || (warp == 3 && dual)) {
        // =============================================================== MMA issuer r (accumulator set r in dual mode)
        const int r = warp >> 1;
        if (elect_one()) {
            // everything the loop needs lives in registers: no divisions, no parameter re-loads per MMA
            const int mode = p.mode, na = ring_n, nbuf = p.nbuf, n_simg = p.n_simg, G = p.G, gp = p.gp, SG = p.SG;
            const bool lone_last = p.lone_last != 0, resident = p.resident != 0;
            uint32_t a_s = 0, a_ph = 0, b_s = 0, b_ph = 0;
            uint64_t* const full_r = full_a + r * ring_n;
            uint64_t* const empty_r = empty_a + r * ring_n;
            const uint32_t sA16 = smem_u32(sA + r * ring_n * p.stage_bytes) >> 4, sB16 = smem_u32(sB) >> 4;
            const uint32_t stage16 = static_cast<uint32_t>(p.stage_bytes) >> 4;
            const uint32_t bbuf16 = static_cast<uint32_t>(p.bbuf_bytes) >> 4;
            const uint32_t b_lbo = static_cast<uint32_t>(p.NB) << 16;   // LBO = NB * 16 bytes
            const uint32_t bstep16 = 2 * p.NB;                          // one step of a B image, in 16-byte units
            const uint32_t bimg16 = static_cast<uint32_t>(p.bimg_stride) >> 4;   // second image of a super-group slot
            for (int q = dual ? r : 0; q < n_seq; q += dual ? 2 : 1) {
                int tile, pass;
                if (!unit_of(q, tile, pass)) {
                    // the partner issuer has a tile in this round, this one has none: release the streamed weight
                    // images it would have consumed, so that the shared ring keeps turning
                    if (!resident) {
                        for (int bi = 0; bi < n_simg; ++bi) {
                            mbar_wait(&full_b[b_s], b_ph);
                            mbar_arrive(&empty_b[b_s]);
                            if (++b_s == static_cast<uint32_t>(nbuf)) {
                                b_s = 0;
                                b_ph ^= 1;
                            }
                        }
                    }
                    continue;
                }
                {
                    if (resident) b_s = pass * n_simg;     // image slot = (pass, super image)
                    const uint32_t set = kSets == 2 ? (q & 1) : 0;
                    mbar_wait(&acc_empty[set], ((kSets == 2 ? (q >> 1) : q) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tacc = tmem + set * 256;
                    int sg = 0, outer = 0;   // si = outer * SG + sg
                    for (int si = 0; si < n_simg; ++si) {
                        const int g0 = gp * sg;
                        const int cnt = min(gp, G - g0);
                        const bool last_lone = lone_last && g0 + cnt == G && mode != B200SEG_TC_K3T;
                        int zi0 = 0, zstep = 1, pp = 0;
                        if (mode == B200SEG_TC_DOWN) {
                            zi0 = outer >> 2;
                            zstep = 2;
                            pp = outer & 3;
                        } else if (mode == B200SEG_TC_UP) {
                            pp = pass;
                        }
                        if (++sg == SG) {
                            sg = 0;
                            ++outer;
                        }
                        // 0 full | 1 lone | 2 full + full | 3 full + lone   (per visit: one or two chunk pairs)
                        const int combo = cnt == 1 ? (last_lone ? 1 : 0) : (last_lone ? 3 : 2);
                        const int kind = mode == B200SEG_TC_K3T ? 8 : (mode == B200SEG_TC_K3 ? 0 : 4) + combo;
                        const uint32_t pbase = parity_base(mode, pp);
                        const uint32_t bs = b_s;
                        mbar_wait(&full_b[bs], resident ? 0u : b_ph);   // resident images complete phase 0 once
                        tc_fence_after();
                        if (++b_s == static_cast<uint32_t>(nbuf)) {
                            b_s = 0;
                            b_ph ^= 1;
                        }
                        const uint32_t b16 = sB16 + bs * bbuf16;
                        const PlaneTab* pt = plane_tab + zi0;
                        const bool first_image = si == 0;
#define B200SEG_RUN(K, K2)                                                                                             \
    run_image<K, K2>(first_image, n_img_planes, zstep, pt, full_r, empty_r, a_s, a_ph, na, sA16, pbase, b16, b_lbo,     \
                     bstep16, tacc, stage16, bimg16)
                        switch (kind) {
                            case 0: B200SEG_RUN(kK3Full, -1); break;
                            case 1: B200SEG_RUN(kK3Lone, -1); break;
                            case 2: B200SEG_RUN(kK3Full, kK3Full); break;
                            case 3: B200SEG_RUN(kK3Full, kK3Lone); break;
                            case 4: B200SEG_RUN(kS2Full, -1); break;
                            case 5: B200SEG_RUN(kS2Lone, -1); break;
                            case 6: B200SEG_RUN(kS2Full, kS2Full); break;
                            case 7: B200SEG_RUN(kS2Full, kS2Lone); break;
                            default:
                                run_image_t(G, lone_last, n_img_planes, pt, full_r, empty_r, a_s, a_ph, na, sA16, b16, b_lbo,
                                            bstep16, tacc, stage16);
                                break;
                        }
#undef B200SEG_RUN
                        if (!resident) umma_commit(&empty_b[bs]);
                    }
                    umma_commit(&acc_full[set]);
                }
            }
        }
    } else if (warp >= kEpiWarp0) {
        // =============================================================== epilogue
        const int lg = warp & 3;            // TMEM lane quarter this warp may read (= warp id % 4)
        const int half = (warp - kEpiWarp0) >> 2;   // with 8 warps, two share a lane quarter and split the items
        constexpr int kHalves = kEpiWarps / 4;
        const int m = lg * 32 + lane;       // accumulator row
        const int my = m >> 3, mx = m & 7;
        const DEpilogue& e = p.epi;
        const int Cpad = p.Cpad, c8 = p.Cpad / 8, split_c8 = e.split_c8;
        const bool has_res = e.residual.data != nullptr, slope01 = e.slope01 != 0;
        // destination geometry hoisted into registers (dst0 / dst1 / residual share the output's spatial extent)
        const int o_x = e.dst0.x;
        const long long o_plane = static_cast<long long>(e.dst0.y) * e.dst0.x;
        const long long o_cs = e.dst0.chunk_stride;
        uint4* const d0_base = reinterpret_cast<uint4*>(e.dst0.data) + e.dst0.c8_off * o_cs;
        const long long d0_ss = e.dst0.sample_stride;
        uint4* const d1_base = reinterpret_cast<uint4*>(e.dst1.data) + (e.dst1.c8_off - split_c8) * o_cs;
        const long long d1_ss = e.dst1.sample_stride;
        const uint4* const r_base = reinterpret_cast<const uint4*>(e.residual.data) + e.residual.c8_off * o_cs;
        const long long r_ss = e.residual.sample_stride;
        for (int q = 0; q < n_seq; ++q) {
            int tile, pass;
            if (!unit_of(q, tile, pass)) continue;
            int t = tile;
            const int tx = t % p.tiles_x; t /= p.tiles_x;
            const int ty = t % p.tiles_y; t /= p.tiles_y;
            const int tz = t % p.tiles_z;
            const int n = t / p.tiles_z;
            const int x0 = tx * p.tile_sx, y0 = ty * p.tile_sy, z0 = tz * p.TZ;
            {
                const uint32_t unit = static_cast<uint32_t>(q);
                const uint32_t set = kSets == 2 ? (unit & 1) : 0;
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
                if (has_res && valid) {
                    // The residual does not depend on the accumulators: pull this warp's share of it into L2 while the
                    // MMAs of the unit are still running, so that the loads in the item loop see L2 latency, not DRAM.
                    const int nq = min(p.TZ, p.out_z - z0);
                    const uint4* pr = r_base + n * r_ss + static_cast<long long>(z0) * o_plane + oy * o_x + ox;
                    for (int q = half; q < nq; q += kHalves)
                        for (int cc = 0; cc < split_c8; ++cc) prefetch_l2(pr + cc * o_cs + q * o_plane);
                }
                mbar_wait(&acc_full[set], (kSets == 2 ? (unit >> 1) : unit) & 1);
                tc_fence_after();
                const uint32_t tbase = tmem + set * 256 + (static_cast<uint32_t>(lg * 32) << 16);
                if (kEpi == 5) {
                    // Final layer with tiny Cout: accumulator row m holds, for halo voxel m of the tile, the products of
                    // all nine in-plane taps (column (dy*3+dx)*cout + co).  The rows go through shared memory; every
                    // interior voxel sums the nine columns of its nine neighbours, then bias / softmax / fp32 store.
                    const int tcout = p.tcout, pitch = p.tpitch, ncol = 9 * p.tcout, nld = p.Cpad >> 3;
                    float* const xbuf = t_xbuf + half * (128 * pitch);
                    const int yT = y0 - 1 + my, xT = x0 - 1 + mx;
                    const bool mine = my >= 1 && my <= 14 && mx >= 1 && mx <= 6 && yT < p.out_y && xT < p.out_x;
                    const int nq = min(p.TZ, p.out_z - z0);
                    const long long vox = 1LL * p.out_z * p.out_y * p.out_x;
                    if (tcout == 2 && e.softmax) {
                        // the shipped heads (2 classes + softmax): straight-line code, 64-bit shared-memory accesses at
                        // immediate offsets (row pitch 26 floats), two-class softmax with one exponential
                        constexpr int kP = 26;
                        float2* const row2 = reinterpret_cast<float2*>(xbuf + m * kP);
                        const float* const nb = xbuf + (m - 9) * kP;       // neighbour (dy, dx) = (0, 0)
                        const float sc0 = s_par[0], sc1 = s_par[1], sh0 = s_par[8], sh1 = s_par[9];
                        const float sl0 = s_par[16], sl1 = s_par[17];
                        float* dst = e.out_ncdhw + static_cast<long long>(n) * 2 * vox +
                                     (static_cast<long long>(z0) * p.out_y + yT) * p.out_x + xT;
                        const long long plane = static_cast<long long>(p.out_y) * p.out_x;
                        for (int q = half; q < nq; q += kHalves) {
                            uint32_t r[3][8];
#pragma unroll
                            for (int j = 0; j < 3; ++j) tmem_ld8(tbase + q * 24 + j * 8, r[j]);
                            tmem_ld_wait();
#pragma unroll
                            for (int t = 0; t < 9; ++t)
                                row2[t] = make_float2(__uint_as_float(r[(2 * t) >> 3][(2 * t) & 7]),
                                                      __uint_as_float(r[(2 * t + 1) >> 3][(2 * t + 1) & 7]));
                            if (half == 0) named_bar_sync<1>(128); else named_bar_sync<2>(128);
                            if (mine) {
                                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                                    for (int dx = 0; dx < 3; ++dx) {
                                        const float2 t2 = *reinterpret_cast<const float2*>(
                                            nb + (dy * 8 + dx) * kP + (dy * 3 + dx) * 2);
                                        a0 += t2.x;
                                        a1 += t2.y;
                                    }
                                float v0 = fmaf(a0, sc0, sh0), v1 = fmaf(a1, sc1, sh1);
                                v0 = v0 > 0.f ? v0 : v0 * sl0;
                                v1 = v1 > 0.f ? v1 : v1 * sl1;
                                // softmax of two logits: the larger one contributes exp(0) = 1 exactly
                                const float ex = expf(-fabsf(v0 - v1)), sum = 1.f + ex;
                                const float hi = 1.f / sum, lo = ex / sum;
                                const bool first = v0 >= v1;
                                float* d = dst + q * plane;
                                d[0] = first ? hi : lo;
                                d[vox] = first ? lo : hi;
                            }
                            if (half == 0) named_bar_sync<1>(128); else named_bar_sync<2>(128);
                        }
                    } else
                    for (int q = half; q < nq; q += kHalves) {
                        uint32_t r[5][8];
#pragma unroll
                        for (int j = 0; j < 5; ++j)
                            if (j < nld) tmem_ld8(tbase + q * p.Cpad + j * 8, r[j]);
                        tmem_ld_wait();
                        float* const row = xbuf + m * pitch;
#pragma unroll
                        for (int j = 0; j < 5; ++j)
                            if (j < nld) {
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    if (j * 8 + k < ncol) row[j * 8 + k] = __uint_as_float(r[j][k]);
                            }
                        if (half == 0) named_bar_sync<1>(128); else named_bar_sync<2>(128);
                        if (mine) {
                            float v[4];
#pragma unroll
                            for (int co = 0; co < 4; ++co) {
                                if (co < tcout) {
                                    float acc = 0.f;
#pragma unroll
                                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                                        for (int dx = 0; dx < 3; ++dx)
                                            acc += xbuf[((my + dy - 1) * 8 + (mx + dx - 1)) * pitch + (dy * 3 + dx) * tcout + co];
                                    const float tv = fmaf(acc, s_par[co], s_par[8 + co]);
                                    v[co] = tv > 0.f ? tv : tv * s_par[16 + co];
                                }
                            }
                            float* dst = e.out_ncdhw + static_cast<long long>(n) * tcout * vox +
                                         (static_cast<long long>(z0 + q) * p.out_y + yT) * p.out_x + xT;
                            if (e.softmax) {
                                float mx_ = -INFINITY, sum = 0.f;
#pragma unroll
                                for (int co = 0; co < 4; ++co)
                                    if (co < tcout) mx_ = fmaxf(mx_, v[co]);
#pragma unroll
                                for (int co = 0; co < 4; ++co)
                                    if (co < tcout) {
                                        v[co] = expf(v[co] - mx_);
                                        sum += v[co];
                                    }
#pragma unroll
                                for (int co = 0; co < 4; ++co)
                                    if (co < tcout) dst[co * vox] = v[co] / sum;
                            } else {
#pragma unroll
                                for (int co = 0; co < 4; ++co)
                                    if (co < tcout) dst[co * vox] = v[co];
                            }
                        }
                        // the buffer is rewritten by the next plane of this half
                        if (half == 0) named_bar_sync<1>(128); else named_bar_sync<2>(128);
                    }
                } else if (kEpi != 0 || e.out_ncdhw == nullptr) {
                    if constexpr (kEpi != 0 && kEpi != 5) {
                        // specialised epilogues: 10 chunks (the two warps of a lane quarter take 5 chunks each of
                        // every plane) or 5 chunks (they take alternate planes), with or without a residual;
                        // 6 / 7: 5 / 10 chunks with the identity epilogue
                        constexpr bool kRes = kEpi == 2 || kEpi == 4;
                        constexpr bool kByPlane = kEpi == 3 || kEpi == 4 || kEpi == 6;
                        constexpr bool kId = kEpi == 6 || kEpi == 7;
                        EpiArgs a;
                        a.tbase = tbase;
                        a.Cpad = Cpad;
                        a.nq = min(p.TZ, p.out_z - z0);
                        a.valid = valid;
                        a.slope01 = slope01;
                        const long long spatial = static_cast<long long>(z0) * o_plane + oy * o_x + ox;
                        a.p0 = d0_base + n * d0_ss + spatial;
                        a.p1 = d1_base + n * d1_ss + spatial;          // already offset by -split_c8 chunks
                        a.pr = r_base + n * r_ss + spatial;
                        a.cs = static_cast<uint32_t>(o_cs);
                        a.plane = static_cast<uint32_t>(o_plane);
                        a.split_c8 = split_c8;
                        a.s_par = s_par;
                        for (int h = kHalves == 2 ? half : 0; h < 2; h += kHalves) {
                            if (kByPlane) epi_planes<5, kRes, kId>(a, 0, h, 2);
                            else epi_planes<5, kRes, kId>(a, 5 * h, 0, 1);
                        }
                    } else {
                        // Work items of this warp: (plane q, round of kR chunks).  Software pipelined: the TMEM and
                        // residual loads of the next item are in flight while the current one is converted and stored.
                        constexpr int kR = 3;
                        const int nq = min(p.TZ, p.out_z - z0);
                        const int rounds = (c8 + kR - 1) / kR;
                        // per-unit base pointers (16-byte vectors): element (chunk cc, plane q) = base[cc * cs + q * plane]
                        const long long spatial = static_cast<long long>(z0) * o_plane + oy * o_x + ox;
                        uint4* const p0 = d0_base + n * d0_ss + spatial;
                        uint4* const p1 = d1_base + n * d1_ss + spatial;          // already offset by -split_c8 chunks
                        const uint4* const pr = r_base + n * r_ss + spatial;
                        auto load_item = [&](int q, int c0, uint32_t (&r)[kR][8], uint4 (&res)[kR]) {
                            const uint32_t taddr = tbase + q * Cpad;
    #pragma unroll
                            for (int j = 0; j < kR; ++j)
                                if (c0 + j < c8) tmem_ld8(taddr + (c0 + j) * 8, r[j]);
                            if (valid && has_res) {
    #pragma unroll
                                for (int j = 0; j < kR; ++j)
                                    if (c0 + j < c8 && c0 + j < split_c8)
                                        res[j] = __ldg(pr + (c0 + j) * o_cs + static_cast<long long>(q) * o_plane);
                            }
                        };
                        auto finish_item = [&](int q, int c0, uint32_t (&r)[kR][8], uint4 (&res)[kR]) {
                            if (!valid) return;
                            const long long qoff = static_cast<long long>(q) * o_plane;
    #pragma unroll
                            for (int j = 0; j < kR; ++j) {
                                const int cc = c0 + j;
                                if (cc >= c8) continue;
                                float v[8];
                                {
                                    // per-channel parameters: 6 x 128-bit broadcast loads per chunk
                                    const float4* ps = reinterpret_cast<const float4*>(s_par + cc * 24);
                                    const float4* ph = ps + 2;
                                    const float4* pl = ps + 4;
                                    const float4 s0 = ps[0], s1 = ps[1], h0 = ph[0], h1 = ph[1], l0 = pl[0], l1 = pl[1];
                                    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                                    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                                    const float sl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                                    if (slope01) {
                                        // 0 <= slope <= 1:  v > 0 ? v : v*slope  ==  max(v, v*slope)
    #pragma unroll
                                        for (int k = 0; k < 8; ++k) {
                                            float tv = fmaf(__uint_as_float(r[j][k]), sc[k], sh[k]);
                                            v[k] = fmaxf(tv, tv * sl[k]);
                                        }
                                    } else {
    #pragma unroll
                                        for (int k = 0; k < 8; ++k) {
                                            float tv = fmaf(__uint_as_float(r[j][k]), sc[k], sh[k]);
                                            v[k] = tv > 0.f ? tv : tv * sl[k];
                                        }
                                    }
                                }
                                const bool to0 = cc < split_c8;
                                if (to0 && has_res) {
                                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&res[j]);
    #pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        float2 f = __bfloat1622float2(h[k]);
                                        v[2 * k] += f.x;
                                        v[2 * k + 1] += f.y;
                                    }
                                }
                                uint4 o;
                                __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
    #pragma unroll
                                for (int k = 0; k < 4; ++k) oh[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                                (to0 ? p0 : p1)[cc * o_cs + qoff] = o;
                            }
                        };
                        // item cursor (q, rr): this warp takes every kHalves-th item of the (q-major) item list
                        auto advance = [&](int& q, int& rr) {
                            rr += kHalves;
                            while (rr >= rounds) {
                                rr -= rounds;
                                ++q;
                            }
                        };
                        uint32_t rA[kR][8], rB[kR][8];
                        uint4 resA[kR], resB[kR];
                        int q = 0, rr = kHalves == 2 ? half : 0;
                        while (rr >= rounds) {
                            rr -= rounds;
                            ++q;
                        }
                        if (q < nq) load_item(q, rr * kR, rA, resA);
                        while (q < nq) {
                            tmem_ld_wait();
                            int q2 = q, rr2 = rr;
                            advance(q2, rr2);
                            if (q2 < nq) load_item(q2, rr2 * kR, rB, resB);
                            finish_item(q, rr * kR, rA, resA);
                            q = q2;
                            rr = rr2;
                            if (q >= nq) break;
                            tmem_ld_wait();
                            advance(q2, rr2);
                            if (q2 < nq) load_item(q2, rr2 * kR, rA, resA);
                            finish_item(q, rr * kR, rB, resB);
                            q = q2;
                            rr = rr2;
                        }
                    }
                } else {
                    // final layer: affine (+bias), optional channel softmax, fp32 NCDHW store (cout <= 16)
                    for (int q = half; q < p.TZ; q += kHalves) {
                        const int oz = z0 + q;
                        if (oz >= p.out_z) break;
                        const uint32_t taddr = tbase + q * p.Cpad;
                        uint32_t r0[8], r1[8];
                        tmem_ld8(taddr, r0);
                        if (c8 > 1) tmem_ld8(taddr + 8, r1);
                        tmem_ld_wait();
                        if (valid) {
                            float v[16];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float tv = fmaf(__uint_as_float(r0[j]), s_par[j], s_par[8 + j]);
                                v[j] = tv > 0.f ? tv : tv * s_par[16 + j];
                            }
                            if (c8 > 1) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    float tv = fmaf(__uint_as_float(r1[j]), s_par[24 + j], s_par[32 + j]);
                                    v[8 + j] = tv > 0.f ? tv : tv * s_par[40 + j];
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
                                float sum = 0.f;
#pragma unroll
                                for (int c = 0; c < 16; ++c)
                                    if (c < e.cout) {
                                        v[c] = expf(v[c] - mx_);
                                        sum += v[c];
                                    }
#pragma unroll
                                for (int c = 0; c < 16; ++c)
                                    if (c < e.cout) dst[c * vox] = v[c] / sum;
                            } else {
#pragma unroll
                                for (int c = 0; c < 16; ++c)
                                    if (c < e.cout) dst[c * vox] = v[c];
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[set]);
            }
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
    B200SEG_CHECK_ARG(mode >= 0 && mode <= 3, "conv3d_tc: bad mode %d", mode);
    B200SEG_CHECK_ARG(cout >= 1 && cout <= 120, "conv3d_tc: cout %d not in [1,120] (split wider layers)", cout);
    B200SEG_CHECK_ARG(mode != B200SEG_TC_K3T || cout <= 4, "conv3d_tc K3T: cout %d not in [1,4]", cout);
    B200SEG_CHECK_ARG(mode != B200SEG_TC_K3T || cin_chunks <= 12, "conv3d_tc K3T: at most 96 input channels");
    B200SEG_CHECK_ARG(cin_chunks >= 1, "conv3d_tc: no input chunks");
    // K3T: the 9 in-plane taps are columns of the accumulator (9 * cout per plane), summed in the epilogue
    g->Cpad = mode == B200SEG_TC_K3T ? (9 * cout + 7) / 8 * 8 : (cout + 7) / 8 * 8;
    g->blocks = (mode == B200SEG_TC_K3 || mode == B200SEG_TC_K3T) ? 3 : (mode == B200SEG_TC_DOWN ? 2 : 4);
    g->NB = g->blocks * g->Cpad + 16;
    g->G = (cin_chunks + 1) / 2;
    // K3T: ONE image per unit whose steps are the chunk groups
    g->steps_full = mode == B200SEG_TC_K3 ? 9 : (mode == B200SEG_TC_K3T ? g->G : 4);
    g->steps_lone = mode == B200SEG_TC_K3 ? 5 : (mode == B200SEG_TC_K3T ? g->G : 2);
    g->lone_last = cin_chunks & 1;
    g->n_pass = mode == B200SEG_TC_UP ? 4 : 1;
    g->n_bimg = mode == B200SEG_TC_DOWN ? 8 * g->G : (mode == B200SEG_TC_K3T ? 1 : g->G);
    g->bimg_stride = g->steps_full * 2 * g->NB * 16;
    g->TZmax = (256 - 16) / g->Cpad;   // one accumulator set = 256 TMEM columns
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
    if (mode == B200SEG_TC_K3 || mode == B200SEG_TC_K3T) {
        oz = in.z; oy = in.y; ox = in.x;
    } else if (mode == B200SEG_TC_DOWN) {
        B200SEG_CHECK_ARG(in.z % 2 == 0 && in.y % 2 == 0 && in.x % 2 == 0, "conv3d_tc DOWN: input extent must be even");
        oz = in.z / 2; oy = in.y / 2; ox = in.x / 2;
    } else {
        oz = in.z * 2; oy = in.y * 2; ox = in.x * 2;
    }
    B200SEG_CHECK_ARG(epi->out_ncdhw == nullptr || mode == B200SEG_TC_K3 || mode == B200SEG_TC_K3T,
                      "conv3d_tc: out_ncdhw only in K3 / K3T mode");
    B200SEG_CHECK_ARG(mode != B200SEG_TC_K3T || epi->out_ncdhw != nullptr, "conv3d_tc K3T: needs out_ncdhw (final layer)");
    DEpilogue de;
    rc = make_depilogue(epi, cout, in.n, oz, oy, ox, B200SEG_BF16, &de);
    if (rc) return rc;
    // the epilogue indexes (chunk, plane) inside one sample with 32-bit arithmetic on 16-byte vectors
    B200SEG_CHECK_ARG(epi->out_ncdhw != nullptr ||
                          static_cast<long long>(de.dst0.chunk_stride) * 16 < (1LL << 32),
                      "conv3d_tc: output sample too large (chunk stride %lld voxels)",
                      static_cast<long long>(de.dst0.chunk_stride));

    int dev = 0, sms = 148;
    B200SEG_CHECK_CUDA(cudaGetDevice(&dev));
    B200SEG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // variant 2 (default): one CTA per SM, two accumulator sets; variant 1: two CTAs per SM, one set each
    static const int variant = [] {
        const char* v = getenv("B200SEG_TC_VARIANT");
        return (v && v[0] == '1') ? 1 : 2;
    }();

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
    p.batch = in.n;
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
        // small grids: thinner z tiles until every CTA slot has work (or TZ bottoms out)
        const int plane_tiles = (mode == B200SEG_TC_UP    ? ((in.x + 7) / 8) * ((in.y + 15) / 16)
                                 : mode == B200SEG_TC_K3T ? ((ox + 5) / 6) * ((oy + 13) / 14)
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
    } else if (mode == B200SEG_TC_K3T) {
        // a tile computes the tap products of its 16 x 8 halo-anchored voxels and owns the 14 x 6 interior
        p.tile_sx = 6;
        p.tile_sy = 14;
        p.tiles_x = (ox + 5) / 6;
        p.tiles_y = (oy + 13) / 14;
        p.out_y = oy;
        p.out_x = ox;
        p.zin_count = TZ + 2;
        p.tcout = cout;
        p.tpitch = 9 * cout + 8;
    } else {
        p.tiles_x = (ox + 7) / 8;
        p.tiles_y = (oy + 15) / 16;
        p.out_y = oy;
        p.out_x = ox;
        p.zin_count = mode == B200SEG_TC_K3 ? TZ + 2 : 2 * TZ + 2;
    }
    if (mode != B200SEG_TC_K3T) {
        p.tile_sx = 8;
        p.tile_sy = 16;
    }
    p.out_z = oz;
    B200SEG_CHECK_ARG(p.zin_count <= kMaxZin, "conv3d_tc: tile walks %d input planes (max %d)", p.zin_count, kMaxZin);
    B200SEG_CHECK_ARG(TZ * g.Cpad + 16 <= 256, "conv3d_tc: accumulator set exceeds 256 TMEM columns");
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
        if (mode == B200SEG_TC_K3T) {   // one box with every chunk of the plane
            cuuint32_t box[4] = {kHX * 8, kHY, 1, static_cast<cuuint32_t>(cin_chunks)};
            rc = encode_map(&maps.m[2], 4, in.data, dims, strides, box);
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
    // ---- shared memory: weight image ring (double-buffered when it fits) + A plane ring
    const size_t misc = (2 * kMaxA + 2 * kMaxB + 4) * 8 + 16 + sizeof(PlaneTab) * kMaxZin +
                        3 * static_cast<size_t>(g.Cpad) * 4 + 256 +
                        (mode == B200SEG_TC_K3T ? 2 * 128 * static_cast<size_t>(p.tpitch) * 4 : 0);
    const size_t budget = variant == 2 ? 224 * 1024 : 112 * 1024;
    p.cin_chunks = cin_chunks;
    // Super groups (gp = 2): one visit of the MMA issuer handles TWO chunk pairs of the plane -- the A stage holds both
    // halo boxes, the B slot two consecutive weight images.  Every layer whose MMAs are short is bound by the issuing
    // thread's per-visit work (barrier round trip, ring bookkeeping, descriptor set-up: ~700 cycles for 2-9 MMAs, ncu
    // profiles/r02_ncu_upsampling0_b8_*), so halving the visits is worth the coarser rings.  Taken when the operand
    // still fits resident, or as a ring of >= 2 slots, next to >= 8 of the doubled A stages.
    static const int gp_forced = [] {
        const char* v = getenv("B200SEG_TC_GP");           // A/B hook: 1 = never, 2 = whenever it fits (default)
        return (v && (v[0] == '1' || v[0] == '2')) ? v[0] - '0' : 0;
    }();
    p.gp = 1;
    if (mode != B200SEG_TC_K3T && g.G >= 2 && gp_forced != 1) {
        const size_t bbuf2 = (2 * static_cast<size_t>(g.bimg_stride) + 127) & ~static_cast<size_t>(127);
        const size_t stage2 = 2 * kAStageBytes;
        const int sg2 = (g.G + 1) / 2;
        const int n_img2 = g.n_pass * (g.n_bimg / g.G) * sg2;
        const bool resident2 = n_img2 <= kMaxB && misc + n_img2 * bbuf2 + 8 * stage2 <= budget;
        const bool ring2 = misc + 2 * bbuf2 + 8 * stage2 <= budget;
        if (resident2 || ring2) p.gp = 2;
    }
    p.SG = mode == B200SEG_TC_K3T ? 1 : (g.G + p.gp - 1) / p.gp;
    p.n_simg = mode == B200SEG_TC_K3T ? 1 : (g.n_bimg / g.G) * p.SG;
    p.bbuf_bytes = (p.gp * g.bimg_stride + 127) & ~127;
    p.stage_bytes = mode == B200SEG_TC_K3T ? g.G * kAStageBytes : p.gp * kAStageBytes;
    const size_t stage = static_cast<size_t>(p.stage_bytes);
    // weights stay resident (one slot per image, loaded once per CTA) when the whole packed operand fits next to
    // at least 8 A stages; otherwise the images stream through a ring (double-buffered when possible)
    const int n_images = g.n_pass * p.n_simg;
    p.resident = (n_images <= kMaxB &&
                  misc + static_cast<size_t>(n_images) * p.bbuf_bytes + 8 * stage <= budget) ? 1 : 0;
    p.nbuf = p.resident ? n_images
                        : ((misc + 2 * static_cast<size_t>(p.bbuf_bytes) + 4 * stage <= budget) ? 2 : 1);
    {
        // streamed weight images: a ring of up to 4 slots when at least 8 A stages still fit -- two issuers share
        // every streamed image, and a 2-deep ring keeps them in lock step.  B200SEG_TC_NBUF overrides (A/B hook).
        static const int forced = [] {
            const char* v = getenv("B200SEG_TC_NBUF");
            return v ? atoi(v) : 0;
        }();
        const int want = forced > 0 ? forced : 4;
        if (!p.resident && p.nbuf == 2) {
            for (int nb = want; nb > 2; --nb) {
                if (nb <= kMaxB && nb <= n_images && misc + static_cast<size_t>(nb) * p.bbuf_bytes + 8 * stage <= budget) {
                    p.nbuf = nb;
                    break;
                }
            }
        }
    }
    B200SEG_CHECK_ARG(misc + static_cast<size_t>(p.nbuf) * p.bbuf_bytes + 3 * stage <= budget,
                      "conv3d_tc: weight image of %d bytes does not fit shared memory", p.bbuf_bytes);
    long na = static_cast<long>((budget - misc - static_cast<size_t>(p.nbuf) * p.bbuf_bytes) / stage);
    if (na > kMaxA) na = kMaxA;
    p.na = static_cast<int>(na);
    // two MMA issuers (one per accumulator set, half of the A ring each) when every CTA has at least two tiles
    // Measured on config 2 (profiles/r01_issuer_modes.log): dual issue helps every layer except the transposed convolution
    // with 40 output channels, whose units are epilogue-heavy and lose more to the lock-step on the shared weight ring.
    // B200SEG_TC_ISSUERS (test / profiling hook): 1 = single issuer everywhere, 2 = dual only with resident weights,
    // 3 = dual wherever possible, unset = the default rule.
    static const int issuers = [] {
        const char* v = getenv("B200SEG_TC_ISSUERS");
        return (v && v[0] >= '1' && v[0] <= '4') ? v[0] - '0' : 0;
    }();
    const long long tiles_all = 1LL * p.tiles_x * p.tiles_y * p.tiles_z * in.n;
    const bool dual_ok = variant == 2 && p.na >= 8 && tiles_all >= 2LL * sms;
    // Round 1 excluded the 40-channel transposed convolution (streamed weights) from dual issue; with the identity
    // epilogue and a 4-deep weight ring it now gains 13-17 % from the second issuer (same-box A/B), so the
    // default is dual wherever the tile count allows.  4 = the round-1 rule (A/B hook).
    const bool dual_rule = issuers == 1 ? false
                         : issuers == 2 ? p.resident != 0
                         : issuers == 4 ? (mode != B200SEG_TC_UP || p.resident || g.Cpad >= 80)
                                        : true;
    p.dual = (dual_ok && dual_rule) ? 1 : 0;
    const size_t smem = misc + static_cast<size_t>(p.nbuf) * p.bbuf_bytes + static_cast<size_t>(p.na) * stage;
    // ---- launch (persistent)
    const long long n_tiles = 1LL * p.tiles_x * p.tiles_y * p.tiles_z * in.n;
    const long long max_ctas = 1LL * sms * (variant == 2 ? 1 : 2);
    dim3 grid(static_cast<unsigned>(n_tiles < max_ctas ? n_tiles : max_ctas));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static const bool generic_epilogue = [] {
        const char* v = getenv("B200SEG_TC_GENERIC_EPILOGUE");   // test hook: force the run-time epilogue
        return v && v[0] == '1';
    }();
    void (*kernel)(const TcMaps, const TcParams) = nullptr;
    if (variant == 2) {
        const bool res = de.residual.data != nullptr;
        const int c8 = p.Cpad / 8;
        const bool ident = epi->slope01 == 2 && !res && de.dst1.data == nullptr;
        if (mode == B200SEG_TC_K3T) kernel = conv_tc_kernel<2, 5>;
        else if (epi->out_ncdhw != nullptr || generic_epilogue || (c8 != 10 && c8 != 5)) kernel = conv_tc_kernel<2, 0>;
        else if (ident) kernel = c8 == 10 ? conv_tc_kernel<2, 7> : conv_tc_kernel<2, 6>;
        else if (c8 == 10) kernel = res ? conv_tc_kernel<2, 2> : conv_tc_kernel<2, 1>;
        else kernel = res ? conv_tc_kernel<2, 4> : conv_tc_kernel<2, 3>;
    } else {
        const bool res = de.residual.data != nullptr;
        const int c8 = p.Cpad / 8;
        if (mode == B200SEG_TC_K3T) kernel = conv_tc_kernel<1, 5>;
        else if (epi->out_ncdhw != nullptr || generic_epilogue || (c8 != 10 && c8 != 5)) kernel = conv_tc_kernel<1, 0>;
        else if (c8 == 10) kernel = res ? conv_tc_kernel<1, 2> : conv_tc_kernel<1, 1>;
        else kernel = res ? conv_tc_kernel<1, 4> : conv_tc_kernel<1, 3>;
    }
    B200SEG_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kernel<<<grid, variant == 2 ? 384 : 224, smem, s>>>(maps, p);
    return check_launch("conv3d_tc");
}
