// HBM-bound kernels of the sliding-window path: layout conversion, pooling / trilinear resampling, channel
// softmax, grid patch extraction, overlap-add aggregation, finalize (+argmax) and the confusion histogram.
// All of them are pure streaming kernels: one 16/32-byte vector per thread access, x (the contiguous axis)
// mapped to threadIdx.x, grids sized from the element count.
#include <cstdlib>

#include "common.cuh"

namespace b200seg {

static constexpr int kThreads = 256;
static inline unsigned blocks_for(long long n, int per_block = kThreads) {
    return static_cast<unsigned>((n + per_block - 1) / per_block);
}

// =========================================================================================== pack / unpack
template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_ncdhw_kernel(const float* __restrict__ src, DView dst, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const long long vox = dst.chunk_stride;
    const int c8 = (dst.c + 7) / 8;
    long long v = t % vox;
    int cc = static_cast<int>((t / vox) % c8);
    int n = static_cast<int>(t / (vox * c8));
    Vec8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = cc * 8 + j;
        r.v[j] = c < dst.c ? __ldg(src + (static_cast<long long>(n) * dst.c + c) * vox + v) : 0.f;
    }
    store_vec8<T>(dst.data, n * dst.sample_stride + (dst.c8_off + cc) * dst.chunk_stride + v, r);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
unpack_ncdhw_kernel(DView src, float* __restrict__ dst, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const long long vox = src.chunk_stride;
    const int c8 = (src.c + 7) / 8;
    long long v = t % vox;
    int cc = static_cast<int>((t / vox) % c8);
    int n = static_cast<int>(t / (vox * c8));
    Vec8 r = load_vec8<T>(src.data, n * src.sample_stride + (src.c8_off + cc) * src.chunk_stride + v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = cc * 8 + j;
        if (c < src.c) dst[(static_cast<long long>(n) * src.c + c) * vox + v] = r.v[j];
    }
}

// =========================================================================================== avgpool / upsample / copy
template <typename T>
__global__ void __launch_bounds__(kThreads)
avgpool2_kernel(DView in, DView out, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const int c8 = (out.c + 7) / 8;
    int x = static_cast<int>(t % out.x);
    long long r = t / out.x;
    int y = static_cast<int>(r % out.y);
    r /= out.y;
    int z = static_cast<int>(r % out.z);
    r /= out.z;
    int cc = static_cast<int>(r % c8);
    int n = static_cast<int>(r / c8);
    Vec8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                Vec8 v = load_vec8<T>(in.data, vox_index(in, n, cc, 2 * z + dz, 2 * y + dy, 2 * x + dx));
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] += v.v[j];
            }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = acc.v[j] / 8.0f;
    store_vec8<T>(out.data, vox_index(out, n, cc, z, y, x), acc);
}

__device__ __forceinline__ void lin_coeff(int dst, int in_size, int out_size, int& i0, int& i1, float& l0,
                                          float& l1) {
    // align_corners=True: src = dst * (in-1)/(out-1)
    float scale = out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
    float src = scale * static_cast<float>(dst);
    i0 = static_cast<int>(src);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - static_cast<float>(i0);
    l0 = 1.f - l1;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
upsample_trilinear2_kernel(DView in, DView out, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const int c8 = (out.c + 7) / 8;
    int x = static_cast<int>(t % out.x);
    long long r = t / out.x;
    int y = static_cast<int>(r % out.y);
    r /= out.y;
    int z = static_cast<int>(r % out.z);
    r /= out.z;
    int cc = static_cast<int>(r % c8);
    int n = static_cast<int>(r / c8);
    int z0, z1, y0, y1, x0, x1;
    float lz0, lz1, ly0, ly1, lx0, lx1;
    lin_coeff(z, in.z, out.z, z0, z1, lz0, lz1);
    lin_coeff(y, in.y, out.y, y0, y1, ly0, ly1);
    lin_coeff(x, in.x, out.x, x0, x1, lx0, lx1);
    Vec8 p000 = load_vec8<T>(in.data, vox_index(in, n, cc, z0, y0, x0));
    Vec8 p001 = load_vec8<T>(in.data, vox_index(in, n, cc, z0, y0, x1));
    Vec8 p010 = load_vec8<T>(in.data, vox_index(in, n, cc, z0, y1, x0));
    Vec8 p011 = load_vec8<T>(in.data, vox_index(in, n, cc, z0, y1, x1));
    Vec8 p100 = load_vec8<T>(in.data, vox_index(in, n, cc, z1, y0, x0));
    Vec8 p101 = load_vec8<T>(in.data, vox_index(in, n, cc, z1, y0, x1));
    Vec8 p110 = load_vec8<T>(in.data, vox_index(in, n, cc, z1, y1, x0));
    Vec8 p111 = load_vec8<T>(in.data, vox_index(in, n, cc, z1, y1, x1));
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        o.v[j] = lz0 * (ly0 * (lx0 * p000.v[j] + lx1 * p001.v[j]) + ly1 * (lx0 * p010.v[j] + lx1 * p011.v[j])) +
                 lz1 * (ly0 * (lx0 * p100.v[j] + lx1 * p101.v[j]) + ly1 * (lx0 * p110.v[j] + lx1 * p111.v[j]));
    }
    store_vec8<T>(out.data, vox_index(out, n, cc, z, y, x), o);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
copy_view_kernel(DView in, DView out, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const int c8 = (out.c + 7) / 8;
    long long v = t % out.chunk_stride;
    int cc = static_cast<int>((t / out.chunk_stride) % c8);
    int n = static_cast<int>(t / (out.chunk_stride * c8));
    Vec8 r = load_vec8<T>(in.data, n * in.sample_stride + (in.c8_off + cc) * in.chunk_stride + v);
    store_vec8<T>(out.data, n * out.sample_stride + (out.c8_off + cc) * out.chunk_stride + v, r);
}

// =========================================================================================== test-time augmentation
// EnsembleFlips / EnsembleOrientations (models/ensemble.py:50-103): member e sees x.permute(0, 1, *perm).flip(dims) and
// its output is flipped / permuted back before the reduction over members.  Both index shuffles are folded into the
// kernels that touch the data anyway: the NCDHW -> blocked packing of the member's input, and the accumulation of the
// member's output into the mean / vote buffers -- no flipped copies, no (E, N, C, ...) stack.
struct TtaXform {
    int perm[3];   // transformed spatial axis k is source axis perm[k] (0..2)
    int flip[3];   // transformed axis k is reversed
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
pack_ncdhw_tta_kernel(const float* __restrict__ src, DView dst, int w0, int w1, int w2, TtaXform xf, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const long long vox = dst.chunk_stride;
    const int c8 = (dst.c + 7) / 8;
    long long v = t % vox;
    int cc = static_cast<int>((t / vox) % c8);
    int n = static_cast<int>(t / (vox * c8));
    int u[3];
    u[2] = static_cast<int>(v % dst.x);
    u[1] = static_cast<int>((v / dst.x) % dst.y);
    u[0] = static_cast<int>(v / (1LL * dst.x * dst.y));
    const int S[3] = {dst.z, dst.y, dst.x};
    int w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[xf.perm[k]] = xf.flip[k] ? S[k] - 1 - u[k] : u[k];
    const long long svox = 1LL * w0 * w1 * w2;
    const long long sv = (1LL * w[0] * w1 + w[1]) * w2 + w[2];
    Vec8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int c = cc * 8 + j;
        r.v[j] = c < dst.c ? __ldg(src + (static_cast<long long>(n) * dst.c + c) * svox + sv) : 0.f;
    }
    store_vec8<T>(dst.data, n * dst.sample_stride + (dst.c8_off + cc) * dst.chunk_stride + v, r);
}

// member: fp32 (N, C, S0, S1, S2) in the member's (transformed) space; one thread per ORIGINAL-space voxel.
// mean: acc (N, C, W0, W1, W2) += member un-transformed.  majority: votes[n][argmax_c member][w] += 1 (uint8).
__global__ void __launch_bounds__(kThreads)
tta_accumulate_kernel(const float* __restrict__ member, int C, int w0, int w1, int w2, TtaXform xf,
                      float* __restrict__ acc, uint8_t* __restrict__ votes, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const long long vox = 1LL * w0 * w1 * w2;
    const long long v = t % vox;
    const long long n = t / vox;
    int w[3];
    w[2] = static_cast<int>(v % w2);
    w[1] = static_cast<int>((v / w2) % w1);
    w[0] = static_cast<int>(v / (1LL * w2 * w1));
    const int W[3] = {w0, w1, w2};
    int S[3], u[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        S[k] = W[xf.perm[k]];
        const int vk = w[xf.perm[k]];
        u[k] = xf.flip[k] ? S[k] - 1 - vk : vk;
    }
    const float* m = member + n * C * vox + (1LL * u[0] * S[1] + u[1]) * S[2] + u[2];
    if (acc != nullptr) {
        float* a = acc + n * C * vox + v;
        for (int c = 0; c < C; ++c) a[c * vox] += __ldg(m + c * vox);
    } else {
        // torch.argmax: first maximum wins, NaN counts as the maximum
        float best = __ldg(m);
        int arg = 0;
        for (int c = 1; c < C; ++c) {
            const float x = __ldg(m + c * vox);
            if (!(best != best) && (x > best || x != x)) {
                best = x;
                arg = c;
            }
        }
        votes[(n * C + arg) * vox + v] += 1;
    }
}

// mean: acc *= 1 / E (ATen's mean multiplies the sum by the reciprocal).  majority: one-hot int64 of the label with
// the most votes, smallest label on ties (torch.mode).
__global__ void __launch_bounds__(kThreads)
tta_finalize_kernel(float* __restrict__ acc, const uint8_t* __restrict__ votes, long long* __restrict__ onehot, int C,
                    long long vox, float inv_members, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    if (acc != nullptr) {
        acc[t] *= inv_members;       // total = N * C * vox
        return;
    }
    const long long v = t % vox, n = t / vox;   // total = N * vox
    const uint8_t* p = votes + n * C * vox + v;
    int best = -1, arg = 0;
    for (int c = 0; c < C; ++c) {
        const int k = p[c * vox];
        if (k > best) {
            best = k;
            arg = c;
        }
    }
    long long* o = onehot + n * C * vox + v;
    for (int c = 0; c < C; ++c) o[c * vox] = c == arg ? 1 : 0;
}

// =========================================================================================== softmax (NCDHW fp32, in place)
__global__ void __launch_bounds__(kThreads)
softmax_ncdhw_kernel(float* __restrict__ data, int channels, long long voxels, int sm_channels, float diag_bias,
                     long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    if (sm_channels <= 0) {
        long long v = t % voxels;
        long long n = t / voxels;
        float* p = data + n * channels * voxels + v;
        float m = -INFINITY;
        for (int c = 0; c < channels; ++c) m = fmaxf(m, p[c * voxels]);
        float s = 0.f;
        for (int c = 0; c < channels; ++c) s += expf(p[c * voxels] - m);
        for (int c = 0; c < channels; ++c) p[c * voxels] = expf(p[c * voxels] - m) / s;
    } else {
        // StochasticMatrix: channel k = i * C + j ; softmax over i for every j ; +diag_bias where i == j
        const int C = sm_channels;
        long long v = t % voxels;
        long long r = t / voxels;
        int j = static_cast<int>(r % C);
        long long n = r / C;
        float* p = data + n * channels * voxels + v;
        float m = -INFINITY;
        for (int i = 0; i < C; ++i) m = fmaxf(m, p[(i * C + j) * voxels] + (i == j ? diag_bias : 0.f));
        float s = 0.f;
        for (int i = 0; i < C; ++i) s += expf(p[(i * C + j) * voxels] + (i == j ? diag_bias : 0.f) - m);
        for (int i = 0; i < C; ++i) {
            float* q = p + (i * C + j) * voxels;
            *q = expf(*q + (i == j ? diag_bias : 0.f) - m) / s;
        }
    }
}

// =========================================================================================== grid extraction
struct LocBatch {
    int count;
    int loc[64][6];
};

// grid: x = (j, k) plane of the patch, y = i, z = (patch, chunk) -- no 64-bit divisions per thread
template <typename T>
__global__ void __launch_bounds__(kThreads)
grid_extract_kernel(const float* __restrict__ vol, int C, int W, int H, int D, LocBatch lb, int b0, int bw, int bh,
                    int bd, int pad_mode, float pad_value, DView dst) {
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= dst.y * dst.x) return;
    const int c8 = (dst.c + 7) / 8;
    const int j = t / dst.x, k = t - j * dst.x;
    const int i = blockIdx.y;
    const int b = blockIdx.z / c8, cc = blockIdx.z - b * c8;
    int si = lb.loc[b][0] + i - bw, sj = lb.loc[b][1] + j - bh, sk = lb.loc[b][2] + k - bd;
    bool inside = si >= 0 && si < W && sj >= 0 && sj < H && sk >= 0 && sk < D;
    if (pad_mode == 1) {
        si = min(max(si, 0), W - 1);
        sj = min(max(sj, 0), H - 1);
        sk = min(max(sk, 0), D - 1);
        inside = true;
    }
    Vec8 o;
    const long long vox = 1LL * W * H * D;
    const long long off = (static_cast<long long>(si) * H + sj) * D + sk;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        int c = cc * 8 + q;
        o.v[q] = (c < C) ? (inside ? __ldg(vol + c * vox + off) : pad_value) : 0.f;
    }
    store_vec8<T>(dst.data, vox_index(dst, b0 + b, cc, i, j, k), o);
}

// Four consecutive k per thread (patch depth multiple of 4): one 128-bit load per channel where the four source
// voxels are inside the volume and 16-byte aligned (the interior of every patch), scalar clamped loads at the padded
// border; 4 x 16 (bf16) or 4 x 32 (fp32) contiguous bytes stored per thread.
template <typename T>
__global__ void __launch_bounds__(kThreads)
grid_extract_kernel4(const float* __restrict__ vol, int C, int W, int H, int D, LocBatch lb, int b0, int bw, int bh,
                     int bd, int pad_mode, float pad_value, DView dst) {
    const int xq = dst.x >> 2;
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= dst.y * xq) return;
    const int c8 = (dst.c + 7) / 8;
    const int j = t / xq, k = (t - j * xq) * 4;
    const int i = blockIdx.y;
    const int b = blockIdx.z / c8, cc = blockIdx.z - b * c8;
    int si = lb.loc[b][0] + i - bw, sj = lb.loc[b][1] + j - bh;
    const int sk0 = lb.loc[b][2] + k - bd;
    bool row_inside = si >= 0 && si < W && sj >= 0 && sj < H;
    if (pad_mode == 1) {
        si = min(max(si, 0), W - 1);
        sj = min(max(sj, 0), H - 1);
        row_inside = true;
    }
    const long long vox = 1LL * W * H * D;
    const long long row = (static_cast<long long>(si) * H + sj) * D;
    const bool fast = row_inside && sk0 >= 0 && sk0 + 3 < D && (((row + sk0) & 3) == 0) && ((vox & 3) == 0);
    Vec8 o[4];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int c = cc * 8 + q;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < C) {
            if (fast) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(vol + c * vox + row + sk0));
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int sk = sk0 + u;
                    bool inside = row_inside && sk >= 0 && sk < D;
                    if (pad_mode == 1) {
                        sk = min(max(sk, 0), D - 1);
                        inside = true;
                    }
                    v[u] = inside ? __ldg(vol + c * vox + row + sk) : pad_value;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) o[u].v[q] = v[u];
    }
    const long long idx = vox_index(dst, b0 + b, cc, i, j, k);
#pragma unroll
    for (int u = 0; u < 4; ++u) store_vec8<T>(dst.data, idx + u, o[u]);
}

// =========================================================================================== overlap-add
// Owner-computes gather: thread (c, i, j, k4) of the batch bounding box sums, in batch order, every patch of
// the batch that covers its voxels.  VEC = 4 uses 128-bit accesses (needs k extents/offsets multiple of 4).
// grid: x = (j, k/VEC) plane of the bounding box, y = i, z = channel.  Each block first selects, with two warp
// ballots, the patches that can touch its (i, j-range) at all, then every thread walks only those (in order).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
overlap_add_kernel(float* __restrict__ out, int C, int PW, int PH, int PD, const float* __restrict__ patches,
                   LocBatch lb, int p0, int p1, int p2, int bi0, int bj0, int bk0, int bw, int bh, int bd) {
    __shared__ unsigned int cand[2];
    const int bdv = bd / VEC;
    const int i = blockIdx.y + bi0;
    const int c = blockIdx.z;
    const int t0 = blockIdx.x * kThreads;
    {
        // j range covered by this block
        const int tl = min(t0 + kThreads - 1, bh * bdv - 1);
        const int jlo = t0 / bdv + bj0, jhi = tl / bdv + bj0;
        if (threadIdx.x < 64) {
            const int b = threadIdx.x;
            const bool hit = b < lb.count && i >= lb.loc[b][0] && i < lb.loc[b][3] && jhi >= lb.loc[b][1] &&
                             jlo < lb.loc[b][4];
            const unsigned int m = __ballot_sync(0xffffffffu, hit);
            if ((threadIdx.x & 31) == 0) cand[threadIdx.x >> 5] = m;
        }
        __syncthreads();
    }
    const int t = t0 + threadIdx.x;
    if (t >= bh * bdv) return;
    const int jj = t / bdv;
    const int kk = (t - jj * bdv) * VEC + bk0;
    const int j = jj + bj0;
    float* o = out + ((static_cast<long long>(c) * PW + i) * PH + j) * PD + kk;
    float acc[VEC];
    bool touched = false, loaded = false;
    const long long pvox = 1LL * p0 * p1 * p2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        unsigned int m = cand[half];
        while (m) {
            const int b = (__ffs(m) - 1) + half * 32;
            m &= m - 1;
            const int i0 = lb.loc[b][0], j0 = lb.loc[b][1], k0 = lb.loc[b][2];
            if (j < j0 || j >= lb.loc[b][4] || kk < k0 || kk >= lb.loc[b][5]) continue;
            if (!loaded) {
                if constexpr (VEC == 4) {
                    float4 v = *reinterpret_cast<const float4*>(o);
                    acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[VEC - 1] = v.w;
                } else {
                    acc[0] = *o;
                }
                loaded = true;
            }
            const float* p = patches + (static_cast<long long>(b) * C + c) * pvox +
                             (static_cast<long long>(i - i0) * p1 + (j - j0)) * p2 + (kk - k0);
            if constexpr (VEC == 4) {
                float4 v = __ldg(reinterpret_cast<const float4*>(p));
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[VEC - 1] += v.w;
            } else {
                acc[0] += __ldg(p);
            }
            touched = true;
        }
    }
    if (!touched) return;
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[VEC - 1]);
    } else {
        *o = acc[0];
    }
}

// Row-owner variant of the 128-bit path (round 2).  ncu on the kernel above: ~110 instructions per 16-byte patch load
// (every thread re-tests every candidate patch of its block against its own j and k, reading the location table from
// the parameter bank) -- issue slots 57 % busy at 32 % DRAM.  Here one WARP owns one accumulator row (c, i, j): its
// lanes test the batch's patches against (i, j) once (ballot + ordered compaction into shared memory, with the
// patch-row base pointer and k range precomputed), then every lane walks the short list for its two 16-byte vectors:
// one broadcast LDS.128, two compares, one load, four adds per candidate.  Batch order is preserved, so the sums are
// bit-identical to the sequential CPU +=.
struct RowCand {
    int k0, k1;            // covered k range of the row, in floats
    long long off;         // element offset of the patch row such that patches[off + k] is the value for out k
};

__global__ void __launch_bounds__(kThreads)
overlap_add_rows_kernel(float* __restrict__ out, int C, int PW, int PH, int PD, const float* __restrict__ patches,
                        LocBatch lb, int p0, int p1, int p2, int bi0, int bj0, int bk0, int bh, int bd) {
    __shared__ RowCand cand[kThreads / 32][64];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int jrow = blockIdx.y * (kThreads / 32) + warp;
    if (jrow >= bh) return;
    const int j = jrow + bj0;
    const int c = blockIdx.z % C;
    const int i = blockIdx.z / C + bi0;
    int n = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int b = lane + 32 * half;
        const bool hit = b < lb.count && i >= lb.loc[b][0] && i < lb.loc[b][3] && j >= lb.loc[b][1] && j < lb.loc[b][4];
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            RowCand rc;
            rc.k0 = lb.loc[b][2];
            rc.k1 = lb.loc[b][5];
            rc.off = ((static_cast<long long>(b) * C + c) * p0 + (i - lb.loc[b][0])) * p1 * p2 +
                     static_cast<long long>(j - lb.loc[b][1]) * p2 - rc.k0;
            cand[warp][n + __popc(m & ((1u << lane) - 1u))] = rc;
        }
        n += __popc(m);
    }
    __syncwarp();
    if (n == 0) return;
    float* orow = out + ((static_cast<long long>(c) * PW + i) * PH + j) * PD;
    const int kend = bk0 + bd;
    for (int kb = bk0 + blockIdx.x * 256; kb < kend; kb += gridDim.x * 256) {      // 64 vectors per warp pass
        const int ka = kb + lane * 4, kc = ka + 128;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b4 = a;
        bool ta = false, tb = false;
        // the accumulator values are needed only if some patch covers the vector: load them up front (independent
        // of the gathers), use them when touched
        const bool ina = ka < kend, inb = kc < kend;
        float4 oa = a, ob = a;
        if (ina) oa = *reinterpret_cast<const float4*>(orow + ka);
        if (inb) ob = *reinterpret_cast<const float4*>(orow + kc);
        a = oa;
        b4 = ob;
        for (int q = 0; q < n; ++q) {
            const RowCand rc = cand[warp][q];
            if (ina && ka >= rc.k0 && ka < rc.k1) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(patches + rc.off + ka));
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                ta = true;
            }
            if (inb && kc >= rc.k0 && kc < rc.k1) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(patches + rc.off + kc));
                b4.x += v.x; b4.y += v.y; b4.z += v.z; b4.w += v.w;
                tb = true;
            }
        }
        if (ta) *reinterpret_cast<float4*>(orow + ka) = a;
        if (tb) *reinterpret_cast<float4*>(orow + kc) = b4;
    }
}

// 'hann' mode (torchio >= 0.19 GridAggregator): every patch is multiplied by the separable Hann window before the
// overlap-add and the sum is divided by the summed windows.  The window product is built the way torchio builds its
// 3-D window, ((w0[i] * w1[j]) * w2[k]), so the weighted values carry the same rounding.
__global__ void __launch_bounds__(kThreads)
window_patches_kernel(float* __restrict__ patches, int p0, int p1, int p2, const float* __restrict__ w0,
                      const float* __restrict__ w1, const float* __restrict__ w2, long long total4) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total4) return;
    const int k = static_cast<int>(t % (p2 / 4)) * 4;
    long long r = t / (p2 / 4);
    const int j = static_cast<int>(r % p1);
    const int i = static_cast<int>((r / p1) % p0);
    const float wij = __ldg(w0 + i) * __ldg(w1 + j);
    float4 v = reinterpret_cast<float4*>(patches)[t];
    v.x *= wij * __ldg(w2 + k);
    v.y *= wij * __ldg(w2 + k + 1);
    v.z *= wij * __ldg(w2 + k + 2);
    v.w *= wij * __ldg(w2 + k + 3);
    reinterpret_cast<float4*>(patches)[t] = v;
}

// out[c][i][j][k] /= (s0[i] * s1[j]) * s2[k]: the summed windows of a product grid are separable
__global__ void __launch_bounds__(kThreads)
divide_separable_kernel(float* __restrict__ out, int PW, int PH, int PD, const float* __restrict__ s0,
                        const float* __restrict__ s1, const float* __restrict__ s2, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    const int k = static_cast<int>(t % PD);
    long long r = t / PD;
    const int j = static_cast<int>(r % PH);
    const int i = static_cast<int>((r / PH) % PW);
    out[t] = out[t] / ((__ldg(s0 + i) * __ldg(s1 + j)) * __ldg(s2 + k));
}

// 'crop' mode: every patch assigns the centre crop of itself; patches are processed in order, later wins.
__global__ void __launch_bounds__(kThreads)
overlap_crop_kernel(float* __restrict__ out, int C, int PW, int PH, int PD, const float* __restrict__ patches,
                    LocBatch lb, int b, int p0, int p1, int p2, int ci0, int cj0, int ck0, int cw, int chh, int cd,
                    int li, int lj, int lk, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    int k = static_cast<int>(t % cd);
    long long r = t / cd;
    int j = static_cast<int>(r % chh);
    r /= chh;
    int i = static_cast<int>(r % cw);
    int c = static_cast<int>(r / cw);
    const long long pvox = 1LL * p0 * p1 * p2;
    float v = __ldg(patches + (static_cast<long long>(b) * C + c) * pvox +
                    (static_cast<long long>(li + i) * p1 + (lj + j)) * p2 + (lk + k));
    out[((static_cast<long long>(c) * PW + ci0 + i) * PH + cj0 + j) * PD + ck0 + k] = v;
}

// =========================================================================================== finalize (+ argmax)
// Block = 256 consecutive (row j, k-vector) pairs of one plane i; grid = (pair blocks, planes): one 32-bit division per
// thread, no 64-bit div / mod, and no idle lanes when D / VEC is not a multiple of 64 (config 2: 48 vectors per row --
// the earlier 64 x 4 tiling left a quarter of every block idle).  The channel
// loop loads up to 4 channels' vectors before touching them, so every thread keeps 4 x 16 bytes in flight (round 1
// loaded, divided and compared one channel at a time and sat at 22 % of the copy bandwidth with 2 channels).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
finalize_kernel(const float* __restrict__ out, int C, int PW, int PH, int PD, const int* __restrict__ cw,
                const int* __restrict__ ch, const int* __restrict__ cd, int b0, int b1, int b2, int W, int H, int D,
                float* __restrict__ probs, long long* __restrict__ lab64, uint8_t* __restrict__ lab8) {
    const int kvs = (D + VEC - 1) / VEC;                       // k-vectors per row
    const int pair = blockIdx.x * kThreads + threadIdx.x;
    const int j = pair / kvs;
    const int i = blockIdx.y;
    const int k = (pair - j * kvs) * VEC;
    if (j >= H) return;
    const long long pvox = 1LL * PW * PH * PD;
    const long long vox = 1LL * W * H * D;
    const long long src = (static_cast<long long>(i + b0) * PH + (j + b1)) * PD + (k + b2);
    const long long dst = (static_cast<long long>(i) * H + j) * D + k;
    // ncu: the round-1 kernel was INSTRUCTION bound (issue slots 66 % busy at 21 % DRAM) by the IEEE divisions.  The
    // coverage count is a product of per-axis counts, i.e. a power of two whenever the overlap is <= 50 % (counts 1 or 2
    // per axis): then x / cnt == x * (1 / cnt) exactly (scaling by 2^-k is exact, and both sides round a subnormal
    // result identically), one multiply instead of a division sequence.  Other counts keep the division.
    float cnt[VEC], rcp[VEC];
    bool pow2 = true;
    if (cw != nullptr) {
        int cij = __ldg(cw + i + b0) * __ldg(ch + j + b1);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            const int n = cij * __ldg(cd + k + b2 + q);
            cnt[q] = static_cast<float>(n);
            rcp[q] = __frcp_rn(cnt[q]);
            pow2 = pow2 && ((n & (n - 1)) == 0);
        }
    }
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        best[q] = -INFINITY;
        arg[q] = 0;
    }
    for (int c0 = 0; c0 < C; c0 += 4) {
        float v[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c0 + u < C) {
                if constexpr (VEC == 4) {
                    const float4 f = __ldg(reinterpret_cast<const float4*>(out + (c0 + u) * pvox + src));
                    v[u][0] = f.x; v[u][1] = f.y; v[u][2] = f.z; v[u][VEC - 1] = f.w;
                } else {
                    v[u][0] = __ldg(out + (c0 + u) * pvox + src);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c0 + u < C) {
                const int c = c0 + u;
                if (cw != nullptr) {
                    if (pow2) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) v[u][q] = v[u][q] * rcp[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) v[u][q] = v[u][q] / cnt[q];
                    }
                }
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    // torch.argmax: first maximal index; NaN counts as maximal
                    if (v[u][q] > best[q] || (v[u][q] != v[u][q] && best[q] == best[q])) {
                        best[q] = v[u][q];
                        arg[q] = c;
                    }
                }
                if (probs != nullptr) {
                    if constexpr (VEC == 4) {
                        *reinterpret_cast<float4*>(probs + c * vox + dst) = make_float4(v[u][0], v[u][1], v[u][2], v[u][VEC - 1]);
                    } else {
                        probs[c * vox + dst] = v[u][0];
                    }
                }
            }
        }
    }
    if (lab64 != nullptr) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) lab64[dst + q] = arg[q];
    }
    if (lab8 != nullptr) {
        if constexpr (VEC == 4) {
            *reinterpret_cast<uchar4*>(lab8 + dst) = make_uchar4(arg[0], arg[1], arg[2], arg[VEC - 1]);
        } else {
            lab8[dst] = static_cast<uint8_t>(arg[0]);
        }
    }
}

// =========================================================================================== confusion histogram
// One private histogram per WARP in shared memory (bins = nc * nc counters: 400 bytes for 10 classes, so every SM runs
// at full occupancy; round 1 kept a column per LANE, 32 x the memory, and ran at 8-16 warps per SM).  Label maps are
// piecewise constant, so the common case is a 16-byte vector whose voxels all fall into one bin: the lanes of the warp
// that hold the same bin are grouped with __match_any_sync, the group's voxel count is summed with __reduce_add_sync
// and ONE lane adds it (a single conflict-free shared-memory atomic per distinct bin).  Vectors that straddle a label
// boundary take the run-length path, one shared-memory atomic per run.  One int64 global atomic per non-empty bin per block at the end.
template <typename L>
__global__ void __launch_bounds__(kThreads)
confusion_kernel(const L* __restrict__ pred, const L* __restrict__ targ, long long voxels, int nc,
                 unsigned long long* __restrict__ cm) {
    extern __shared__ unsigned int hist[];  // [warp][bin]
    const int bins = nc * nc;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    constexpr int kWarps = kThreads / 32;
    for (int i = threadIdx.x; i < kWarps * bins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    unsigned int* mine = hist + warp * bins;
    constexpr int PER = 16 / sizeof(L);  // labels per 16-byte load
    const long long nvec = voxels / PER;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    // values outside [0, nc) fall into the LAST class ("other"): nothing is dropped, so row / column sums are the
    // true marginals (the unsigned compare also catches negative int64 labels)
    auto bin_of = [&](L p, L t) -> int {
        const unsigned long long pu = static_cast<unsigned long long>(p), tu = static_cast<unsigned long long>(t);
        const int pi = pu < static_cast<unsigned long long>(nc) ? static_cast<int>(pu) : nc - 1;
        const int ti = tu < static_cast<unsigned long long>(nc) ? static_cast<int>(tu) : nc - 1;
        return ti * nc + pi;
    };
    // called by ALL 32 lanes together (lanes past the end pass valid = false)
    auto count_vec = [&](const uint4& pr, const uint4& tr, bool valid) {
        const L* pp = reinterpret_cast<const L*>(&pr);
        const L* tp = reinterpret_cast<const L*>(&tr);
        bool uniform = true;
        int bin0 = 0;
        if (valid) {
            // all labels of the vector equal <=> every 32-bit word equals the first one rotated ... cheap exact test:
            bin0 = bin_of(pp[0], tp[0]);
#pragma unroll
            for (int q = 1; q < PER; ++q) uniform = uniform && (pp[q] == pp[0]) && (tp[q] == tp[0]);
        }
        const unsigned fast = __ballot_sync(0xffffffffu, valid && uniform);
        if (valid && uniform) {
            const unsigned peers = __match_any_sync(fast, bin0);
            if (lane == __ffs(peers) - 1) atomicAdd(mine + bin0, static_cast<unsigned>(PER * __popc(peers)));
        } else if (valid) {
            int cur = bin0;
            unsigned run = 1;
#pragma unroll
            for (int q = 1; q < PER; ++q) {
                const int b = bin_of(pp[q], tp[q]);
                if (b == cur) {
                    ++run;
                } else {
                    atomicAdd(mine + cur, run);
                    cur = b;
                    run = 1;
                }
            }
            atomicAdd(mine + cur, run);
        }
    };
    // four vector pairs (128 bytes) in flight per thread
    const uint4* pv = reinterpret_cast<const uint4*>(pred);
    const uint4* tv = reinterpret_cast<const uint4*>(targ);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    const long long warp_base = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) - lane;   // first vector of this warp
    for (long long w0 = warp_base; w0 < nvec; w0 += 4 * stride) {      // warp-uniform trip count
        uint4 pr[4], tr[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long v = w0 + lane + u * stride;
            ok[u] = v < nvec;
            pr[u] = ok[u] ? __ldg(pv + v) : zero;
            tr[u] = ok[u] ? __ldg(tv + v) : zero;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) count_vec(pr[u], tr[u], ok[u]);
    }
    // tail (voxels not a multiple of PER): first thread of the grid
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long v = nvec * PER; v < voxels; ++v) atomicAdd(mine + bin_of(pred[v], targ[v]), 1u);
    }
    __syncthreads();
    for (int bin = threadIdx.x; bin < bins; bin += blockDim.x) {
        unsigned long long s = 0;
        for (int w = 0; w < kWarps; ++w) s += hist[w * bins + bin];
        if (s) atomicAdd(cm + bin, s);
    }
}

// argmax over channels of [C][V] fp32
template <int VEC>
__global__ void __launch_bounds__(kThreads)
argmax_kernel(const float* __restrict__ probs, int C, long long voxels, long long* __restrict__ lab64,
              uint8_t* __restrict__ lab8, long long total) {
    long long t = blockIdx.x * 1LL * kThreads + threadIdx.x;
    if (t >= total) return;
    long long v0 = t * VEC;
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        best[q] = -INFINITY;
        arg[q] = 0;
    }
    for (int c = 0; c < C; ++c) {
        float v[VEC];
        if constexpr (VEC == 4) {
            float4 f = __ldg(reinterpret_cast<const float4*>(probs + c * voxels + v0));
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[VEC - 1] = f.w;
        } else {
            v[0] = __ldg(probs + c * voxels + v0);
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q)
            if (v[q] > best[q] || (v[q] != v[q] && best[q] == best[q])) {
                best[q] = v[q];
                arg[q] = c;
            }
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        if (lab64) lab64[v0 + q] = arg[q];
        if (lab8) lab8[v0 + q] = static_cast<uint8_t>(arg[q]);
    }
}

}  // namespace b200seg

using namespace b200seg;

#define DISPATCH_DTYPE(dtype, ...)                          \
    do {                                                    \
        if ((dtype) == B200SEG_F32) {                       \
            using T = float;                                \
            __VA_ARGS__;                                    \
        } else {                                            \
            using T = __nv_bfloat16;                        \
            __VA_ARGS__;                                    \
        }                                                   \
    } while (0)

extern "C" {

int b200seg_pack_ncdhw(const float* src, b200seg_view dst, void* stream) {
    B200SEG_CHECK_ARG(src != nullptr, "pack_ncdhw: null src");
    int rc = validate_view(dst, "pack_ncdhw dst");
    if (rc) return rc;
    DView d = make_dview(dst);
    long long total = 1LL * d.n * ((d.c + 7) / 8) * d.chunk_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_DTYPE(dst.dtype, (pack_ncdhw_kernel<T><<<blocks_for(total), kThreads, 0, s>>>(src, d, total)));
    return check_launch("pack_ncdhw");
}

int b200seg_unpack_ncdhw(b200seg_view src, float* dst, void* stream) {
    B200SEG_CHECK_ARG(dst != nullptr, "unpack_ncdhw: null dst");
    int rc = validate_view(src, "unpack_ncdhw src");
    if (rc) return rc;
    DView d = make_dview(src);
    long long total = 1LL * d.n * ((d.c + 7) / 8) * d.chunk_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_DTYPE(src.dtype, (unpack_ncdhw_kernel<T><<<blocks_for(total), kThreads, 0, s>>>(d, dst, total)));
    return check_launch("unpack_ncdhw");
}

static int check_pair(const b200seg_view& in, const b200seg_view& out, const char* what) {
    int rc = validate_view(in, what);
    if (rc) return rc;
    rc = validate_view(out, what);
    if (rc) return rc;
    B200SEG_CHECK_ARG(in.dtype == out.dtype, "%s: dtype mismatch", what);
    B200SEG_CHECK_ARG(in.n == out.n && in.c == out.c, "%s: n/c mismatch (%d,%d) vs (%d,%d)", what, in.n, in.c, out.n,
                      out.c);
    return B200SEG_OK;
}

int b200seg_avgpool2(b200seg_view in, b200seg_view out, void* stream) {
    int rc = check_pair(in, out, "avgpool2");
    if (rc) return rc;
    B200SEG_CHECK_ARG(out.z == in.z / 2 && out.y == in.y / 2 && out.x == in.x / 2, "avgpool2: out extent must be in/2");
    DView di = make_dview(in), dout = make_dview(out);
    long long total = 1LL * dout.n * ((dout.c + 7) / 8) * dout.chunk_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_DTYPE(in.dtype, (avgpool2_kernel<T><<<blocks_for(total), kThreads, 0, s>>>(di, dout, total)));
    return check_launch("avgpool2");
}

int b200seg_upsample_trilinear2(b200seg_view in, b200seg_view out, void* stream) {
    int rc = check_pair(in, out, "upsample_trilinear2");
    if (rc) return rc;
    B200SEG_CHECK_ARG(out.z == in.z * 2 && out.y == in.y * 2 && out.x == in.x * 2,
                      "upsample_trilinear2: out extent must be 2*in");
    DView di = make_dview(in), dout = make_dview(out);
    long long total = 1LL * dout.n * ((dout.c + 7) / 8) * dout.chunk_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_DTYPE(in.dtype, (upsample_trilinear2_kernel<T><<<blocks_for(total), kThreads, 0, s>>>(di, dout, total)));
    return check_launch("upsample_trilinear2");
}

int b200seg_copy_view(b200seg_view in, b200seg_view out, void* stream) {
    int rc = check_pair(in, out, "copy_view");
    if (rc) return rc;
    B200SEG_CHECK_ARG(out.z == in.z && out.y == in.y && out.x == in.x, "copy_view: extent mismatch");
    DView di = make_dview(in), dout = make_dview(out);
    long long total = 1LL * dout.n * ((dout.c + 7) / 8) * dout.chunk_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_DTYPE(in.dtype, (copy_view_kernel<T><<<blocks_for(total), kThreads, 0, s>>>(di, dout, total)));
    return check_launch("copy_view");
}

static int check_xform(const int32_t perm[3], const int32_t flip[3], TtaXform* xf, const char* what) {
    int seen = 0;
    for (int k = 0; k < 3; ++k) {
        B200SEG_CHECK_ARG(perm[k] >= 0 && perm[k] <= 2, "%s: perm[%d] = %d", what, k, perm[k]);
        seen |= 1 << perm[k];
        xf->perm[k] = perm[k];
        xf->flip[k] = flip[k] ? 1 : 0;
    }
    B200SEG_CHECK_ARG(seen == 7, "%s: perm is not a permutation of (0, 1, 2)", what);
    return B200SEG_OK;
}

int b200seg_pack_ncdhw_tta(const float* src, int32_t w0, int32_t w1, int32_t w2, const int32_t perm[3],
                           const int32_t flip[3], b200seg_view dst, void* stream) {
    B200SEG_CHECK_ARG(src != nullptr, "pack_ncdhw_tta: null source");
    int rc = validate_view(dst, "pack_ncdhw_tta dst");
    if (rc) return rc;
    TtaXform xf;
    rc = check_xform(perm, flip, &xf, "pack_ncdhw_tta");
    if (rc) return rc;
    const int W[3] = {w0, w1, w2};
    B200SEG_CHECK_ARG(dst.z == W[perm[0]] && dst.y == W[perm[1]] && dst.x == W[perm[2]],
                      "pack_ncdhw_tta: dst extent (%d, %d, %d) is not the permuted source extent", dst.z, dst.y, dst.x);
    DView dd = make_dview(dst);
    long long total = 1LL * dd.n * ((dd.c + 7) / 8) * dd.chunk_stride;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DISPATCH_DTYPE(dst.dtype, (pack_ncdhw_tta_kernel<T><<<blocks_for(total), kThreads, 0, s>>>(src, dd, w0, w1, w2, xf, total)));
    return check_launch("pack_ncdhw_tta");
}

int b200seg_tta_accumulate(const float* member, int64_t n, int32_t c, int32_t w0, int32_t w1, int32_t w2,
                           const int32_t perm[3], const int32_t flip[3], float* acc, uint8_t* votes, void* stream) {
    B200SEG_CHECK_ARG(member != nullptr && n > 0 && c > 0 && w0 > 0 && w1 > 0 && w2 > 0, "tta_accumulate: bad arguments");
    B200SEG_CHECK_ARG((acc != nullptr) != (votes != nullptr), "tta_accumulate: exactly one of acc / votes");
    TtaXform xf;
    int rc = check_xform(perm, flip, &xf, "tta_accumulate");
    if (rc) return rc;
    long long total = n * w0 * w1 * w2;
    tta_accumulate_kernel<<<blocks_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(member, c, w0, w1, w2, xf,
                                                                                                 acc, votes, total);
    return check_launch("tta_accumulate");
}

int b200seg_tta_finalize(float* acc, const uint8_t* votes, int64_t* onehot, int64_t n, int32_t c, int64_t voxels,
                         int32_t members, void* stream) {
    B200SEG_CHECK_ARG(n > 0 && c > 0 && voxels > 0 && members > 0 && members <= 255, "tta_finalize: bad arguments");
    B200SEG_CHECK_ARG((acc != nullptr) != (votes != nullptr && onehot != nullptr), "tta_finalize: acc, or votes + onehot");
    long long total = acc != nullptr ? n * c * voxels : n * voxels;
    tta_finalize_kernel<<<blocks_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        acc, votes, reinterpret_cast<long long*>(onehot), c, voxels, 1.0f / static_cast<float>(members), total);
    return check_launch("tta_finalize");
}

int b200seg_copy_planes(const float* patch, int32_t c, int32_t p0, int32_t p1, int32_t p2, int32_t plane_lo,
                        int32_t plane_hi, float* dst, void* stream) {
    B200SEG_CHECK_ARG(patch != nullptr && dst != nullptr && c > 0 && p0 > 0 && p1 > 0 && p2 > 0, "copy_planes: bad arguments");
    B200SEG_CHECK_ARG(plane_lo >= 0 && plane_lo < plane_hi && plane_hi <= p0, "copy_planes: planes [%d, %d) of %d", plane_lo,
                      plane_hi, p0);
    // c rows of (plane_hi - plane_lo) * p1 * p2 contiguous floats: a pitched copy on the copy engine, no kernel
    const size_t width = static_cast<size_t>(plane_hi - plane_lo) * p1 * p2 * sizeof(float);
    const size_t spitch = static_cast<size_t>(p0) * p1 * p2 * sizeof(float);
    B200SEG_CHECK_CUDA(cudaMemcpy2DAsync(dst, width, patch + static_cast<size_t>(plane_lo) * p1 * p2, spitch, width,
                                         static_cast<size_t>(c), cudaMemcpyDeviceToDevice,
                                         static_cast<cudaStream_t>(stream)));
    return B200SEG_OK;
}

int b200seg_softmax_ncdhw(float* data, int64_t n, int32_t channels, int64_t voxels, int32_t sm_channels,
                          float diag_bias, void* stream) {
    B200SEG_CHECK_ARG(data != nullptr && n > 0 && channels > 0 && voxels > 0, "softmax_ncdhw: bad arguments");
    B200SEG_CHECK_ARG(sm_channels <= 0 || sm_channels * sm_channels == channels,
                      "softmax_ncdhw: StochasticMatrix needs channels == C*C (C=%d, channels=%d)", sm_channels, channels);
    long long total = n * voxels * (sm_channels > 0 ? sm_channels : 1);
    softmax_ncdhw_kernel<<<blocks_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        data, channels, voxels, sm_channels, diag_bias, total);
    return check_launch("softmax_ncdhw");
}

int b200seg_grid_extract(const float* volume, int32_t c, int32_t w, int32_t h, int32_t d,
                         const int32_t* locations_host, int32_t count, const int32_t border[3], int32_t pad_mode,
                         float pad_value, b200seg_view dst, void* stream) {
    B200SEG_CHECK_ARG(volume != nullptr && locations_host != nullptr && count > 0, "grid_extract: bad arguments");
    int rc = validate_view(dst, "grid_extract dst");
    if (rc) return rc;
    B200SEG_CHECK_ARG(dst.n >= count && dst.c == c, "grid_extract: dst holds %d samples / %d channels, need %d / %d", dst.n,
                      dst.c, count, c);
    B200SEG_CHECK_ARG(pad_mode >= 0 && pad_mode <= 2, "grid_extract: pad_mode %d", pad_mode);
    int bw = pad_mode ? border[0] : 0, bh = pad_mode ? border[1] : 0, bd = pad_mode ? border[2] : 0;
    DView dd = make_dview(dst);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int b0 = 0; b0 < count; b0 += 64) {
        LocBatch lb;
        lb.count = count - b0 < 64 ? count - b0 : 64;
        for (int b = 0; b < lb.count; ++b) {
            for (int q = 0; q < 6; ++q) lb.loc[b][q] = locations_host[(b0 + b) * 6 + q];
            B200SEG_CHECK_ARG(lb.loc[b][3] - lb.loc[b][0] == dst.z && lb.loc[b][4] - lb.loc[b][1] == dst.y &&
                                  lb.loc[b][5] - lb.loc[b][2] == dst.x,
                              "grid_extract: location %d extent differs from the patch view", b0 + b);
            if (pad_mode == 0) {
                B200SEG_CHECK_ARG(lb.loc[b][0] >= 0 && lb.loc[b][3] <= w && lb.loc[b][1] >= 0 && lb.loc[b][4] <= h &&
                                      lb.loc[b][2] >= 0 && lb.loc[b][5] <= d,
                                  "grid_extract: location %d outside the volume", b0 + b);
            }
        }
        if (dst.x % 4 == 0 && (reinterpret_cast<uintptr_t>(volume) & 15) == 0) {
            dim3 grid(blocks_for(1LL * dst.y * (dst.x / 4)), static_cast<unsigned>(dst.z),
                      static_cast<unsigned>(lb.count * ((c + 7) / 8)));
            DISPATCH_DTYPE(dst.dtype, (grid_extract_kernel4<T><<<grid, kThreads, 0, s>>>(
                                          volume, c, w, h, d, lb, b0, bw, bh, bd, pad_mode, pad_value, dd)));
        } else {
            dim3 grid(blocks_for(1LL * dst.y * dst.x), static_cast<unsigned>(dst.z),
                      static_cast<unsigned>(lb.count * ((c + 7) / 8)));
            DISPATCH_DTYPE(dst.dtype, (grid_extract_kernel<T><<<grid, kThreads, 0, s>>>(
                                          volume, c, w, h, d, lb, b0, bw, bh, bd, pad_mode, pad_value, dd)));
        }
        rc = check_launch("grid_extract");
        if (rc) return rc;
    }
    return B200SEG_OK;
}

int b200seg_overlap_add(float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const float* patches,
                        const int32_t* locations_host, int32_t count, void* stream) {
    B200SEG_CHECK_ARG(out && patches && locations_host && count > 0 && c > 0, "overlap_add: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int p0 = locations_host[3] - locations_host[0], p1 = locations_host[4] - locations_host[1],
              p2 = locations_host[5] - locations_host[2];
    for (int b0 = 0; b0 < count; b0 += 64) {
        LocBatch lb;
        lb.count = count - b0 < 64 ? count - b0 : 64;
        int bb[6] = {1 << 30, 1 << 30, 1 << 30, 0, 0, 0};
        bool vec = (pd % 4 == 0) && (p2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(patches) & 15) == 0);
        for (int b = 0; b < lb.count; ++b) {
            const int32_t* l = locations_host + (b0 + b) * 6;
            for (int q = 0; q < 6; ++q) lb.loc[b][q] = l[q];
            B200SEG_CHECK_ARG(l[3] - l[0] == p0 && l[4] - l[1] == p1 && l[5] - l[2] == p2,
                              "overlap_add: patches of one call must share one size");
            B200SEG_CHECK_ARG(l[0] >= 0 && l[1] >= 0 && l[2] >= 0 && l[3] <= pw && l[4] <= ph && l[5] <= pd,
                              "overlap_add: location %d outside the output", b0 + b);
            for (int q = 0; q < 3; ++q) {
                bb[q] = l[q] < bb[q] ? l[q] : bb[q];
                bb[q + 3] = l[q + 3] > bb[q + 3] ? l[q + 3] : bb[q + 3];
            }
            vec = vec && (l[2] % 4 == 0);
        }
        const int bw = bb[3] - bb[0], bh = bb[4] - bb[1], bd = bb[5] - bb[2];
        const float* pp = patches + 1LL * b0 * c * p0 * p1 * p2;
        static const bool legacy = [] {
            const char* v = getenv("B200SEG_OVERLAP_ADD_V1");      // test / A-B hook: the round-1 kernel
            return v && v[0] == '1';
        }();
        if (vec && !legacy && 1LL * bw * c <= 65535) {
            dim3 grid(static_cast<unsigned>((bd + 255) / 256), static_cast<unsigned>((bh + kThreads / 32 - 1) / (kThreads / 32)),
                      static_cast<unsigned>(bw * c));
            overlap_add_rows_kernel<<<grid, kThreads, 0, s>>>(out, c, pw, ph, pd, pp, lb, p0, p1, p2, bb[0], bb[1], bb[2],
                                                              bh, bd);
        } else if (vec) {
            dim3 grid(blocks_for(1LL * bh * (bd / 4)), static_cast<unsigned>(bw), static_cast<unsigned>(c));
            overlap_add_kernel<4><<<grid, kThreads, 0, s>>>(out, c, pw, ph, pd, pp, lb, p0, p1, p2, bb[0], bb[1], bb[2],
                                                            bw, bh, bd);
        } else {
            dim3 grid(blocks_for(1LL * bh * bd), static_cast<unsigned>(bw), static_cast<unsigned>(c));
            overlap_add_kernel<1><<<grid, kThreads, 0, s>>>(out, c, pw, ph, pd, pp, lb, p0, p1, p2, bb[0], bb[1], bb[2],
                                                            bw, bh, bd);
        }
        int rc = check_launch("overlap_add");
        if (rc) return rc;
    }
    return B200SEG_OK;
}

int b200seg_window_patches(float* patches, int32_t count, int32_t c, int32_t p0, int32_t p1, int32_t p2, const float* w0,
                           const float* w1, const float* w2, void* stream) {
    B200SEG_CHECK_ARG(patches && w0 && w1 && w2 && count > 0 && c > 0 && p0 > 0 && p1 > 0 && p2 > 0,
                      "window_patches: bad arguments");
    B200SEG_CHECK_ARG(p2 % 4 == 0 && (reinterpret_cast<uintptr_t>(patches) & 15) == 0,
                      "window_patches: the last patch axis must be a multiple of 4 and the buffer 16-byte aligned");
    const long long total4 = 1LL * count * c * p0 * p1 * (p2 / 4);
    window_patches_kernel<<<blocks_for(total4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(patches, p0, p1, p2, w0,
                                                                                                 w1, w2, total4);
    return check_launch("window_patches");
}

int b200seg_divide_separable(float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const float* s0, const float* s1,
                             const float* s2, void* stream) {
    B200SEG_CHECK_ARG(out && s0 && s1 && s2 && c > 0 && pw > 0 && ph > 0 && pd > 0, "divide_separable: bad arguments");
    const long long total = 1LL * c * pw * ph * pd;
    divide_separable_kernel<<<blocks_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(out, pw, ph, pd, s0, s1,
                                                                                                  s2, total);
    return check_launch("divide_separable");
}

int b200seg_overlap_crop(float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const float* patches,
                         const int32_t* locations_host, int32_t count, const int32_t border[3],
                         int32_t volume_padded, void* stream) {
    B200SEG_CHECK_ARG(out && patches && locations_host && count > 0 && c > 0, "overlap_crop: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int size[3] = {pw, ph, pd};
    for (int b = 0; b < count; ++b) {
        const int32_t* l = locations_host + b * 6;
        int ini[3], fin[3], left[3], crop[3], psz[3];
        for (int q = 0; q < 3; ++q) {
            int b_ini = border[q], b_fin = border[q];
            if (!volume_padded) {
                if (l[q] == 0) b_ini = 0;
                if (l[q + 3] == size[q]) b_fin = 0;
            }
            ini[q] = l[q] + b_ini;
            fin[q] = l[q + 3] - b_fin;
            psz[q] = l[q + 3] - l[q];
            crop[q] = fin[q] - ini[q];
            left[q] = (psz[q] - crop[q]) / 2;  // centre crop, as GridAggregator.crop_batch
            B200SEG_CHECK_ARG(crop[q] > 0, "overlap_crop: empty crop");
        }
        LocBatch lb;
        lb.count = 0;
        long long total = 1LL * c * crop[0] * crop[1] * crop[2];
        overlap_crop_kernel<<<blocks_for(total), kThreads, 0, s>>>(out, c, pw, ph, pd, patches, lb, b, psz[0], psz[1],
                                                                   psz[2], ini[0], ini[1], ini[2], crop[0], crop[1],
                                                                   crop[2], left[0], left[1], left[2], total);
        int rc = check_launch("overlap_crop");
        if (rc) return rc;
    }
    return B200SEG_OK;
}

int b200seg_finalize(const float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const int32_t* cw,
                     const int32_t* ch, const int32_t* cd, const int32_t border[3], float* probs,
                     int64_t* labels_i64, uint8_t* labels_u8, void* stream) {
    B200SEG_CHECK_ARG(pw > 0 && ph > 0 && pd > 0, "finalize: bad arguments");
    const int32_t extent[3] = {pw - 2 * border[0], ph - 2 * border[1], pd - 2 * border[2]};
    return b200seg_finalize_region(out, c, pw, ph, pd, cw, ch, cd, border, extent, probs, labels_i64, labels_u8, stream);
}

int b200seg_finalize_region(const float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const int32_t* cw,
                            const int32_t* ch, const int32_t* cd, const int32_t offset[3], const int32_t extent[3],
                            float* probs, int64_t* labels_i64, uint8_t* labels_u8, void* stream) {
    B200SEG_CHECK_ARG(out && c > 0 && pw > 0 && ph > 0 && pd > 0, "finalize: bad arguments");
    B200SEG_CHECK_ARG((cw == nullptr) == (ch == nullptr) && (cw == nullptr) == (cd == nullptr),
                      "finalize: give all three count arrays or none");
    B200SEG_CHECK_ARG(labels_u8 == nullptr || c <= 256, "finalize: uint8 labels need <= 256 classes");
    const int32_t* border = offset;
    const int W = extent[0], H = extent[1], D = extent[2];
    B200SEG_CHECK_ARG(W > 0 && H > 0 && D > 0, "finalize: empty region");
    B200SEG_CHECK_ARG(offset[0] >= 0 && offset[1] >= 0 && offset[2] >= 0 && offset[0] + W <= pw && offset[1] + H <= ph &&
                          offset[2] + D <= pd,
                      "finalize: region exceeds the accumulator");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    bool vec = (D % 4 == 0) && (pd % 4 == 0) && (border[2] % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
               (probs == nullptr || (reinterpret_cast<uintptr_t>(probs) & 15) == 0) &&
               (labels_u8 == nullptr || (reinterpret_cast<uintptr_t>(labels_u8) & 3) == 0);
    B200SEG_CHECK_ARG(W <= 65535, "finalize: extent %d exceeds the launch grid", W);
    if (vec) {
        dim3 grid(static_cast<unsigned>((1LL * H * (D / 4) + kThreads - 1) / kThreads), static_cast<unsigned>(W));
        finalize_kernel<4><<<grid, kThreads, 0, s>>>(out, c, pw, ph, pd, cw, ch, cd, border[0], border[1], border[2], W, H, D,
                                                    probs, reinterpret_cast<long long*>(labels_i64), labels_u8);
    } else {
        dim3 grid(static_cast<unsigned>((1LL * H * D + kThreads - 1) / kThreads), static_cast<unsigned>(W));
        finalize_kernel<1><<<grid, kThreads, 0, s>>>(out, c, pw, ph, pd, cw, ch, cd, border[0], border[1], border[2], W, H, D,
                                                    probs, reinterpret_cast<long long*>(labels_i64), labels_u8);
    }
    return check_launch("finalize");
}

int b200seg_argmax(const float* probs, int32_t c, int64_t voxels, int64_t* labels_i64, uint8_t* labels_u8,
                   void* stream) {
    B200SEG_CHECK_ARG(probs && c > 0 && voxels > 0 && (labels_i64 || labels_u8), "argmax: bad arguments");
    B200SEG_CHECK_ARG(labels_u8 == nullptr || c <= 256, "argmax: uint8 labels need <= 256 classes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (voxels % 4 == 0 && (reinterpret_cast<uintptr_t>(probs) & 15) == 0) {
        long long total = voxels / 4;
        argmax_kernel<4><<<blocks_for(total), kThreads, 0, s>>>(probs, c, voxels,
                                                               reinterpret_cast<long long*>(labels_i64), labels_u8,
                                                               total);
    } else {
        argmax_kernel<1><<<blocks_for(voxels), kThreads, 0, s>>>(probs, c, voxels,
                                                                reinterpret_cast<long long*>(labels_i64), labels_u8,
                                                                voxels);
    }
    return check_launch("argmax");
}

int b200seg_confusion(const void* pred, const void* target, int32_t label_bytes, int64_t voxels,
                      int32_t num_classes, int64_t* cm, void* stream) {
    B200SEG_CHECK_ARG(pred && target && cm && voxels > 0, "confusion: bad arguments");
    B200SEG_CHECK_ARG(label_bytes == 1 || label_bytes == 8, "confusion: label_bytes must be 1 or 8");
    B200SEG_CHECK_ARG(num_classes >= 1 && num_classes <= 40, "confusion: num_classes %d not in [1,40]", num_classes);
    B200SEG_CHECK_ARG((reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0,
                      "confusion: label maps must be 16-byte aligned");
    int dev = 0, sms = 148;
    B200SEG_CHECK_CUDA(cudaGetDevice(&dev));
    B200SEG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int bins = num_classes * num_classes;
    const size_t smem = static_cast<size_t>(kThreads / 32) * bins * sizeof(unsigned int);   // <= 51 KB
    const long long per = label_bytes == 1 ? 16 : 2;
    // a warp's counter holds up to 2^32 voxels: one CTA sees voxels / grid of them
    long long want = (voxels / per + 4LL * kThreads - 1) / (4LL * kThreads);
    const int per_sm = smem > 24 * 1024 ? 4 : 8;
    unsigned grid = static_cast<unsigned>(want < 1LL * sms * per_sm ? (want > 0 ? want : 1) : 1LL * sms * per_sm);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (label_bytes == 1) {
        B200SEG_CHECK_CUDA(cudaFuncSetAttribute(confusion_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(smem)));
        confusion_kernel<uint8_t><<<grid, kThreads, smem, s>>>(static_cast<const uint8_t*>(pred),
                                                                static_cast<const uint8_t*>(target), voxels,
                                                                num_classes,
                                                                reinterpret_cast<unsigned long long*>(cm));
    } else {
        B200SEG_CHECK_CUDA(cudaFuncSetAttribute(confusion_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(smem)));
        confusion_kernel<long long><<<grid, kThreads, smem, s>>>(static_cast<const long long*>(pred),
                                                                  static_cast<const long long*>(target), voxels,
                                                                  num_classes,
                                                                  reinterpret_cast<unsigned long long*>(cm));
    }
    return check_launch("confusion");
}

}  // extern "C"
