// Training-step kernels (fp32, blocked activations) -- the device side of the reference's trainer step,
// segmentation_trainer.py:162-180: model.train() forward (BatchNorm3d with BATCH statistics), loss, backward.
// The forward convolutions and the data gradients (dgrad) are b200seg_conv3d_direct launches (a dgrad IS a convolution
// with re-arranged weights: flipped taps for the 3x3x3 layers, the transposed geometry for the strided ones and vice
// versa); this file holds what is new for training:
//   wgrad            G[tap][a][b] = sum over (n, pos) of A[a](pos) * B[b](stride * pos + tap - pad)
//   channel_moments  per-channel mean / biased variance over (N, Z, Y, X)            (BatchNorm3d training statistics)
//   affine_act       act(scale * z + shift) (+ residual)                             (BN apply + activation + res add)
//   bn_backward_*    the two reductions and the data gradient of BN + activation
//   softmax_backward dlogits = p * (dp - sum_c dp_c p_c), NCDHW -> blocked
// All reductions are two-stage with per-block partials summed in a fixed order (deterministic, no atomics); the
// per-channel sums are carried in double.
#include <cstdlib>

#include "common.cuh"

namespace b200seg {

constexpr int kTrThreads = 256;

template <typename T>
__device__ __forceinline__ Vec8 ldv(const DView& v, long long idx) { return load_vec8<T>(v.data, idx); }

#define TRAIN_DISPATCH(dtype, ...)        \
    do {                                   \
        if ((dtype) == B200SEG_F32) {      \
            using T = float;               \
            __VA_ARGS__;                   \
        } else {                           \
            using T = __nv_bfloat16;       \
            __VA_ARGS__;                   \
        }                                  \
    } while (0)

// ------------------------------------------------------------------------------------------- two-sum reductions
// MODE 0: (x, x^2) of view a.   MODE 1: (g, g * xhat) with g = dy * act'(scale * z + shift), xhat = (z - mean) * rstd;
// a = dy, b = z.
struct ReduceParams {
    const float* scale;
    const float* shift;
    const float* slope;
    const float* mean;
    const float* rstd;
};

template <typename T, int MODE>
__global__ void __launch_bounds__(kTrThreads)
chan_reduce_partial_kernel(DView a, DView b, ReduceParams p, double* __restrict__ partial, int nblk) {
    __shared__ double red[kTrThreads / 32][16];
    const int cc = blockIdx.y, blk = blockIdx.x;
    const long long total = 1LL * a.n * a.chunk_stride;
    const long long per = (total + nblk - 1) / nblk;
    const long long t0 = blk * per, t1 = min(t0 + per, total);
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    float sc[8], sh[8], sl[8], mu[8], rs[8];
    if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = __ldg(p.scale + cc * 8 + j);
            sh[j] = __ldg(p.shift + cc * 8 + j);
            sl[j] = __ldg(p.slope + cc * 8 + j);
            mu[j] = __ldg(p.mean + cc * 8 + j);
            rs[j] = __ldg(p.rstd + cc * 8 + j);
        }
    }
    for (long long t = t0 + threadIdx.x; t < t1; t += kTrThreads) {
        const long long n = t / a.chunk_stride, v = t - n * a.chunk_stride;
        const Vec8 x = ldv<T>(a, n * a.sample_stride + (a.c8_off + cc) * a.chunk_stride + v);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s1[j] += x.v[j];
                s2[j] = fmaf(x.v[j], x.v[j], s2[j]);
            }
        } else {
            const Vec8 z = ldv<T>(b, n * b.sample_stride + (b.c8_off + cc) * b.chunk_stride + v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float y = fmaf(z.v[j], sc[j], sh[j]);
                const float g = y > 0.f ? x.v[j] : x.v[j] * sl[j];
                s1[j] += g;
                s2[j] = fmaf(g, (z.v[j] - mu[j]) * rs[j], s2[j]);
            }
        }
    }
    double d[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        d[j] = s1[j];
        d[8 + j] = s2[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) red[warp][j] = d[j];
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double s = 0.0;
        for (int w = 0; w < kTrThreads / 32; ++w) s += red[w][threadIdx.x];
        partial[(static_cast<long long>(cc) * nblk + blk) * 16 + threadIdx.x] = s;
    }
}

// moments != 0: out0 = s1 / count (mean), out1 = s2 / count - mean^2 (biased variance, >= 0); else the raw sums.
// One warp per channel: lanes stride over the per-block partials, then a fixed-order shuffle tree (the first version
// walked up to 256 partials serially in one thread: 25 us per launch, 56 launches per training step).
__global__ void __launch_bounds__(32)
chan_reduce_finish_kernel(const double* __restrict__ partial, int nblk, int channels, double count, int moments,
                          float* __restrict__ out0, float* __restrict__ out1) {
    const int c = blockIdx.x;
    if (c >= channels) return;
    const int cc = c / 8, j = c % 8, lane = threadIdx.x;
    double s1 = 0.0, s2 = 0.0;
    for (int b = lane; b < nblk; b += 32) {
        s1 += partial[(static_cast<long long>(cc) * nblk + b) * 16 + j];
        s2 += partial[(static_cast<long long>(cc) * nblk + b) * 16 + 8 + j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane != 0) return;
    if (moments) {
        const double mean = s1 / count;
        double var = s2 / count - mean * mean;
        out0[c] = static_cast<float>(mean);
        out1[c] = static_cast<float>(var > 0.0 ? var : 0.0);
    } else {
        out0[c] = static_cast<float>(s1);
        out1[c] = static_cast<float>(s2);
    }
}

// ------------------------------------------------------------------------------------------- elementwise
template <typename T>
__global__ void __launch_bounds__(kTrThreads)
affine_act_kernel(DView src, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ slope, DView residual, DView dst, int c8n, long long total) {
    const long long t = blockIdx.x * 1LL * kTrThreads + threadIdx.x;
    if (t >= total) return;
    const long long v = t % src.chunk_stride;
    const long long r = t / src.chunk_stride;
    const int cc = static_cast<int>(r % c8n);
    const long long n = r / c8n;
    Vec8 a = ldv<T>(src, n * src.sample_stride + (src.c8_off + cc) * src.chunk_stride + v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float y = fmaf(a.v[j], __ldg(scale + cc * 8 + j), __ldg(shift + cc * 8 + j));
        a.v[j] = y > 0.f ? y : y * __ldg(slope + cc * 8 + j);
    }
    if (residual.data != nullptr) {
        const Vec8 q = ldv<T>(residual, n * residual.sample_stride + (residual.c8_off + cc) * residual.chunk_stride + v);
#pragma unroll
        for (int j = 0; j < 8; ++j) a.v[j] += q.v[j];
    }
    store_vec8<T>(dst.data, n * dst.sample_stride + (dst.c8_off + cc) * dst.chunk_stride + v, a);
}

// nn.Dropout3d (forward and backward are the same map): dst = src * mask[n][c], mask = 0 or 1 / (1 - p) per (sample, channel)
template <typename T>
__global__ void __launch_bounds__(kTrThreads)
channel_scale_kernel(DView src, const float* __restrict__ mask, int mask_stride, DView dst, int c8n, long long total) {
    const long long t = blockIdx.x * 1LL * kTrThreads + threadIdx.x;
    if (t >= total) return;
    const long long v = t % src.chunk_stride;
    const long long r = t / src.chunk_stride;
    const int cc = static_cast<int>(r % c8n);
    const long long n = r / c8n;
    Vec8 a = ldv<T>(src, n * src.sample_stride + (src.c8_off + cc) * src.chunk_stride + v);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] *= __ldg(mask + n * mask_stride + cc * 8 + j);
    store_vec8<T>(dst.data, n * dst.sample_stride + (dst.c8_off + cc) * dst.chunk_stride + v, a);
}

// dz = has_norm ? scale * (g - c1 - xhat * c2) : g,   g = dy * act'(scale * z + shift); c1 = sum g / M, c2 = sum g xhat / M
template <typename T>
__global__ void __launch_bounds__(kTrThreads)
bn_backward_apply_kernel(DView dy, DView z, ReduceParams p, const float* __restrict__ sum_g,
                         const float* __restrict__ sum_gx, float inv_count, int has_norm, DView dst, int c8n,
                         long long total) {
    const long long t = blockIdx.x * 1LL * kTrThreads + threadIdx.x;
    if (t >= total) return;
    const long long v = t % dy.chunk_stride;
    const long long r = t / dy.chunk_stride;
    const int cc = static_cast<int>(r % c8n);
    const long long n = r / c8n;
    const Vec8 d = ldv<T>(dy, n * dy.sample_stride + (dy.c8_off + cc) * dy.chunk_stride + v);
    const Vec8 zz = ldv<T>(z, n * z.sample_stride + (z.c8_off + cc) * z.chunk_stride + v);
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cc * 8 + j;
        const float sc = __ldg(p.scale + c);
        const float y = fmaf(zz.v[j], sc, __ldg(p.shift + c));
        const float g = y > 0.f ? d.v[j] : d.v[j] * __ldg(p.slope + c);
        if (has_norm) {
            const float xh = (zz.v[j] - __ldg(p.mean + c)) * __ldg(p.rstd + c);
            o.v[j] = sc * (g - __ldg(sum_g + c) * inv_count - xh * (__ldg(sum_gx + c) * inv_count));
        } else {
            o.v[j] = g;
        }
    }
    store_vec8<T>(dst.data, n * dst.sample_stride + (dst.c8_off + cc) * dst.chunk_stride + v, o);
}

// probs / dprobs: fp32 NCDHW [N][C][vox]; dst: blocked view with C channels (padding channels written as 0)
template <typename T>
__global__ void __launch_bounds__(kTrThreads)
softmax_backward_kernel(const float* __restrict__ probs, const float* __restrict__ dprobs, int C, long long vox,
                        int softmax, DView dst, long long total) {
    const long long t = blockIdx.x * 1LL * kTrThreads + threadIdx.x;
    if (t >= total) return;
    const long long n = t / vox, v = t - n * vox;
    const float* pp = probs + n * C * vox + v;
    const float* dp = dprobs + n * C * vox + v;
    float dot = 0.f;
    if (softmax)
        for (int c = 0; c < C; ++c) dot = fmaf(__ldg(dp + c * vox), __ldg(pp + c * vox), dot);
    const int c8n = (C + 7) / 8;
    for (int cc = 0; cc < c8n; ++cc) {
        Vec8 o;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cc * 8 + j;
            float val = 0.f;
            if (c < C) val = softmax ? __ldg(pp + c * vox) * (__ldg(dp + c * vox) - dot) : __ldg(dp + c * vox);
            o.v[j] = val;
        }
        store_vec8<T>(dst.data, n * dst.sample_stride + (dst.c8_off + cc) * dst.chunk_stride + v, o);
    }
}

// ------------------------------------------------------------------------------------------- pooling / upsampling adjoints
// nn.AvgPool3d(2) backward: dx(v) = dy(v / 2) / 8 (+ add(v): the skip connection's gradient)
template <typename T>
__global__ void __launch_bounds__(kTrThreads)
avgpool2_backward_kernel(DView dy, DView add, DView dx, int c8n, long long total) {
    const long long t = blockIdx.x * 1LL * kTrThreads + threadIdx.x;
    if (t >= total) return;
    const int x = static_cast<int>(t % dx.x);
    long long r = t / dx.x;
    const int y = static_cast<int>(r % dx.y);
    r /= dx.y;
    const int z = static_cast<int>(r % dx.z);
    r /= dx.z;
    const int cc = static_cast<int>(r % c8n);
    const int n = static_cast<int>(r / c8n);
    Vec8 o;
    if ((z >> 1) < dy.z && (y >> 1) < dy.y && (x >> 1) < dy.x) {
        o = ldv<T>(dy, vox_index(dy, n, cc, z >> 1, y >> 1, x >> 1));
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = o.v[j] / 8.0f;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] = 0.f;
    }
    if (add.data != nullptr) {
        const Vec8 q = ldv<T>(add, vox_index(add, n, cc, z, y, x));
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] += q.v[j];
    }
    store_vec8<T>(dx.data, vox_index(dx, n, cc, z, y, x), o);
}

// Same interpolation coefficients as upsample_trilinear2_kernel (hbm_kernels.cu): align_corners = True
__device__ __forceinline__ void tr_lin_coeff(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
    const float scale = out_size > 1 ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
    const float src = scale * static_cast<float>(dst);
    i0 = static_cast<int>(src);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - static_cast<float>(i0);
    l0 = 1.f - l1;
}

constexpr int kMaxAdj = 8;
// outputs o of one axis that read input index i, with their weights: w = [i0(o) == i] l0 + [i1(o) == i] l1
__device__ __forceinline__ int adjoint_taps(int i, int in_size, int out_size, int (&o_idx)[kMaxAdj], float (&w)[kMaxAdj]) {
    int count = 0;
    const float inv = in_size > 1 ? static_cast<float>(out_size - 1) / static_cast<float>(in_size - 1) : 0.f;
    int lo = static_cast<int>(floorf((i - 1) * inv)) - 1, hi = static_cast<int>(ceilf((i + 1) * inv)) + 1;
    if (in_size == 1) {
        lo = 0;
        hi = out_size - 1;
    }
    lo = max(lo, 0);
    hi = min(hi, out_size - 1);
    for (int o = lo; o <= hi; ++o) {
        int i0, i1;
        float l0, l1;
        tr_lin_coeff(o, in_size, out_size, i0, i1, l0, l1);
        const float ww = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
        if (ww != 0.f && count < kMaxAdj) {
            o_idx[count] = o;
            w[count] = ww;
            ++count;
        }
    }
    return count;
}

// adjoint of nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True): dx(i) = sum_o W(o, i) dy(o), as a gather
template <typename T>
__global__ void __launch_bounds__(kTrThreads)
upsample_trilinear2_backward_kernel(DView dy, DView dx, int c8n, long long total) {
    const long long t = blockIdx.x * 1LL * kTrThreads + threadIdx.x;
    if (t >= total) return;
    const int x = static_cast<int>(t % dx.x);
    long long r = t / dx.x;
    const int y = static_cast<int>(r % dx.y);
    r /= dx.y;
    const int z = static_cast<int>(r % dx.z);
    r /= dx.z;
    const int cc = static_cast<int>(r % c8n);
    const int n = static_cast<int>(r / c8n);
    int oz[kMaxAdj], oy[kMaxAdj], ox[kMaxAdj];
    float wz[kMaxAdj], wy[kMaxAdj], wx[kMaxAdj];
    const int nz = adjoint_taps(z, dx.z, dy.z, oz, wz);
    const int ny = adjoint_taps(y, dx.y, dy.y, oy, wy);
    const int nx = adjoint_taps(x, dx.x, dy.x, ox, wx);
    Vec8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    for (int a = 0; a < nz; ++a)
        for (int b = 0; b < ny; ++b) {
            const float wzy = wz[a] * wy[b];
            for (int c = 0; c < nx; ++c) {
                const Vec8 v = ldv<T>(dy, vox_index(dy, n, cc, oz[a], oy[b], ox[c]));
                const float ww = wzy * wx[c];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(ww, v.v[j], acc.v[j]);
            }
        }
    store_vec8<T>(dx.data, vox_index(dx, n, cc, z, y, x), acc);
}

// ------------------------------------------------------------------------------------------- wgrad
// One block = the K*K in-plane taps of one tz (one warp per tap) for one (A chunk, B chunk) pair; lanes = 32
// consecutive positions of the flattened (y, x) plane of the A grid, so both operand loads of a warp are contiguous
// 1 KB segments (A) / contiguous up to row wraps (B); the K*K warps of a block read the same A vectors and
// neighbouring B vectors, which L1 serves.  Each thread carries the 8 x 8 outer-product accumulator of its tap.
struct WgradGeom {
    int stride, pad;
    int pz, py, px;        // A (position) grid
    int plane_blocks;      // ceil(py * px / 32)
    int items;             // n * pz * plane_blocks
    int cb8n;              // B chunks
    float inv_px;
};

template <typename T, int K, int H>
__global__ void __launch_bounds__(K * K * 32 * H, (K == 3 && H == 1) ? 2 : 1)
wgrad_partial_kernel(DView A, DView B, WgradGeom g, float* __restrict__ partial) {
    // H groups of K*K tap-warps: group h takes every H-th trip over the plane and writes its own partial slice
    const int half = threadIdx.x / (K * K * 32);
    const int warp = (threadIdx.x / 32) % (K * K), lane = threadIdx.x % 32;
    const int ty = warp / K, tx = warp % K, tz = blockIdx.z;
    const int ca = blockIdx.y / g.cb8n, cb = blockIdx.y % g.cb8n;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int plane = g.py * g.px;
    // A block owns a contiguous range of (n, z) rows of the position grid and walks each plane 2 x 32 positions per
    // trip: both operand pairs are loaded before the first FMA (latency), and (y, x) of a lane advance incrementally
    // (first version: a division and 64-bit index math per item cost twice the FMAs -- ncu: 33 % of the instructions
    // were FFMA, issue slots 32 % busy).
    constexpr int U = 2;
    const int rows = A.n * g.pz;
    const int per = (rows + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * per, r1 = min(r0 + per, rows);
    for (int r = r0; r < r1; ++r) {
        const int n = r / g.pz, z = r - n * g.pz;
        const int bz = g.stride * z + tz - g.pad;
        if (bz < 0 || bz >= B.z) continue;                       // block-uniform
        const long long arow = vox_index(A, n, ca, z, 0, 0), bplane = vox_index(B, n, cb, bz, 0, 0);
        int py[U], px[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = (half * U + u) * 32 + lane;
            py[u] = p / g.px;
            px[u] = p - py[u] * g.px;
        }
        for (int p0 = half * 32 * U; p0 < plane; p0 += 32 * U * H) {
            Vec8 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = p0 + u * 32 + lane;
                const int by = g.stride * py[u] + ty - g.pad, bx = g.stride * px[u] + tx - g.pad;
                const bool ok = p < plane && by >= 0 && by < B.y && bx >= 0 && bx < B.x;
                if (ok) {
                    a[u] = ldv<T>(A, arow + p);
                    b[u] = ldv<T>(B, bplane + by * B.x + bx);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[u].v[i] = b[u].v[i] = 0.f;
                }
                px[u] += 32 * U * H;
                while (px[u] >= g.px) {
                    px[u] -= g.px;
                    ++py[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[u].v[i], b[u].v[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], o);
    if (lane == 0) {
        const int tap = (tz * K + ty) * K + tx;
        float* dst = partial + ((static_cast<long long>(blockIdx.x * H + half) * (K * K * K) + tap) * gridDim.y + blockIdx.y) * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[i * 8 + j] = acc[i][j];
    }
}

// ------------------------------------------------------------------------------------------- wgrad on the tensor cores
// bf16 operands (mixed-precision training): the same sum as wgrad_partial_kernel as warp-level MMAs
// (mma.sync.m16n8k16, bf16 x bf16 -> fp32).  Per tap it is a GEMM  G_tap[a][b] = sum_pos A[a](pos) * B[b](s*pos + tap - pad)
// with K = positions.  A block owns one tz, a group of <= 48 A channels x <= 40 B channels and a contiguous range of
// (n, z, y) rows; one warp per in-plane tap holds the whole 48 x 40 fp32 accumulator of its tap (60 registers).  Per
// row the block stages the A row and the K halo rows of B in shared memory as [chunk][position] 16-byte vectors --
// exactly the 8 x 8 b16 tiles ldmatrix reads -- and .trans hands every lane the (channel, 2 positions) pairs the
// fragments want, so no transposed copy of the activations is ever made.  Structural zeros (halo, channel padding,
// positions past the row end) are zero vectors in shared memory.
constexpr int kMmaA = 48, kMmaB = 40;      // channels per block: 3 m16 tiles x 5 n8 tiles
constexpr int kMmaMaxX = 256;

struct WgradMmaGeom {
    int stride, pad;
    int pz, py, px;     // A (position) grid
    int xp;             // px rounded up to 16
    int bw;             // B positions staged per halo row: stride * (xp - 1) + K
    int a_groups, b_groups;
    int a_chunks, b_chunks;   // total chunks of A / B
    int rows;           // n * pz * py
};

__device__ __forceinline__ void ldmatrix_x4_trans(unsigned (&r)[4], const void* smem_row) {
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(smem_row));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT / NT: m16 / n8 tiles actually computed (3 x 5 in general; 1 when the A / B group holds at most 2 / 1 chunks -- the
// first layer's 2-channel input and the 2-class logits would otherwise spend 2/3 resp. 4/5 of their MMAs on zeros)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(addr), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <int K, int MT, int NT>
__global__ void __launch_bounds__(K * K * 32, K == 3 ? 2 : 1)
wgrad_mma_kernel(DView A, DView B, WgradMmaGeom g, float* __restrict__ partial) {
    constexpr int S = K == 3 ? 1 : 2;            // the two geometries: 3^3 stride 1, 4^3 stride 2
    constexpr int RING = K + S;                  // halo rows of row y AND the S new rows of row y + 1
    extern __shared__ uint4 wg_smem[];
    uint4* sA = wg_smem;                         // [2 buffers][6 chunks][xp]
    uint4* sB = wg_smem + 12 * g.xp;             // [RING rows][6 chunk slots][bw]  (5 used; slot 5 = zeros for the x4 loads)
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int ty = warp / K, tx = warp % K, tz = blockIdx.z;
    const int ag = blockIdx.y / g.b_groups, bg = blockIdx.y % g.b_groups;
    const int ca0 = ag * 6, cb0 = bg * 5;
    float acc[MT][NT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[m][n][e] = 0.f;
    const int per = (g.rows + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * per, r1 = min(r0 + per, g.rows);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4* Adata = reinterpret_cast<const uint4*>(A.data);
    const uint4* Bdata = reinterpret_cast<const uint4*>(B.data);
    const int mat = lane >> 3, jrow = lane & 7;
    const int nwarps = K * K;
    // Everything that is structurally zero (halo columns, positions past the row end, missing chunks, the sixth B slot)
    // is written once; per row only the data sections [0, px) of A and [pad, pad + B.x) of the NEW halo rows of B are
    // copied, one contiguous global line per warp trip (first version: per-vector index arithmetic in the staging loop
    // cost twice the instructions of the MMA loop).  The halo rows live in a ring of K + S slots (row by in slot
    // (by + pad) % RING) and the A row is double-buffered, so the copies of row y + 1 (cp.async) run under the MMAs of
    // row y and there is ONE barrier per row (second version: two barriers and synchronous copies -- ncu: tensor pipe
    // 41 % active, 4.2 barrier + 2.4 scoreboard stall cycles per issue).
    for (int i = threadIdx.x; i < 12 * g.xp + RING * 6 * g.bw; i += blockDim.x) wg_smem[i] = zero;
    const int na = min(6, g.a_chunks - ca0), nb = min(5, g.b_chunks - cb0);

    // copies of one row: the A row into buffer ab, the halo rows [first, last] of B into their ring slots
    auto issue = [&](int n, int z, int y, int bz, int ab, int first, int last) {
        const int lines = na + (last - first + 1) * nb;
        for (int line = warp; line < lines; line += nwarps) {
            if (line < na) {
                const uint4* src = Adata + vox_index(A, n, ca0 + line, z, y, 0);
                uint4* dst = sA + (ab * 6 + line) * g.xp;
                for (int x = lane; x < g.px; x += 32) cp_async16(dst + x, src + x);
            } else {
                const int q = line - na;
                const int rr = q / nb, c = q - rr * nb;
                const int by = first + rr;
                uint4* dst = sB + (((by + g.pad) % RING) * 6 + c) * g.bw + g.pad;
                if (by >= 0 && by < B.y) {
                    const uint4* src = Bdata + vox_index(B, n, cb0 + c, bz, by, 0);
                    for (int x = lane; x < B.x; x += 32) cp_async16(dst + x, src + x);
                } else {
                    for (int x = lane; x < B.x; x += 32) dst[x] = zero;
                }
            }
        }
    };

    int r = r0;
    while (r < r1) {
        const int nz = r / g.py, yb = r - nz * g.py;
        const int yend = min(g.py, yb + (r1 - r));                // rows [yb, yend) of plane nz belong to this block
        const int z = nz % g.pz, n = nz / g.pz;
        const int bz = S * z + tz - g.pad;
        r += yend - yb;
        if (bz < 0 || bz >= B.z) continue;                        // block-uniform
        __syncthreads();                                           // every warp is done with the previous plane
        issue(n, z, yb, bz, 0, S * yb - g.pad, S * yb - g.pad + K - 1);
        cp_async_wait_all();
        __syncthreads();
        for (int y = yb; y < yend; ++y) {
            const int ab = (y - yb) & 1;
            const int hi = S * y - g.pad + K - 1;
            if (y + 1 < yend) issue(n, z, y + 1, bz, ab ^ 1, hi + 1, hi + S);
            const uint4* arow = sA + ab * 6 * g.xp;
            const uint4* brow = sB + static_cast<long long>((S * y + ty) % RING) * 6 * g.bw;
            for (int ks = 0; ks < g.xp; ks += 16) {
                unsigned af[MT][4], bf[(NT + 1) / 2][4];
                // matrices of one x4 load: (chunk 2m, pos 0-7), (chunk 2m+1, pos 0-7), (chunk 2m, pos 8-15), (chunk 2m+1, pos 8-15)
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    ldmatrix_x4_trans(af[m], arow + (2 * m + (mat & 1)) * g.xp + ks + (mat >> 1) * 8 + jrow);
                // B: (chunk 2q, pos 0-7), (chunk 2q, pos 8-15), (chunk 2q+1, pos 0-7), (chunk 2q+1, pos 8-15)
#pragma unroll
                for (int q = 0; q < (NT + 1) / 2; ++q)
                    ldmatrix_x4_trans(bf[q], brow + (2 * q + (mat >> 1)) * g.bw + S * (ks + (mat & 1) * 8 + jrow) + tx);
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int nn = 0; nn < NT; ++nn)
                        mma_bf16_16816(acc[m][nn], af[m], bf[nn >> 1][(nn & 1) * 2], bf[nn >> 1][(nn & 1) * 2 + 1]);
            }
            cp_async_wait_all();
            __syncthreads();
        }
    }
    // accumulator fragment: c0,c1 = (row g, cols 2t, 2t+1); c2,c3 = (row g + 8, same cols)
    const int tap = (tz * K + ty) * K + tx;
    float* dst = partial + ((static_cast<long long>(blockIdx.x) * (K * K * K) + tap) * gridDim.y + blockIdx.y) * (kMmaA * kMmaB);
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int m = 0; m < 3; ++m)
#pragma unroll
        for (int nn = 0; nn < 5; ++nn) {
            const int a = m * 16 + gq, b = nn * 8 + 2 * tq;
            const bool live = m < MT && nn < NT;
            dst[a * kMmaB + b] = live ? acc[m < MT ? m : 0][nn < NT ? nn : 0][0] : 0.f;
            dst[a * kMmaB + b + 1] = live ? acc[m < MT ? m : 0][nn < NT ? nn : 0][1] : 0.f;
            dst[(a + 8) * kMmaB + b] = live ? acc[m < MT ? m : 0][nn < NT ? nn : 0][2] : 0.f;
            dst[(a + 8) * kMmaB + b + 1] = live ? acc[m < MT ? m : 0][nn < NT ? nn : 0][3] : 0.f;
        }
}

// grad[tap][a][b] (rows a_pad, b_pad) = sum over slices of the per-block 48 x 40 tiles
__global__ void wgrad_mma_finish_kernel(const float* __restrict__ partial, int slices, int taps, int a_groups, int b_groups,
                                        int a_pad, int b_pad, float* __restrict__ grad, long long total) {
    const long long t = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = static_cast<int>(t % b_pad);
    long long r = t / b_pad;
    const int a = static_cast<int>(r % a_pad);
    const int tap = static_cast<int>(r / a_pad);
    const int ag = a / kMmaA, bg = b / kMmaB;
    const int groups = a_groups * b_groups;
    const int e = (a - ag * kMmaA) * kMmaB + (b - bg * kMmaB);
    float s = 0.f;
#pragma unroll 8
    for (int sl = 0; sl < slices; ++sl)
        s += __ldg(partial + ((static_cast<long long>(sl) * taps + tap) * groups + ag * b_groups + bg) * (kMmaA * kMmaB) + e);
    grad[t] = s;
}

// G[tap][ca * 8 + i][cb * 8 + j] = sum over slices, in slice order
__global__ void wgrad_finish_kernel(const float* __restrict__ partial, int slices, int taps, int pairs, int cb8n,
                                    float* __restrict__ grad, long long total) {
    const long long t = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int e = static_cast<int>(t % 64);
    const long long r = t / 64;
    const int pair = static_cast<int>(r % pairs);
    const int tap = static_cast<int>(r / pairs);
    float s = 0.f;
    for (int sl = 0; sl < slices; ++sl) s += partial[((static_cast<long long>(sl) * taps + tap) * pairs + pair) * 64 + e];
    const int ca = pair / cb8n, cb = pair % cb8n;
    const int ca8n = pairs / cb8n;
    grad[(static_cast<long long>(tap) * ca8n * 8 + ca * 8 + e / 8) * (cb8n * 8) + cb * 8 + e % 8] = s;
}

static int check_view(const b200seg_view& v, const char* name) {     // fp32 or bf16
    return validate_view(v, name);
}

static bool same_extent(const b200seg_view& a, const b200seg_view& b) {
    return a.n == b.n && a.z == b.z && a.y == b.y && a.x == b.x && (a.c + 7) / 8 == (b.c + 7) / 8;
}

static int reduce_blocks(const b200seg_view& v) {
    const long long total = 1LL * v.n * v.z * v.y * v.x;
    long long nblk = (total + 4 * kTrThreads - 1) / (4 * kTrThreads);
    return static_cast<int>(nblk < 1 ? 1 : (nblk > 256 ? 256 : nblk));
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int64_t b200seg_train_scratch_bytes(int32_t channels) {
    // doubles: [c8][256 blocks][16]
    return static_cast<int64_t>((channels + 7) / 8) * 256 * 16 * 8;
}

extern "C" int b200seg_channel_moments(b200seg_view x, void* scratch, float* mean, float* var, void* stream) {
    int rc = check_view(x, "channel_moments x");
    if (rc) return rc;
    B200SEG_CHECK_ARG(scratch && mean && var, "channel_moments: null argument");
    const int c8n = (x.c + 7) / 8, nblk = reduce_blocks(x);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DView dx = make_dview(x);
    ReduceParams p{};
    TRAIN_DISPATCH(x.dtype, (chan_reduce_partial_kernel<T, 0><<<dim3(nblk, c8n), kTrThreads, 0, s>>>(
                                dx, dx, p, static_cast<double*>(scratch), nblk)));
    const double count = 1.0 * x.n * x.z * x.y * x.x;
    chan_reduce_finish_kernel<<<c8n * 8, 32, 0, s>>>(static_cast<const double*>(scratch), nblk, c8n * 8, count, 1, mean, var);
    return check_launch("channel_moments");
}

extern "C" int b200seg_affine_act(b200seg_view src, const float* scale, const float* shift, const float* slope,
                                  b200seg_view residual, b200seg_view dst, void* stream) {
    int rc = check_view(src, "affine_act src");
    if (rc) return rc;
    rc = check_view(dst, "affine_act dst");
    if (rc) return rc;
    B200SEG_CHECK_ARG(scale && shift && slope && same_extent(src, dst), "affine_act: bad arguments");
    DView dr = null_dview();
    if (residual.data != nullptr) {
        rc = check_view(residual, "affine_act residual");
        if (rc) return rc;
        B200SEG_CHECK_ARG(same_extent(src, residual), "affine_act: residual extent differs");
        dr = make_dview(residual);
    }
    const int c8n = (src.c + 7) / 8;
    const long long total = 1LL * src.n * c8n * src.z * src.y * src.x;
    const unsigned blocks = static_cast<unsigned>((total + kTrThreads - 1) / kTrThreads);
    B200SEG_CHECK_ARG(src.dtype == dst.dtype && (residual.data == nullptr || residual.dtype == src.dtype),
                      "affine_act: views must share one dtype");
    TRAIN_DISPATCH(src.dtype, (affine_act_kernel<T><<<blocks, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                  make_dview(src), scale, shift, slope, dr, make_dview(dst), c8n, total)));
    return check_launch("affine_act");
}

extern "C" int b200seg_bn_backward(b200seg_view dy, b200seg_view z, const float* scale, const float* shift,
                                   const float* slope, const float* mean, const float* rstd, int32_t has_norm,
                                   void* scratch, float* sum_g, float* sum_gx, b200seg_view dz, void* stream) {
    int rc = check_view(dy, "bn_backward dy");
    if (rc) return rc;
    rc = check_view(z, "bn_backward z");
    if (rc) return rc;
    rc = check_view(dz, "bn_backward dz");
    if (rc) return rc;
    B200SEG_CHECK_ARG(scale && shift && slope && mean && rstd && scratch && sum_g && sum_gx, "bn_backward: null argument");
    B200SEG_CHECK_ARG(same_extent(dy, z) && same_extent(dy, dz), "bn_backward: extents differ");
    const int c8n = (dy.c + 7) / 8, nblk = reduce_blocks(dy);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ReduceParams p{scale, shift, slope, mean, rstd};
    DView ddy = make_dview(dy), dzv = make_dview(z);
    B200SEG_CHECK_ARG(dy.dtype == z.dtype && dy.dtype == dz.dtype, "bn_backward: views must share one dtype");
    TRAIN_DISPATCH(dy.dtype, (chan_reduce_partial_kernel<T, 1><<<dim3(nblk, c8n), kTrThreads, 0, s>>>(
                                 ddy, dzv, p, static_cast<double*>(scratch), nblk)));
    chan_reduce_finish_kernel<<<c8n * 8, 32, 0, s>>>(static_cast<const double*>(scratch), nblk, c8n * 8, 1.0, 0, sum_g, sum_gx);
    const double count = 1.0 * dy.n * dy.z * dy.y * dy.x;
    const long long total = 1LL * dy.n * c8n * dy.z * dy.y * dy.x;
    const unsigned blocks = static_cast<unsigned>((total + kTrThreads - 1) / kTrThreads);
    TRAIN_DISPATCH(dy.dtype, (bn_backward_apply_kernel<T><<<blocks, kTrThreads, 0, s>>>(
                                 ddy, dzv, p, sum_g, sum_gx, static_cast<float>(1.0 / count), has_norm, make_dview(dz), c8n,
                                 total)));
    return check_launch("bn_backward");
}

extern "C" int b200seg_softmax_backward(const float* probs, const float* dprobs, int32_t n, int32_t c, int32_t softmax,
                                        b200seg_view dst, void* stream) {
    int rc = check_view(dst, "softmax_backward dst");
    if (rc) return rc;
    B200SEG_CHECK_ARG(probs && dprobs && n == dst.n && c == dst.c, "softmax_backward: bad arguments");
    const long long vox = 1LL * dst.z * dst.y * dst.x, total = vox * n;
    const unsigned blocks = static_cast<unsigned>((total + kTrThreads - 1) / kTrThreads);
    TRAIN_DISPATCH(dst.dtype, (softmax_backward_kernel<T><<<blocks, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                  probs, dprobs, c, vox, softmax, make_dview(dst), total)));
    return check_launch("softmax_backward");
}

static long long simt_slices(long long pairs, int ksize) {
    // at most three full waves of resident blocks (2 per SM for the 9-warp 3^3 kernel, 1 per SM for the 16-warp 4^3
    // one); rounding DOWN keeps a nearly empty fourth wave from forming (6 * 148 / (pairs * k), rounded up, did that)
    const long long slots = (ksize == 3 ? 2 : 1) * 148;
    const long long slices = 3 * slots / (pairs * ksize);
    return slices < 1 ? 1 : slices;
}

static int mma_slices(int groups, int ksize) {
    // one full wave of resident blocks: 2 per SM for the 9-warp 3^3 kernel, 1 per SM for the 16-warp 4^3 kernel
    // (the first choice, 3 * 148 blocks, ran the 3^3 kernel as 1.5 waves: the second wave half empty)
    const int slots = (ksize == 3 ? 2 : 1) * 148;
    int slices = slots / (groups * ksize);
    return slices < 1 ? 1 : slices;
}

extern "C" int64_t b200seg_wgrad_scratch_floats(int32_t a_channels, int32_t b_channels, int32_t ksize) {
    // the larger of the two kernels' needs (CUDA-core partials / tensor-core 48 x 40 tiles)
    const int64_t pairs = static_cast<int64_t>((a_channels + 7) / 8) * ((b_channels + 7) / 8);
    const int64_t slices = simt_slices(pairs, ksize);
    const int64_t simt = slices * ksize * ksize * ksize * pairs * 64;
    const int groups = ((a_channels + kMmaA - 1) / kMmaA) * ((b_channels + kMmaB - 1) / kMmaB);
    const int64_t mma = static_cast<int64_t>(mma_slices(groups, ksize)) * ksize * ksize * ksize * groups * kMmaA * kMmaB;
    return simt > mma ? simt : mma;
}

static int launch_wgrad_mma(const b200seg_view& a, const b200seg_view& b, int ksize, int stride, int pad, float* scratch,
                            float* grad, cudaStream_t s) {
    WgradMmaGeom g;
    g.stride = stride;
    g.pad = pad;
    g.pz = a.z;
    g.py = a.y;
    g.px = a.x;
    g.xp = (a.x + 15) / 16 * 16;
    g.bw = stride * (g.xp - 1) + ksize;
    g.a_chunks = (a.c + 7) / 8;
    g.b_chunks = (b.c + 7) / 8;
    g.a_groups = (g.a_chunks + 5) / 6;
    g.b_groups = (g.b_chunks + 4) / 5;
    g.rows = a.n * a.z * a.y;
    const int groups = g.a_groups * g.b_groups;
    int slices = mma_slices(groups, ksize);
    if (slices > g.rows) slices = g.rows;
    const int ring = ksize + stride;
    const size_t smem = (12 * static_cast<size_t>(g.xp) + static_cast<size_t>(ring) * 6 * g.bw) * sizeof(uint4);
    dim3 grid(slices, groups, ksize);
    const bool one_m = g.a_chunks <= 2, one_n = g.b_chunks <= 1;      // a single group then, too
#define B200SEG_WGRAD_MMA(KK, MT, NT)                                                                                   \
    do {                                                                                                               \
        static bool configured = false;                                                                                \
        if (!configured) {                                                                                             \
            B200SEG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_mma_kernel<KK, MT, NT>,                                       \
                                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));         \
            configured = true;                                                                                         \
        }                                                                                                              \
        wgrad_mma_kernel<KK, MT, NT><<<grid, KK * KK * 32, smem, s>>>(make_dview(a), make_dview(b), g, scratch);        \
    } while (0)
    if (ksize == 3) {
        if (one_m) B200SEG_WGRAD_MMA(3, 1, 5);
        else if (one_n) B200SEG_WGRAD_MMA(3, 3, 1);
        else B200SEG_WGRAD_MMA(3, 3, 5);
    } else {
        if (one_m) B200SEG_WGRAD_MMA(4, 1, 5);
        else if (one_n) B200SEG_WGRAD_MMA(4, 3, 1);
        else B200SEG_WGRAD_MMA(4, 3, 5);
    }
#undef B200SEG_WGRAD_MMA
    int rc = check_launch("wgrad (mma partial)");
    if (rc) return rc;
    const int taps = ksize * ksize * ksize;
    const int a_pad = g.a_chunks * 8, b_pad = g.b_chunks * 8;
    const long long total = 1LL * taps * a_pad * b_pad;
    wgrad_mma_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(scratch, slices, taps, g.a_groups,
                                                                                       g.b_groups, a_pad, b_pad, grad, total);
    return check_launch("wgrad (mma finish)");
}

extern "C" int b200seg_wgrad(b200seg_view a, b200seg_view b, int32_t ksize, int32_t stride, int32_t pad, float* scratch,
                             float* grad, void* stream) {
    int rc = check_view(a, "wgrad a");
    if (rc) return rc;
    rc = check_view(b, "wgrad b");
    if (rc) return rc;
    B200SEG_CHECK_ARG((ksize == 3 || ksize == 4) && stride >= 1 && stride <= 2 && pad >= 0 && scratch && grad && a.n == b.n,
                      "wgrad: bad arguments");
    static const bool no_mma = getenv("B200SEG_WGRAD_SIMT") != nullptr;      // A/B switch
    if (a.dtype == B200SEG_BF16 && b.dtype == B200SEG_BF16 && a.x <= kMmaMaxX && !no_mma &&
        stride == (ksize == 3 ? 1 : 2))
        return launch_wgrad_mma(a, b, ksize, stride, pad, scratch, grad, static_cast<cudaStream_t>(stream));
    const int ca8n = (a.c + 7) / 8, cb8n = (b.c + 7) / 8, pairs = ca8n * cb8n;
    WgradGeom g;
    g.stride = stride;
    g.pad = pad;
    g.pz = a.z;
    g.py = a.y;
    g.px = a.x;
    g.plane_blocks = (a.y * a.x + 31) / 32;
    B200SEG_CHECK_ARG(1LL * a.n * a.z * g.plane_blocks < (1LL << 30), "wgrad: tensor too large");
    g.items = a.n * a.z * g.plane_blocks;
    g.cb8n = cb8n;
    g.inv_px = 1.0f / static_cast<float>(a.x);
    long long slices = simt_slices(pairs, ksize);
    if (slices > 1LL * a.n * a.z) slices = 1LL * a.n * a.z;      // a block owns whole (n, z) rows
    B200SEG_CHECK_ARG(pairs <= 65535, "wgrad: too many channel chunk pairs");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(slices), pairs, ksize);
    const int halves = 1;      // (two tap-warp groups per block measured slower for K = 3: 14.0 vs 15.5 TFLOP/s)
    B200SEG_CHECK_ARG(a.dtype == b.dtype, "wgrad: operands must share one dtype");
    if (ksize == 3)
        TRAIN_DISPATCH(a.dtype, (wgrad_partial_kernel<T, 3, 1><<<grid, 9 * 32, 0, s>>>(make_dview(a), make_dview(b), g, scratch)));
    else
        TRAIN_DISPATCH(a.dtype, (wgrad_partial_kernel<T, 4, 1><<<grid, 16 * 32, 0, s>>>(make_dview(a), make_dview(b), g, scratch)));
    rc = check_launch("wgrad (partial)");
    if (rc) return rc;
    const int taps = ksize * ksize * ksize;
    const long long total = 1LL * taps * pairs * 64;
    wgrad_finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(scratch, static_cast<int>(slices) * halves, taps,
                                                                                   pairs, cb8n, grad, total);
    return check_launch("wgrad (finish)");
}

extern "C" int b200seg_avgpool2_backward(b200seg_view dy, b200seg_view add, b200seg_view dx, void* stream) {
    int rc = check_view(dy, "avgpool2_backward dy");
    if (rc) return rc;
    rc = check_view(dx, "avgpool2_backward dx");
    if (rc) return rc;
    B200SEG_CHECK_ARG(dy.n == dx.n && (dy.c + 7) / 8 == (dx.c + 7) / 8 && dy.z == dx.z / 2 && dy.y == dx.y / 2 && dy.x == dx.x / 2,
                      "avgpool2_backward: dy must be the pooled extent of dx");
    DView dadd = null_dview();
    if (add.data != nullptr) {
        rc = check_view(add, "avgpool2_backward add");
        if (rc) return rc;
        B200SEG_CHECK_ARG(same_extent(add, dx), "avgpool2_backward: add extent differs");
        dadd = make_dview(add);
    }
    const int c8n = (dx.c + 7) / 8;
    const long long total = 1LL * dx.n * c8n * dx.z * dx.y * dx.x;
    const unsigned blocks = static_cast<unsigned>((total + kTrThreads - 1) / kTrThreads);
    B200SEG_CHECK_ARG(dy.dtype == dx.dtype && (add.data == nullptr || add.dtype == dx.dtype),
                      "avgpool2_backward: views must share one dtype");
    TRAIN_DISPATCH(dx.dtype, (avgpool2_backward_kernel<T><<<blocks, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                 make_dview(dy), dadd, make_dview(dx), c8n, total)));
    return check_launch("avgpool2_backward");
}

extern "C" int b200seg_upsample_trilinear2_backward(b200seg_view dy, b200seg_view dx, void* stream) {
    int rc = check_view(dy, "upsample_trilinear2_backward dy");
    if (rc) return rc;
    rc = check_view(dx, "upsample_trilinear2_backward dx");
    if (rc) return rc;
    B200SEG_CHECK_ARG(dy.n == dx.n && (dy.c + 7) / 8 == (dx.c + 7) / 8 && dy.z == 2 * dx.z && dy.y == 2 * dx.y && dy.x == 2 * dx.x,
                      "upsample_trilinear2_backward: dy must be twice the extent of dx");
    const int c8n = (dx.c + 7) / 8;
    const long long total = 1LL * dx.n * c8n * dx.z * dx.y * dx.x;
    const unsigned blocks = static_cast<unsigned>((total + kTrThreads - 1) / kTrThreads);
    B200SEG_CHECK_ARG(dy.dtype == dx.dtype, "upsample_trilinear2_backward: views must share one dtype");
    TRAIN_DISPATCH(dx.dtype, (upsample_trilinear2_backward_kernel<T><<<blocks, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                 make_dview(dy), make_dview(dx), c8n, total)));
    return check_launch("upsample_trilinear2_backward");
}

extern "C" int b200seg_channel_scale(b200seg_view src, const float* mask, b200seg_view dst, void* stream) {
    int rc = check_view(src, "channel_scale src");
    if (rc) return rc;
    rc = check_view(dst, "channel_scale dst");
    if (rc) return rc;
    B200SEG_CHECK_ARG(mask && same_extent(src, dst) && src.dtype == dst.dtype, "channel_scale: bad arguments");
    const int c8n = (src.c + 7) / 8;
    const long long total = 1LL * src.n * c8n * src.z * src.y * src.x;
    const unsigned blocks = static_cast<unsigned>((total + kTrThreads - 1) / kTrThreads);
    TRAIN_DISPATCH(src.dtype, (channel_scale_kernel<T><<<blocks, kTrThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                  make_dview(src), mask, c8n * 8, make_dview(dst), c8n, total)));
    return check_launch("channel_scale");
}
