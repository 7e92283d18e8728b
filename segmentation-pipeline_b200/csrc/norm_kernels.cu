// InstanceNorm3d on blocked activations, in place, fused with the activation and the residual add of Block3d.
// TWO streaming passes over the tensor (round 1: three).  Pass 0: every block reduces its strip of voxels to
// (mean_b, M2_b) per channel in ONE read, with the shifted-data form -- deviations from the strip's first voxel K,
// sum d and sum d^2, so that M2_b = sum d^2 - (sum d)^2 / n_b does not cancel -- reduced with warp shuffles.  Pass 1:
// every block merges the <= 64 block statistics of its (n, channel chunk) in a FIXED order with Chan's pairwise update
// (deterministic, no atomics; as safe as the mean-then-deviations form), then applies scale / shift, activation and
// residual.
#include "common.cuh"

namespace b200seg {

static constexpr int kNormThreads = 256;

__device__ __forceinline__ void block_reduce8(float (&v)[8], float* smem /* [8 warps][8] */) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    }
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) smem[warp * 8 + j] = v[j];
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float s = 0.f;
        for (int w = 0; w < kNormThreads / 32; ++w) s += smem[w * 8 + threadIdx.x];
        smem[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = smem[j];
    __syncthreads();
}

// partial layout: [n][c8][nb][8]
// Chan et al. merge of the per-block (mean, M2) in block order; block b holds n_b = min(per, vox - b * per) voxels
__device__ __forceinline__ void merge_partials(const float* part_mean, const float* part_m2, int nb, long long per,
                                               long long vox, float (&mean)[8], float (&m2)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) mean[j] = m2[j] = 0.f;
    float n = 0.f;
    for (int b = 0; b < nb; ++b) {
        const long long v0 = b * per;
        const float nbk = static_cast<float>(min(per, vox - v0));
        if (nbk <= 0.f) break;
        const float tot = n + nbk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float d = __ldg(part_mean + b * 8 + j) - mean[j];
            mean[j] += d * (nbk / tot);
            m2[j] += __ldg(part_m2 + b * 8 + j) + d * d * (n * nbk / tot);
        }
        n = tot;
    }
}

template <typename T, int PASS>
__global__ void __launch_bounds__(kNormThreads)
instnorm_kernel(DView x, float* __restrict__ part_mean, float* __restrict__ part_m2, int nb, float inv_count,
                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float slope,
                DView residual) {
    __shared__ float red[(kNormThreads / 32) * 8];
    const int b = blockIdx.x, cc = blockIdx.y, n = blockIdx.z;
    const long long vox = x.chunk_stride;
    const long long per = (vox + nb - 1) / nb;
    const long long v0 = b * per, v1 = min(v0 + per, vox);
    const long long base = n * x.sample_stride + (x.c8_off + cc) * x.chunk_stride;
    const long long pidx = (static_cast<long long>(n) * gridDim.y + cc) * nb;
    if (PASS == 0) {
        if (v0 >= v1) return;
        const Vec8 k = load_vec8<T>(x.data, base + v0);      // shift: the strip's first voxel
        float s1[8], s2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
        for (long long v = v0 + threadIdx.x; v < v1; v += kNormThreads) {
            Vec8 a = load_vec8<T>(x.data, base + v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = a.v[j] - k.v[j];
                s1[j] += d;
                s2[j] = fmaf(d, d, s2[j]);
            }
        }
        block_reduce8(s1, red);
        block_reduce8(s2, red);
        if (threadIdx.x < 8) {
            const int j = threadIdx.x;
            const float nbk = static_cast<float>(v1 - v0);
            part_mean[(pidx + b) * 8 + j] = k.v[j] + s1[j] / nbk;
            part_m2[(pidx + b) * 8 + j] = fmaxf(s2[j] - s1[j] * s1[j] / nbk, 0.f);
        }
    } else {
        float mean[8], m2[8], sc[8], sh[8];
        merge_partials(part_mean + pidx * 8, part_m2 + pidx * 8, nb, per, vox, mean, m2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cc * 8 + j;
            const float rstd = rsqrtf(m2[j] * inv_count + eps);
            const float g = (gamma != nullptr && c < x.c) ? __ldg(gamma + c) : 1.f;
            const float bt = (beta != nullptr && c < x.c) ? __ldg(beta + c) : 0.f;
            sc[j] = c < x.c ? rstd * g : 0.f;     // padding channels stay exactly zero
            sh[j] = c < x.c ? bt - mean[j] * rstd * g : 0.f;
        }
        const bool has_res = residual.data != nullptr;
        const long long rbase = has_res ? n * residual.sample_stride + (residual.c8_off + cc) * residual.chunk_stride : 0;
        for (long long v = v0 + threadIdx.x; v < v1; v += kNormThreads) {
            Vec8 a = load_vec8<T>(x.data, base + v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = fmaf(a.v[j], sc[j], sh[j]);
                a.v[j] = t > 0.f ? t : t * slope;
            }
            if (has_res) {
                Vec8 r = load_vec8<T>(residual.data, rbase + v);
#pragma unroll
                for (int j = 0; j < 8; ++j) a.v[j] += r.v[j];
            }
            store_vec8<T>(x.data, base + v, a);
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int64_t b200seg_instnorm_scratch_bytes(b200seg_view x) {
    const long long c8 = (x.c + 7) / 8;
    return 2LL * x.n * c8 * 64 * 8 * static_cast<long long>(sizeof(float));
}

extern "C" int b200seg_instnorm(b200seg_view x, const float* gamma, const float* beta, float eps, float slope,
                                b200seg_view residual, void* scratch, int64_t scratch_bytes, void* stream) {
    int rc = validate_view(x, "instnorm x");
    if (rc) return rc;
    B200SEG_CHECK_ARG(scratch != nullptr && scratch_bytes >= b200seg_instnorm_scratch_bytes(x),
                      "instnorm: scratch too small (%lld bytes needed)",
                      static_cast<long long>(b200seg_instnorm_scratch_bytes(x)));
    DView dx = make_dview(x);
    DView dr = null_dview();
    if (residual.data != nullptr) {
        rc = validate_view(residual, "instnorm residual");
        if (rc) return rc;
        B200SEG_CHECK_ARG(residual.dtype == x.dtype && residual.n == x.n && residual.c == x.c && residual.z == x.z &&
                              residual.y == x.y && residual.x == x.x,
                          "instnorm: residual does not match x");
        dr = make_dview(residual);
    }
    const int c8 = (x.c + 7) / 8;
    const long long vox = dx.chunk_stride;
    int dev = 0, sms = 148;
    B200SEG_CHECK_CUDA(cudaGetDevice(&dev));
    B200SEG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long nb = (4LL * sms + 1LL * x.n * c8 - 1) / (1LL * x.n * c8);
    if (nb > 64) nb = 64;
    if (nb > (vox + 1023) / 1024) nb = (vox + 1023) / 1024;
    if (nb < 1) nb = 1;
    float* part_sum = static_cast<float*>(scratch);
    float* part_m2 = part_sum + 1LL * x.n * c8 * 64 * 8;
    const float inv_count = 1.0f / static_cast<float>(vox);
    dim3 grid(static_cast<unsigned>(nb), static_cast<unsigned>(c8), static_cast<unsigned>(x.n));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH_NORM(T)                                                                                              \
    do {                                                                                                            \
        instnorm_kernel<T, 0><<<grid, kNormThreads, 0, s>>>(dx, part_sum, part_m2, static_cast<int>(nb), inv_count, \
                                                            gamma, beta, eps, slope, dr);                           \
        instnorm_kernel<T, 1><<<grid, kNormThreads, 0, s>>>(dx, part_sum, part_m2, static_cast<int>(nb), inv_count, \
                                                            gamma, beta, eps, slope, dr);                           \
    } while (0)
    if (x.dtype == B200SEG_F32) LAUNCH_NORM(float);
    else LAUNCH_NORM(__nv_bfloat16);
#undef LAUNCH_NORM
    return check_launch("instnorm");
}
