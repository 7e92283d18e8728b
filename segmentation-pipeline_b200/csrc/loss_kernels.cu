// HybridLogisticDiceLoss (criterions/hybrid_logistic_dice_loss.py:13-43) on the device.
//
// The reference makes five full-tensor passes (prediction * target, target^2, prediction^2, log, mean) with one
// temporary each; here ONE pass over prediction and target produces, per (n, c), the four sums everything else follows
// from:  S_pt = sum p t,  S_pp = sum p^2 (or sum p),  S_tt = sum t^2 (or sum t),  S_tl = sum t log((p + eps) / (1 + eps)).
// Reduction order is fixed (per-thread strided partial -> warp shuffle -> block -> per-(n, c) partial slots summed in
// index order by the finishing kernel), so the result is deterministic.  fp64 accumulation of the block partials keeps
// it within 1e-6 of the reference's fp32 torch.sum.  The backward pass is one elementwise kernel from the same sums.
#include "common.cuh"

namespace b200seg {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 64;   // partial slots per (n, c)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// partials: [n*c][block][4] doubles
__global__ void __launch_bounds__(kLossThreads)
loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ targ, long long vox, int square_dice,
                    double* __restrict__ partials) {
    __shared__ double red[4][kLossThreads / 32];
    const long long nc = blockIdx.y;
    const float* p = pred + nc * vox;
    const float* t = targ + nc * vox;
    const long long per = (vox + gridDim.x - 1) / gridDim.x;
    const long long v0 = blockIdx.x * per, v1 = min(v0 + per, vox);
    const float eps = 1e-8f;
    float s_pt = 0.f, s_pp = 0.f, s_tt = 0.f, s_tl = 0.f;
    double d_pt = 0., d_pp = 0., d_tt = 0., d_tl = 0.;
    int k = 0;
    for (long long v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
        const float a = __ldg(p + v), b = __ldg(t + v);
        s_pt = fmaf(a, b, s_pt);
        s_pp += square_dice ? a * a : a;
        s_tt += square_dice ? b * b : b;
        s_tl = fmaf(b, logf((a + eps) / (1.f + eps)), s_tl);
        if (++k == 64) {      // bound the fp32 run length
            d_pt += s_pt; d_pp += s_pp; d_tt += s_tt; d_tl += s_tl;
            s_pt = s_pp = s_tt = s_tl = 0.f;
            k = 0;
        }
    }
    d_pt += s_pt; d_pp += s_pp; d_tt += s_tt; d_tl += s_tl;
    double vals[4] = {warp_sum(d_pt), warp_sum(d_pp), warp_sum(d_tt), warp_sum(d_tl)};
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (lane == 0)
        for (int q = 0; q < 4; ++q) red[q][warp] = vals[q];
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.;
        for (int w = 0; w < kLossThreads / 32; ++w) s += red[threadIdx.x][w];
        partials[(nc * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = s;
    }
}

// sums[n*c][4] (fp32) and the three scalars out[3] = {loss, dice_loss, logistic_loss}
__global__ void loss_finish_kernel(const double* __restrict__ partials, int blocks, int n, int c, long long vox,
                                   float dice_weight, const float* __restrict__ class_weights, float* __restrict__ sums,
                                   float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float eps = 1e-8f;
    float dice_acc = 0.f, log_acc = 0.f;
    for (int i = 0; i < n * c; ++i) {
        double s[4] = {0., 0., 0., 0.};
        for (int b = 0; b < blocks; ++b)
            for (int q = 0; q < 4; ++q) s[q] += partials[(static_cast<long long>(i) * blocks + b) * 4 + q];
        const float s_pt = static_cast<float>(s[0]), s_pp = static_cast<float>(s[1]), s_tt = static_cast<float>(s[2]);
        const float s_tl = static_cast<float>(s[3]);
        for (int q = 0; q < 4; ++q) sums[i * 4 + q] = static_cast<float>(s[q]);
        const float dice = 2.f * s_pt / ((s_tt + s_pp) + eps);
        float logistic = s_tl / static_cast<float>(vox);
        if (class_weights != nullptr) logistic *= class_weights[i % c];
        dice_acc += 1.f - dice;
        log_acc += -logistic;
    }
    const float dice_loss = dice_acc / static_cast<float>(n * c), logistic_loss = log_acc / static_cast<float>(n * c);
    out[0] = (1.f - dice_weight) * logistic_loss + dice_weight * dice_loss;
    out[1] = dice_loss;
    out[2] = logistic_loss;
}

// d loss / d prediction, scaled by the incoming gradient of 'loss'
__global__ void __launch_bounds__(kLossThreads)
loss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ targ, const float* __restrict__ sums,
                     long long vox, int c, int nc_total, int square_dice, float dice_weight,
                     const float* __restrict__ class_weights, const float* __restrict__ grad_loss,
                     float* __restrict__ grad_pred, long long total) {
    long long i = blockIdx.x * 1LL * kLossThreads + threadIdx.x;
    if (i >= total) return;
    const long long nc = i / vox;
    const float eps = 1e-8f;
    const float s_pt = sums[nc * 4], s_pp = sums[nc * 4 + 1], s_tt = sums[nc * 4 + 2];
    const float denom = (s_tt + s_pp) + eps;
    const float p = __ldg(pred + i), t = __ldg(targ + i);
    // dice = 2 S_pt / denom:  d/dp = 2 t / denom - 2 S_pt * (2 p | 1) / denom^2
    const float ddice = 2.f * t / denom - 2.f * s_pt * (square_dice ? 2.f * p : 1.f) / (denom * denom);
    // logistic = mean_v t log((p + eps) / (1 + eps)) * w_c:  d/dp = w_c t / ((p + eps) V)
    const float w = class_weights != nullptr ? class_weights[nc % c] : 1.f;
    const float dlog = w * t / ((p + eps) * static_cast<float>(vox));
    const float inv = 1.f / static_cast<float>(nc_total);
    grad_pred[i] = grad_loss[0] * inv * (-(1.f - dice_weight) * dlog - dice_weight * ddice);
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int64_t b200seg_hybrid_loss_scratch_bytes(int64_t n, int32_t c) {
    return n * c * kLossMaxBlocks * 4 * static_cast<int64_t>(sizeof(double));
}

extern "C" int b200seg_hybrid_loss_forward(const float* prediction, const float* target, int64_t n, int32_t c,
                                           int64_t voxels, float dice_weight, const float* class_weights,
                                           int32_t square_dice, void* scratch, int64_t scratch_bytes, float* sums,
                                           float* out3, void* stream) {
    B200SEG_CHECK_ARG(prediction && target && sums && out3 && n > 0 && c > 0 && voxels > 0, "hybrid_loss_forward: bad arguments");
    B200SEG_CHECK_ARG(n * c <= 65535, "hybrid_loss_forward: n * c = %lld exceeds 65535", static_cast<long long>(n * c));
    B200SEG_CHECK_ARG(scratch != nullptr && scratch_bytes >= b200seg_hybrid_loss_scratch_bytes(n, c),
                      "hybrid_loss_forward: scratch too small");
    int dev = 0, sms = 148;
    B200SEG_CHECK_CUDA(cudaGetDevice(&dev));
    B200SEG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (8LL * sms + n * c - 1) / (n * c);
    if (blocks > kLossMaxBlocks) blocks = kLossMaxBlocks;
    if (blocks > (voxels + 4095) / 4096) blocks = (voxels + 4095) / 4096;
    if (blocks < 1) blocks = 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(n * c));
    loss_partial_kernel<<<grid, kLossThreads, 0, s>>>(prediction, target, voxels, square_dice, static_cast<double*>(scratch));
    int rc = check_launch("hybrid_loss_forward");
    if (rc) return rc;
    loss_finish_kernel<<<1, 32, 0, s>>>(static_cast<const double*>(scratch), static_cast<int>(blocks), static_cast<int>(n),
                                        c, voxels, dice_weight, class_weights, sums, out3);
    return check_launch("hybrid_loss_forward (finish)");
}

extern "C" int b200seg_hybrid_loss_backward(const float* prediction, const float* target, const float* sums, int64_t n,
                                            int32_t c, int64_t voxels, float dice_weight, const float* class_weights,
                                            int32_t square_dice, const float* grad_loss, float* grad_prediction,
                                            void* stream) {
    B200SEG_CHECK_ARG(prediction && target && sums && grad_loss && grad_prediction && n > 0 && c > 0 && voxels > 0,
                      "hybrid_loss_backward: bad arguments");
    const long long total = n * c * voxels;
    loss_backward_kernel<<<static_cast<unsigned>((total + kLossThreads - 1) / kLossThreads), kLossThreads, 0,
                           static_cast<cudaStream_t>(stream)>>>(prediction, target, sums, voxels, c, static_cast<int>(n * c),
                                                                square_dice, dice_weight, class_weights, grad_loss,
                                                                grad_prediction, total);
    return check_launch("hybrid_loss_backward");
}
