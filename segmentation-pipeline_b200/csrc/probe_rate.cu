// Developer probe: tcgen05.mma issue/execute rate for small N with a tight, unrolled issue loop.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "ptx_sm100.cuh"
using namespace b200seg::ptx;
#define CK(x) do { cudaError_t e_=(x); if(e_!=cudaSuccess){printf("CUDA error %s line %d\n",cudaGetErrorString(e_),__LINE__);exit(2);} } while(0)

// N: MMA N ; R: number of distinct accumulators rotated ; ASTEP: A start-address step between MMAs (bytes)
template <int N, int R, int DSTRIDE>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* cycles, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (threadIdx.x / 32 == 0) tmem_alloc<512>(&tmem_slot);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x / 32 == 0 && elect_one()) {
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 64 * 1024;
        const uint64_t da0 = make_desc_kmajor_noswz(sa, 2880, 160);
        const uint64_t db0 = make_desc_kmajor_noswz(sb, 256 * 16, 128);
        const uint32_t idesc = make_idesc_bf16(128, N);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // A start advances by one voxel (16 B => +1 in descriptor units), B by 512 B (+32 units)
                umma_bf16(tmem + (i % R) * DSTRIDE, da0 + uint64_t(i), db0 + uint64_t(i * 32), idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        *cycles = t1 - t0;
    }
    __syncthreads();
    tc_fence_before(); __syncthreads();
    if (threadIdx.x / 32 == 0) tmem_dealloc<512>(tmem);
}

template <int N, int R, int DSTRIDE>
void run(const char* what) {
    long long* dc; CK(cudaMalloc(&dc, 8));
    CK(cudaFuncSetAttribute(rate_kernel<N, R, DSTRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    const int iters = 512;
    rate_kernel<N, R, DSTRIDE><<<1, 128, 160 * 1024>>>(dc, iters);
    CK(cudaDeviceSynchronize());
    rate_kernel<N, R, DSTRIDE><<<1, 128, 160 * 1024>>>(dc, iters);
    CK(cudaDeviceSynchronize());
    long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    printf("RATE N=%3d R=%d dstride=%3d %-28s: %.1f cyc/MMA (floor N/2=%.0f, smem-bound (4096+32N)/128=%.0f)\n", N, R, DSTRIDE, what,
           double(c) / (iters * 8), N / 2.0, (4096 + 32.0 * N) / 128);
    cudaFree(dc);
}

int main() {
    run<48, 1, 0>("same accumulator");
    run<48, 8, 48>("8 disjoint");
    run<80, 1, 0>("same accumulator");
    run<80, 4, 80>("4 disjoint");
    run<128, 1, 0>("same accumulator");
    run<128, 4, 128>("4 disjoint");
    run<128, 4, 40>("overlapping windows s=40");
    run<160, 1, 0>("same accumulator");
    run<160, 2, 160>("2 disjoint");
    run<240, 1, 0>("same accumulator");
    run<240, 2, 240>("2 disjoint");
    run<240, 4, 80>("overlapping windows s=80");
    run<256, 1, 0>("same accumulator");
    run<256, 2, 256>("2 disjoint");
    return 0;
}
