// Direct (CUDA-core) 3-D convolution with fp32 accumulation on blocked activations.
// One thread owns one output voxel and walks taps x input chunks; weights are read through the read-only path
// (every lane of a warp reads the same address -> one broadcast transaction).  This is the precision path
// (fp32 in / fp32 out, logits within 1e-5 of the reference) and the catch-all for geometries the tensor-core
// engine does not take; it is not the throughput path.
#include "common.cuh"

namespace b200seg {

struct DirectGeom {
    int k, stride, pad, transposed;
    int cin8;      // input chunks
    int cout_pad;  // round_up(cout, 8)
    int oz, oy, ox;
};

template <typename T>
__global__ void __launch_bounds__(128)
conv_direct_kernel(DView in, const float* __restrict__ w, DirectGeom g, DEpilogue e, long long total) {
    long long t = blockIdx.x * 128LL + threadIdx.x;
    if (t >= total) return;
    int x = static_cast<int>(t % g.ox);
    long long r = t / g.ox;
    int y = static_cast<int>(r % g.oy);
    r /= g.oy;
    int z = static_cast<int>(r % g.oz);
    int n = static_cast<int>(r / g.oz);
    const int cout8 = g.cout_pad / 8;
    float keep[16];  // raw outputs for the NCDHW / softmax epilogue (cout <= 16)
    for (int oc = 0; oc < cout8; ++oc) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int tz = 0; tz < g.k; ++tz) {
            int iz;
            if (!g.transposed) {
                iz = z * g.stride + tz - g.pad;
            } else {
                int num = z + g.pad - tz;
                if (num < 0 || num % g.stride) continue;
                iz = num / g.stride;
            }
            if (iz < 0 || iz >= in.z) continue;
            for (int ty = 0; ty < g.k; ++ty) {
                int iy;
                if (!g.transposed) {
                    iy = y * g.stride + ty - g.pad;
                } else {
                    int num = y + g.pad - ty;
                    if (num < 0 || num % g.stride) continue;
                    iy = num / g.stride;
                }
                if (iy < 0 || iy >= in.y) continue;
                for (int tx = 0; tx < g.k; ++tx) {
                    int ix;
                    if (!g.transposed) {
                        ix = x * g.stride + tx - g.pad;
                    } else {
                        int num = x + g.pad - tx;
                        if (num < 0 || num % g.stride) continue;
                        ix = num / g.stride;
                    }
                    if (ix < 0 || ix >= in.x) continue;
                    const int tap = (tz * g.k + ty) * g.k + tx;
                    const float* wt = w + (static_cast<long long>(tap) * g.cin8 * 8) * g.cout_pad + oc * 8;
                    for (int ic = 0; ic < g.cin8; ++ic) {
                        Vec8 v = load_vec8<T>(in.data, vox_index(in, n, ic, iz, iy, ix));
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci) {
                            const float4* wp = reinterpret_cast<const float4*>(wt + (ic * 8 + ci) * g.cout_pad);
                            float4 a = __ldg(wp), b = __ldg(wp + 1);
                            float xv = v.v[ci];
                            acc[0] = fmaf(xv, a.x, acc[0]);
                            acc[1] = fmaf(xv, a.y, acc[1]);
                            acc[2] = fmaf(xv, a.z, acc[2]);
                            acc[3] = fmaf(xv, a.w, acc[3]);
                            acc[4] = fmaf(xv, b.x, acc[4]);
                            acc[5] = fmaf(xv, b.y, acc[5]);
                            acc[6] = fmaf(xv, b.z, acc[6]);
                            acc[7] = fmaf(xv, b.w, acc[7]);
                        }
                    }
                }
            }
        }
        Vec8 a;
#pragma unroll
        for (int j = 0; j < 8; ++j) a.v[j] = acc[j];
        if (e.out_ncdhw == nullptr) {
            epi_store_chunk<T>(e, oc, n, z, y, x, a);
        } else {
            epi_affine_act(e, oc * 8, a);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (oc < 2) keep[oc * 8 + j] = a.v[j];
        }
    }
    if (e.out_ncdhw != nullptr) {
        const long long vox = 1LL * g.oz * g.oy * g.ox;
        const long long o = (static_cast<long long>(z) * g.oy + y) * g.ox + x;
        float* dst = e.out_ncdhw + static_cast<long long>(n) * e.cout * vox + o;
        if (e.softmax) {
            float m = -INFINITY;
            for (int c = 0; c < e.cout; ++c) m = fmaxf(m, keep[c]);
            float s = 0.f;
            for (int c = 0; c < e.cout; ++c) s += expf(keep[c] - m);
            for (int c = 0; c < e.cout; ++c) dst[c * vox] = expf(keep[c] - m) / s;
        } else {
            for (int c = 0; c < e.cout; ++c) dst[c * vox] = keep[c];
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_conv3d_direct(b200seg_view in, const float* weight, int32_t cout, int32_t ksize,
                                     int32_t stride, int32_t pad, int32_t transposed, const b200seg_epilogue* epi,
                                     void* stream) {
    int rc = validate_view(in, "conv3d_direct in");
    if (rc) return rc;
    B200SEG_CHECK_ARG(weight != nullptr && cout > 0, "conv3d_direct: bad weight/cout");
    B200SEG_CHECK_ARG(ksize >= 1 && ksize <= 7 && stride >= 1 && stride <= 4 && pad >= 0, "conv3d_direct: bad geometry");
    B200SEG_CHECK_ARG(epi != nullptr, "conv3d_direct: null epilogue");
    int oz, oy, ox;
    if (epi->out_ncdhw == nullptr) {
        B200SEG_CHECK_ARG(epi->dst0.data != nullptr, "conv3d_direct: no destination");
        oz = epi->dst0.z;
        oy = epi->dst0.y;
        ox = epi->dst0.x;
    } else {
        B200SEG_CHECK_ARG(!transposed && stride == 1 && 2 * pad == ksize - 1,
                          "conv3d_direct: out_ncdhw path needs a 'same' convolution");
        oz = in.z;
        oy = in.y;
        ox = in.x;
    }
    if (!transposed) {
        B200SEG_CHECK_ARG(oz == (in.z + 2 * pad - ksize) / stride + 1 && oy == (in.y + 2 * pad - ksize) / stride + 1 &&
                              ox == (in.x + 2 * pad - ksize) / stride + 1,
                          "conv3d_direct: output extent (%d,%d,%d) does not match the geometry", oz, oy, ox);
    } else {
        // output_padding in [0, stride) is implied by the destination extent
        int bz = (in.z - 1) * stride - 2 * pad + ksize, by = (in.y - 1) * stride - 2 * pad + ksize,
            bx = (in.x - 1) * stride - 2 * pad + ksize;
        B200SEG_CHECK_ARG(oz >= bz && oz < bz + stride && oy >= by && oy < by + stride && ox >= bx && ox < bx + stride,
                          "conv3d_direct: transposed output extent (%d,%d,%d) does not match the geometry", oz, oy, ox);
    }
    DEpilogue de;
    rc = make_depilogue(epi, cout, in.n, oz, oy, ox, in.dtype, &de);
    if (rc) return rc;
    DirectGeom g;
    g.k = ksize;
    g.stride = stride;
    g.pad = pad;
    g.transposed = transposed;
    g.cin8 = (in.c + 7) / 8;
    g.cout_pad = (cout + 7) / 8 * 8;
    g.oz = oz;
    g.oy = oy;
    g.ox = ox;
    DView di = make_dview(in);
    long long total = 1LL * in.n * oz * oy * ox;
    unsigned blocks = static_cast<unsigned>((total + 127) / 128);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (in.dtype == B200SEG_F32)
        conv_direct_kernel<float><<<blocks, 128, 0, s>>>(di, weight, g, de, total);
    else
        conv_direct_kernel<__nv_bfloat16><<<blocks, 128, 0, s>>>(di, weight, g, de, total);
    return check_launch("conv3d_direct");
}
