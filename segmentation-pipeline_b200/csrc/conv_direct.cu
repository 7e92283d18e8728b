// Direct (CUDA-core) 3-D convolution with fp32 accumulation on blocked activations.
// One thread owns one output voxel and walks taps x input chunks; weights are read through the read-only path
// (every lane of a warp reads the same address -> one broadcast transaction).  This is the precision path
// (fp32 in / fp32 out, logits within 1e-5 of the reference) and the catch-all for geometries the tensor-core
// engine does not take; it is not the throughput path.
//
// The three geometries of the networks (3x3x3 'same', blurred 4x4x4 stride-2, blurred 4x4x4 transposed stride-2) take
// conv_tiled_kernel instead: a register-tiled, shared-memory staged CUDA-core kernel (8 voxels x 8 output channels
// per thread, 64 FMAs per ~5 shared loads).
#include <cstdlib>

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

// ------------------------------------------------------------------------------------------- tiled kernels
// Register-tiled, shared-memory staged CUDA-core convolution for the three geometries of the networks:
//   MODE 0  3x3x3, stride 1, pad 1                                   (Block3d / NestedResUNet convolutions)
//   MODE 1  4x4x4, stride 2, pad 1   (BlurConv3d, components.py:111-121 with the blur folded into the weights)
//   MODE 2  4x4x4 transposed, stride 2, pad 1, output = 2 x input    (BlurConvTranspose3d, components.py:144-154)
// The strided geometries are parity-decomposed into dense 2x2x2 convolutions: MODE 1 sums, per input chunk, over the
// 8 parity classes of the input (in(2(o+e)+p), e in {0,1} for p = 0 and {-1,0} for p = 1, tap t = 2e+p+1); MODE 2 runs
// one output parity class q per blockIdx.z (out(2m+q) = sum_e in(m+e) w[t], e in {-1,0}, t = 1-2e for q = 0 and
// e in {0,1}, t = 2-2e for q = 1).  So every stage of every mode is "dense KT^3 taps over a halo tile".
//
// Block = 64 * ng threads: a 32 x 8 x 2 (x, y, z) tile of the tiled domain times ng groups of 8 output channels.  Per
// stage the block puts the (32+KT-1) x (8+KT-1) x (2+KT-1) halo tile channel-major in shared memory (row stride 37
// floats: lane (ox, y) -> bank 5y + 8ox, a bijection, so the scalar row loads are conflict-free) and the
// KT^3 x 8 x (ng * 8) weight slab; a thread owns 8 consecutive x positions and one group of 8 output channels:
// per (channel, dz, dy) it loads 8+KT-1 row values and KT x 8 weights (broadcast float4s) for KT * 64 FMAs.
constexpr int kTX = 32, kTY = 8, kTZ = 2, kRS = 37;
constexpr int kMaxGroups = 5;

__host__ __device__ constexpr int tiled_kt(int mode) { return mode == 0 ? 3 : 2; }
__host__ __device__ constexpr int tiled_plane(int mode) { return (kTZ + tiled_kt(mode) - 1) * (kTY + tiled_kt(mode) - 1) * kRS; }
__host__ __device__ constexpr int tiled_wrows(int mode) { return tiled_kt(mode) * tiled_kt(mode) * tiled_kt(mode) * 8; }

struct TiledGeom {
    int cin8, cout_pad, ng;
    int dz, dy, dx;          // tiled domain: output extent (modes 0, 1) or input extent (mode 2)
    int tiles_x, tiles_y, tiles_z;
};

template <typename T, int MODE>
__global__ void __launch_bounds__(64 * kMaxGroups, 2)
conv_tiled_kernel(DView in, const float* __restrict__ w, TiledGeom g, DEpilogue e) {
    constexpr int KT = tiled_kt(MODE);
    constexpr int HX = kTX + KT - 1, HY = kTY + KT - 1, HZ = kTZ + KT - 1;
    constexpr int PLANE = tiled_plane(MODE);
    constexpr int WROWS = tiled_wrows(MODE);
    constexpr int K = MODE == 0 ? 3 : 4;               // taps per axis of the weight tensor
    extern __shared__ float4 smem4[];
    float* s_in = reinterpret_cast<float*>(smem4);
    float* s_w = s_in + 8 * PLANE;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    int bt = blockIdx.x;
    const int x0 = (bt % g.tiles_x) * kTX;
    bt /= g.tiles_x;
    const int y0 = (bt % g.tiles_y) * kTY;
    bt /= g.tiles_y;
    const int z0 = (bt % g.tiles_z) * kTZ;
    const int n = bt / g.tiles_z;
    const int cout8 = g.cout_pad / 8;
    const int ng = g.ng;
    const int g0 = blockIdx.y * ng;
    const int ng_here = min(ng, cout8 - g0);
    const int grp = tid >> 6, v = tid & 63;
    const int ox = v & 3, yy = (v >> 2) & 7, zz = v >> 5;
    const int q = MODE == 2 ? blockIdx.z : 0;          // output parity class (mode 2)

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int stages = MODE == 1 ? g.cin8 * 8 : g.cin8;
    for (int st = 0; st < stages; ++st) {
        const int ic = MODE == 1 ? st >> 3 : st;
        const int par = MODE == 1 ? st & 7 : q;        // (pz, py, px) bits 2, 1, 0
        const int pz = (par >> 2) & 1, py = (par >> 1) & 1, px = par & 1;
        // halo origin in the (sub-sampled) input and the tap of the weight tensor per kernel position d
        int offz, offy, offx;
        if (MODE == 0) {
            offz = offy = offx = -1;
        } else if (MODE == 1) {
            offz = pz ? -1 : 0; offy = py ? -1 : 0; offx = px ? -1 : 0;
        } else {
            offz = pz ? 0 : -1; offy = py ? 0 : -1; offx = px ? 0 : -1;
        }
        __syncthreads();
        for (int i = tid; i < HZ * HY * HX; i += nthreads) {
            const int hx = i % HX, hy = (i / HX) % HY, hz = i / (HX * HY);
            int gz = z0 + hz + offz, gy = y0 + hy + offy, gx = x0 + hx + offx;
            if (MODE == 1) {
                gz = 2 * gz + pz; gy = 2 * gy + py; gx = 2 * gx + px;
            }
            Vec8 val;
            if (gz >= 0 && gz < in.z && gy >= 0 && gy < in.y && gx >= 0 && gx < in.x) {
                val = load_vec8<T>(in.data, vox_index(in, n, ic, gz, gy, gx));
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) val.v[c] = 0.f;
            }
            float* dst = s_in + (hz * HY + hy) * kRS + hx;
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c * PLANE] = val.v[c];
        }
        const int row_vecs = ng * 2;
        for (int i = tid; i < WROWS * row_vecs; i += nthreads) {
            const int row = i / row_vecs, r = i - row * row_vecs;
            const int d = row >> 3, c = row & 7;
            const int ddx = d % KT, ddy = (d / KT) % KT, ddz = d / (KT * KT);
            int tz, ty, tx;
            if (MODE == 0) {
                tz = ddz; ty = ddy; tx = ddx;
            } else if (MODE == 1) {
                tz = 2 * ddz + (pz ? 0 : 1); ty = 2 * ddy + (py ? 0 : 1); tx = 2 * ddx + (px ? 0 : 1);
            } else {
                tz = (pz ? 2 : 3) - 2 * ddz; ty = (py ? 2 : 3) - 2 * ddy; tx = (px ? 2 : 3) - 2 * ddx;
            }
            const int tap = (tz * K + ty) * K + tx;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < ng_here * 2)
                val = __ldg(reinterpret_cast<const float4*>(w + (static_cast<long long>(tap) * g.cin8 * 8 + ic * 8 + c) * g.cout_pad + g0 * 8) + r);
            reinterpret_cast<float4*>(s_w)[i] = val;
        }
        __syncthreads();
        if (grp < ng_here) {
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
#pragma unroll
                for (int dz = 0; dz < KT; ++dz) {
#pragma unroll
                    for (int dy = 0; dy < KT; ++dy) {
                        const float* rowp = s_in + c * PLANE + ((zz + dz) * HY + (yy + dy)) * kRS + ox * 8;
                        float xv[8 + KT - 1];
#pragma unroll
                        for (int i = 0; i < 8 + KT - 1; ++i) xv[i] = rowp[i];
#pragma unroll
                        for (int dx = 0; dx < KT; ++dx) {
                            const float4* wp = reinterpret_cast<const float4*>(s_w + ((((dz * KT + dy) * KT + dx) * 8 + c) * ng + grp) * 8);
                            const float4 a = wp[0], b = wp[1];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float x = xv[i + dx];
                                acc[i][0] = fmaf(x, a.x, acc[i][0]);
                                acc[i][1] = fmaf(x, a.y, acc[i][1]);
                                acc[i][2] = fmaf(x, a.z, acc[i][2]);
                                acc[i][3] = fmaf(x, a.w, acc[i][3]);
                                acc[i][4] = fmaf(x, b.x, acc[i][4]);
                                acc[i][5] = fmaf(x, b.y, acc[i][5]);
                                acc[i][6] = fmaf(x, b.z, acc[i][6]);
                                acc[i][7] = fmaf(x, b.w, acc[i][7]);
                            }
                        }
                    }
                }
            }
        }
    }
    if (grp >= ng_here) return;
    const int mz = z0 + zz, my = y0 + yy;
    if (mz >= g.dz || my >= g.dy) return;
    const int cc = g0 + grp;
    const int z = MODE == 2 ? 2 * mz + ((q >> 2) & 1) : mz, y = MODE == 2 ? 2 * my + ((q >> 1) & 1) : my;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int mx = x0 + ox * 8 + i;
        if (mx >= g.dx) break;
        const int x = MODE == 2 ? 2 * mx + (q & 1) : mx;
        Vec8 a;
#pragma unroll
        for (int j = 0; j < 8; ++j) a.v[j] = acc[i][j];
        if (MODE != 0 || e.out_ncdhw == nullptr) {
            epi_store_chunk<T>(e, cc, n, z, y, x, a);
        } else {  // cout <= 8: the whole channel vector of the voxel is in this thread
            epi_affine_act(e, 0, a);
            const long long vox = 1LL * in.z * in.y * in.x;
            float* dst = e.out_ncdhw + static_cast<long long>(n) * e.cout * vox + (static_cast<long long>(z) * in.y + y) * in.x + x;
            if (e.softmax) {
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < e.cout) m = fmaxf(m, a.v[c]);
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < e.cout) sum += expf(a.v[c] - m);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < e.cout) dst[c * vox] = expf(a.v[c] - m) / sum;
            } else {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < e.cout) dst[c * vox] = a.v[c];
            }
        }
    }
}

template <typename T, int MODE>
static int launch_tiled(const DView& di, const float* weight, const DirectGeom& dg, const DEpilogue& de, cudaStream_t s) {
    TiledGeom g;
    g.cin8 = dg.cin8;
    g.cout_pad = dg.cout_pad;
    const int cout8 = dg.cout_pad / 8;
    g.ng = cout8 < kMaxGroups ? cout8 : kMaxGroups;
    if (MODE == 2) {
        g.dz = di.z; g.dy = di.y; g.dx = di.x;
    } else {
        g.dz = dg.oz; g.dy = dg.oy; g.dx = dg.ox;
    }
    g.tiles_x = (g.dx + kTX - 1) / kTX;
    g.tiles_y = (g.dy + kTY - 1) / kTY;
    g.tiles_z = (g.dz + kTZ - 1) / kTZ;
    const size_t smem = (8 * tiled_plane(MODE) + tiled_wrows(MODE) * g.ng * 8) * sizeof(float);
    static bool configured = false;   // per instantiation
    if (!configured) {
        B200SEG_CHECK_CUDA(cudaFuncSetAttribute(conv_tiled_kernel<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (8 * tiled_plane(MODE) + tiled_wrows(MODE) * kMaxGroups * 8) * sizeof(float)));
        configured = true;
    }
    dim3 grid(static_cast<unsigned>(1LL * di.n * g.tiles_z * g.tiles_y * g.tiles_x), (cout8 + g.ng - 1) / g.ng,
              MODE == 2 ? 8 : 1);
    conv_tiled_kernel<T, MODE><<<grid, 64 * g.ng, smem, s>>>(di, weight, g, de);
    return check_launch("conv3d_direct (tiled)");
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
    static const bool legacy = std::getenv("B200SEG_DIRECT_V1") != nullptr;   // A/B switch: the one-thread-per-voxel kernel
    const bool f32 = in.dtype == B200SEG_F32;
    if (!legacy && ksize == 3 && stride == 1 && pad == 1 && !transposed && (epi->out_ncdhw == nullptr || cout <= 8))
        return f32 ? launch_tiled<float, 0>(di, weight, g, de, s) : launch_tiled<__nv_bfloat16, 0>(di, weight, g, de, s);
    if (!legacy && ksize == 4 && stride == 2 && pad == 1 && !transposed && epi->out_ncdhw == nullptr)
        return f32 ? launch_tiled<float, 1>(di, weight, g, de, s) : launch_tiled<__nv_bfloat16, 1>(di, weight, g, de, s);
    if (!legacy && ksize == 4 && stride == 2 && pad == 1 && transposed && epi->out_ncdhw == nullptr && oz == 2 * in.z &&
        oy == 2 * in.y && ox == 2 * in.x)
        return f32 ? launch_tiled<float, 2>(di, weight, g, de, s) : launch_tiled<__nv_bfloat16, 2>(di, weight, g, de, s);
    if (in.dtype == B200SEG_F32)
        conv_direct_kernel<float><<<blocks, 128, 0, s>>>(di, weight, g, de, total);
    else
        conv_direct_kernel<__nv_bfloat16><<<blocks, 128, 0, s>>>(di, weight, g, de, total);
    return check_launch("conv3d_direct");
}
