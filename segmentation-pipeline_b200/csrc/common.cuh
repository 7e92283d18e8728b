// Shared host/device helpers for libb200seg: error reporting, blocked-layout indexing, epilogue.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/b200seg.h"

namespace b200seg {

// ------------------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);

#define B200SEG_CHECK_ARG(cond, ...)         \
    do {                                     \
        if (!(cond)) {                       \
            ::b200seg::set_error(__VA_ARGS__); \
            return B200SEG_ERR_ARG;          \
        }                                    \
    } while (0)

#define B200SEG_CHECK_CUDA(expr)                                                                  \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess) {                                                               \
            ::b200seg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                                 __LINE__);                                                       \
            return B200SEG_ERR_CUDA;                                                              \
        }                                                                                         \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s launch failed: %s", what, cudaGetErrorString(err));
        return B200SEG_ERR_CUDA;
    }
    return B200SEG_OK;
}

inline int validate_view(const b200seg_view& v, const char* name) {
    B200SEG_CHECK_ARG(v.data != nullptr, "%s: null data", name);
    B200SEG_CHECK_ARG(v.dtype == B200SEG_F32 || v.dtype == B200SEG_BF16, "%s: bad dtype %d", name, v.dtype);
    B200SEG_CHECK_ARG(v.n > 0 && v.c > 0 && v.z > 0 && v.y > 0 && v.x > 0, "%s: empty extent", name);
    B200SEG_CHECK_ARG(v.c8_off >= 0 && v.c8_off + (v.c + 7) / 8 <= v.c8_total,
                      "%s: chunk range [%d,+%d) exceeds c8_total %d", name, v.c8_off, (v.c + 7) / 8, v.c8_total);
    B200SEG_CHECK_ARG((reinterpret_cast<uintptr_t>(v.data) & 15) == 0, "%s: data not 16-byte aligned", name);
    return B200SEG_OK;
}

// ------------------------------------------------------------------------------------------- device view
// Device-side copy of b200seg_view with precomputed strides (in units of one 8-channel voxel vector).
struct DView {
    void* data;
    int n, c, c8_total, c8_off, z, y, x;
    long long chunk_stride;   // z*y*x
    long long sample_stride;  // c8_total * z*y*x
};

inline DView make_dview(const b200seg_view& v) {
    DView d;
    d.data = v.data;
    d.n = v.n;
    d.c = v.c;
    d.c8_total = v.c8_total;
    d.c8_off = v.c8_off;
    d.z = v.z;
    d.y = v.y;
    d.x = v.x;
    d.chunk_stride = 1LL * v.z * v.y * v.x;
    d.sample_stride = d.chunk_stride * v.c8_total;
    return d;
}
inline DView null_dview() {
    DView d{};
    d.data = nullptr;
    return d;
}

// index (in 8-channel vectors) of voxel (n, chunk cc of the view, z, y, x)
__device__ __forceinline__ long long vox_index(const DView& v, int n, int cc, int z, int y, int x) {
    return n * v.sample_stride + (v.c8_off + cc) * v.chunk_stride + (static_cast<long long>(z) * v.y + y) * v.x + x;
}

// 8 channels of one voxel
struct Vec8 {
    float v[8];
};

template <typename T>
__device__ __forceinline__ Vec8 load_vec8(const void* base, long long idx);
template <>
__device__ __forceinline__ Vec8 load_vec8<float>(const void* base, long long idx) {
    const float4* p = reinterpret_cast<const float4*>(base) + idx * 2;
    float4 a = p[0], b = p[1];
    Vec8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <>
__device__ __forceinline__ Vec8 load_vec8<__nv_bfloat16>(const void* base, long long idx) {
    uint4 raw = reinterpret_cast<const uint4*>(base)[idx];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    Vec8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
template <typename T>
__device__ __forceinline__ void store_vec8(void* base, long long idx, const Vec8& r);
template <>
__device__ __forceinline__ void store_vec8<float>(void* base, long long idx, const Vec8& r) {
    float4* p = reinterpret_cast<float4*>(base) + idx * 2;
    p[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    p[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <>
__device__ __forceinline__ void store_vec8<__nv_bfloat16>(void* base, long long idx, const Vec8& r) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    reinterpret_cast<uint4*>(base)[idx] = raw;
}

// ------------------------------------------------------------------------------------------- epilogue
struct DEpilogue {
    const float* scale;
    const float* shift;
    const float* slope;
    DView dst0, dst1, residual;
    int split_c8;     // chunks routed to dst0
    float* out_ncdhw;
    int softmax;
    int slope01;
    int cout;
};

int make_depilogue(const b200seg_epilogue* e, int cout, int n, int z, int y, int x, int act_dtype, DEpilogue* out);

// scale/shift/activation of one chunk of 8 channels (channel base c0)
__device__ __forceinline__ void epi_affine_act(const DEpilogue& e, int c0, Vec8& a) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = fmaf(a.v[j], __ldg(e.scale + c0 + j), __ldg(e.shift + c0 + j));
        float s = __ldg(e.slope + c0 + j);
        a.v[j] = v > 0.f ? v : v * s;
    }
}

// Full epilogue for chunk cc of an output voxel when writing blocked activations.
template <typename T>
__device__ __forceinline__ void epi_store_chunk(const DEpilogue& e, int cc, int n, int z, int y, int x, Vec8 a) {
    epi_affine_act(e, cc * 8, a);
    if (cc < e.split_c8) {
        if (e.residual.data != nullptr) {
            Vec8 r = load_vec8<T>(e.residual.data, vox_index(e.residual, n, cc, z, y, x));
#pragma unroll
            for (int j = 0; j < 8; ++j) a.v[j] += r.v[j];
        }
        store_vec8<T>(e.dst0.data, vox_index(e.dst0, n, cc, z, y, x), a);
    } else {
        store_vec8<T>(e.dst1.data, vox_index(e.dst1, n, cc - e.split_c8, z, y, x), a);
    }
}

}  // namespace b200seg
