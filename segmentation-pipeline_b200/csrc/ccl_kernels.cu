// Connected-component labelling of a 3-D binary mask and the instance-overlap histogram, on the device -- the voxel work
// of InstanceSegmentationEvaluator (evaluators/instance_segmentation_evaluator.py:103-129): skimage.morphology.label of
// prediction > 0 and target > 0, then the (N + 1) x (M + 1) table of voxel counts per (target component, predicted
// component) pair that the reference builds with target + prediction * 10^6 -> torch.unique.
//
// Labelling: union-find over linear voxel indices (parent array in global memory, atomicMin hooks roots onto smaller
// indices -- the "label equivalence" scheme of Playne & Hawick / Komura).  Every foreground voxel is united with its
// already-scanned neighbours (3 / 9 / 13 of them for connectivity 1 / 2 / 3), then every voxel is pointed at its root.
// A root is the SMALLEST linear index of its component, i.e. its first voxel in C-order scan, so numbering the roots in
// increasing index order reproduces skimage's / scipy's numbering exactly (component k = k-th component met by a raster
// scan).  The result is deterministic although the hooks race: the final partition and the roots do not depend on order.
#include "common.cuh"

namespace b200seg {

constexpr int kCclThreads = 256;

__device__ __forceinline__ int ccl_find(const int* __restrict__ parent, int x) {
    int p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}

// find while other threads are still hooking roots: volatile loads (no stale L1 / read-only-cache lines); a stale answer
// would only cost extra iterations (parents only ever decrease and the atomicMin below sees the truth), but there is no
// reason to pay for it
__device__ __forceinline__ int ccl_find_live(const volatile int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        x = p;
        p = parent[x];
    }
    return x;
}

__device__ __forceinline__ void ccl_union(int* parent, int a, int b) {
    for (;;) {
        a = ccl_find_live(parent, a);
        b = ccl_find_live(parent, b);
        if (a == b) return;
        if (a > b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(parent + b, a);     // hook the larger root onto the smaller index
        if (old == b) return;
        b = old;                                      // somebody re-parented b meanwhile: continue from there
    }
}

template <typename L>
__global__ void __launch_bounds__(kCclThreads)
ccl_init_kernel(const L* __restrict__ src, long long vox, int* __restrict__ parent, int by_value) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v >= vox) return;
    // binary mode: foreground = values > 0 (the evaluator's `data > 0`); by-value mode: skimage.measure.label of an
    // integer image -- background 0, two voxels are connected when they are neighbours AND carry the same value
    // mode 2: the INVERTED mask (values <= 0) -- the holes remove_small_holes looks for (post_processing.py:56)
    const bool fg = by_value == 1 ? src[v] != 0 : (by_value == 2 ? !(src[v] > 0) : src[v] > 0);
    parent[v] = fg ? static_cast<int>(v) : -1;
}

template <typename L>
__global__ void __launch_bounds__(kCclThreads)
ccl_merge_kernel(const L* __restrict__ src, int* __restrict__ parent, int w, int h, int d, int connectivity, int by_value) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    const long long vox = 1LL * w * h * d;
    if (v >= vox || reinterpret_cast<const volatile int*>(parent)[v] < 0) return;
    const int k = static_cast<int>(v % d);
    const int j = static_cast<int>((v / d) % h);
    const int i = static_cast<int>(v / (1LL * d * h));
    // already-scanned half of the neighbourhood: offsets (di, dj, dk) that precede (0, 0, 0) in raster order and have
    // at most `connectivity` non-zero components
    for (int di = -1; di <= 0; ++di)
        for (int dj = -1; dj <= 1; ++dj)
            for (int dk = -1; dk <= 1; ++dk) {
                if (di == 0 && (dj > 0 || (dj == 0 && dk >= 0))) continue;
                if ((di != 0) + (dj != 0) + (dk != 0) > connectivity) continue;
                const int ni = i + di, nj = j + dj, nk = k + dk;
                if (ni < 0 || nj < 0 || nj >= h || nk < 0 || nk >= d) continue;
                const long long n = (1LL * ni * h + nj) * d + nk;
                if (reinterpret_cast<const volatile int*>(parent)[n] < 0) continue;
                if (by_value == 1 && src[n] != src[v]) continue;
                ccl_union(parent, static_cast<int>(v), static_cast<int>(n));
            }
}

// parent -> root for every foreground voxel; roots are counted and listed (unordered)
__global__ void __launch_bounds__(kCclThreads)
ccl_flatten_kernel(int* __restrict__ parent, long long vox, int* __restrict__ roots, int max_roots,
                   int* __restrict__ n_roots) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v >= vox || parent[v] < 0) return;
    const int r = ccl_find(parent, static_cast<int>(v));
    if (r == v) {
        const int slot = atomicAdd(n_roots, 1);
        if (slot < max_roots) roots[slot] = r;
    }
}

// after every root is known: label = 1 + rank of the voxel's root among the SORTED roots; background 0
__global__ void __launch_bounds__(kCclThreads)
ccl_relabel_kernel(const int* __restrict__ parent, long long vox, const int* __restrict__ sorted_roots, int n_roots,
                   int* __restrict__ labels) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v >= vox) return;
    if (parent[v] < 0) {
        labels[v] = 0;
        return;
    }
    const int r = ccl_find(parent, static_cast<int>(v));
    int lo = 0, hi = n_roots - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sorted_roots[mid] < r) lo = mid + 1;
        else hi = mid;
    }
    labels[v] = lo + 1;
}

// dst[v] = lut[src[v]]  (sort_by_size / unsort_by_size of post_processing.py:5-26 as one gather)
__global__ void __launch_bounds__(kCclThreads)
relabel_lut_kernel(const int* __restrict__ src, long long vox, const int* __restrict__ lut, int n_lut, int* __restrict__ dst) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v >= vox) return;
    const int x = src[v];
    dst[v] = (x >= 0 && x < n_lut) ? __ldg(lut + x) : x;
}

// grey dilation with the 3-D cross (skimage.morphology.dilation default footprint, connectivity 1); border voxels see
// their in-volume neighbours only (scipy's 'reflect' mode mirrors the border voxel itself, which changes nothing)
__global__ void __launch_bounds__(kCclThreads)
dilate_cross_kernel(const int* __restrict__ src, int w, int h, int d, int* __restrict__ dst) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    const long long vox = 1LL * w * h * d;
    if (v >= vox) return;
    const int k = static_cast<int>(v % d);
    const int j = static_cast<int>((v / d) % h);
    const int i = static_cast<int>(v / (1LL * d * h));
    int m = src[v];
    if (k > 0) m = max(m, __ldg(src + v - 1));
    if (k < d - 1) m = max(m, __ldg(src + v + 1));
    if (j > 0) m = max(m, __ldg(src + v - d));
    if (j < h - 1) m = max(m, __ldg(src + v + d));
    if (i > 0) m = max(m, __ldg(src + v - 1LL * d * h));
    if (i < w - 1) m = max(m, __ldg(src + v + 1LL * d * h));
    dst[v] = m;
}

// out[v] = (mask[v] && D != dil_src[v]) ? D : pass_src[v], D = cross dilation of dil_src at v.  One kernel for
// `img[small_holes] = dilation(img)[small_holes]` (post_processing.py:62; dil_src = pass_src = img) and for
// `change = (dilated != to_dilate) & remove; sorted_img[change] = dilated[change]` (post_processing.py:44-46).
__global__ void __launch_bounds__(kCclThreads)
dilate_where_kernel(const int* __restrict__ dil_src, const int* __restrict__ mask, const int* __restrict__ pass_src,
                    int w, int h, int d, int* __restrict__ dst) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    const long long vox = 1LL * w * h * d;
    if (v >= vox) return;
    int out = pass_src[v];
    if (mask[v] != 0) {
        const int k = static_cast<int>(v % d);
        const int j = static_cast<int>((v / d) % h);
        const int i = static_cast<int>(v / (1LL * d * h));
        const int c = dil_src[v];
        int m = c;
        if (k > 0) m = max(m, __ldg(dil_src + v - 1));
        if (k < d - 1) m = max(m, __ldg(dil_src + v + 1));
        if (j > 0) m = max(m, __ldg(dil_src + v - d));
        if (j < h - 1) m = max(m, __ldg(dil_src + v + d));
        if (i > 0) m = max(m, __ldg(dil_src + v - 1LL * d * h));
        if (i < w - 1) m = max(m, __ldg(dil_src + v + 1LL * d * h));
        if (m != c) out = m;
    }
    dst[v] = out;
}

// dst[v] = keep_lut[comp[v]] ? lut[img[v]] : 0   (`to_dilate = sorted_img * keep`, post_processing.py:42)
__global__ void __launch_bounds__(kCclThreads)
relabel_masked_kernel(const int* __restrict__ img, const int* __restrict__ lut, int n_lut, const int* __restrict__ comp,
                      const int* __restrict__ keep_lut, int n_keep, long long vox, int* __restrict__ dst) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v >= vox) return;
    const int c = comp[v];
    const bool keep = c >= 0 && c < n_keep && __ldg(keep_lut + c) != 0;
    const int x = img[v];
    dst[v] = keep ? ((x >= 0 && x < n_lut) ? __ldg(lut + x) : x) : 0;
}

// dst[v] = src[v] == value   /   dst[v] = mask[v] ? value : dst[v]   (post_processing.py:69,71)
__global__ void __launch_bounds__(kCclThreads)
label_equals_kernel(const int* __restrict__ src, long long vox, int value, int* __restrict__ dst) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v < vox) dst[v] = src[v] == value ? 1 : 0;
}
__global__ void __launch_bounds__(kCclThreads)
mask_assign_kernel(int* __restrict__ dst, const int* __restrict__ mask, long long vox, int value) {
    long long v = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    if (v < vox && mask[v] != 0) dst[v] = value;
}

// hist[t * m1 + p] += 1 per voxel (pred may be null: plain label counts); lanes of a warp that hold the same pair are grouped (label maps are piecewise
// constant, the background pair dominates), one int64 atomic per distinct pair per warp
__global__ void __launch_bounds__(kCclThreads)
overlap_histogram_kernel(const int* __restrict__ target, const int* __restrict__ pred, long long vox, int m1,
                         unsigned long long* __restrict__ hist) {
    const long long stride = static_cast<long long>(gridDim.x) * kCclThreads;
    const long long base = blockIdx.x * 1LL * kCclThreads + threadIdx.x;
    const int lane = threadIdx.x % 32;
    for (long long v0 = base - lane; v0 < vox; v0 += stride) {           // warp-uniform trip count
        const long long v = v0 + lane;
        const bool ok = v < vox;
        const long long key = ok ? static_cast<long long>(__ldg(target + v)) * m1 + (pred != nullptr ? __ldg(pred + v) : 0) : -1;
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned peers = __match_any_sync(active, key);
            if (lane == __ffs(peers) - 1) atomicAdd(hist + key, static_cast<unsigned long long>(__popc(peers)));
        }
    }
}

}  // namespace b200seg

using namespace b200seg;

extern "C" int b200seg_ccl3d_roots(const void* mask, int32_t label_bytes, int32_t w, int32_t h, int32_t d,
                                   int32_t connectivity, int32_t by_value, int32_t* parent, int32_t* roots,
                                   int32_t max_roots, int32_t* n_roots, void* stream) {
    B200SEG_CHECK_ARG(mask && parent && roots && n_roots && w > 0 && h > 0 && d > 0, "ccl3d_roots: bad arguments");
    B200SEG_CHECK_ARG(label_bytes == 1 || label_bytes == 4 || label_bytes == 8, "ccl3d_roots: label_bytes must be 1, 4 or 8");
    B200SEG_CHECK_ARG(connectivity >= 1 && connectivity <= 3, "ccl3d_roots: connectivity %d not in [1,3]", connectivity);
    const long long vox = 1LL * w * h * d;
    B200SEG_CHECK_ARG(vox < (1LL << 31), "ccl3d_roots: volume too large for int32 voxel indices");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned blocks = static_cast<unsigned>((vox + kCclThreads - 1) / kCclThreads);
    B200SEG_CHECK_CUDA(cudaMemsetAsync(n_roots, 0, sizeof(int32_t), s));
#define B200SEG_CCL(T)                                                                                                \
    do {                                                                                                              \
        ccl_init_kernel<T><<<blocks, kCclThreads, 0, s>>>(static_cast<const T*>(mask), vox, parent, by_value);         \
        ccl_merge_kernel<T><<<blocks, kCclThreads, 0, s>>>(static_cast<const T*>(mask), parent, w, h, d, connectivity,  \
                                                           by_value);                                                  \
    } while (0)
    if (label_bytes == 1) B200SEG_CCL(uint8_t);
    else if (label_bytes == 4) B200SEG_CCL(int);
    else B200SEG_CCL(long long);
#undef B200SEG_CCL
    ccl_flatten_kernel<<<blocks, kCclThreads, 0, s>>>(parent, vox, roots, max_roots, n_roots);
    return check_launch("ccl3d_roots");
}

extern "C" int b200seg_ccl3d_relabel(const int32_t* parent, int64_t voxels, const int32_t* sorted_roots, int32_t n_roots,
                                     int32_t* labels, void* stream) {
    B200SEG_CHECK_ARG(parent && labels && voxels > 0 && n_roots >= 0 && (n_roots == 0 || sorted_roots), "ccl3d_relabel: bad arguments");
    const unsigned blocks = static_cast<unsigned>((voxels + kCclThreads - 1) / kCclThreads);
    ccl_relabel_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(parent, voxels, sorted_roots, n_roots, labels);
    return check_launch("ccl3d_relabel");
}

extern "C" int b200seg_overlap_histogram(const int32_t* target, const int32_t* pred, int64_t voxels, int32_t n_target,
                                         int32_t n_pred, int64_t* hist, void* stream) {
    B200SEG_CHECK_ARG(target && hist && voxels > 0 && n_target >= 0 && n_pred >= 0, "overlap_histogram: bad arguments");
    B200SEG_CHECK_ARG(pred != nullptr || n_pred == 0, "overlap_histogram: pred may be null only with n_pred == 0");
    int dev = 0, sms = 148;
    B200SEG_CHECK_CUDA(cudaGetDevice(&dev));
    B200SEG_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long want = (voxels + kCclThreads - 1) / kCclThreads;
    const unsigned blocks = static_cast<unsigned>(want < 8LL * sms ? want : 8LL * sms);
    overlap_histogram_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        target, pred, voxels, n_pred + 1, reinterpret_cast<unsigned long long*>(hist));
    return check_launch("overlap_histogram");
}

extern "C" int b200seg_relabel_lut(const int32_t* src, int64_t voxels, const int32_t* lut, int32_t n_lut, int32_t* dst,
                                   void* stream) {
    B200SEG_CHECK_ARG(src && dst && lut && voxels > 0 && n_lut > 0, "relabel_lut: bad arguments");
    const unsigned blocks = static_cast<unsigned>((voxels + kCclThreads - 1) / kCclThreads);
    relabel_lut_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, voxels, lut, n_lut, dst);
    return check_launch("relabel_lut");
}

extern "C" int b200seg_dilate_cross(const int32_t* src, int32_t w, int32_t h, int32_t d, int32_t* dst, void* stream) {
    B200SEG_CHECK_ARG(src && dst && src != dst && w > 0 && h > 0 && d > 0, "dilate_cross: bad arguments");
    const long long vox = 1LL * w * h * d;
    const unsigned blocks = static_cast<unsigned>((vox + kCclThreads - 1) / kCclThreads);
    dilate_cross_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, w, h, d, dst);
    return check_launch("dilate_cross");
}

extern "C" int b200seg_dilate_where(const int32_t* dil_src, const int32_t* mask, const int32_t* pass_src, int32_t w,
                                    int32_t h, int32_t d, int32_t* dst, void* stream) {
    B200SEG_CHECK_ARG(dil_src && mask && pass_src && dst && dst != dil_src && w > 0 && h > 0 && d > 0,
                      "dilate_where: bad arguments");
    const long long vox = 1LL * w * h * d;
    const unsigned blocks = static_cast<unsigned>((vox + kCclThreads - 1) / kCclThreads);
    dilate_where_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(dil_src, mask, pass_src, w, h, d, dst);
    return check_launch("dilate_where");
}

extern "C" int b200seg_relabel_masked(const int32_t* img, const int32_t* lut, int32_t n_lut, const int32_t* comp,
                                      const int32_t* keep_lut, int32_t n_keep, int64_t voxels, int32_t* dst,
                                      void* stream) {
    B200SEG_CHECK_ARG(img && lut && comp && keep_lut && dst && voxels > 0 && n_lut > 0 && n_keep > 0,
                      "relabel_masked: bad arguments");
    const unsigned blocks = static_cast<unsigned>((voxels + kCclThreads - 1) / kCclThreads);
    relabel_masked_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(img, lut, n_lut, comp, keep_lut,
                                                                                         n_keep, voxels, dst);
    return check_launch("relabel_masked");
}

extern "C" int b200seg_label_equals(const int32_t* src, int64_t voxels, int32_t value, int32_t* dst, void* stream) {
    B200SEG_CHECK_ARG(src && dst && voxels > 0, "label_equals: bad arguments");
    const unsigned blocks = static_cast<unsigned>((voxels + kCclThreads - 1) / kCclThreads);
    label_equals_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, voxels, value, dst);
    return check_launch("label_equals");
}

extern "C" int b200seg_mask_assign(int32_t* dst, const int32_t* mask, int64_t voxels, int32_t value, void* stream) {
    B200SEG_CHECK_ARG(dst && mask && voxels > 0, "mask_assign: bad arguments");
    const unsigned blocks = static_cast<unsigned>((voxels + kCclThreads - 1) / kCclThreads);
    mask_assign_kernel<<<blocks, kCclThreads, 0, static_cast<cudaStream_t>(stream)>>>(dst, mask, voxels, value);
    return check_launch("mask_assign");
}
