// Stand-alone hardware probe (developer tool, not part of the library): pins down, on a real B200,
// the tcgen05 / TMA behaviours the convolution engine relies on
//   1. no-swizzle K-major descriptors: which of LBO / SBO is the K stride, arbitrary (non-dense) strides,
//      16-byte-aligned start addresses, overlapping core matrices (LBO = 16 B),
//   2. accumulator column offsets that are not multiples of N, N padded with zero rows ("spill"),
//   3. TMA 4-D tiled loads without swizzle: negative coordinates, boxes larger than the tensor,
//   4. MMA issue rate versus N for shared-memory operands (is the A re-read the bound?).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma probe_umma.cu
// Run  :  ./probe_umma          (prints PASS/FAIL per case)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include "ptx_sm100.cuh"

using namespace b200seg::ptx;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                   \
        }                                                                              \
    } while (0)

struct MmaOp {
    uint32_t a_off, a_lbo, a_sbo;
    uint32_t b_off, b_lbo, b_sbo;
    uint32_t d_col, n, acc;
};
struct ProbeParams {
    int n_ops;
    MmaOp ops[8];
    int a_bytes, b_bytes, dump_cols, repeat;
};

constexpr int A_REGION = 64 * 1024;
constexpr int B_REGION = 96 * 1024;

__global__ void __launch_bounds__(128, 1)
probe_mma(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, ProbeParams p,
          float* __restrict__ d_out, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sa = smem;
    uint8_t* sb = smem + A_REGION;
    for (int i = threadIdx.x; i < p.a_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < p.b_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    fence_proxy_async();
    const int warp = threadIdx.x / 32;
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int r = 0; r < p.repeat; ++r) {
            for (int i = 0; i < p.n_ops; ++i) {
                const MmaOp& o = p.ops[i];
                uint64_t da = make_desc_kmajor_noswz(smem_u32(sa) + o.a_off, o.a_lbo, o.a_sbo);
                uint64_t db = make_desc_kmajor_noswz(smem_u32(sb) + o.b_off, o.b_lbo, o.b_sbo);
                umma_bf16(tmem + o.d_col, da, db, make_idesc_bf16(128, o.n), (r > 0) ? 1u : o.acc);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (threadIdx.x == 0) {
        t1 = clock64();
        if (cycles) *cycles = t1 - t0;
    }
    tc_fence_after();
    // dump: thread t = lane t of TMEM
    for (int c = 0; c < p.dump_cols; c += 8) {
        uint32_t r[8];
        tmem_ld8(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 8; ++j) d_out[threadIdx.x * p.dump_cols + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------
// Host-side model of the layout hypothesis:  element (row r, k) of a K=16 operand lives at
//   off + (r / 8) * SBO + (r % 8) * 16 + (k / 8) * LBO + (k % 8) * 2
static inline float bf(const __nv_bfloat16& v) { return __bfloat162float(v); }

struct Case {
    const char* name;
    ProbeParams p;
    std::vector<uint8_t> a_img, b_img;
};

static float rd(const std::vector<uint8_t>& img, size_t off) {
    __nv_bfloat16 v;
    if (off + 2 > img.size()) return 0.f;  // swapped-hypothesis reads may fall outside the image
    memcpy(&v, img.data() + off, 2);
    return bf(v);
}

static void expect_seq(const Case& c, std::vector<float>& D, bool swapped) {
    // D: 128 x dump_cols, initial NaN-free marker value 777
    for (int i = 0; i < c.p.n_ops; ++i) {
        const MmaOp& o = c.p.ops[i];
        uint32_t a_k = swapped ? o.a_sbo : o.a_lbo, a_m = swapped ? o.a_lbo : o.a_sbo;
        uint32_t b_k = swapped ? o.b_sbo : o.b_lbo, b_n = swapped ? o.b_lbo : o.b_sbo;
        for (int m = 0; m < 128; ++m)
            for (uint32_t n = 0; n < o.n; ++n) {
                float s = 0.f;
                for (int k = 0; k < 16; ++k) {
                    float av = rd(c.a_img, o.a_off + (m / 8) * a_m + (m % 8) * 16 + (k / 8) * a_k + (k % 8) * 2);
                    float bv = rd(c.b_img, o.b_off + (n / 8) * b_n + (n % 8) * 16 + (k / 8) * b_k + (k % 8) * 2);
                    s += av * bv;
                }
                float& d = D[m * c.p.dump_cols + o.d_col + n];
                d = o.acc ? d + s : s;
            }
    }
}

static uint32_t rng_state = 12345u;
static int rnd_small() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return static_cast<int>((rng_state >> 24) % 7) - 3;
}
static void fill_rand(std::vector<uint8_t>& img) {
    for (size_t i = 0; i + 1 < img.size(); i += 2) {
        __nv_bfloat16 v = __float2bfloat16(static_cast<float>(rnd_small()));
        memcpy(img.data() + i, &v, 2);
    }
}

static bool run_case(Case& c, bool timing_only = false) {
    uint8_t *da, *db;
    float* dd;
    long long* dc;
    CK(cudaMalloc(&da, A_REGION));
    CK(cudaMalloc(&db, B_REGION));
    CK(cudaMalloc(&dd, 128 * 512 * 4));
    CK(cudaMalloc(&dc, 8));
    CK(cudaMemset(dd, 0, 128 * 512 * 4));
    CK(cudaMemcpy(da, c.a_img.data(), c.a_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, c.b_img.data(), c.b_img.size(), cudaMemcpyHostToDevice));
    c.p.a_bytes = static_cast<int>(c.a_img.size());
    c.p.b_bytes = static_cast<int>(c.b_img.size());
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, A_REGION + B_REGION));
    probe_mma<<<1, 128, A_REGION + B_REGION>>>(da, db, c.p, dd, dc);
    CK(cudaDeviceSynchronize());
    long long cyc = 0;
    CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    bool ok = true;
    if (timing_only) {
        int nm = c.p.n_ops * c.p.repeat;
        printf("TIMING %-44s : %lld cycles / %d MMAs = %.1f cyc/MMA\n", c.name, cyc, nm, double(cyc) / nm);
    } else {
        std::vector<float> got(128 * c.p.dump_cols), e1(128 * c.p.dump_cols, 0.f), e2(128 * c.p.dump_cols, 0.f);
        CK(cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost));
        expect_seq(c, e1, false);
        expect_seq(c, e2, true);
        // only compare columns that the op list fully defines (first op touching a column must be acc=0)
        std::vector<char> defined(c.p.dump_cols, 0);
        for (int i = 0; i < c.p.n_ops; ++i) {
            const MmaOp& o = c.p.ops[i];
            for (uint32_t n = 0; n < o.n; ++n) {
                if (!o.acc) defined[o.d_col + n] = 1;
            }
        }
        // columns whose first touch was acc=1 are undefined
        std::vector<char> first_acc(c.p.dump_cols, 0), seen(c.p.dump_cols, 0);
        for (int i = 0; i < c.p.n_ops; ++i) {
            const MmaOp& o = c.p.ops[i];
            for (uint32_t n = 0; n < o.n; ++n) {
                int col = o.d_col + n;
                if (!seen[col]) {
                    seen[col] = 1;
                    first_acc[col] = o.acc ? 1 : 0;
                }
            }
        }
        long bad1 = 0, bad2 = 0, cnt = 0;
        for (int m = 0; m < 128; ++m)
            for (int col = 0; col < c.p.dump_cols; ++col) {
                if (!seen[col] || first_acc[col]) continue;
                ++cnt;
                if (got[m * c.p.dump_cols + col] != e1[m * c.p.dump_cols + col]) ++bad1;
                if (got[m * c.p.dump_cols + col] != e2[m * c.p.dump_cols + col]) ++bad2;
            }
        ok = (bad1 == 0);
        printf("%s %-44s : hypothesis(LBO=K,SBO=MN) mismatches %ld/%ld ; swapped mismatches %ld/%ld\n",
               ok ? "PASS" : "FAIL", c.name, bad1, cnt, bad2, cnt);
        if (!ok) {
            printf("   sample got/exp row0: ");
            for (int col = 0; col < 8 && col < c.p.dump_cols; ++col)
                printf("%g/%g ", got[col], e1[col]);
            printf("\n");
        }
    }
    cudaFree(da);
    cudaFree(db);
    cudaFree(dd);
    cudaFree(dc);
    return ok;
}

static MmaOp op(uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_off, uint32_t b_lbo, uint32_t b_sbo,
                uint32_t d_col, uint32_t n, uint32_t acc) {
    MmaOp o{a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, d_col, n, acc};
    return o;
}

// ---------------------------------------------------------------------------------------------
// TMA probe: 4-D tensor (X*8 elements, Y, Z, NC) of bf16; box (HX*8, HY, 1, 2).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe_tma(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, int bytes,
                          uint8_t* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = 0xEE;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

static bool run_tma_case(const char* name, int X, int Y, int Z, int NC, int HX, int HY, int c_x, int c_y, int c_z,
                         int c_n) {
    EncodeTiledFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &qres));
    if (!encode) {
        printf("FAIL %s: no cuTensorMapEncodeTiled\n", name);
        return false;
    }
    size_t n_el = size_t(NC) * Z * Y * X * 8;
    std::vector<__nv_bfloat16> h(n_el);
    for (size_t i = 0; i < n_el; ++i) h[i] = __float2bfloat16(float(i % 251) + 1.0f);
    __nv_bfloat16* d;
    CK(cudaMalloc(&d, n_el * 2));
    CK(cudaMemcpy(d, h.data(), n_el * 2, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t gdim[4] = {cuuint64_t(X) * 8, cuuint64_t(Y), cuuint64_t(Z), cuuint64_t(NC)};
    cuuint64_t gstr[3] = {cuuint64_t(X) * 16, cuuint64_t(Y) * X * 16, cuuint64_t(Z) * Y * X * 16};
    cuuint32_t box[4] = {cuuint32_t(HX * 8), cuuint32_t(HY), 1, 2};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("FAIL %-44s : cuTensorMapEncodeTiled -> %d\n", name, int(r));
        cudaFree(d);
        return false;
    }
    int bytes = HX * 8 * HY * 2 * 2;
    uint8_t* dout;
    CK(cudaMalloc(&dout, bytes));
    probe_tma<<<1, 128, bytes>>>(tm, c_x * 8, c_y, c_z, c_n, bytes, dout);
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> got(bytes / 2);
    CK(cudaMemcpy(got.data(), dout, bytes, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int ch = 0; ch < 2; ++ch)
        for (int y = 0; y < HY; ++y)
            for (int x = 0; x < HX; ++x)
                for (int e = 0; e < 8; ++e) {
                    int gx = c_x + x, gy = c_y + y, gz = c_z, gn = c_n + ch;
                    float exp = 0.f;
                    if (gx >= 0 && gx < X && gy >= 0 && gy < Y && gz >= 0 && gz < Z && gn >= 0 && gn < NC) {
                        size_t idx = (((size_t(gn) * Z + gz) * Y + gy) * X + gx) * 8 + e;
                        exp = bf(h[idx]);
                    }
                    float g = bf(got[((size_t(ch) * HY + y) * HX + x) * 8 + e]);
                    if (g != exp) ++bad;
                }
    printf("%s %-44s : mismatches %ld/%d\n", bad == 0 ? "PASS" : "FAIL", name, bad, bytes / 2);
    cudaFree(d);
    cudaFree(dout);
    return bad == 0;
}

int main() {
    int fails = 0;
    // ---- case 1: dense operands, single MMA, N=48
    {
        Case c{"dense A/B, N=48", {}, std::vector<uint8_t>(2 * 2048), std::vector<uint8_t>(2 * 48 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 1;
        c.p.ops[0] = op(0, 2048, 128, 0, 48 * 16, 128, 0, 48, 0);
        c.p.dump_cols = 48;
        c.p.repeat = 1;
        fails += !run_case(c);
    }
    // ---- case 2: halo-style A (SBO=160, LBO=2880, start offset (dy*10+dx)*16), N=80
    {
        Case c{"halo A (SBO=160,LBO=2880,off=(1*10+2)*16), N=80", {}, std::vector<uint8_t>(5760 + 512),
               std::vector<uint8_t>(2 * 80 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 1;
        c.p.ops[0] = op((1 * 10 + 2) * 16, 2880, 160, 0, 80 * 16, 128, 0, 80, 0);
        c.p.dump_cols = 80;
        c.p.repeat = 1;
        fails += !run_case(c);
    }
    // ---- case 3: accumulator offsets + zero-row spill + accumulate chain (the dz-fused schedule)
    {
        // B: 3 blocks of 40 rows + 16 zero rows = 136 rows, K=16 -> [k2][136][8]
        const int NB = 136;
        Case c{"dz-fused schedule: cols 0/40/80, N=48/80/128 w/ spill", {}, std::vector<uint8_t>(5760 * 2),
               std::vector<uint8_t>(2 * NB * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        for (int k2 = 0; k2 < 2; ++k2)
            memset(c.b_img.data() + (k2 * NB + 120) * 16, 0, 16 * 16);
        const uint32_t lb = NB * 16;
        c.p.n_ops = 6;
        // plane0 first touch (block 2), spills zeros into cols 40..47
        c.p.ops[0] = op(0, 2880, 160, 80 * 16, lb, 128, 0, 48, 0);
        // zi=1: plane0 += block1 (N=48, spill reads block2 rows 80..87 -> adds garbage to cols 40..47, later overwritten)
        c.p.ops[1] = op(5760, 2880, 160, 40 * 16, lb, 128, 0, 48, 1);
        //        plane1 first touch (block 2)
        c.p.ops[2] = op(5760, 2880, 160, 80 * 16, lb, 128, 40, 48, 0);
        // zi=2: planes 0,1 += blocks 0,1 (N=80) ; plane 2 first touch
        c.p.ops[3] = op(16, 2880, 160, 0, lb, 128, 0, 80, 1);
        c.p.ops[4] = op(16, 2880, 160, 80 * 16, lb, 128, 80, 48, 0);
        // full 3-plane update with N=128 (120 + 8 zero rows) spilling zeros into cols 120..127
        c.p.ops[5] = op(32, 2880, 160, 0, lb, 128, 0, 128, 1);
        c.p.dump_cols = 128;
        c.p.repeat = 1;
        fails += !run_case(c);
    }
    // ---- case 4: LBO = 16 B (second K core matrix = first shifted by one voxel), N=48
    {
        Case c{"tap-pair A (LBO=16 B), N=48", {}, std::vector<uint8_t>(5760), std::vector<uint8_t>(2 * 48 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 1;
        c.p.ops[0] = op(160, 16, 160, 0, 48 * 16, 128, 0, 48, 0);
        c.p.dump_cols = 48;
        c.p.repeat = 1;
        fails += !run_case(c);
    }
    // ---- case 5: N=240 and N=256
    for (int n : {240, 256}) {
        char nm[64];
        snprintf(nm, sizeof nm, "halo A, N=%d", n);
        Case c{strdup(nm), {}, std::vector<uint8_t>(5760), std::vector<uint8_t>(2 * n * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 1;
        c.p.ops[0] = op(0, 2880, 160, 0, n * 16, 128, 0, n, 0);
        c.p.dump_cols = n;
        c.p.repeat = 1;
        fails += !run_case(c);
    }
    // ---- case 6: D column offset 8 (not 16/32-aligned), N=16 minimal
    {
        Case c{"D col offset 8, N=16", {}, std::vector<uint8_t>(5760), std::vector<uint8_t>(2 * 16 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 1;
        c.p.ops[0] = op(0, 2880, 160, 0, 16 * 16, 128, 8, 16, 0);
        c.p.dump_cols = 24;
        c.p.repeat = 1;
        fails += !run_case(c);
    }
    // ---- TMA cases
    fails += !run_tma_case("TMA interior box", 24, 40, 3, 4, 10, 18, 3, 5, 1, 1);
    fails += !run_tma_case("TMA negative x/y coords (zero fill)", 24, 40, 3, 4, 10, 18, -1, -1, 0, 2);
    fails += !run_tma_case("TMA high edge overrun", 24, 40, 3, 4, 10, 18, 15, 31, 2, 0);
    fails += !run_tma_case("TMA z out of range (-1)", 24, 40, 3, 4, 10, 18, 0, 0, -1, 0);
    fails += !run_tma_case("TMA box larger than tensor (3x3)", 3, 3, 3, 4, 10, 18, -1, -1, 1, 1);

    // ---- timing: cycles per MMA for several N (same operands, accumulate chain)
    for (int n : {48, 80, 128, 160, 240, 256}) {
        char nm[64];
        snprintf(nm, sizeof nm, "halo A (SBO=160), N=%d", n);
        Case c{strdup(nm), {}, std::vector<uint8_t>(5760 * 2), std::vector<uint8_t>(2 * 256 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 8;
        for (int i = 0; i < 8; ++i)
            c.p.ops[i] = op((i % 3) * 16 + (i / 3) * 160, 2880, 160, 0, 256 * 16, 128, 0, n, i ? 1 : 0);
        c.p.dump_cols = 8;
        c.p.repeat = 256;
        run_case(c, true);
    }
    for (int n : {48, 80, 128, 240}) {
        char nm[64];
        snprintf(nm, sizeof nm, "dense A (SBO=128), N=%d", n);
        Case c{strdup(nm), {}, std::vector<uint8_t>(4096 * 2), std::vector<uint8_t>(2 * 256 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 8;
        for (int i = 0; i < 8; ++i) c.p.ops[i] = op(0, 2048, 128, 0, 256 * 16, 128, 0, n, i ? 1 : 0);
        c.p.dump_cols = 8;
        c.p.repeat = 256;
        run_case(c, true);
    }
    // ---- timing 2: rotate over R disjoint accumulators (is the 124-cycle figure a dependent-chain latency?)
    for (int n : {48, 80, 128, 240}) {
        for (int R : {1, 2, 4, 8}) {
            if (n * R > 512) continue;
            char nm[64];
            snprintf(nm, sizeof nm, "halo A, N=%d, %d disjoint accumulators", n, R);
            Case c{strdup(nm), {}, std::vector<uint8_t>(5760 * 2), std::vector<uint8_t>(2 * 256 * 16)};
            fill_rand(c.a_img);
            fill_rand(c.b_img);
            c.p.n_ops = 8;
            for (int i = 0; i < 8; ++i)
                c.p.ops[i] = op((i % 3) * 16 + (i / 3) * 160, 2880, 160, 0, 256 * 16, 128, (i % R) * n, n, i >= R ? 1 : 0);
            c.p.dump_cols = 8;
            c.p.repeat = 256;
            run_case(c, true);
        }
    }
    // ---- timing 3: overlapping-but-shifted accumulators (plane windows zi-2..zi as in the dz-fused schedule)
    for (int cp : {40, 80}) {
        char nm[64];
        snprintf(nm, sizeof nm, "dz-fused windows, Cpad=%d (N=%d), stride Cpad", cp, cp == 40 ? 128 : 240);
        Case c{strdup(nm), {}, std::vector<uint8_t>(5760 * 2), std::vector<uint8_t>(2 * 256 * 16)};
        fill_rand(c.a_img);
        fill_rand(c.b_img);
        c.p.n_ops = 6;
        int order[6] = {0, 3, 1, 4, 2, 5};
        for (int i = 0; i < 6; ++i)
            c.p.ops[i] = op((i % 3) * 16, 2880, 160, 0, 256 * 16, 128, (cp == 40 ? order[i] : (order[i] % 3)) * cp,
                            cp == 40 ? 128 : 240, 1);
        c.p.dump_cols = 8;
        c.p.repeat = 256;
        run_case(c, true);
    }
    printf("probe finished: %d failing case(s)\n", fails);
    return fails ? 1 : 0;
}
