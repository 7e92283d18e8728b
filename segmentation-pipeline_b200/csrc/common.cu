// Error state, version / device queries and epilogue validation for libb200seg.
#include "common.cuh"

namespace b200seg {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

static int check_dst(const b200seg_view& d, int c_expected, int n, int z, int y, int x, int dtype, const char* nm) {
    int rc = validate_view(d, nm);
    if (rc) return rc;
    B200SEG_CHECK_ARG(d.dtype == dtype, "%s: dtype %d differs from the activation dtype %d", nm, d.dtype, dtype);
    B200SEG_CHECK_ARG(d.n == n && d.z == z && d.y == y && d.x == x, "%s: extent (%d,%d,%d,%d) != (%d,%d,%d,%d)", nm,
                      d.n, d.z, d.y, d.x, n, z, y, x);
    B200SEG_CHECK_ARG(d.c == c_expected, "%s: %d channels, expected %d", nm, d.c, c_expected);
    return B200SEG_OK;
}

int make_depilogue(const b200seg_epilogue* e, int cout, int n, int z, int y, int x, int act_dtype, DEpilogue* out) {
    B200SEG_CHECK_ARG(e != nullptr, "epilogue: null");
    B200SEG_CHECK_ARG(e->scale && e->shift && e->slope, "epilogue: scale/shift/slope must be given");
    DEpilogue d{};
    d.scale = e->scale;
    d.shift = e->shift;
    d.slope = e->slope;
    d.cout = cout;
    d.softmax = e->softmax;
    d.slope01 = e->slope01;
    d.out_ncdhw = e->out_ncdhw;
    d.dst0 = null_dview();
    d.dst1 = null_dview();
    d.residual = null_dview();
    const int cout8 = (cout + 7) / 8;
    if (e->out_ncdhw != nullptr) {
        B200SEG_CHECK_ARG(e->dst0.data == nullptr && e->dst1.data == nullptr && e->residual.data == nullptr,
                          "epilogue: out_ncdhw excludes dst0/dst1/residual");
        B200SEG_CHECK_ARG(cout <= 16, "epilogue: out_ncdhw path supports cout <= 16 (got %d)", cout);
        d.split_c8 = cout8;
    } else {
        B200SEG_CHECK_ARG(e->softmax == 0, "epilogue: softmax needs out_ncdhw");
        B200SEG_CHECK_ARG(e->dst0.data != nullptr, "epilogue: no destination");
        if (e->dst1.data != nullptr) {
            B200SEG_CHECK_ARG(e->split > 0 && e->split < cout && e->split % 8 == 0,
                              "epilogue: split %d must be a multiple of 8 in (0,%d)", e->split, cout);
            int rc = check_dst(e->dst0, e->split, n, z, y, x, act_dtype, "dst0");
            if (rc) return rc;
            rc = check_dst(e->dst1, cout - e->split, n, z, y, x, act_dtype, "dst1");
            if (rc) return rc;
            d.dst1 = make_dview(e->dst1);
            d.split_c8 = e->split / 8;
        } else {
            int rc = check_dst(e->dst0, cout, n, z, y, x, act_dtype, "dst0");
            if (rc) return rc;
            d.split_c8 = cout8;
        }
        d.dst0 = make_dview(e->dst0);
        if (e->residual.data != nullptr) {
            int rc = check_dst(e->residual, e->dst0.c, n, z, y, x, act_dtype, "residual");
            if (rc) return rc;
            d.residual = make_dview(e->residual);
        }
    }
    *out = d;
    return B200SEG_OK;
}

}  // namespace b200seg

extern "C" {

const char* b200seg_last_error(void) { return b200seg::g_error; }

int b200seg_version(void) { return B200SEG_VERSION; }

int b200seg_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    B200SEG_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    B200SEG_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (prop.major != 10) {
        b200seg::set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major,
                           prop.minor);
        return B200SEG_ERR_DEVICE;
    }
    return B200SEG_OK;
}

}  // extern "C"
