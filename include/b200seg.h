/*
 * b200seg.h -- C ABI of libb200seg.so, the sm_100a (B200) implementation of the sliding-window 3D U-Net
 * inference path of efirdc/Segmentation-Pipeline.
 *
 * The reference has NO FFI for this path: everything is Python calling ATen.  The entry
 * points below are therefore the operator set a binding for that path needs; each one names the reference
 * call it replaces (paths relative to /root/reference/segmentation_pipeline/).  The reference-side binding
 * (a ctypes stub) is shown in INTEGRATION.md; the host-side mirror that uses it lives in
 * segmentation-pipeline_b200/segmentation_pipeline/.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - the caller (PyTorch) owns every buffer; the library allocates nothing that outlives a call;
 *   - every launch is asynchronous on the cudaStream_t passed as `stream` (a void* to keep this header free
 *     of CUDA headers);
 *   - return value: 0 = ok, negative = error; b200seg_last_error() gives the message (thread local);
 *   - there is NO CPU fallback: unsupported shapes / dtypes return an error.
 *
 * Activation layout ("blocked"):  [N][C8][Z][Y][X][8]  -- channels in chunks of 8 (one 16-byte bf16 vector per
 * voxel and chunk), spatial dims in the reference's (W, H, D) order named (Z, Y, X) here, X contiguous.
 * A b200seg_view selects the chunk range [c8_off, c8_off + ceil(c/8)) of a buffer holding c8_total chunks per
 * sample, which is how torch.cat along channels (models/modular_unet.py:97,
 * models/nested_residual_unet.py:92-101) is expressed without a copy.
 */
#ifndef B200SEG_H_
#define B200SEG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SEG_VERSION 100

enum { B200SEG_F32 = 0, B200SEG_BF16 = 1 };

enum {
    B200SEG_OK = 0,
    B200SEG_ERR_ARG = -1,      /* invalid argument / unsupported configuration */
    B200SEG_ERR_CUDA = -2,     /* a CUDA runtime or driver call failed */
    B200SEG_ERR_DEVICE = -3    /* not an sm_100 device */
};

typedef struct b200seg_view {
    void* data;        /* base of the whole buffer [N][c8_total][Z][Y][X][8] */
    int32_t dtype;     /* B200SEG_F32 or B200SEG_BF16 */
    int32_t n;         /* samples */
    int32_t c;         /* logical channels in this view */
    int32_t c8_total;  /* chunks per sample in the underlying buffer */
    int32_t c8_off;    /* first chunk of this view */
    int32_t z, y, x;   /* spatial extent */
} b200seg_view;

/* Fused epilogue of every convolution:
 *   v   = acc * scale[c] + shift[c]                         (folded BatchNorm3d eval / conv bias)
 *   v   = v > 0 ? v : v * slope[c]                          (ReLU: 0, LeakyReLU: negative_slope, none: 1)
 *   v  += residual[c]                                       (Block3d: `res_conv(x_in) + x`, components.py:67-68)
 * Channels [0, split) go to dst0, channels [split, cout) to dst1 (used to run conv0 and res_conv, which read
 * the same input, as ONE contraction with N = 2*Cout).  With softmax != 0 the channel softmax of
 * modular_unet.py:100 / nested_residual_unet.py:104 is applied and fp32 NCDHW probabilities are written to
 * out_ncdhw instead (cout <= 16).  scale/shift/slope are fp32 device arrays of length >= round_up(cout, 8). */
typedef struct b200seg_epilogue {
    const float* scale;
    const float* shift;
    const float* slope;
    b200seg_view dst0;
    b200seg_view dst1;       /* data == NULL when unused */
    int32_t split;           /* channels routed to dst0 (multiple of 8 when dst1 is used) */
    b200seg_view residual;   /* data == NULL when unused; added to dst0 channels only */
    float* out_ncdhw;        /* softmax / final output, fp32 [N][cout][Z][Y][X]; NULL when unused */
    int32_t softmax;         /* 1: softmax over channels before writing out_ncdhw; 0: raw values */
    int32_t slope01;         /* 1: the caller guarantees 0 <= slope[c] <= 1 (ReLU / LeakyReLU / none), which lets the
                                tensor-core epilogue use max(v, v*slope); 0: general v > 0 ? v : v*slope;
                                2: the caller guarantees scale == 1, shift == 0, slope == 1 for every channel (the blur
                                convolutions): the tensor-core epilogue is a plain fp32 -> bf16 conversion */
} b200seg_epilogue;

/* ------------------------------------------------------------------------------------------------ misc */
const char* b200seg_last_error(void);
int b200seg_version(void);
/* Fills sm_count / compute capability of the current device; returns B200SEG_ERR_DEVICE if it is not sm_100. */
int b200seg_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------------------------------------ layout
 * collate_subjects' torch.stack + .to(device) result (utils/utils.py:75-85), fp32 NCDHW, into the blocked
 * layout (pad channels of the last chunk are zero filled), and back. */
int b200seg_pack_ncdhw(const float* src, b200seg_view dst, void* stream);
int b200seg_unpack_ncdhw(b200seg_view src, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------ convolutions
 * Direct (CUDA-core, fp32 accumulate) 3-D convolution for any kernel size / stride / padding, regular or
 * transposed; fp32 or bf16 blocked activations.  This is the fp32 path (logit tolerance 1e-5) and the path for
 * layer shapes the tensor-core engine does not take.  Replaces nn.Conv3d (components.py:46,51;
 * modular_unet.py:83), F.conv3d with the blurred 4^3 weight (components.py:119) and F.conv_transpose3d
 * (components.py:152).
 *   weight: fp32 [k^3][cin_phys][cout_pad]  with cin_phys = 8 * ceil(in.c / 8) ... see pack_direct_weight in
 *   segmentation_pipeline/models/_plan.py; for transposed convs the taps are already those of the gather form. */
int b200seg_conv3d_direct(b200seg_view in, const float* weight, int32_t cout, int32_t ksize, int32_t stride,
                          int32_t pad, int32_t transposed, const b200seg_epilogue* epi, void* stream);

/* Tensor-core engine (tcgen05 + TMEM + TMA), bf16 activations, fp32 accumulate.  mode selects the geometry:
 *   B200SEG_TC_K3   : kernel 3, stride 1, padding 1            (nn.Conv3d of Block3d / out_conv)
 *   B200SEG_TC_DOWN : kernel 4, stride 2, padding 1            (BlurConv3d with the blur folded, :111-121)
 *   B200SEG_TC_UP   : transposed kernel 4, stride 2, padding 1 (BlurConvTranspose3d folded, :144-154)
 *   B200SEG_TC_K3T  : kernel 3, stride 1, padding 1 for a FINAL layer with cout <= 4 (out_conv, modular_unet.py:81-84):
 *                     same arithmetic as K3 but the nine in-plane taps are accumulator columns summed in the epilogue,
 *                     so one MMA per chunk pair and plane replaces nine; needs epi.out_ncdhw; own weight packing.
 * wpacked is the bf16 operand image produced by pack_tc_weight (segmentation_pipeline/models/_plan.py), laid out
 * exactly as the kernel stages it in shared memory; wpacked_bytes is checked against the geometry.
 * cout <= 80 per call (the host splits wider layers). */
enum { B200SEG_TC_K3 = 0, B200SEG_TC_DOWN = 1, B200SEG_TC_UP = 2, B200SEG_TC_K3T = 3 };
int b200seg_conv3d_tc(int32_t mode, b200seg_view in, const void* wpacked, int64_t wpacked_bytes, int32_t cout,
                      const b200seg_epilogue* epi, void* stream);
/* Bytes of the packed operand image the engine expects for (mode, cin chunks, cout). */
int64_t b200seg_conv3d_tc_wbytes(int32_t mode, int32_t cin_chunks, int32_t cout);

/* ------------------------------------------------------------------------------------------------ normalisation
 * nn.InstanceNorm3d (normalization_class of Block3d, components.py:53) in place on a blocked view, fused with the
 * block's activation (slope: ReLU 0, LeakyReLU negative_slope, none 1) and the `res_conv(x_in) + x` add
 * (residual.data == NULL when unused).  Statistics: biased variance over the spatial extent per (sample, channel),
 * two-pass (mean, then squared deviations), warp-shuffle + fixed-order partial sums (deterministic).
 * gamma / beta: fp32 device arrays of x.c entries or NULL (affine=False).  scratch: device memory of at least
 * b200seg_instnorm_scratch_bytes(x) bytes. */
int64_t b200seg_instnorm_scratch_bytes(b200seg_view x);
int b200seg_instnorm(b200seg_view x, const float* gamma, const float* beta, float eps, float slope,
                     b200seg_view residual, void* scratch, int64_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ resampling
 * nn.AvgPool3d(2, 2, count_include_pad=False) (modular_unet.py:40-41, nested_residual_unet.py:67) and
 * nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True) (modular_unet.py:38-39, :68), writing
 * into a chunk range of the consumer's concat buffer. */
int b200seg_avgpool2(b200seg_view in, b200seg_view out, void* stream);
int b200seg_upsample_trilinear2(b200seg_view in, b200seg_view out, void* stream);
/* Copy a view into a chunk range of another buffer (concat member that was not produced in place). */
int b200seg_copy_view(b200seg_view in, b200seg_view out, void* stream);
/* ------------------------------------------------------------------------------------------------ test-time augmentation
 * EnsembleFlips / EnsembleOrientations / EnsembleModels + apply_strategy (models/ensemble.py:16-103).  Member e is
 * evaluated on x.permute(0, 1, *perm + 2).flip(dims); perm[k] in {0, 1, 2} names the source spatial axis of transformed
 * axis k and flip[k] != 0 reverses transformed axis k.
 *   pack_ncdhw_tta   fp32 (N, C, w0, w1, w2) -> blocked dst in the member's space (dst extent = permuted extent)
 *   tta_accumulate   member output fp32 (N, C, S0, S1, S2), un-transformed on the fly, into EITHER the fp32 sum
 *                    acc (N, C, w0, w1, w2) ('mean') OR the uint8 vote counts votes (N, C, w0, w1, w2) of the member's
 *                    per-voxel argmax ('majority'; first maximum wins like torch.argmax)
 *   tta_finalize     'mean': acc *= 1 / members in place; 'majority': int64 one-hot (N, C, voxels) of the label with the
 *                    most votes, smallest label on ties (torch.mode). */
int b200seg_pack_ncdhw_tta(const float* src, int32_t w0, int32_t w1, int32_t w2, const int32_t perm[3],
                           const int32_t flip[3], b200seg_view dst, void* stream);
int b200seg_tta_accumulate(const float* member, int64_t n, int32_t c, int32_t w0, int32_t w1, int32_t w2,
                           const int32_t perm[3], const int32_t flip[3], float* acc, uint8_t* votes, void* stream);
int b200seg_tta_finalize(float* acc, const uint8_t* votes, int64_t* onehot, int64_t n, int32_t c, int64_t voxels,
                         int32_t members, void* stream);

/* Planes [plane_lo, plane_hi) of one fp32 (c, p0, p1, p2) patch -> a dense (c, plane_hi - plane_lo, p1, p2) block:
 * the staging step of the z-slab exchange (every output plane of a patch goes to the rank that owns it). */
int b200seg_copy_planes(const float* patch, int32_t c, int32_t p0, int32_t p1, int32_t p2, int32_t plane_lo,
                        int32_t plane_hi, float* dst, void* stream);
/* Channel softmax / StochasticMatrix softmax (components.py:170-185) on fp32 NCDHW data, in place.
 * groups = 1 for nn.Softmax(dim=1); for StochasticMatrix groups = C, diag_bias added to the diagonal. */
int b200seg_softmax_ncdhw(float* data, int64_t n, int32_t channels, int64_t voxels, int32_t sm_channels,
                          float diag_bias, void* stream);

/* ------------------------------------------------------------------------------------------------ grid sampling
 * tio.GridSampler.__getitem__ + collate (prediction.py:132-138): crops `count` patches of the fp32 volume
 * [C][W][H][D] at padded-grid locations (int32 rows i0,j0,k0,i1,j1,k1) straight into the
 * blocked layout.  Padding (torchio Pad by overlap//2) is never materialised: border = per-axis pad width,
 * pad_mode 0 = none (border ignored), 1 = 'edge' (clamp), 2 = constant pad_value.
 * locations_host is a HOST array (it travels as a kernel argument, no device copy needed). */
int b200seg_grid_extract(const float* volume, int32_t c, int32_t w, int32_t h, int32_t d,
                         const int32_t* locations_host, int32_t count, const int32_t border[3], int32_t pad_mode,
                         float pad_value, b200seg_view dst, void* stream);

/* tio.GridAggregator.add_batch, 'average' mode (prediction.py:141): out[:, i0:i1, j0:j1, k0:k1] += patch for the
 * `count` patches of one batch, IN BATCH ORDER for every voxel (owner-computes gather over the batch's bounding
 * box -- no atomics, bit-identical to the sequential CPU loop).  out is fp32 [C][PW][PH][PD] (padded extent).
 * patches: fp32 [count][C][p0][p1][p2]; locations_host: HOST int32 [count][6]. */
int b200seg_overlap_add(float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const float* patches,
                        const int32_t* locations_host, int32_t count, void* stream);
/* 'hann' mode (the weighted overlap-add of newer torchio GridAggregators, reached through the same
 * `overlap_mode` argument of PatchPredict, prediction.py:115,134): window_patches multiplies the `count` patches in
 * place by the separable window ((w0[i] * w1[j]) * w2[k]) before b200seg_overlap_add; divide_separable divides the
 * accumulator by the summed windows ((s0[i] * s1[j]) * s2[k]) -- separable because the patch grid is a product grid --
 * before b200seg_finalize is called without counts. */
int b200seg_window_patches(float* patches, int32_t count, int32_t c, int32_t p0, int32_t p1, int32_t p2, const float* w0,
                           const float* w1, const float* w2, void* stream);
int b200seg_divide_separable(float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const float* s0, const float* s1,
                             const float* s2, void* stream);
/* 'crop' mode: assign the centre crop of each patch (GridAggregator.crop_batch). */
int b200seg_overlap_crop(float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const float* patches,
                         const int32_t* locations_host, int32_t count, const int32_t border[3],
                         int32_t volume_padded, void* stream);

/* tio.GridAggregator.get_output_tensor (prediction.py:143) fused with CustomArgMax
 * (transforms/custom_label_transforms.py:267): probs = out / count (count = cw[i]*ch[j]*cd[k], the separable
 * per-axis coverage, int32 device arrays of the padded extent; pass NULL for crop mode), cropped by `border`;
 * writes fp32 probs [C][W][H][D] (may be NULL), int64 labels [W][H][D] (may be NULL) and uint8 labels
 * (may be NULL).  argmax ties resolve to the lowest index. */
int b200seg_finalize(const float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const int32_t* cw,
                     const int32_t* ch, const int32_t* cd, const int32_t border[3], float* probs,
                     int64_t* labels_i64, uint8_t* labels_u8, void* stream);

/* Same, for an arbitrary sub-region of the accumulator: output voxel (i,j,k) of extent[] reads out[:, offset+ijk].
 * Used by the z-slab multi-GPU mode, where each rank finalises only the planes it owns (count pointers are
 * passed pre-offset to the rank's first plane). */
int b200seg_finalize_region(const float* out, int32_t c, int32_t pw, int32_t ph, int32_t pd, const int32_t* cw,
                            const int32_t* ch, const int32_t* cd, const int32_t offset[3], const int32_t extent[3],
                            float* probs, int64_t* labels_i64, uint8_t* labels_u8, void* stream);

/* torch.argmax(data, dim=0, keepdim=True) on fp32 [C][V] -> int64 [V] and/or uint8 [V]. */
int b200seg_argmax(const float* probs, int32_t c, int64_t voxels, int64_t* labels_i64, uint8_t* labels_u8,
                   void* stream);

/* ------------------------------------------------------------------------------------------------ evaluator
 * Single pass over two label maps -> num_classes x num_classes joint histogram cm[target][prediction] (int64,
 * ACCUMULATED into cm so cohorts can be summed), from which TP/FP/TN/FN of segmentation_evaluator.py:69-77
 * and the volumes of label_map_evaluator.py:77-81 follow exactly.  label_bytes = 1 (uint8) or 8 (int64).
 * Values outside [0, num_classes) are counted in the LAST class (num_classes - 1) -- nothing is dropped, so callers
 * that want the reference's semantics for labels they do not list ((target == v) & (pred != v) for ANY pred value,
 * segmentation_evaluator.py:72-77) pass num_classes = max listed value + 2 and read the last row / column as
 * "other".  num_classes <= 40 (per-lane histogram columns must fit shared memory). */
int b200seg_confusion(const void* pred, const void* target, int32_t label_bytes, int64_t voxels,
                      int32_t num_classes, int64_t* cm, void* stream);

/* ------------------------------------------------------------------------------------------------ instance evaluation
 * InstanceSegmentationEvaluator.__call__ (evaluators/instance_segmentation_evaluator.py:103-129): connected components
 * of `labels > 0` (skimage.morphology.label, connectivity 1 / 2 / 3 = 6 / 18 / 26 neighbours) and the table of voxel
 * counts per (target component, predicted component) pair.
 *   ccl3d_roots        union-find over voxel indices; parent[v] = root index (smallest index of the component) or -1
 *                      for background; the roots are listed (unordered) in `roots`, their number in *n_roots
 *   ccl3d_relabel      labels[v] = 1 + rank of v's root among `sorted_roots` (ascending) -- the numbering of a raster
 *                      scan, i.e. skimage's; 0 = background.  The caller sorts the (few) roots between the two calls.
 *   overlap_histogram  hist[t * (n_pred + 1) + p] += #voxels with target component t and predicted component p
 *                      (pred == NULL with n_pred == 0: plain per-label voxel counts)
 * by_value == 1 labels an INTEGER image the way skimage.measure.label does (background 0, neighbours connect only when
 * they carry the same value) -- post_processing.py:31 keep_components; by_value == 2 labels the INVERTED mask
 * (values <= 0), the holes skimage's remove_small_holes measures -- post_processing.py:56.
 * Post-processing (post_processing.py:5-73 sort_by_size / unsort_by_size / keep_components / remove_holes /
 * remove_small_components; research/msseg2/competition/ms-inference.py:47-50, research/dmri_hippo/hippo_inference.py:40-44):
 *   relabel_lut        dst[v] = lut[src[v]] (values outside [0, n_lut) are copied)
 *   dilate_cross       grey dilation with the 3-D cross, skimage.morphology.dilation's default footprint
 *   dilate_where       dst = (mask && D != dil_src) ? D : pass_src with D = dilate_cross(dil_src), fused
 *                      (post_processing.py:44-46 and :62)
 *   relabel_masked     dst = keep_lut[comp] ? lut[img] : 0 (post_processing.py:42 `sorted_img * keep`)
 *   label_equals       dst = (src == value);  mask_assign  dst[mask != 0] = value (post_processing.py:69-71) */
int b200seg_ccl3d_roots(const void* mask, int32_t label_bytes, int32_t w, int32_t h, int32_t d, int32_t connectivity,
                        int32_t by_value, int32_t* parent, int32_t* roots, int32_t max_roots, int32_t* n_roots,
                        void* stream);
int b200seg_ccl3d_relabel(const int32_t* parent, int64_t voxels, const int32_t* sorted_roots, int32_t n_roots,
                          int32_t* labels, void* stream);
int b200seg_overlap_histogram(const int32_t* target, const int32_t* pred, int64_t voxels, int32_t n_target,
                              int32_t n_pred, int64_t* hist, void* stream);
int b200seg_relabel_lut(const int32_t* src, int64_t voxels, const int32_t* lut, int32_t n_lut, int32_t* dst, void* stream);
int b200seg_dilate_cross(const int32_t* src, int32_t w, int32_t h, int32_t d, int32_t* dst, void* stream);
int b200seg_dilate_where(const int32_t* dil_src, const int32_t* mask, const int32_t* pass_src, int32_t w, int32_t h,
                         int32_t d, int32_t* dst, void* stream);
int b200seg_relabel_masked(const int32_t* img, const int32_t* lut, int32_t n_lut, const int32_t* comp,
                           const int32_t* keep_lut, int32_t n_keep, int64_t voxels, int32_t* dst, void* stream);
int b200seg_label_equals(const int32_t* src, int64_t voxels, int32_t value, int32_t* dst, void* stream);
int b200seg_mask_assign(int32_t* dst, const int32_t* mask, int64_t voxels, int32_t value, void* stream);

/* ------------------------------------------------------------------------------------------------ criterion
 * HybridLogisticDiceLoss.forward (criterions/hybrid_logistic_dice_loss.py:13-43): prediction / target fp32
 * (N, C, voxels) contiguous.  ONE pass produces sums[n*c][4] = {sum p t, sum p^2 | p, sum t^2 | t,
 * sum t log((p + 1e-8) / (1 + 1e-8))} and out3 = {loss, dice_loss, logistic_loss}; class_weights may be NULL.
 * The backward call writes d loss / d prediction * grad_loss[0] from the saved sums (one elementwise pass). */
int64_t b200seg_hybrid_loss_scratch_bytes(int64_t n, int32_t c);
int b200seg_hybrid_loss_forward(const float* prediction, const float* target, int64_t n, int32_t c, int64_t voxels,
                                float dice_weight, const float* class_weights, int32_t square_dice, void* scratch,
                                int64_t scratch_bytes, float* sums, float* out3, void* stream);
int b200seg_hybrid_loss_backward(const float* prediction, const float* target, const float* sums, int64_t n, int32_t c,
                                 int64_t voxels, float dice_weight, const float* class_weights, int32_t square_dice,
                                 const float* grad_loss, float* grad_prediction, void* stream);

/* ------------------------------------------------------------------------------------------------ training step
 * The device side of the reference's trainer step (segmentation_trainer.py:162-180: model.train(), forward, loss,
 * backward; BatchNorm3d with batch statistics, components.py:53).  fp32 blocked views (the precision path) or bf16 ones
 * (mixed precision: bf16 storage, fp32 arithmetic; all views of a call share one dtype).  Forward convolutions and
 * data gradients are b200seg_conv3d_direct (fp32) / b200seg_conv3d_tc (bf16) launches (a dgrad is a convolution with re-arranged weights); these entry
 * points add the rest.  All reductions are deterministic (per-block partials summed in a fixed order).
 *   train_scratch_bytes  size of the double scratch the reductions need for `channels` channels
 *   channel_moments      mean[c], var[c] (biased) over (N, Z, Y, X)               -- nn.BatchNorm3d training statistics
 *   affine_act           dst = act(scale[c] * src + shift[c]) (+ residual)        -- BN apply + activation + res add
 *   bn_backward          g = dy * act'(scale*z+shift); sum_g[c] = sum g (= d beta), sum_gx[c] = sum g*xhat (= d gamma),
 *                        dz = scale[c] * (g - sum_g/M - xhat * sum_gx/M) with xhat = (z - mean) * rstd; has_norm == 0:
 *                        dz = g (activation only; sum_g is then the bias gradient)
 *   softmax_backward     dlogits = p * (dp - sum_c dp_c p_c) (softmax != 0) or dp, fp32 NCDHW -> blocked view
 *   wgrad                grad[tap][a_ch][b_ch] = sum_{n,pos} A[a_ch](pos) * B[b_ch](stride*pos + tap - pad), zero outside
 *                        B; tap = (tz*k + ty)*k + tx, rows padded to multiples of 8.  3x3x3 layers: A = dz, B = x,
 *                        stride 1, pad 1; BlurConv3d: A = dz, B = x, k 4, stride 2, pad 1; BlurConvTranspose3d: A = x,
 *                        B = dy, k 4, stride 2, pad 1.  scratch: wgrad_scratch_floats() floats.
 *   avgpool2_backward / upsample_trilinear2_backward   adjoints of nn.AvgPool3d(2) (dx = dy(v/2)/8 + add) and of
 *                        nn.Upsample(scale_factor=2, 'trilinear', align_corners=True) (modular_unet.py:38-41), the
 *                        latter as a deterministic gather with the forward kernel's coefficients
 *   channel_scale        dst = src * mask[n][c] (mask: fp32 [N][round_up(C, 8)]): nn.Dropout3d forward / backward with the
 *                        mask the host drew (components.py:70-71, nested_residual_unet.py:44-45) */
int64_t b200seg_train_scratch_bytes(int32_t channels);
int b200seg_channel_scale(b200seg_view src, const float* mask, b200seg_view dst, void* stream);
int b200seg_avgpool2_backward(b200seg_view dy, b200seg_view add, b200seg_view dx, void* stream);
int b200seg_upsample_trilinear2_backward(b200seg_view dy, b200seg_view dx, void* stream);
int b200seg_channel_moments(b200seg_view x, void* scratch, float* mean, float* var, void* stream);
int b200seg_affine_act(b200seg_view src, const float* scale, const float* shift, const float* slope,
                       b200seg_view residual, b200seg_view dst, void* stream);
int b200seg_bn_backward(b200seg_view dy, b200seg_view z, const float* scale, const float* shift, const float* slope,
                        const float* mean, const float* rstd, int32_t has_norm, void* scratch, float* sum_g,
                        float* sum_gx, b200seg_view dz, void* stream);
int b200seg_softmax_backward(const float* probs, const float* dprobs, int32_t n, int32_t c, int32_t softmax,
                             b200seg_view dst, void* stream);
int64_t b200seg_wgrad_scratch_floats(int32_t a_channels, int32_t b_channels, int32_t ksize);
int b200seg_wgrad(b200seg_view a, b200seg_view b, int32_t ksize, int32_t stride, int32_t pad, float* scratch,
                  float* grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SEG_H_ */
