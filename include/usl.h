/* libusl.so -- C ABI of the B200-native stereo-uncertainty loss path.
 *
 * This is the drop-in boundary.  The reference (Probabilistic-Surgical-Vision/
 * uncertainty-model) is pure Python on top of ATen and has no FFI of its own;
 * each entry point below replaces the ATen call sequence of one reference
 * function (cited as file:line under /root/reference) and is bound from Python
 * with ctypes (uncertainty_model_b200/_lib.py; the stub a reference maintainer
 * would add is shown in INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (unless stated), owned by the
 *     caller (PyTorch's allocator); the library never allocates, frees or
 *     retains caller memory;
 *   - image-like tensors are NCHW with contiguous h*w planes; `*_bs` / `*_cs`
 *     are the batch / channel strides in elements, so channel slices of a
 *     larger tensor (prediction[:, :2]) are passed without a copy;
 *   - all work is enqueued on `stream` (a cudaStream_t of the device that owns
 *     the tensors); nothing synchronises the host.  The tensors need not live
 *     on the caller's current device: every call makes their device current
 *     for its duration and restores the caller's (the reference's DDP launcher
 *     never calls torch.cuda.set_device, parallel_main.py:152-160);
 *   - return value: USL_OK or a negative UslError; never throws, never exits.
 *   - re-entrant.  State: tables that are written once and immutable
 *     afterwards (tuning knobs read from the environment at first use, SM
 *     counts), a launch counter (usl_launch_count) and, per host thread and
 *     device, a small pool of side streams and events for the concurrent
 *     per-scale launches.
 */
#ifndef USL_H_
#define USL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define USL_VERSION 200
#define USL_MAX_SCALES 8
#define USL_NUM_TERMS 6

typedef enum UslError {
    USL_OK = 0,
    USL_ERR_ARG = -1,         /* null pointer / non-positive size / bad enum */
    USL_ERR_CUDA = -2,        /* launch failed (cudaGetLastError) */
    USL_ERR_UNSUPPORTED = -3, /* shape outside what the kernels handle */
    USL_ERR_WORKSPACE = -4    /* workspace too small */
} UslError;

int usl_version(void);
const char* usl_strerror(int rc);
/* kernel launches this process has issued through the library so far */
long long usl_launch_count(void);
/* Profiling aid.  slots: device buffer of USL_TIMELINE_SLOTS uint64 (or NULL to
 * switch it off).  While set, one-thread marker kernels write %globaltimer
 * (ns) into it around the launches of a training step: 0/1 before/after the
 * pyramid kernel, 2+i after scale i's fused kernel, 6+i after its transposed
 * warp, 10 after the reduction, 11 after the combination, 12 after the
 * rescale; 13 / 14 = the start and the last stamp of the step before.  Works inside a CUDA-graph replay (tools/step_timeline.py). */
#define USL_TIMELINE_SLOTS 16
int usl_debug_timeline(void* slots);

/* ---- train/utils.py:27-50  scale_pyramid ---------------------------------
 * src (B,C,H,W).  dst[i], i = 1..scales-1, contiguous (B,C,H>>i,W>>i);
 * dst[0] is ignored (level 0 is the input itself, aliased by the host). */
int usl_pyramid(const float* src, int B, int C, int H, int W, long long src_bs,
                long long src_cs, int scales, float* const* dst, void* stream);

/* ---- train/utils.py:65-109  reconstruct / reconstruct_{left,right}_image --
 * out[b,c] = grid_sample(image[b,c], base + sign*disp[b]) ; sign = -1 for the
 * left reconstruction, +1 for the right one. */
int usl_warp_fwd(const float* disp, long long disp_bs, float sign,
                 const float* image, long long img_bs, long long img_cs, int B,
                 int C, int h, int w, float* out, long long out_bs,
                 long long out_cs, void* stream);
/* gradient of the above w.r.t. disp (gather; deterministic). */
int usl_warp_bwd_disp(const float* disp, long long disp_bs, float sign,
                      const float* image, long long img_bs, long long img_cs,
                      const float* grad_out, long long go_bs, long long go_cs,
                      int B, int C, int h, int w, float* grad_disp,
                      long long gd_bs, void* stream);

/* gradient of the above w.r.t. the SAMPLED image (the reference's op is
 * differentiable there too, utils.py:96-97: ATen's grid_sampler_2d_backward with
 * float atomics; here deterministic, in two launches).  workspace:
 * usl_warp_bwd_image_workspace_bytes(B, C, h, w) bytes, contents irrelevant. */
long long usl_warp_bwd_image_workspace_bytes(int B, int C, int h, int w);
int usl_warp_bwd_image(const float* disp, long long disp_bs, float sign,
                       const float* grad_out, long long go_bs, long long go_cs,
                       int B, int C, int h, int w, void* workspace,
                       float* grad_image, long long gi_bs, long long gi_cs,
                       void* stream);

/* ---- train/loss.py  fused per-scale loss ----------------------------------
 * Term bits (UslLossConfig.terms) and the index of their raw sum in `sums`: */
#define USL_TERM_REPROJ 1u   /* [0] WeightedSSIMLoss            loss.py:133-151 */
#define USL_TERM_CONS_D 2u   /* [1] ConsistencyLoss(disp)       loss.py:167-188 */
#define USL_TERM_SMOOTH_D 4u /* [2] SmoothnessLoss(disp, img)   loss.py:248-264 */
#define USL_TERM_UNC 8u      /* [3] l1|bayesian|log_bayesian    loss.py:389-403 */
#define USL_TERM_SMOOTH_U 16u/* [4] SmoothnessLoss(unc, img)    loss.py:428-429 */
#define USL_TERM_CONS_U 32u  /* [5] ConsistencyLoss(unc, disp)  loss.py:430-431 */

#define USL_LOSS_L1 0
#define USL_LOSS_BAYESIAN 1
#define USL_LOSS_LOG_BAYESIAN 2

typedef struct UslLossConfig {
    uint32_t terms;
    int32_t loss_type;          /* USL_LOSS_*  (loss.py:364-375) */
    float alpha, c1, c2;        /* loss.py:30-32 (c1 = k1^2, c2 = k2^2) */
    /* loss contribution of raw sum k is coef[k] * sums[k]: the host folds the
     * config weights (loss.py:560-566), 1/N means and the 1/2^i of
     * loss.py:546 into it.  Terms 0-2 feed output 0 (disparity loss), terms
     * 3-5 output 1 (error loss). */
    float coef[USL_NUM_TERMS];
} UslLossConfig;

/* keep this call on the general strip kernels (tests: same result either way) */
#define USL_SCALE_GENERAL_KERNELS 1

typedef struct UslLossScale {
    int32_t B, h, w;
    int32_t flags;              /* USL_SCALE_* (0 = default) */
    const float* images;  int64_t img_bs, img_cs;    /* (B,6,h,w) L_rgb,R_rgb */
    const float* disp;    int64_t disp_bs, disp_cs;  /* (B,2,h,w) d_L,d_R     */
    const float* unc;     int64_t unc_bs, unc_cs;    /* (B,2,h,w) u_L,u_R     */
    const float* recon_in; int64_t rin_bs, rin_cs;   /* optional (B,6,h,w): use
                              this reconstruction instead of warping in-kernel */
    const float* err_in;  int64_t ein_bs, ein_cs;    /* optional (B,2,h,w): use
                              this error map (ReprojectionErrorLoss stand-alone) */
    float* recon_out;            /* optional, contiguous (B,6,h,w) [forward]  */
    float* err_out;              /* optional, contiguous (B,2,h,w) [forward]  */
    const float* grad_recon_in;  /* optional, contiguous (B,6,h,w) [backward]:
                              extra dL/d(recon_out), e.g. from a discriminator */
    float* grad_disp;     int64_t gd_bs, gd_cs;      /* (B,2,h,w) [backward]  */
    float* grad_unc;      int64_t gu_bs, gu_cs;      /* (B,2,h,w) [backward]  */
    float* grad_recon_out;       /* contiguous (B,6,h,w) [backward, recon_in] */
    float* scatter_ws;           /* optional workspace of usl_loss_grad /
                              usl_loss_bwd, 32*B*h*w bytes, 16-byte aligned:
                              the fused kernels leave {d, u, signed coefficient
                              of either consistency term} of every pixel there
                              and the transposed warp runs on the lane-per-row
                              kernel fed from it (otherwise on the warp-per-row
                              kernel, which recomputes the warp) */
} UslLossScale;

/* All pyramid scales of a step are processed by ONE launch: `cfgs` and
 * `scales` are HOST arrays of n_scales entries (largest scale first).
 *
 * Two kernel families sit behind these entry points.  The column-marching
 * kernels (csrc/col_core.cuh) take every call whose scales all warp in-kernel
 * (no recon_in / err_in) and have the reprojection term on -- the training
 * step; everything else (given reconstructions or error maps, stand-alone
 * terms, rows too small) runs on the general strip kernels
 * (csrc/loss_core.cuh).
 *
 * usl_loss_plan: row offsets into `partials` (USL_NUM_TERMS floats per row) of
 * every scale for a launch in `mode`: cta_starts[n_scales + 1] (HOST).
 * USL_MODE_GRAD is the one-pass sums+gradient launch (usl_loss_grad); it
 * returns USL_ERR_UNSUPPORTED when the call does not qualify for it. */
#define USL_MODE_FWD 0
#define USL_MODE_GRAD 1
int usl_loss_plan(const UslLossConfig* cfgs, const UslLossScale* scales,
                  int n_scales, int mode, int* cta_starts);
/* (legacy) rows the general forward kernel uses for one scale; <0 on error. */
int usl_loss_fwd_ctas(const UslLossScale* s);
/* forward: partial sums of the enabled terms -> partials (scale-major, scale i
 * starting at row cta_starts[i] of usl_loss_plan(USL_MODE_FWD)). */
int usl_loss_fwd(const UslLossConfig* cfgs, const UslLossScale* scales,
                 int n_scales, float* partials, void* stream);
/* fixed-order fp64 reduction of the partials -> sums[n_scales][6] (device).
 * cta_starts: HOST array of n_scales+1 row offsets into `partials`. */
int usl_loss_reduce(const float* partials, const int* cta_starts, int n_scales,
                    double* sums, void* stream);
/* *out_disp = sum_s sum_{k<3} coef[s][k]*sums[s][k]; *out_err likewise for
 * k>=3.  All device pointers (coef: fp32 [n_scales][6]). */
int usl_loss_combine(const double* sums, const float* coef, int n_scales,
                     float* out_disp, float* out_err, void* stream);
/* backward: gout_* = device scalars, the upstream gradients of the two
 * outputs (NULL = that output does not take part in the backward).  Writes
 * grad_disp / grad_unc (and grad_recon_out when recon_in is given) of every
 * scale.  Two launches: the deterministic transposed warp of the consistency
 * terms, then the fused stencil backward. */
int usl_loss_bwd(const UslLossConfig* cfgs, const UslLossScale* scales,
                 int n_scales, const float* gout_disp, const float* gout_err,
                 int stages, void* stream);
/* `stages`: which of the two backward launches to enqueue (both for a real
 * backward; bench.py times them one at a time). */
#define USL_BWD_STAGE_SCATTER 1
#define USL_BWD_STAGE_MAIN 2
#define USL_BWD_STAGE_ALL 3
/* forward AND backward in one pass over the inputs (the backward needs every
 * forward intermediate anyway): writes the partial sums (rows per
 * usl_loss_plan(USL_MODE_GRAD); `partials` may be NULL) and grad_disp /
 * grad_unc of every scale, scaled by the upstream gradients gout_* (device
 * scalars; NULL = 1).  The training loop back-propagates disp_loss +
 * error_loss (reference train.py:126-128), i.e. both upstream gradients are 1:
 * the host runs this in `forward` with gout = NULL and, in `backward`, calls it
 * again with the real upstream gradients and USL_GRAD_SKIP_IF_UNIT -- every
 * CTA then returns at once if they are both 1 (nothing to redo), and
 * recomputes the gradients otherwise.  No host synchronisation either way. */
#define USL_GRAD_SKIP_IF_UNIT 1
/* The launch sequence in two halves, so that a caller can put work that only
 * needs the partial sums (their reduction, an all-reduce across ranks) between
 * them: NO_SCATTER = the fused kernels only (sums and gradients but for the
 * transposed warp of the consistency terms), ONLY_SCATTER = that missing part,
 * added to grad_disp.  Neither flag = both, in this order. */
#define USL_GRAD_NO_SCATTER 2
#define USL_GRAD_ONLY_SCATTER 4
/* ... or, keeping the transposed warp of every scale but the largest right
 * behind that scale's fused kernel (where it is hidden): DEFER_SCATTER0 =
 * everything but the largest scale's transposed warp, ONLY_SCATTER0 = that. */
#define USL_GRAD_DEFER_SCATTER0 8
#define USL_GRAD_ONLY_SCATTER0 16
int usl_loss_grad(const UslLossConfig* cfgs, const UslLossScale* scales,
                  int n_scales, const float* gout_disp, const float* gout_err,
                  float* partials, int flags, void* stream);

/* Batch sharded over ranks (reference parallel_main.py: one process per GPU,
 * loss.py:560-566 evaluated per rank): usl_loss_grad for unit upstream
 * gradients, AND usl_loss_reduce(partials, cta_starts, n_scales, sums)
 * enqueued on `reduce_stream` behind the column kernels only -- not behind the
 * transposed warps, which need another ~25 % of the step.  The caller enqueues
 * its exchange of `sums` between ranks and usl_loss_combine behind
 * `reduce_stream` and joins it to `stream`; both then run beside the
 * transposed warp (worth it on one GPU too: the reduction and the combination
 * leave the critical path).  reduce_stream: a cudaStream_t of the same device, different
 * from `stream` (it is made to wait for work forked from `stream`, so it takes
 * part in a stream capture of `stream`).  USL_ERR_UNSUPPORTED where
 * usl_loss_grad would not run its per-scale transposed warps (the caller then
 * uses usl_loss_grad + usl_loss_reduce). */
int usl_loss_grad_sharded(const UslLossConfig* cfgs, const UslLossScale* scales,
                          int n_scales, float* partials, const int* cta_starts,
                          double* sums, void* reduce_stream, void* stream);

/* The gradients usl_loss_grad wrote are those for unit upstream gradients, and
 * they are linear in the upstream pair.  When BOTH outputs receive the SAME
 * upstream gradient g (the training loop back-propagates disp_loss +
 * error_loss, reference train.py:126-128; a loss scaler multiplies that sum)
 * the backward is this one launch: every buffer is multiplied by *g in place,
 * and nothing at all is touched when *g == 1 (decided on the device: no host
 * synchronisation).  buffers / counts: HOST arrays of n device pointers and
 * element counts. */
#define USL_MAX_RESCALE 16
int usl_grad_rescale(const float* g, float* const* buffers,
                     const long long* counts, int n, void* stream);

/* ---- model/layers/decoder.py:239-246  disparity head ---------------------
 * pred = scale * sigmoid(logits) for every pyramid level in one launch, and
 * its backward grad_logits = grad_pred * pred * (1 - pred / scale).  HOST
 * arrays of `levels` device pointers / element counts. */
int usl_head_fwd(const float* const* logits, float* const* pred,
                 const long long* counts, int levels, float scale, void* stream);
int usl_head_bwd(const float* const* grad_pred, const float* const* pred,
                 float* const* grad_logits, const long long* counts, int levels,
                 float scale, void* stream);

/* ---- train/utils.py:53-62,138-140,248-273  discriminator input ------------
 * out (2B,6,h,w) of every level = [image level ; its reconstruction] along the
 * batch axis, one launch for all levels: the reconstruction half is warped on
 * the fly from `pred` (channels 0, 1 = d_L, d_R; utils.py:112-135) -- the
 * reconstruction pyramid is never materialised -- or, with pred = NULL,
 * copied from a materialised `recon` level. */
typedef struct UslDiscLevel {
    int32_t B, h, w, reserved;
    const float* images; int64_t img_bs, img_cs;   /* (B,6,h,w) */
    const float* pred;   int64_t pred_bs, pred_cs; /* (B,>=2,h,w) or NULL */
    const float* recon;  int64_t rec_bs, rec_cs;   /* (B,6,h,w) when pred is NULL */
    float* out;                                    /* contiguous (2B,6,h,w) */
} UslDiscLevel;
int usl_disc_input(const UslDiscLevel* levels, int n_levels, void* stream);

/* A gradient arriving at the reconstructions of every level (from the
 * discriminator, train/loss.py:552-558) -> the disparity channels 0/1 of the
 * prediction gradients, stored or added (`accumulate`): the transpose of the
 * warp of utils.py:65-135 w.r.t. its shift, one launch for all levels.
 * levels[i].images / pred as above (out, recon unused); grad_recon[i]
 * contiguous (B,6,h,w); grad_pred[i] with batch / channel strides gp_bs[i],
 * gp_cs[i].  HOST arrays. */
int usl_recon_bwd(const UslDiscLevel* levels, const float* const* grad_recon,
                  float* const* grad_pred, const long long* gp_bs,
                  const long long* gp_cs, int n_levels, int accumulate,
                  void* stream);

/* ---- train/evaluate.py:142-146  torchmetrics SSIM ---------------------------
 * structural_similarity_index_measure(preds, target, gaussian_kernel=True,
 * sigma, kernel_size=k, data_range, k1, k2): per-image values -> per_image[B]
 * (device fp32; reduction='sum' is their sum).  (B,C,H,W) with contiguous
 * planes; k odd, <= 15.  See csrc/ssim.cu for the restated algorithm. */
size_t usl_ssim_workspace_bytes(int B, int C, int H, int W, int k);
int usl_ssim_gauss(const float* pred, long long p_bs, long long p_cs,
                   const float* target, long long t_bs, long long t_cs, int B,
                   int C, int H, int W, int k, float sigma, float data_range,
                   float k1, float k2, float* per_image, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ---- train/utils.py:199-245  combine_disparity ----------------------------
 * left, right: contiguous fp32 (planes,h,w); out fp64 like the numpy original. */
int usl_combine_disparity(const float* left, const float* right, int planes,
                          int h, int w, double alpha, double beta, double* out,
                          void* stream);
/* ---- train/utils.py:177-196  to_heatmap -----------------------------------
 * x: n fp32 values; lut: device fp64 [entries][3] (the colour map's table);
 * out fp64 (3, n).  matplotlib's Colormap.__call__ indexing. */
int usl_heatmap(const float* x, long long n, int inverse, const double* lut,
                int entries, double* out, void* stream);

/* 3x3 valid mean (loss.py:386-387, `pooling=True`) and its transpose. */
int usl_pool3_fwd(const float* x, long long x_bs, long long x_cs, int B, int C,
                  int h, int w, float* out, void* stream);
int usl_pool3_bwd(const float* grad_out, int B, int C, int h, int w,
                  float* grad_x, void* stream);

/* ---- train/sparsification.py  curve / ause --------------------------------
 * rows = frames*2 maps of (H,W); k = pooling kernel; n = (H-k+1)*(W-k+1). */
#define USL_MAX_STEPS 512
#define USL_SPARS_MAX_KERNEL 15
size_t usl_spars_workspace_bytes(int rows, int H, int W, int k, int with_order);
/* cuts: HOST array of steps+1 ints, removed_k then n (sparsification.py:26-27).
 * row_norm_sum: device fp64[steps], sum over rows of the normalised tail means.
 * order_out: optional device int32[rows*n], the stable descending permutation
 * of the pooled predicted error; pooled_*_out optional device fp32[rows*n]. */
int usl_spars_curve(const float* oracle, const float* predicted, int rows,
                    int H, int W, int k, const int* cuts, int steps,
                    double* row_norm_sum, int32_t* order_out,
                    float* pooled_oracle_out, float* pooled_pred_out,
                    void* workspace, size_t workspace_bytes, void* stream);
/* curve[k] = fp32(row_norm_sum[k] / total_rows)  (sparsification.py:32-36). */
int usl_spars_finish(const double* row_norm_sum, int steps, long long total_rows,
                     float* curve, void* stream);
/* sparsification.py:46-57: fp32( sum_k fp32(pred_k - oracle_k) / steps ). */
int usl_spars_ause(const float* oracle_curve, const float* pred_curve,
                   int steps, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* USL_H_ */
