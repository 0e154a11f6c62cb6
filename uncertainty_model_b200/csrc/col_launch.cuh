// Launch interface of the column-marching loss kernels (col_kernels.cu), used
// by the C-ABI entry points in loss_kernels.cu.
#pragma once

#include "loss_core.cuh"
#include "usl_common.cuh"

namespace usl {

constexpr int COL_MAX_THREADS = 512;
// the term set of the training configuration (config.yml): everything but the
// smoothness of the uncertainty -- compiled with the term tests folded away
constexpr int COL_HOT_TERMS = TERM_REPROJ | TERM_CONS_D | TERM_SMOOTH_D | TERM_UNC | TERM_CONS_U;

// Block-size classes: a unit of nv * LW columns runs in the smallest class that
// holds it; the class fixes the (compile-time) shared-memory row stride.
__host__ __device__ constexpr int col_class_srow(int cls) { return cls + 16; }
__host__ __device__ constexpr int col_class_threads(int srow) { return srow - 16; }

struct ColPlan {
    LossParams P[USL_MAX_SCALES];
    int tiles_x[USL_MAX_SCALES], strips[USL_MAX_SCALES];
    int nv[USL_MAX_SCALES];              // views per unit (1 or 2)
    int units[USL_MAX_SCALES];           // CTAs: B * strips * (2 / nv) * tiles_x
    int cls[USL_MAX_SCALES];             // block-size class (64 .. 512)
    int threads[USL_MAX_SCALES];         // blockDim.x
    int mode[USL_MAX_SCALES];            // plain / masked / tiled
    size_t smem[USL_MAX_SCALES];         // dynamic shared memory per CTA
    int row_start[USL_MAX_SCALES + 1];   // partial-sum row offsets (one per unit)
    int n;
};

// True when every scale can run on the column kernels: warp in-kernel (no
// given reconstruction / error map), reprojection term on.
bool col_eligible(const UslLossConfig* cfgs, const UslLossScale* scales, int n);

// Fills the plan (tiling, unit counts, row offsets); `P[i]` must already hold
// the tensors and configuration of scale i (fill_params in loss_kernels.cu).
int col_plan(ColPlan* M, bool grad);

// One scale.  grad = false: per-unit partial sums only.  grad = true: partial
// sums (when P[i].partials is set) and the gradients, in one pass.
// `skip_if_unit`: every CTA returns at once when both upstream gradients are 1.
int col_launch_scale(const ColPlan* M, int i, bool grad, int skip_if_unit,
                     cudaStream_t st);

// Every scale of the plan: scale 0 on `st`, the others forked onto side streams
// (they run concurrently and join `st` again); USL_COL_SERIAL=1 keeps all of
// them on `st`.
// `after` (optional): called right behind the launch of every scale, on that
// scale's stream -- work that depends on one scale only (its transposed warp)
// then runs while the other scales are still busy.
typedef int (*ColAfter)(void* ctx, int scale, cudaStream_t stream);
int col_launch(const ColPlan* M, bool grad, int skip_if_unit, cudaStream_t st,
               ColAfter after = nullptr, void* after_ctx = nullptr);

}  // namespace usl
