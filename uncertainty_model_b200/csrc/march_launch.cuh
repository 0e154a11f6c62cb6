// Launch interface of the register-marching loss kernels (march_kernels.cu),
// used by the C-ABI entry points in loss_kernels.cu.
#pragma once

#include "loss_core.cuh"
#include "usl_common.cuh"

namespace usl {

struct MarchPlan {
    LossParams P[USL_MAX_SCALES];
    int cta_start[USL_MAX_SCALES + 1];   // CTA (= partials row) offsets
    int tiles_x[USL_MAX_SCALES], strips[USL_MAX_SCALES];
    int nsub[USL_MAX_SCALES];            // independent sub-blocks per CTA
    int units[USL_MAX_SCALES];           // B * strips * tiles_x
    int n;
    int threads;                         // blockDim.x
    size_t smem;                         // dynamic shared memory per CTA
};

// True when every scale can run on the marching kernels: warp in-kernel (no
// given reconstruction / error map), reprojection term on, even width.
bool march_eligible(const UslLossConfig* cfgs, const UslLossScale* scales, int n);

// Fills the plan (tiling, CTA offsets); `P[i]` must already hold the tensors
// and configuration of scale i (see fill_params in loss_kernels.cu).
int march_plan(MarchPlan* M, bool grad);

// grad = false: per-CTA partial sums only.  grad = true: partial sums (when
// P[i].partials is set) and the gradients, in one pass.  `skip_if_unit`: every
// CTA returns at once when both upstream gradients equal 1 (the speculative
// forward already produced exactly these gradients).
int march_launch(const MarchPlan* M, bool grad, int skip_if_unit, cudaStream_t st);

}  // namespace usl
