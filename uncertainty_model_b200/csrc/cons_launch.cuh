// Launch interface of the consistency-scatter kernels.
#pragma once

#include "cons_core.cuh"
#include "usl_common.cuh"

namespace usl {

struct MultiCons {
    ConsParams P[USL_MAX_SCALES];
    int cta_start[USL_MAX_SCALES + 1];
    int strips[USL_MAX_SCALES];
    int lanes[USL_MAX_SCALES];   // (lane-per-row kernel) private rows per CTA
    int n;
    int skip_if_unit;   // return at once when both upstream gradients are 1
};

// Warp-per-row scatter (cons_kernels.cu): fills strips / cta_start of `C`
// (tensors, terms and coefficients of every scale must be set) and launches.
int cons_scatter2_launch(MultiCons* C, cudaStream_t st);

// Lane-per-row scatter (cons_rows.cu) fed by the column kernels: every
// ConsParams::scat must be set.  USL_ERR_UNSUPPORTED for rows too wide for it.
int cons_rows_launch(MultiCons* C, cudaStream_t st);

}  // namespace usl
