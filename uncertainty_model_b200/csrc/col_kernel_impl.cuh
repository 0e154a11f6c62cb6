// Kernels (1) and (2), hot form: the column-marching fused loss (see
// col_core.cuh for the algorithm).  One CTA = one unit (sample, row strip, view
// set, column tile).  Every pyramid scale is its own launch -- block size, row
// stride, geometry and configuration are then CTA-uniform kernel parameters at
// fixed constant-bank addresses.  This header holds the kernel template and
// the per-class launcher; col_inst_*.cu instantiate one block-size class each
// (so that they compile in parallel), col_kernels.cu plans and dispatches.
#pragma once

#include "col_core.cuh"
#include "col_launch.cuh"

namespace usl {

using namespace ck;

struct ColArgs {
    LossParams P;
    int tiles_x, strips, nv;
    int skip_if_unit;
};

// One step, entered with P1(r) done and its barrier still to come:
//   barrier | P2(r), V(r+1) | barrier | P3(r), ring request, P1(r+1)
// P3(r) and P1(r+1) share an interval: what P1 writes is either private to the
// thread (history, registers) or read only after the next barrier.
template <class C, int PAR>
__device__ __forceinline__ void col_step(const LossParams& P, const CGeo& G,
                                         const CRings& S, int r, int r_last,
                                         int ring_last, CState& T) {
    if (C::MODE != MODE_PLAIN) c_fill_wait();   // row r+2, requested a step ago
    __syncthreads();
    c_p2<C, PAR>(P, G, S, r, T);
    if (C::STEADY || r + 1 <= r_last)
        c_pV<C::SROW, C::MODE, C::STEADY>(P, G, S, r, threadIdx.x, blockDim.x, T);
    __syncthreads();
    c_p3<C, PAR>(P, G, S, r, T);
    if (C::STEADY || r + 1 <= r_last) {
        // row r+3 replaces row r in the ring (its last readers were P2 / V above)
        if (C::MODE == MODE_PLAIN) {
            if (threadIdx.x < 32) {
                const int row = S.RW[c_step_index(G, r + 1)].issue;
                if (C::STEADY || row >= 0) c_ring_issue(P, G, S, C::SROW, row, threadIdx.x);
            }
        } else if (r + 3 <= ring_last) {
            c_ring_fill<C::SROW, C::MODE>(P, G, T, r + 3);
        }
        c_p1<C>(P, G, S, r + 1, T);
    }
}

template <int SROW, bool GRAD, int MODE, int TERMS>
__global__ void __launch_bounds__(col_class_threads(SROW), 512 / col_class_threads(SROW))
col_kernel(const __grid_constant__ ColArgs A) {
    using CG = Cfg<SROW, GRAD, MODE, TERMS, false>;
    using CS = Cfg<SROW, GRAD, MODE, TERMS, true>;
    extern __shared__ float4 smem_raw[];
    __shared__ float red[col_class_threads(SROW) / 32][NUM_ACC];

    const LossParams& P = A.P;
    CGeo G;
    G.gd_up = 1.0f; G.ge_up = 1.0f;
    if (GRAD) {
        if (P.gout_d) G.gd_up = __ldg(P.gout_d);
        if (P.gout_e) G.ge_up = __ldg(P.gout_e);
        if (A.skip_if_unit && G.gd_up == 1.0f && G.ge_up == 1.0f) return;
    }
    {
        int u = blockIdx.x;
        const int tx = u % A.tiles_x; u /= A.tiles_x;
        const int nvs = 2 / A.nv;
        const int vs = u % nvs; u /= nvs;
        const int st = u % A.strips;
        G.b = u / A.strips;
        G.nv = A.nv; G.v0 = vs * A.nv;
        G.xa = tx * P.TW; G.xb = min(P.w, G.xa + P.TW);
        G.cbeg = MODE == MODE_TILED ? max(G.xa - 2, 0) : 0;
        G.LW = (MODE == MODE_TILED ? min(G.xb + 2, P.w) : P.w) - G.cbeg;
        G.ya = st * P.R; G.yb = min(P.h, G.ya + P.R);
        G.qlo = max(G.ya - 2, 0);
        G.sH = ac_scale(P.h - 2, P.h); G.sW = ac_scale(P.w - 2, P.w);
    }
    const CRings S = c_carve(reinterpret_cast<float*>(smem_raw), SROW, P.w,
                             G.nv, P.R, GRAD);
    const int tid = threadIdx.x;
    CState T;
    c_thread_init<SROW, GRAD>(P, G, S, tid, T);
    c_init_unit<SROW, GRAD, MODE>(P, G, S, tid, blockDim.x);
    const int r0 = c_first_row(G), r1 = c_last_row(G);
    const int ring_last = c_last_ring_row(P, G);
    __syncthreads();
    // the ring starts one row above the first step: V(r0) may sample it
    if (MODE == MODE_PLAIN) {
        if (tid < 32)
            for (int row = r0 - 1; row <= r0 + 2 && row <= ring_last; ++row)
                c_ring_issue(P, G, S, SROW, row, tid);
    } else {
        for (int row = r0 - 1; row <= r0 + 2 && row <= ring_last; ++row)
            c_ring_fill<SROW, MODE>(P, G, T, row);
        c_fill_wait();
        __syncthreads();
    }
    c_pV<SROW, MODE, false>(P, G, S, r0 - 1, tid, blockDim.x, T);
    __syncthreads();
    c_p1<CG>(P, G, S, r0, T);
    // r0 = ya - 2 is even (strip heights are even): PAR is the row parity.
    // Interior steps (see Cfg::STEADY; the step also runs P1 of the row after)
    // use the specialised body.
    const int s_lo = G.ya + 2, s_hi = min(G.yb - 1, P.h - 3) - 1;
    int r = r0;
    for (; r < s_lo && r <= r1; r += 2) {
        col_step<CG, 0>(P, G, S, r, r1, ring_last, T);
        col_step<CG, 1>(P, G, S, r + 1, r1, ring_last, T);
    }
    for (; r + 1 <= s_hi; r += 2) {
        col_step<CS, 0>(P, G, S, r, r1, ring_last, T);
        col_step<CS, 1>(P, G, S, r + 1, r1, ring_last, T);
    }
    for (; r <= r1; r += 2) {
        col_step<CG, 0>(P, G, S, r, r1, ring_last, T);
        col_step<CG, 1>(P, G, S, r + 1, r1, ring_last, T);
    }
    if (P.partials) {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int k = 0; k < NUM_ACC; ++k) {
            const float v = warp_sum(T.acc[k]);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (tid < NUM_ACC) {
            float t = 0.0f;
            const int nw = (blockDim.x + 31) >> 5;
            for (int i = 0; i < nw; ++i) t += red[i][tid];
            P.partials[(long long)blockIdx.x * NUM_ACC + tid] = t;
        }
    }
}

template <int SROW, bool GRAD, int MODE, int TERMS>
static int col_launch_one(const ColArgs& A, int grid, int threads, size_t smem,
                          cudaStream_t st) {
    if (cudaFuncSetAttribute(col_kernel<SROW, GRAD, MODE, TERMS>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return USL_ERR_CUDA;
    col_kernel<SROW, GRAD, MODE, TERMS><<<grid, threads, smem, st>>>(A);
    return check_launch();
}

// One scale in block-size class CLS.
template <int CLS>
int col_launch_class(const ColPlan* M, int i, bool grad, int skip_if_unit,
                     cudaStream_t st) {
    constexpr int SROW = col_class_srow(CLS);
    ColArgs A;
    A.P = M->P[i];
    A.tiles_x = M->tiles_x[i]; A.strips = M->strips[i]; A.nv = M->nv[i];
    A.skip_if_unit = skip_if_unit;
    const int grid = M->units[i], nt = M->threads[i], mode = M->mode[i];
    const size_t smem = M->smem[i];
    // the hot variant has no optional inputs / outputs compiled in
    const bool hot = (A.P.terms == COL_HOT_TERMS) && !A.P.grad_recon_in &&
                     !A.P.err_out && !A.P.recon_out;
#define USL_COL_GO(GRAD, MODE, TERMS) \
    return col_launch_one<SROW, GRAD, MODE, TERMS>(A, grid, nt, smem, st)
    if (mode == MODE_PLAIN) {
        if (grad) { if (hot) USL_COL_GO(true, MODE_PLAIN, COL_HOT_TERMS); USL_COL_GO(true, MODE_PLAIN, -1); }
        if (hot) USL_COL_GO(false, MODE_PLAIN, COL_HOT_TERMS);
        USL_COL_GO(false, MODE_PLAIN, -1);
    }
    if (mode == MODE_MASKED) {
        if (grad) USL_COL_GO(true, MODE_MASKED, -1);
        USL_COL_GO(false, MODE_MASKED, -1);
    }
    if constexpr (CLS == COL_MAX_THREADS) {      // column tiles: widest class only
        if (grad) { if (hot) USL_COL_GO(true, MODE_TILED, COL_HOT_TERMS); USL_COL_GO(true, MODE_TILED, -1); }
        if (hot) USL_COL_GO(false, MODE_TILED, COL_HOT_TERMS);
        USL_COL_GO(false, MODE_TILED, -1);
    }
#undef USL_COL_GO
    return USL_ERR_UNSUPPORTED;
}

}  // namespace usl
