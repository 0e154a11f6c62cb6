// Kernels (1) and (2), hot form: the column-marching fused loss (see
// col_core.cuh for the algorithm).  This file owns the unit scheduling, the
// shared-memory arena and the reductions.
//
// One CTA = one unit (sample, row strip, view set, column tile).  Every pyramid
// scale is its own launch -- block size, row stride and all of the geometry /
// configuration are then CTA-uniform kernel parameters at fixed constant-bank
// addresses -- and the launches of a step run concurrently (col_launch forks
// them over side streams), largest scale first, so the short CTAs of the small
// scales fill the tail of the large one.
#include <stdlib.h>

#include "col_core.cuh"
#include "col_launch.cuh"

namespace usl {

using namespace ck;

struct ColArgs {
    LossParams P;
    int tiles_x, strips, nv;
    int skip_if_unit;
};

enum { MODE_PLAIN = 0, MODE_MASKED = 1, MODE_TILED = 2 };

template <int SROW, bool GRAD, bool TILED, bool MASKED, int PAR>
__device__ __forceinline__ void col_step(const LossParams& P, const CGeo& G,
                                         const CRings& S, int r, int r_last,
                                         CState& T) {
    if (r + 1 <= r_last) c_load_row<MASKED>(P, T, r + 1, T.xn, T.dn, T.un);
    c_p1<SROW, GRAD, MASKED>(P, G, r, T);
    __syncthreads();
    c_p2<SROW, GRAD, MASKED, PAR>(P, G, S, r, T);
    __syncthreads();
    c_p3<SROW, GRAD, MASKED, PAR>(P, G, S, r, T);
    if (r + 1 <= r_last)
        c_pV<TILED, MASKED>(P, G, S, r + 1, threadIdx.x, blockDim.x, T);
    c_advance(T);
    __syncthreads();
}

template <int SROW, bool GRAD, int MODE>
__global__ void __launch_bounds__(col_class_threads(SROW), 512 / col_class_threads(SROW))
col_kernel(const __grid_constant__ ColArgs A) {
    constexpr bool TILED = MODE == MODE_TILED;
    constexpr bool MASKED = MODE != MODE_PLAIN;
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
        G.cbeg = TILED ? max(G.xa - 2, 0) : 0;
        G.LW = (TILED ? min(G.xb + 2, P.w) : P.w) - G.cbeg;
        G.ya = st * P.R; G.yb = min(P.h, G.ya + P.R);
        G.qlo = max(G.ya - 2, 0);
        G.sH = ac_scale(P.h - 2, P.h); G.sW = ac_scale(P.w - 2, P.w);
    }
    const CRings S = c_carve(reinterpret_cast<float*>(smem_raw), SROW, P.w,
                             G.nv, P.R, GRAD);
    const int tid = threadIdx.x;
    CState T;
    c_thread_init<GRAD>(P, G, S, tid, T);
    c_init_unit<SROW, GRAD>(P, G, S, tid, blockDim.x);
    const int r0 = c_first_row(G), r1 = c_last_row(G);
    c_load_row<MASKED>(P, T, r0, T.x, T.d, T.u);
    __syncthreads();
    c_pV<TILED, MASKED>(P, G, S, r0, tid, blockDim.x, T);
    __syncthreads();
    // r0 = ya - 2 is even (strip heights are even): PAR is the row parity
    for (int r = r0; r <= r1; r += 2) {
        col_step<SROW, GRAD, TILED, MASKED, 0>(P, G, S, r, r1, T);
        col_step<SROW, GRAD, TILED, MASKED, 1>(P, G, S, r + 1, r1, T);
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

// ------------------------------------------------------------------ host ---
static int env_int3(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return x > 0 ? x : dflt;
}

bool col_eligible(const UslLossConfig* cfgs, const UslLossScale* scales, int n) {
    if (getenv("USL_NO_COL")) return false;
    for (int i = 0; i < n; ++i) {
        const UslLossScale& s = scales[i];
        const unsigned t = cfgs[i].terms;
        if (!(t & TERM_REPROJ) || s.recon_in || s.err_in) return false;
        if (!s.images || !s.disp) return false;
        if ((t & (TERM_UNC | TERM_SMOOTH_U | TERM_CONS_U)) && !s.unc) return false;
        if (s.w < 4 || s.h < 4) return false;
        // 32-bit element offsets inside every tensor
        const long long lim = 0x7fffffffLL;
        if ((long long)s.B * s.img_bs > lim || (long long)s.B * s.disp_bs > lim ||
            (long long)s.B * s.unc_bs > lim || (long long)s.B * s.gd_bs > lim ||
            (long long)s.B * s.gu_bs > lim)
            return false;
    }
    return true;
}

int col_plan(ColPlan* M, bool grad) {
    const int maxT = COL_MAX_THREADS;
    const int R0 = env_int3("USL_COL_R0", 32), R1 = env_int3("USL_COL_R", 16);
    long long rows = 0;
    for (int i = 0; i < M->n; ++i) {
        LossParams& p = M->P[i];
        int nv, tiles = 1, TW = p.w, LW = p.w;
        if (2 * p.w <= maxT) nv = 2;
        else if (p.w <= maxT) nv = 1;
        else {
            nv = 1;
            tiles = (p.w + (maxT - 4) - 1) / (maxT - 4);
            TW = (p.w + tiles - 1) / tiles;
            tiles = (p.w + TW - 1) / TW;
            LW = TW + 4;
        }
        // strip height: long strips for the big scale (less halo work), short
        // ones for the small scales (they fill the tail of the step)
        int wantR = (i == 0) ? R0 : R1;
        int strips = (p.h + wantR - 1) / wantR;
        int R = (((p.h + strips - 1) / strips) + 1) & ~1;
        strips = (p.h + R - 1) / R;
        p.TW = TW; p.R = R; p.LW = LW;
        M->nv[i] = nv; M->tiles_x[i] = tiles; M->strips[i] = strips;
        M->units[i] = tiles * strips * p.B * (2 / nv);
        const int nt = nv * LW;
        int cls = 64;
        while (cls < nt) cls *= 2;
        if (cls > maxT) return USL_ERR_UNSUPPORTED;
        M->cls[i] = cls;
        M->threads[i] = (nt + 31) & ~31;
        M->mode[i] = tiles > 1 ? MODE_TILED : ((nt & 31) ? MODE_MASKED : MODE_PLAIN);
        M->smem[i] = ck::c_floats(col_class_srow(cls), p.w, nv, R, grad) * sizeof(float);
        if (M->smem[i] > 220 * 1024) return USL_ERR_UNSUPPORTED;
        M->row_start[i] = (int)rows;
        rows += M->units[i];
    }
    M->row_start[M->n] = (int)rows;
    return USL_OK;
}

template <int SROW, bool GRAD>
static int launch_class(const ColArgs& A, int mode, int grid, int threads,
                        size_t smem, cudaStream_t st) {
#define USL_COL_LAUNCH(MODE)                                                    \
    do {                                                                        \
        if (cudaFuncSetAttribute(col_kernel<SROW, GRAD, MODE>,                  \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem) != cudaSuccess)                     \
            return USL_ERR_CUDA;                                                \
        col_kernel<SROW, GRAD, MODE><<<grid, threads, smem, st>>>(A);           \
    } while (0)
    if (mode == MODE_PLAIN) USL_COL_LAUNCH(MODE_PLAIN);
    else if (mode == MODE_MASKED) USL_COL_LAUNCH(MODE_MASKED);
    else {
        // column tiles only exist in the widest class
        if constexpr (SROW == col_class_srow(COL_MAX_THREADS)) USL_COL_LAUNCH(MODE_TILED);
        else return USL_ERR_UNSUPPORTED;
    }
#undef USL_COL_LAUNCH
    return check_launch();
}

template <bool GRAD>
static int launch_scale(const ColPlan* M, int i, int skip_if_unit, cudaStream_t st) {
    ColArgs A;
    A.P = M->P[i];
    A.tiles_x = M->tiles_x[i]; A.strips = M->strips[i]; A.nv = M->nv[i];
    A.skip_if_unit = skip_if_unit;
    const int grid = M->units[i], nt = M->threads[i];
    const size_t smem = M->smem[i];
    switch (M->cls[i]) {
        case 512: return launch_class<col_class_srow(512), GRAD>(A, M->mode[i], grid, nt, smem, st);
        case 256: return launch_class<col_class_srow(256), GRAD>(A, M->mode[i], grid, nt, smem, st);
        case 128: return launch_class<col_class_srow(128), GRAD>(A, M->mode[i], grid, nt, smem, st);
        case 64: return launch_class<col_class_srow(64), GRAD>(A, M->mode[i], grid, nt, smem, st);
    }
    return USL_ERR_UNSUPPORTED;
}

int col_launch_scale(const ColPlan* M, int i, bool grad, int skip_if_unit,
                     cudaStream_t st) {
    return grad ? launch_scale<true>(M, i, skip_if_unit, st)
                : launch_scale<false>(M, i, skip_if_unit, st);
}

// Side streams for the concurrent per-scale launches: one small pool per host
// thread and device, created on first use and never destroyed (immutable
// afterwards; nothing is shared between host threads).
namespace {
constexpr int POOL_SIDE = USL_MAX_SCALES - 1;
struct StreamPool {
    bool ready = false, failed = false;
    cudaStream_t side[POOL_SIDE];
    cudaEvent_t fork, join[POOL_SIDE];
};
StreamPool* pool_for_current_device() {
    constexpr int MAX_DEV = 64;
    thread_local StreamPool pools[MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    StreamPool& p = pools[dev];
    if (p.failed) return nullptr;
    if (!p.ready) {
        bool ok = cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < POOL_SIDE; ++i)
            ok = cudaStreamCreateWithFlags(&p.side[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { p.failed = true; cudaGetLastError(); return nullptr; }
        p.ready = true;
    }
    return &p;
}
}  // namespace

int col_launch(const ColPlan* M, bool grad, int skip_if_unit, cudaStream_t st) {
    StreamPool* pool = (M->n > 1 && !getenv("USL_COL_SERIAL"))
                           ? pool_for_current_device() : nullptr;
    if (!pool) {
        for (int i = 0; i < M->n; ++i) {
            const int rc = col_launch_scale(M, i, grad, skip_if_unit, st);
            if (rc != USL_OK) return rc;
        }
        return USL_OK;
    }
    // fork: what precedes on `st` (the scatter kernel) is ordered before every
    // scale; scale 0 stays on `st` and is launched first
    if (cudaEventRecord(pool->fork, st) != cudaSuccess) return USL_ERR_CUDA;
    int rc = col_launch_scale(M, 0, grad, skip_if_unit, st);
    for (int i = 1; i < M->n && rc == USL_OK; ++i) {
        cudaStream_t s = pool->side[i - 1];
        if (cudaStreamWaitEvent(s, pool->fork, 0) != cudaSuccess) { rc = USL_ERR_CUDA; break; }
        rc = col_launch_scale(M, i, grad, skip_if_unit, s);
        // join even after a failed launch so that `st` stays well ordered
        if (cudaEventRecord(pool->join[i - 1], s) != cudaSuccess ||
            cudaStreamWaitEvent(st, pool->join[i - 1], 0) != cudaSuccess)
            rc = rc == USL_OK ? USL_ERR_CUDA : rc;
    }
    return rc;
}

}  // namespace usl
