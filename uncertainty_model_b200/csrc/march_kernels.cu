// Kernels (1) and (2), hot form: the register-marching fused loss (see
// march_core.cuh for the algorithm).  This file owns the CTA scheduling, the
// shared-memory arenas and the reductions.
//
// All pyramid scales of a step go out in ONE launch: blockIdx.x indexes the
// concatenation of the per-scale unit lists (largest scale first).  A unit is
// one (sample, row strip, column tile); a CTA of the small scales carries
// several units side by side (`nsub` sub-blocks marching in lock step) so that
// every CTA of the launch has a similar amount of work and thread count.
#include <stdlib.h>

#include "march_core.cuh"
#include "march_launch.cuh"

namespace usl {

using namespace mk;

constexpr int MARCH_MAX_THREADS = 512;

struct MarchDev {
    LossParams P[USL_MAX_SCALES];
    int cta_start[USL_MAX_SCALES + 1];
    int tiles_x[USL_MAX_SCALES], strips[USL_MAX_SCALES];
    int nsub[USL_MAX_SCALES], units[USL_MAX_SCALES];
    int arena[USL_MAX_SCALES];           // floats per sub-block
    int n, skip_if_unit;
};

template <bool GRAD, int PAR>
__device__ __forceinline__ void march_step(const LossParams& P, const Geo& G,
                                           const MRings& S, int r, int r_last,
                                           int ltid, int tsub, bool live,
                                           TState& T) {
    if (live) pB<GRAD, PAR>(P, G, S, r, T);
    __syncthreads();
    if (live) {
        // row r+1 replaces row r-1 (its last reader was pB above)
        if (r + 1 <= r_last) load_row(P, G, T, r + 1, T.in[PAR ^ 1]);
        pC<GRAD, PAR>(P, G, S, r, T);
    }
    __syncthreads();
    if (live) {
        pD<GRAD, PAR>(P, G, S, r, T);
        if (r + 1 <= r_last) pV(P, G, S, r + 1, ltid, tsub);
    }
    __syncthreads();
}

template <bool GRAD>
__global__ void __launch_bounds__(MARCH_MAX_THREADS, 1)
march_kernel(const __grid_constant__ MarchDev M) {
    extern __shared__ float4 smem_raw[];
    __shared__ float red[MARCH_MAX_THREADS / 32][NUM_ACC];

    int s = 0;
    while (s + 1 < M.n && (int)blockIdx.x >= M.cta_start[s + 1]) ++s;
    const LossParams& P = M.P[s];
    float gd_up = 1.0f, ge_up = 1.0f;
    if (GRAD) {
        if (P.gout_d) gd_up = __ldg(P.gout_d);
        if (P.gout_e) ge_up = __ldg(P.gout_e);
        if (M.skip_if_unit && gd_up == 1.0f && ge_up == 1.0f) return;
    }
    const int np = P.LW >> 1;
    const int tsub = (2 * np + 31) & ~31;
    const int tid = threadIdx.x;
    const int sub = tid / tsub, ltid = tid - sub * tsub;
    const int unit = (blockIdx.x - M.cta_start[s]) * M.nsub[s] + sub;
    const bool live = sub < M.nsub[s] && unit < M.units[s];

    Geo G;
    {
        int u = live ? unit : 0;
        const int tx = u % M.tiles_x[s]; u /= M.tiles_x[s];
        const int st = u % M.strips[s];
        G.b = u / M.strips[s];
        G.xa = tx * P.TW; G.xb = min(P.w, G.xa + P.TW);
        G.ya = st * P.R; G.yb = min(P.h, G.ya + P.R);
        G.cbeg = G.xa > 0 ? G.xa - 2 : 0;
        const int cend = min(G.xb + 2, P.w);
        G.npairs = (cend - G.cbeg + 1) >> 1;
        G.np = np;
        G.qlo = max(G.ya - 2, 0);
        G.sH = ac_scale(P.h - 2, P.h); G.sW = ac_scale(P.w - 2, P.w);
        G.gd_up = gd_up; G.ge_up = ge_up;
    }
    const MRings S = march_carve(
        P, reinterpret_cast<float*>(smem_raw) + (size_t)(live ? sub : 0) * M.arena[s], GRAD);
    TState T;
    thread_init<GRAD>(P, G, S, live ? ltid : 2 * np, T);
    const int r0 = first_row(G), r1 = last_row(G);
    if (live) {
        if (GRAD) cta_init_tables(P, G, S, ltid, tsub);
        load_row(P, G, T, r0, T.in[0]);
        pV(P, G, S, r0, ltid, tsub);
    }
    __syncthreads();
    // r0 = ya - 2 is even (strip heights are even): PAR is the row parity
    for (int r = r0; r <= r1; r += 2) {
        march_step<GRAD, 0>(P, G, S, r, r1, ltid, tsub, live, T);
        march_step<GRAD, 1>(P, G, S, r + 1, r1, ltid, tsub, live, T);
    }
    if (P.partials) {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int k = 0; k < NUM_ACC; ++k) {
            const float v = warp_sum(T.acc[k]);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (live && ltid < NUM_ACC) {
            float t = 0.0f;
            const int w0 = (sub * tsub) >> 5, nw = tsub >> 5;
            for (int i = 0; i < nw; ++i) t += red[w0 + i][ltid];
            P.partials[(long long)unit * NUM_ACC + ltid] = t;
        }
    }
}

// ------------------------------------------------------------------ host ---
static int env_int2(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return x > 0 ? x : dflt;
}

bool march_eligible(const UslLossConfig* cfgs, const UslLossScale* scales, int n) {
    if (getenv("USL_NO_MARCH")) return false;
    for (int i = 0; i < n; ++i) {
        const UslLossScale& s = scales[i];
        const unsigned t = cfgs[i].terms;
        if (!(t & TERM_REPROJ) || s.recon_in || s.err_in) return false;
        if (!s.images || !s.disp) return false;
        if ((t & (TERM_UNC | TERM_SMOOTH_U | TERM_CONS_U)) && !s.unc) return false;
        if ((s.w & 1) || s.w < 4 || s.h < 4) return false;
        // 64-bit loads: every plane must start on an 8-byte boundary
        if ((s.img_bs | s.img_cs | s.disp_bs | s.disp_cs | s.unc_bs | s.unc_cs |
             s.gd_bs | s.gd_cs | s.gu_bs | s.gu_cs) & 1) return false;
        const uintptr_t a = (uintptr_t)s.images | (uintptr_t)s.disp |
                            (uintptr_t)s.unc | (uintptr_t)s.grad_disp |
                            (uintptr_t)s.grad_unc;
        if (a & 7) return false;
    }
    return true;
}

int march_plan(MarchPlan* M, bool grad) {
    const int maxTW = env_int2("USL_MARCH_TW", 256) & ~1;
    const int wantR = env_int2("USL_MARCH_R", 16);
    int threads = 32;
    // pass 1: tiling of every scale, widest sub-block decides the block size
    for (int i = 0; i < M->n; ++i) {
        LossParams& p = M->P[i];
        int tiles = 1, TW = p.w, np = p.w / 2;
        if (p.w > maxTW) {
            tiles = (p.w + maxTW - 1) / maxTW;
            TW = (((p.w + tiles - 1) / tiles) + 1) & ~1;
            tiles = (p.w + TW - 1) / TW;
            np = TW / 2 + 2;
        }
        int strips = (p.h + wantR - 1) / wantR;
        int R = (((p.h + strips - 1) / strips) + 1) & ~1;
        strips = (p.h + R - 1) / R;
        p.TW = TW; p.R = R; p.LW = 2 * np;
        M->tiles_x[i] = tiles; M->strips[i] = strips;
        M->units[i] = tiles * strips * p.B;
        const int tsub = (2 * np + 31) & ~31;
        if (tsub > MARCH_MAX_THREADS) return USL_ERR_UNSUPPORTED;
        if (tsub > threads) threads = tsub;
    }
    M->threads = threads;
    M->smem = 0;
    M->cta_start[0] = 0;
    for (int i = 0; i < M->n; ++i) {
        const LossParams& p = M->P[i];
        const int tsub = (p.LW + 31) & ~31;
        const size_t arena = (mk::march_floats(p, grad) + 3) & ~(size_t)3;
        int nsub = threads / tsub;
        // shared memory: keep two CTAs per SM resident when possible
        while (nsub > 1 && nsub * arena * sizeof(float) > 110 * 1024) --nsub;
        if (arena * sizeof(float) > 227 * 1024) return USL_ERR_UNSUPPORTED;
        M->nsub[i] = nsub;
        const size_t bytes = nsub * arena * sizeof(float);
        if (bytes > M->smem) M->smem = bytes;
        M->cta_start[i + 1] = M->cta_start[i] + (M->units[i] + nsub - 1) / nsub;
    }
    return USL_OK;
}

int march_launch(const MarchPlan* M, bool grad, int skip_if_unit, cudaStream_t st) {
    MarchDev D;
    D.n = M->n; D.skip_if_unit = skip_if_unit;
    for (int i = 0; i < M->n; ++i) {
        D.P[i] = M->P[i];
        D.tiles_x[i] = M->tiles_x[i]; D.strips[i] = M->strips[i];
        D.nsub[i] = M->nsub[i]; D.units[i] = M->units[i];
        D.arena[i] = (int)((mk::march_floats(M->P[i], grad) + 3) & ~(size_t)3);
    }
    for (int i = 0; i <= M->n; ++i) D.cta_start[i] = M->cta_start[i];
    if (grad) {
        if (cudaFuncSetAttribute(march_kernel<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)M->smem) != cudaSuccess)
            return USL_ERR_CUDA;
        march_kernel<true><<<M->cta_start[M->n], M->threads, M->smem, st>>>(D);
    } else {
        if (cudaFuncSetAttribute(march_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)M->smem) != cudaSuccess)
            return USL_ERR_CUDA;
        march_kernel<false><<<M->cta_start[M->n], M->threads, M->smem, st>>>(D);
    }
    return check_launch();
}

}  // namespace usl
