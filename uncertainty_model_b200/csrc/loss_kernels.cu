// Kernels (1) and (2): fused multi-scale loss forward / backward launches.
// The per-phase arithmetic lives in loss_core.cuh / cons_core.cuh (shared with
// the CPU emulation harness in tests/emu); this file owns the CTA scheduling,
// the shared-memory arenas, the warp-serial scatter and the reductions.
//
// All pyramid scales of a step go out in ONE launch: blockIdx.x indexes the
// concatenation of the per-scale tile lists (largest scale first, so the long
// CTAs start first and the small scales fill the tail of the wave).
#include <stdlib.h>

#include "col_launch.cuh"
#include "cons_core.cuh"
#include "cons_launch.cuh"
#include "usl_common.cuh"

namespace usl {

struct MultiParams {
    LossParams P[USL_MAX_SCALES];
    int cta_start[USL_MAX_SCALES + 1];
    int tiles_x[USL_MAX_SCALES], strips[USL_MAX_SCALES];
    int n;
};

template <bool BWD>
__global__ void __launch_bounds__(640)
loss_main_kernel(const __grid_constant__ MultiParams M) {
    extern __shared__ float4 smem_raw[];
    __shared__ float red[20][NUM_ACC];

    int s = 0;
    while (s + 1 < M.n && (int)blockIdx.x >= M.cta_start[s + 1]) ++s;
    const LossParams& P = M.P[s];
    int local = blockIdx.x - M.cta_start[s];
    Tile T;
    const int tx = local % M.tiles_x[s]; local /= M.tiles_x[s];
    const int st = local % M.strips[s];
    T.b = local / M.strips[s];
    T.xa = tx * P.TW; T.xb = min(P.w, T.xa + P.TW);
    T.ya = st * P.R; T.yb = min(P.h, T.ya + P.R);
    T.cbeg = T.xa - HALO_L;
    T.cta = blockIdx.x;

    const Rings S = carve(P, reinterpret_cast<float*>(smem_raw), BWD);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int LWp = (P.LW + 31) & ~31;
    float acc[NUM_ACC];
#pragma unroll
    for (int k = 0; k < NUM_ACC; ++k) acc[k] = 0.0f;
    float gd_up = 0.0f, ge_up = 0.0f;
    if (BWD) {
        gd_up = P.gout_d ? __ldg(P.gout_d) : 0.0f;
        ge_up = P.gout_e ? __ldg(P.gout_e) : 0.0f;
        phase_init_bwd(P, T, S, tid, nt);
    }
    const int r1 = last_step(T);
    for (int r = first_step(T); r <= r1; ++r) {
        phase_A(P, T, S, r, tid, nt);
        __syncthreads();
        phase_B<BWD>(P, T, S, r, tid, nt, LWp, acc, gd_up, ge_up);
        __syncthreads();
        phase_C<BWD>(P, T, S, r, tid, nt, LWp, gd_up);
        __syncthreads();
        phase_D<BWD>(P, T, S, r, tid, nt, LWp, acc, gd_up, ge_up);
    }
    if (!BWD) {
        const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
        for (int k = 0; k < NUM_ACC; ++k) {
            const float v = warp_sum(acc[k]);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (tid < NUM_ACC) {
            float t = 0.0f;
            const int nw = (nt + 31) >> 5;
            for (int i = 0; i < nw; ++i) t += red[i][tid];
            P.partials[(long long)(blockIdx.x - M.cta_start[s]) * NUM_ACC + tid] = t;
        }
    }
}

// One warp scatters the sources of one row (one view, both terms) into the
// destination row `Hrow`.  Chunks of 32 columns in order; duplicates inside a
// chunk are folded by the lowest lane of each group in lane order.
__device__ __forceinline__ void scatter_row_warp(float* Hrow, const int* dest,
                                                 const float* c0,
                                                 const float* c1, int w,
                                                 int lane) {
    for (int base = 0; base < w; base += 32) {
        const int x = base + lane;
        const bool valid = x < w;
        const int d = valid ? dest[x] : (-1000000 - lane);
        const float a0 = valid ? c0[x] : 0.0f;
        const float a1 = valid ? c1[x] : 0.0f;
        unsigned grp = __match_any_sync(0xffffffffu, d);
        const bool leader = (__ffs(grp) - 1) == lane;
        float s0 = 0.0f, s1 = 0.0f;
        while (__any_sync(0xffffffffu, grp != 0u)) {
            const int src = grp ? (__ffs(grp) - 1) : lane;
            const float v0 = __shfl_sync(0xffffffffu, a0, src);
            const float v1 = __shfl_sync(0xffffffffu, a1, src);
            if (grp) { s0 += v0; s1 += v1; grp &= grp - 1; }
        }
        if (leader && valid && d >= 0 && d < w) Hrow[d] += s0;
        __syncwarp();
        if (leader && valid && d + 1 >= 0 && d + 1 < w) Hrow[d + 1] += s1;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
cons_scatter_kernel(const __grid_constant__ MultiCons M) {
    extern __shared__ float4 smem_raw[];
    int s = 0;
    while (s + 1 < M.n && (int)blockIdx.x >= M.cta_start[s + 1]) ++s;
    const ConsParams& P = M.P[s];
    const int local = blockIdx.x - M.cta_start[s];
    ConsTile T;
    T.b = local / M.strips[s];
    T.ya = (local % M.strips[s]) * P.R;
    T.yb = min(P.h, T.ya + P.R);
    const ConsRings S = cons_carve(P.w, reinterpret_cast<float*>(smem_raw));
    const int tid = threadIdx.x, nt = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31;
    const float gd_up = P.gout_d ? __ldg(P.gout_d) : P.gout_default;
    const float ge_up = P.gout_e ? __ldg(P.gout_e) : P.gout_default;
    if (M.skip_if_unit && gd_up == 1.0f && ge_up == 1.0f) return;
    const int r1 = cons_last_step(T);
    for (int r = cons_first_step(T); r <= r1; ++r) {
        cons_phase_A(P, T, S, r, tid, nt);
        __syncthreads();
        cons_phase_B(P, T, S, r, tid, nt, gd_up, ge_up);
        __syncthreads();
        if (warp < 2 && r >= 0 && r < P.h) {
            const int v = warp;      // source view; destination is the other
            float* Hrow = S.H + ((size_t)mod4(r) * 2 + (1 - v)) * P.w;
            for (int term = 0; term < 2; ++term) {
                if (!(P.terms & (term ? TERM_CONS_U : TERM_CONS_D))) continue;
                const size_t j = (size_t)(term * 2 + v) * P.w;
                scatter_row_warp(Hrow, S.dest + j, S.c0 + j, S.c1 + j, P.w, lane);
            }
        }
        __syncthreads();
        cons_phase_D(P, T, S, r, tid, nt);
    }
}

struct CtaStarts { int v[USL_MAX_SCALES + 1]; };

__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* partials, const CtaStarts starts_,
                       int n_scales, double* sums) {
    // one block per (scale, term): fixed-order fp64 sum over that scale's rows
    // (strided per-thread sums, then a fixed tree) -- deterministic
    __shared__ double part[256];
    const int i = blockIdx.x;
    const int s = i / NUM_ACC, k = i % NUM_ACC;
    double t = 0.0;
    for (int c = starts_.v[s] + threadIdx.x; c < starts_.v[s + 1]; c += 256)
        t += (double)partials[(long long)c * NUM_ACC + k];
    part[threadIdx.x] = t;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[i] = part[0];
}

__global__ void combine_kernel(const double* sums, const float* coef,
                               int n_scales, float* out_d, float* out_e) {
    if (threadIdx.x >= 2) return;
    const int o = threadIdx.x;
    double t = 0.0;
    for (int s = 0; s < n_scales; ++s)
        for (int k = 0; k < 3; ++k) {
            const int i = s * NUM_ACC + o * 3 + k;
            t += (double)coef[i] * sums[i];
        }
    if (o == 0) *out_d = (float)t; else *out_e = (float)t;
}

// ------------------------------------------------------------------ host ---
static int pos_or(int v, int dflt) { return v > 0 ? v : dflt; }

static void choose_tiling(int h, int w, bool bwd, int* TW, int* R) {
    int tw = pos_or(bwd ? knobs().bwd_tw : knobs().fwd_tw, bwd ? 128 : 256);
    int r = pos_or(bwd ? knobs().bwd_r : knobs().fwd_r, 32);
    if (tw > w) tw = w;
    // even out the column tiles
    const int nx = (w + tw - 1) / tw;
    tw = (w + nx - 1) / nx;
    if (r > h) r = h;
    const int ny = (h + r - 1) / r;
    r = (h + ny - 1) / ny;
    *TW = tw; *R = r;
}

static int fill_params(const UslLossConfig* cfg, const UslLossScale* s,
                       bool bwd, LossParams* P) {
    if (!cfg || !s || s->B <= 0) return USL_ERR_ARG;
    if (s->h < 3 || s->w < 3) return USL_ERR_UNSUPPORTED;
    if (cfg->loss_type < 0 || cfg->loss_type > 2) return USL_ERR_ARG;
    const unsigned t = cfg->terms;
    if ((t & (TERM_REPROJ | TERM_SMOOTH_D | TERM_SMOOTH_U)) && !s->images &&
        !(s->err_in && !(t & (TERM_SMOOTH_D | TERM_SMOOTH_U))))
        return USL_ERR_ARG;
    const bool need_disp = (t & (TERM_CONS_D | TERM_CONS_U | TERM_SMOOTH_D)) ||
                           ((t & TERM_REPROJ) && !s->recon_in && !s->err_in);
    if (need_disp && !s->disp) return USL_ERR_ARG;
    if ((t & (TERM_UNC | TERM_SMOOTH_U | TERM_CONS_U)) && !s->unc)
        return USL_ERR_ARG;
    if ((t & TERM_UNC) && !(t & TERM_REPROJ) && !s->err_in) return USL_ERR_ARG;
    LossParams p = {};
    p.B = s->B; p.h = s->h; p.w = s->w;
    p.img = s->images; p.img_bs = s->img_bs; p.img_cs = s->img_cs;
    p.disp = s->disp; p.d_bs = s->disp_bs; p.d_cs = s->disp_cs;
    p.unc = s->unc; p.u_bs = s->unc_bs; p.u_cs = s->unc_cs;
    p.recon_in = s->recon_in; p.ri_bs = s->rin_bs; p.ri_cs = s->rin_cs;
    p.err_in = s->err_in; p.ei_bs = s->ein_bs; p.ei_cs = s->ein_cs;
    p.recon_out = s->recon_out; p.err_out = s->err_out;
    p.grad_recon_in = s->grad_recon_in;
    p.grad_disp = s->grad_disp; p.gd_bs = s->gd_bs; p.gd_cs = s->gd_cs;
    p.grad_unc = s->grad_unc; p.gu_bs = s->gu_bs; p.gu_cs = s->gu_cs;
    p.grad_recon_out = s->grad_recon_out;
    p.scat = s->scatter_ws;
    p.terms = t; p.loss_type = cfg->loss_type;
    p.recon_given = s->recon_in != nullptr;
    p.err_given = s->err_in != nullptr;
    p.alpha = cfg->alpha; p.c1 = cfg->c1; p.c2 = cfg->c2;
    for (int k = 0; k < NUM_ACC; ++k) p.coef[k] = cfg->coef[k];
    if (bwd && p.recon_given && (t & TERM_REPROJ) && !p.err_given &&
        !p.grad_recon_out)
        return USL_ERR_ARG;
    if (p.recon_out && (p.recon_given || !(t & TERM_REPROJ))) return USL_ERR_ARG;
    choose_tiling(p.h, p.w, bwd, &p.TW, &p.R);
    p.LW = p.TW + HALO_L + HALO_R;
    *P = p;
    return USL_OK;
}

static int plan(const UslLossConfig* cfgs, const UslLossScale* scales, int n,
                bool bwd, MultiParams* M, size_t* smem, int* threads) {
    if (n < 1 || n > USL_MAX_SCALES) return USL_ERR_ARG;
    M->n = n;
    M->cta_start[0] = 0;
    *smem = 0;
    int maxLWp = 32;
    for (int i = 0; i < n; ++i) {
        const int rc = fill_params(&cfgs[i], &scales[i], bwd, &M->P[i]);
        if (rc != USL_OK) return rc;
        const LossParams& p = M->P[i];
        M->tiles_x[i] = (p.w + p.TW - 1) / p.TW;
        M->strips[i] = (p.h + p.R - 1) / p.R;
        M->cta_start[i + 1] =
            M->cta_start[i] + M->tiles_x[i] * M->strips[i] * p.B;
        const size_t bytes = ring_floats(p, bwd) * sizeof(float);
        if (bytes > *smem) *smem = bytes;
        const int LWp = (p.LW + 31) & ~31;
        if (LWp > maxLWp) maxLWp = LWp;
    }
    if (*smem > 227 * 1024) return USL_ERR_UNSUPPORTED;
    int nt = 2 * maxLWp;
    if (nt > 640) nt = maxLWp;       // two passes per step over the item list
    if (nt > 640) nt = 640;
    *threads = nt;
    return USL_OK;
}

}  // namespace usl

using namespace usl;

// Any tensor of the call names the device it runs on (stand-alone terms pass
// no images).
static const void* any_tensor(const UslLossScale* s, int n) {
    for (int i = 0; s && i < n; ++i) {
        const void* c[] = {s[i].images, s[i].disp, s[i].unc, s[i].recon_in,
                           s[i].err_in, s[i].grad_disp, s[i].grad_unc};
        for (const void* p : c)
            if (p) return p;
    }
    return nullptr;
}

static int fill_all(const UslLossConfig* cfgs, const UslLossScale* scales,
                    int n, bool bwd, LossParams* P) {
    if (n < 1 || n > USL_MAX_SCALES) return USL_ERR_ARG;
    for (int i = 0; i < n; ++i) {
        const int rc = fill_params(&cfgs[i], &scales[i], bwd, &P[i]);
        if (rc != USL_OK) return rc;
    }
    return USL_OK;
}

extern "C" int usl_loss_plan(const UslLossConfig* cfgs,
                             const UslLossScale* scales, int n_scales,
                             int mode, int* cta_starts) {
    if (!cfgs || !scales || !cta_starts) return USL_ERR_ARG;
    DeviceGuard guard(any_tensor(scales, n_scales));
    if (mode != USL_MODE_FWD && mode != USL_MODE_GRAD) return USL_ERR_ARG;
    if (col_eligible(cfgs, scales, n_scales)) {
        ColPlan M;
        M.n = n_scales;
        int rc = fill_all(cfgs, scales, n_scales, false, M.P);
        if (rc != USL_OK) return rc;
        rc = col_plan(&M, mode == USL_MODE_GRAD);
        if (rc == USL_OK) {
            // one row of partial sums per unit
            for (int i = 0; i <= n_scales; ++i) cta_starts[i] = M.row_start[i];
            return USL_OK;
        }
        if (rc != USL_ERR_UNSUPPORTED) return rc;
    }
    if (mode == USL_MODE_GRAD) return USL_ERR_UNSUPPORTED;
    MultiParams M;
    size_t smem; int nt;
    const int rc = plan(cfgs, scales, n_scales, false, &M, &smem, &nt);
    if (rc != USL_OK) return rc;
    for (int i = 0; i <= n_scales; ++i) cta_starts[i] = M.cta_start[i];
    return USL_OK;
}

extern "C" int usl_loss_fwd_ctas(const UslLossScale* s) {
    if (!s || s->B <= 0 || s->h < 3 || s->w < 3) return USL_ERR_ARG;
    int TW, R;
    choose_tiling(s->h, s->w, false, &TW, &R);
    return ((s->w + TW - 1) / TW) * ((s->h + R - 1) / R) * s->B;
}

// The column-marching path, if every scale qualifies: 1 = launched (or failed
// with *rc_out set), 0 = not eligible.
static int try_col(const UslLossConfig* cfgs, const UslLossScale* scales,
                   int n, bool grad, float* partials, const float* gout_d,
                   const float* gout_e, int accumulate, int skip_if_unit,
                   cudaStream_t st, int* rc_out, ColAfter after = nullptr,
                   void* after_ctx = nullptr) {
    if (!col_eligible(cfgs, scales, n)) return 0;
    ColPlan M;
    M.n = n;
    int rc = fill_all(cfgs, scales, n, false, M.P);
    if (rc != USL_OK) { *rc_out = rc; return 1; }
    rc = col_plan(&M, grad);
    if (rc == USL_ERR_UNSUPPORTED) return 0;
    if (rc != USL_OK) { *rc_out = rc; return 1; }
    for (int i = 0; i < n; ++i) {
        LossParams& p = M.P[i];
        p.partials = partials ? partials + (long long)M.row_start[i] * NUM_ACC : nullptr;
        p.gout_d = gout_d; p.gout_e = gout_e;
        p.grad_disp_accumulate =
            (accumulate && (p.terms & (TERM_CONS_D | TERM_CONS_U))) ? 1 : 0;
        if (grad && (!p.grad_disp || !p.grad_unc)) { *rc_out = USL_ERR_ARG; return 1; }
    }
    *rc_out = col_launch(&M, grad, skip_if_unit, st, after, after_ctx);
    return 1;
}

// Whether try_col will take this call (eligible AND plannable).
static bool col_ready(const UslLossConfig* cfgs, const UslLossScale* scales, int n,
                      bool grad) {
    if (!col_eligible(cfgs, scales, n)) return false;
    ColPlan M;
    M.n = n;
    if (fill_all(cfgs, scales, n, false, M.P) != USL_OK) return false;
    return col_plan(&M, grad) == USL_OK;
}

extern "C" int usl_loss_fwd(const UslLossConfig* cfgs,
                            const UslLossScale* scales, int n_scales,
                            float* partials, void* stream) {
    if (!partials) return USL_ERR_ARG;
    DeviceGuard guard(partials);
    int rc = USL_OK;
    if (try_col(cfgs, scales, n_scales, false, partials, nullptr, nullptr, 0,
                0, (cudaStream_t)stream, &rc))
        return rc;
    MultiParams M;
    size_t smem; int nt;
    rc = plan(cfgs, scales, n_scales, false, &M, &smem, &nt);
    if (rc != USL_OK) return rc;
    for (int i = 0; i < n_scales; ++i)
        M.P[i].partials = partials + (long long)M.cta_start[i] * NUM_ACC;
    if (cudaFuncSetAttribute(loss_main_kernel<false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return USL_ERR_CUDA;
    loss_main_kernel<false><<<M.cta_start[n_scales], nt, smem,
                              (cudaStream_t)stream>>>(M);
    return check_launch();
}

extern "C" int usl_loss_reduce(const float* partials, const int* cta_starts,
                               int n_scales, double* sums, void* stream) {
    if (!partials || !cta_starts || !sums || n_scales < 1 ||
        n_scales > USL_MAX_SCALES)
        return USL_ERR_ARG;
    DeviceGuard guard(partials);
    CtaStarts st;
    for (int i = 0; i <= n_scales; ++i) st.v[i] = cta_starts[i];
    reduce_partials_kernel<<<n_scales * NUM_ACC, 256, 0, (cudaStream_t)stream>>>(
        partials, st, n_scales, sums);
    mark(10, (cudaStream_t)stream);
    return check_launch();
}

extern "C" int usl_loss_combine(const double* sums, const float* coef,
                                int n_scales, float* out_disp, float* out_err,
                                void* stream) {
    if (!sums || !coef || !out_disp || !out_err || n_scales < 1)
        return USL_ERR_ARG;
    DeviceGuard guard(sums);
    combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, coef, n_scales,
                                                       out_disp, out_err);
    mark(11, (cudaStream_t)stream);
    return check_launch();
}

// The deterministic transposed warp of the consistency terms -> grad_disp
// (pure store).  Returns the number of scales that scatter.
static int launch_scatter(const LossParams* P, int n_scales,
                          const float* gout_disp, const float* gout_err,
                          float gout_default, int skip_if_unit, bool launch,
                          cudaStream_t st, int* rc_out, int accumulate = 0,
                          bool scat_filled = false, int only_scale = -1) {
    MultiCons C;
    C.n = 0; C.cta_start[0] = 0; C.skip_if_unit = skip_if_unit;
    size_t csmem = 0;
    bool use_scat = scat_filled && !knobs().scatter_warp_per_row;
    for (int i = 0; i < n_scales; ++i) {
        const LossParams& p = P[i];
        if (only_scale >= 0 && i != only_scale) continue;
        if (!(p.terms & (TERM_CONS_D | TERM_CONS_U))) continue;
        if (!p.grad_disp) { *rc_out = USL_ERR_ARG; return 0; }
        ConsParams c = {};
        c.B = p.B; c.h = p.h; c.w = p.w;
        c.disp = p.disp; c.d_bs = p.d_bs; c.d_cs = p.d_cs;
        c.unc = p.unc; c.u_bs = p.u_bs; c.u_cs = p.u_cs;
        c.gout_d = gout_disp; c.gout_e = gout_err;
        c.gout_default = gout_default;
        c.grad_disp = p.grad_disp; c.gd_bs = p.gd_bs; c.gd_cs = p.gd_cs;
        c.accumulate = accumulate;
        c.terms = p.terms & (TERM_CONS_D | TERM_CONS_U);
        c.coef_dd = p.coef[ACC_CONS_D]; c.coef_ud = p.coef[ACC_CONS_U];
        c.scat = p.scat;
        use_scat = use_scat && p.scat != nullptr;
        c.R = pos_or(knobs().cons_r, 16);
        if (c.R > c.h) c.R = c.h;
        const int k = C.n++;
        C.P[k] = c;
        C.strips[k] = (c.h + c.R - 1) / c.R;
        C.cta_start[k + 1] = C.cta_start[k] + C.strips[k] * c.B;
        const size_t bytes = cons_ring_floats(c.w) * sizeof(float);
        if (bytes > csmem) csmem = bytes;
    }
    *rc_out = USL_OK;
    if (C.n > 0 && launch && use_scat) {
        // (the column kernels of this call left the per-pixel taps behind)
        const int rc3 = cons_rows_launch(&C, st);
        if (rc3 != USL_ERR_UNSUPPORTED) { *rc_out = rc3; return C.n; }
    }
    if (C.n > 0 && launch && !knobs().scatter_v1) {
        const int rc2 = cons_scatter2_launch(&C, st);
        if (rc2 != USL_ERR_UNSUPPORTED) { *rc_out = rc2; return C.n; }
        // (rows too wide for the warp-per-row arena: the strip kernel below)
        for (int k = 0; k < C.n; ++k) {
            C.P[k].R = pos_or(knobs().cons_r, 16);
            if (C.P[k].R > C.P[k].h) C.P[k].R = C.P[k].h;
            C.strips[k] = (C.P[k].h + C.P[k].R - 1) / C.P[k].R;
            C.cta_start[k + 1] = C.cta_start[k] + C.strips[k] * C.P[k].B;
        }
    }
    if (C.n > 0 && launch) {
        if (csmem > 227 * 1024) { *rc_out = USL_ERR_UNSUPPORTED; return C.n; }
        if (cudaFuncSetAttribute(cons_scatter_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)csmem) != cudaSuccess) {
            *rc_out = USL_ERR_CUDA; return C.n;
        }
        cons_scatter_kernel<<<C.cta_start[C.n], 256, csmem, st>>>(C);
        *rc_out = check_launch();
    }
    return C.n;
}

struct AfterScale {
    const LossParams* P; int n;
    const float* gout_disp; const float* gout_err;
    int skip;
    int defer0;        // leave the largest scale's scatter to a later call
    cudaStream_t reduce_stream;   // if set: made to wait for every column kernel
};
static int scatter_after_scale(void* ctx_, int scale, cudaStream_t st) {
    const AfterScale* c = static_cast<const AfterScale*>(ctx_);
    int rc = USL_OK;
    if (c->reduce_stream) {
        // the partial sums of this scale are complete here, before its scatter
        StreamPool* pool = stream_pool();
        if (!pool || cudaEventRecord(pool->col_done[scale], st) != cudaSuccess ||
            cudaStreamWaitEvent(c->reduce_stream, pool->col_done[scale], 0) != cudaSuccess)
            return USL_ERR_CUDA;
    }
    if (scale == 0 && c->defer0) return rc;
    launch_scatter(c->P, c->n, c->gout_disp, c->gout_err, 1.0f, c->skip, true, st, &rc,
                   1, true, scale);
    return rc;
}

extern "C" int usl_loss_grad(const UslLossConfig* cfgs,
                             const UslLossScale* scales, int n_scales,
                             const float* gout_disp, const float* gout_err,
                             float* partials, int flags, void* stream) {
    if (!cfgs || !scales || n_scales < 1) return USL_ERR_ARG;
    DeviceGuard guard(any_tensor(scales, n_scales));
    if (!col_ready(cfgs, scales, n_scales, true)) return USL_ERR_UNSUPPORTED;
    LossParams P[USL_MAX_SCALES];
    int rc = fill_all(cfgs, scales, n_scales, false, P);
    if (rc != USL_OK) return rc;
    const int skip = (flags & USL_GRAD_SKIP_IF_UNIT) ? 1 : 0;
    // column kernels: they store the gradient, then the scatter adds its part
    // (the read-modify-write sits in the kernel that has warps to spare)
    if (flags & USL_GRAD_ONLY_SCATTER0) {
        launch_scatter(P, n_scales, gout_disp, gout_err, 1.0f, skip, true,
                       (cudaStream_t)stream, &rc, 1, true, 0);
        return rc;
    }
    if (!(flags & (USL_GRAD_ONLY_SCATTER | USL_GRAD_NO_SCATTER)) && !knobs().scatter_after_all) {
        // both halves: the scatter of every scale right behind its own fused
        // kernel, on that scale's stream -- the small scales' finish in the
        // shadow of the largest scale's fused kernel
        bool rows_ok = true;
        for (int i = 0; i < n_scales; ++i)
            rows_ok = rows_ok && (!(P[i].terms & (TERM_CONS_D | TERM_CONS_U)) || P[i].scat);
        if (rows_ok) {
            AfterScale ctx = {P, n_scales, gout_disp, gout_err, skip,
                              (flags & USL_GRAD_DEFER_SCATTER0) ? 1 : 0, nullptr};
            if (!try_col(cfgs, scales, n_scales, true, partials, gout_disp,
                         gout_err, 0, skip, (cudaStream_t)stream, &rc,
                         scatter_after_scale, &ctx))
                return USL_ERR_UNSUPPORTED;
            return rc;
        }
    }
    if (!(flags & USL_GRAD_ONLY_SCATTER)) {
        if (!try_col(cfgs, scales, n_scales, true, partials, gout_disp,
                     gout_err, 0, skip, (cudaStream_t)stream, &rc))
            return USL_ERR_UNSUPPORTED;
        if (rc != USL_OK) return rc;
    }
    if (flags & USL_GRAD_DEFER_SCATTER0) {
        for (int i = 1; i < n_scales && rc == USL_OK; ++i)
            launch_scatter(P, n_scales, gout_disp, gout_err, 1.0f, skip, true,
                           (cudaStream_t)stream, &rc, 1, true, i);
        return rc;
    }
    if (!(flags & USL_GRAD_NO_SCATTER))
        launch_scatter(P, n_scales, gout_disp, gout_err, 1.0f, skip, true,
                       (cudaStream_t)stream, &rc, 1, true);
    return rc;
}

extern "C" int usl_loss_grad_sharded(const UslLossConfig* cfgs,
                                     const UslLossScale* scales, int n_scales,
                                     float* partials, const int* cta_starts,
                                     double* sums, void* reduce_stream,
                                     void* stream) {
    if (!cfgs || !scales || n_scales < 1 || n_scales > USL_MAX_SCALES ||
        !partials || !cta_starts || !sums || !reduce_stream || reduce_stream == stream)
        return USL_ERR_ARG;
    DeviceGuard guard(any_tensor(scales, n_scales));
    if (!col_ready(cfgs, scales, n_scales, true)) return USL_ERR_UNSUPPORTED;
    LossParams P[USL_MAX_SCALES];
    int rc = fill_all(cfgs, scales, n_scales, false, P);
    if (rc != USL_OK) return rc;
    for (int i = 0; i < n_scales; ++i)
        if ((P[i].terms & (TERM_CONS_D | TERM_CONS_U)) && !P[i].scat)
            return USL_ERR_UNSUPPORTED;
    if (knobs().scatter_after_all) return USL_ERR_UNSUPPORTED;
    AfterScale ctx = {P, n_scales, nullptr, nullptr, 0, 0, (cudaStream_t)reduce_stream};
    if (!try_col(cfgs, scales, n_scales, true, partials, nullptr, nullptr, 0, 0,
                 (cudaStream_t)stream, &rc, scatter_after_scale, &ctx))
        return USL_ERR_UNSUPPORTED;
    if (rc != USL_OK) return rc;
    // (reduce_stream now waits for the four column kernels, not for the scatters)
    CtaStarts st;
    for (int i = 0; i <= n_scales; ++i) st.v[i] = cta_starts[i];
    reduce_partials_kernel<<<n_scales * NUM_ACC, 256, 0, (cudaStream_t)reduce_stream>>>(
        partials, st, n_scales, sums);
    mark(10, (cudaStream_t)reduce_stream);
    return check_launch();
}

extern "C" int usl_loss_bwd(const UslLossConfig* cfgs,
                            const UslLossScale* scales, int n_scales,
                            const float* gout_disp, const float* gout_err,
                            int stages, void* stream) {
    if (!(stages & (USL_BWD_STAGE_SCATTER | USL_BWD_STAGE_MAIN)))
        return USL_ERR_ARG;
    DeviceGuard guard(any_tensor(scales, n_scales));
    int rc = USL_OK;
    if (gout_disp && gout_err && col_ready(cfgs, scales, n_scales, true)) {
        LossParams P[USL_MAX_SCALES];
        rc = fill_all(cfgs, scales, n_scales, false, P);
        if (rc != USL_OK) return rc;
        // column kernels store, then the scatter adds its part
        if (stages & USL_BWD_STAGE_MAIN) {
            if (!try_col(cfgs, scales, n_scales, true, nullptr, gout_disp,
                         gout_err, 0, 0, (cudaStream_t)stream, &rc))
                return USL_ERR_UNSUPPORTED;
            if (rc != USL_OK) return rc;
        }
        launch_scatter(P, n_scales, gout_disp, gout_err, 0.0f, 0,
                       (stages & USL_BWD_STAGE_SCATTER) != 0,
                       (cudaStream_t)stream, &rc,
                       (stages & USL_BWD_STAGE_MAIN) ? 1 : 0,
                       (stages & USL_BWD_STAGE_MAIN) != 0);
        return rc;
    }
    // (a NULL upstream gradient means "this output takes no part": the general
    //  kernels below treat it as zero)
    MultiParams M;
    size_t smem; int nt;
    rc = plan(cfgs, scales, n_scales, true, &M, &smem, &nt);
    if (rc != USL_OK) return rc;
    for (int i = 0; i < n_scales; ++i) {
        LossParams& p = M.P[i];
        p.gout_d = gout_disp; p.gout_e = gout_err;
        p.grad_disp_accumulate =
            (p.terms & (TERM_CONS_D | TERM_CONS_U)) ? 1 : 0;
    }
    if (stages & USL_BWD_STAGE_SCATTER) {
        launch_scatter(M.P, n_scales, gout_disp, gout_err, 0.0f, 0, true,
                       (cudaStream_t)stream, &rc);
        if (rc != USL_OK) return rc;
    }
    // 2) everything else, adding to the scattered part
    if (!(stages & USL_BWD_STAGE_MAIN)) return USL_OK;
    if (cudaFuncSetAttribute(loss_main_kernel<true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return USL_ERR_CUDA;
    loss_main_kernel<true><<<M.cta_start[n_scales], nt, smem,
                             (cudaStream_t)stream>>>(M);
    return check_launch();
}
