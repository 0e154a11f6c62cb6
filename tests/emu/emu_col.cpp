// TEST INFRASTRUCTURE ONLY -- CPU emulation of the column-marching loss
// kernels (csrc/col_core.cuh): the product's host+device phase functions
// compiled with g++ and run the way col_kernels.cu runs them -- same units,
// same rings, same step order -- with one CState per emulated thread and the
// threads of a phase executed one after the other.  Never linked into
// libusl.so and never used by the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../uncertainty_model_b200/csrc/col_core.cuh"
#include "../../include/usl.h"

using namespace usl;
using namespace usl::ck;

constexpr int SROW = 528;

static void to_params(const UslLossConfig* cfg, const UslLossScale* s,
                      LossParams* P) {
    LossParams p = {};
    p.B = s->B; p.h = s->h; p.w = s->w;
    p.img = s->images; p.img_bs = s->img_bs; p.img_cs = s->img_cs;
    p.disp = s->disp; p.d_bs = s->disp_bs; p.d_cs = s->disp_cs;
    p.unc = s->unc; p.u_bs = s->unc_bs; p.u_cs = s->unc_cs;
    p.recon_out = s->recon_out; p.err_out = s->err_out;
    p.grad_recon_in = s->grad_recon_in;
    p.grad_disp = s->grad_disp; p.gd_bs = s->gd_bs; p.gd_cs = s->gd_cs;
    p.grad_unc = s->grad_unc; p.gu_bs = s->gu_bs; p.gu_cs = s->gu_cs;
    p.terms = cfg->terms; p.loss_type = cfg->loss_type;
    p.alpha = cfg->alpha; p.c1 = cfg->c1; p.c2 = cfg->c2;
    for (int k = 0; k < NUM_ACC; ++k) p.coef[k] = cfg->coef[k];
    *P = p;
}

template <class C, int PAR>
static void step(const LossParams& P, const CGeo& G, const CRings& S, int r,
                 int r1, int ring_last, int nt, std::vector<CState>& T) {
    // barrier | P2(r), V(r+1) | barrier | P3(r), ring request, P1(r+1)
    for (int t = 0; t < nt; ++t) c_p2<C, PAR>(P, G, S, r, T[t]);
    if (C::STEADY || r + 1 <= r1)
        for (int t = 0; t < nt; ++t)
            c_pV<C::SROW, C::MODE, C::STEADY>(P, G, S, r, t, nt, T[t]);
    for (int t = 0; t < nt; ++t) c_p3<C, PAR>(P, G, S, r, T[t]);
    if (C::STEADY || r + 1 <= r1) {
        if (C::MODE == MODE_PLAIN) {
            const int row = S.RW[c_step_index(G, r + 1)].issue;
            if (C::STEADY || row >= 0)
                for (int l = 0; l < 32; ++l) c_ring_issue(P, G, S, C::SROW, row, l);
        } else if (r + 3 <= ring_last) {
            for (int t = 0; t < nt; ++t) c_ring_fill<C::SROW, C::MODE>(P, G, T[t], r + 3);
        }
        for (int t = 0; t < nt; ++t) c_p1<C>(P, G, S, r + 1, T[t]);
    }
}

template <bool GRAD, int MODE>
static void unit(const LossParams& P, const CGeo& G, int nt, int terms_ct,
                 double* sums) {
    std::vector<float> arena(c_floats(SROW, P.w, G.nv, P.R, GRAD) + 8);
    for (auto& v : arena) v = NAN;   // nothing may be read before it is written
    // 16-byte aligned like the device arena
    float* base = arena.data();
    while ((uintptr_t)base & 15) ++base;
    const CRings S = c_carve(base, SROW, P.w, G.nv, P.R, GRAD);
    std::vector<CState> T(nt);
    for (int t = 0; t < nt; ++t) c_thread_init<SROW, GRAD>(P, G, S, t, T[t]);
    for (int t = 0; t < nt; ++t) c_init_unit<SROW, GRAD, MODE>(P, G, S, t, nt);
    const int r0 = c_first_row(G), r1 = c_last_row(G);
    const int ring_last = c_last_ring_row(P, G);
    for (int row = r0 - 1; row <= r0 + 2 && row <= ring_last; ++row) {
        if (MODE == MODE_PLAIN)
            for (int l = 0; l < 32; ++l) c_ring_issue(P, G, S, SROW, row, l);
        else for (int t = 0; t < nt; ++t) c_ring_fill<SROW, MODE>(P, G, T[t], row);
    }
    for (int t = 0; t < nt; ++t) c_pV<SROW, MODE, false>(P, G, S, r0 - 1, t, nt, T[t]);
    for (int t = 0; t < nt; ++t) c_p1<Cfg<SROW, GRAD, MODE, -1, false>>(P, G, S, r0, T[t]);
    const int s_lo = G.ya + 2;
    const int s_hi = ((G.yb - 1 < P.h - 3) ? G.yb - 1 : P.h - 3) - 1;
    using CG = Cfg<SROW, GRAD, MODE, -1, false>;
    using CS = Cfg<SROW, GRAD, MODE, -1, true>;
    using HG = Cfg<SROW, GRAD, MODE, 47, false>;
    using HS = Cfg<SROW, GRAD, MODE, 47, true>;
    for (int r = r0; r <= r1; r += 2) {
        const bool steady = r >= s_lo && r + 1 <= s_hi;
        if (terms_ct >= 0) {
            if (steady) { step<HS, 0>(P, G, S, r, r1, ring_last, nt, T); step<HS, 1>(P, G, S, r + 1, r1, ring_last, nt, T); }
            else { step<HG, 0>(P, G, S, r, r1, ring_last, nt, T); step<HG, 1>(P, G, S, r + 1, r1, ring_last, nt, T); }
        } else {
            if (steady) { step<CS, 0>(P, G, S, r, r1, ring_last, nt, T); step<CS, 1>(P, G, S, r + 1, r1, ring_last, nt, T); }
            else { step<CG, 0>(P, G, S, r, r1, ring_last, nt, T); step<CG, 1>(P, G, S, r + 1, r1, ring_last, nt, T); }
        }
    }
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < NUM_ACC; ++k) sums[k] += T[t].acc[k];
}

// maxT: widest unit in threads; wantR: strip height (even).
template <bool GRAD>
static int run(const UslLossConfig* cfg, const UslLossScale* s, int maxT,
               int wantR, int accumulate, const float* gout, double* sums) {
    LossParams P;
    to_params(cfg, s, &P);
    if (wantR & 1) return -1;
    int nv, tiles = 1, TW = P.w, LW = P.w;
    if (2 * P.w <= maxT) nv = 2;
    else if (P.w <= maxT) nv = 1;
    else {
        nv = 1;
        tiles = (P.w + (maxT - 4) - 1) / (maxT - 4);
        TW = (P.w + tiles - 1) / tiles;
        tiles = (P.w + TW - 1) / TW;
        LW = TW + 4;
    }
    if (nv * (LW + 2 * SEG_PAD) > SROW) return -2;
    int strips = (P.h + wantR - 1) / wantR;
    int R = (((P.h + strips - 1) / strips) + 1) & ~1;
    strips = (P.h + R - 1) / R;
    P.TW = TW; P.R = R; P.LW = LW;
    P.grad_disp_accumulate = accumulate;
    for (int k = 0; k < NUM_ACC; ++k) sums[k] = 0.0;
    for (int b = 0; b < P.B; ++b)
        for (int st = 0; st < strips; ++st)
            for (int vs = 0; vs < 2 / nv; ++vs)
                for (int tx = 0; tx < tiles; ++tx) {
                    CGeo G;
                    const bool tiled = tiles > 1;
                    G.b = b; G.nv = nv; G.v0 = vs * nv;
                    G.xa = tx * TW; G.xb = G.xa + TW < P.w ? G.xa + TW : P.w;
                    G.cbeg = tiled ? (G.xa - 2 > 0 ? G.xa - 2 : 0) : 0;
                    G.LW = (tiled ? (G.xb + 2 < P.w ? G.xb + 2 : P.w) : P.w) - G.cbeg;
                    G.ya = st * R; G.yb = G.ya + R < P.h ? G.ya + R : P.h;
                    G.qlo = G.ya - 2 > 0 ? G.ya - 2 : 0;
                    G.sH = ac_scale(P.h - 2, P.h); G.sW = ac_scale(P.w - 2, P.w);
                    G.gd_up = gout ? gout[0] : 1.f; G.ge_up = gout ? gout[1] : 1.f;
                    const int n = nv * G.LW;
                    const int nt = (n + 31) & ~31;
                    // bulk row copies need 16-byte aligned rows (see col_plan)
                    const bool al = (P.w % 4 == 0) &&
                        ((((uintptr_t)P.img | (uintptr_t)P.disp | (uintptr_t)P.unc) & 15) == 0) &&
                        (((P.img_bs | P.img_cs | P.d_bs | P.d_cs | P.u_bs | P.u_cs) & 3) == 0);
                    const int hot = ((int)P.terms == 47 && !tiled && !(n & 31) && al && !P.grad_recon_in &&
                                     !P.err_out && !P.recon_out) ? 47 : -1;
                    if (tiled) unit<GRAD, MODE_TILED>(P, G, nt, -1, sums);
                    else if ((n & 31) || !al) unit<GRAD, MODE_MASKED>(P, G, nt, -1, sums);
                    else unit<GRAD, MODE_PLAIN>(P, G, nt, hot, sums);
                }
    return 0;
}

extern "C" int emu_col_fwd(const UslLossConfig* cfg, const UslLossScale* s,
                           int maxT, int R, double* sums) {
    return run<false>(cfg, s, maxT, R, 0, nullptr, sums);
}

extern "C" int emu_col_grad(const UslLossConfig* cfg, const UslLossScale* s,
                            int maxT, int R, int accumulate, const float* gout,
                            double* sums) {
    return run<true>(cfg, s, maxT, R, accumulate, gout, sums);
}
