// TEST INFRASTRUCTURE ONLY -- CPU emulation of the column-marching loss
// kernels (csrc/col_core.cuh): the product's host+device phase functions
// compiled with g++ and run the way col_kernels.cu runs them -- same units,
// same rings, same step order -- with one CState per emulated thread and the
// threads of a phase executed one after the other.  Never linked into
// libusl.so and never used by the product.
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

template <bool GRAD, bool TILED, bool MASKED, int PAR>
static void step(const LossParams& P, const CGeo& G, const CRings& S, int r,
                 int r1, int nt, std::vector<CState>& T) {
    for (int t = 0; t < nt; ++t) {
        if (r + 1 <= r1)
            c_load_row<MASKED>(P, T[t], r + 1, T[t].xn, T[t].dn, T[t].un);
        c_p1<SROW, GRAD, MASKED>(P, G, r, T[t]);
    }
    for (int t = 0; t < nt; ++t) c_p2<SROW, GRAD, MASKED, PAR>(P, G, S, r, T[t]);
    for (int t = 0; t < nt; ++t) {
        c_p3<SROW, GRAD, MASKED, PAR>(P, G, S, r, T[t]);
        if (r + 1 <= r1) c_pV<TILED, MASKED>(P, G, S, r + 1, t, nt, T[t]);
        c_advance(T[t]);
    }
}

template <bool GRAD, bool TILED, bool MASKED>
static void unit(const LossParams& P, const CGeo& G, int nt, double* sums) {
    std::vector<float> arena(c_floats(SROW, P.w, G.nv, P.R, GRAD) + 4);
    for (auto& v : arena) v = NAN;   // nothing may be read before it is written
    const CRings S = c_carve(arena.data(), SROW, P.w, G.nv, P.R, GRAD);
    std::vector<CState> T(nt);
    for (int t = 0; t < nt; ++t) c_thread_init<GRAD>(P, G, S, t, T[t]);
    for (int t = 0; t < nt; ++t) c_init_unit<SROW, GRAD>(P, G, S, t, nt);
    const int r0 = c_first_row(G), r1 = c_last_row(G);
    for (int t = 0; t < nt; ++t)
        c_load_row<MASKED>(P, T[t], r0, T[t].x, T[t].d, T[t].u);
    for (int t = 0; t < nt; ++t) c_pV<TILED, MASKED>(P, G, S, r0, t, nt, T[t]);
    for (int r = r0; r <= r1; r += 2) {
        step<GRAD, TILED, MASKED, 0>(P, G, S, r, r1, nt, T);
        step<GRAD, TILED, MASKED, 1>(P, G, S, r + 1, r1, nt, T);
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
    if (nv * (LW + 4) > SROW) return -2;
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
                    if (tiled) unit<GRAD, true, true>(P, G, nt, sums);
                    else if (n & 31) unit<GRAD, false, true>(P, G, nt, sums);
                    else unit<GRAD, false, false>(P, G, nt, sums);
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
