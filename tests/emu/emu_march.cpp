// TEST INFRASTRUCTURE ONLY -- CPU emulation of the register-marching loss
// kernels (csrc/march_core.cuh): the product's host+device phase functions
// compiled with g++ and run the way march_kernels.cu runs them -- same tiles,
// same rings, same step order -- with one TState per emulated thread and the
// threads of a phase executed one after the other.  Never linked into
// libusl.so and never used by the product.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../uncertainty_model_b200/csrc/march_core.cuh"
#include "../../include/usl.h"

using namespace usl;
using namespace usl::mk;

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

template <bool GRAD, int PAR>
static void step(const LossParams& P, const Geo& G, const MRings& S, int r,
                 int r1, int nt, std::vector<TState>& T) {
    for (int t = 0; t < nt; ++t) pB<GRAD, PAR>(P, G, S, r, T[t]);
    for (int t = 0; t < nt; ++t) {
        if (r + 1 <= r1) load_row(P, G, T[t], r + 1, T[t].in[PAR ^ 1]);
        pC<GRAD, PAR>(P, G, S, r, T[t]);
    }
    for (int t = 0; t < nt; ++t) pD<GRAD, PAR>(P, G, S, r, T[t]);
    if (r + 1 <= r1)
        for (int t = 0; t < nt; ++t) pV(P, G, S, r + 1, t, nt);
}

// maxTW: widest column tile (even), wantR: strip height (even).
template <bool GRAD>
static int run(const UslLossConfig* cfg, const UslLossScale* s, int maxTW,
               int wantR, int accumulate, const float* gout, double* sums) {
    LossParams P;
    to_params(cfg, s, &P);
    if ((P.w & 1) || (maxTW & 1) || (wantR & 1)) return -1;
    int tiles = 1, TW = P.w, np = P.w / 2;
    if (P.w > maxTW) {
        tiles = (P.w + maxTW - 1) / maxTW;
        TW = (((P.w + tiles - 1) / tiles) + 1) & ~1;
        tiles = (P.w + TW - 1) / TW;
        np = TW / 2 + 2;
    }
    int strips = (P.h + wantR - 1) / wantR;
    int R = (((P.h + strips - 1) / strips) + 1) & ~1;
    strips = (P.h + R - 1) / R;
    P.TW = TW; P.R = R; P.LW = 2 * np;
    P.grad_disp_accumulate = accumulate;
    const int nt = (2 * np + 31) & ~31;
    std::vector<float> arena(march_floats(P, GRAD) + 4);
    std::vector<TState> T(nt);
    for (int k = 0; k < NUM_ACC; ++k) sums[k] = 0.0;
    for (int b = 0; b < P.B; ++b)
        for (int st = 0; st < strips; ++st)
            for (int tx = 0; tx < tiles; ++tx) {
                // poison the rings: nothing may be read before it is written
                for (auto& v : arena) v = NAN;
                Geo G;
                G.b = b;
                G.xa = tx * TW; G.xb = G.xa + TW < P.w ? G.xa + TW : P.w;
                G.ya = st * R; G.yb = G.ya + R < P.h ? G.ya + R : P.h;
                G.cbeg = G.xa > 0 ? G.xa - 2 : 0;
                const int cend = G.xb + 2 < P.w ? G.xb + 2 : P.w;
                G.npairs = (cend - G.cbeg + 1) >> 1;
                G.np = np;
                G.qlo = G.ya - 2 > 0 ? G.ya - 2 : 0;
                G.sH = ac_scale(P.h - 2, P.h); G.sW = ac_scale(P.w - 2, P.w);
                G.gd_up = gout ? gout[0] : 1.f; G.ge_up = gout ? gout[1] : 1.f;
                const MRings S = march_carve(P, arena.data(), GRAD);
                for (int t = 0; t < nt; ++t) thread_init<GRAD>(P, G, S, t, T[t]);
                const int r0 = first_row(G), r1 = last_row(G);
                if (GRAD) cta_init_tables(P, G, S, 0, 1);
                for (int t = 0; t < nt; ++t) load_row(P, G, T[t], r0, T[t].in[0]);
                for (int t = 0; t < nt; ++t) pV(P, G, S, r0, t, nt);
                for (int r = r0; r <= r1; r += 2) {
                    step<GRAD, 0>(P, G, S, r, r1, nt, T);
                    step<GRAD, 1>(P, G, S, r + 1, r1, nt, T);
                }
                for (int t = 0; t < nt; ++t)
                    for (int k = 0; k < NUM_ACC; ++k) sums[k] += T[t].acc[k];
            }
    return 0;
}

extern "C" int emu_march_fwd(const UslLossConfig* cfg, const UslLossScale* s,
                             int TW, int R, double* sums) {
    return run<false>(cfg, s, TW, R, 0, nullptr, sums);
}

extern "C" int emu_march_grad(const UslLossConfig* cfg, const UslLossScale* s,
                              int TW, int R, int accumulate, const float* gout,
                              double* sums) {
    return run<true>(cfg, s, TW, R, accumulate, gout, sums);
}
