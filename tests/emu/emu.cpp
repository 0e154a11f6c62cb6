// TEST INFRASTRUCTURE ONLY -- CPU emulation of the fused loss kernels.
//
// Compiles the product's host+device phase functions (csrc/loss_core.cuh,
// csrc/cons_core.cuh) with g++ and runs them the way the CUDA kernels do --
// same tiles, same rings, same step order -- with one sequential "thread".
// tests/test_emu.py compares the result with the oracle, so the kernel logic
// is checked on the CPU before any GPU time is spent.  Never linked into
// libusl.so and never used by the product.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../uncertainty_model_b200/csrc/cons_core.cuh"
#include "../../include/usl.h"

using namespace usl;

static void to_params(const UslLossConfig* cfg, const UslLossScale* s, int TW,
                      int R, LossParams* P) {
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
    p.terms = cfg->terms; p.loss_type = cfg->loss_type;
    p.recon_given = s->recon_in != nullptr;
    p.err_given = s->err_in != nullptr;
    p.alpha = cfg->alpha; p.c1 = cfg->c1; p.c2 = cfg->c2;
    for (int k = 0; k < NUM_ACC; ++k) p.coef[k] = cfg->coef[k];
    if (TW > p.w) TW = p.w;
    if (R > p.h) R = p.h;
    p.TW = TW; p.R = R; p.LW = TW + HALO_L + HALO_R;
    *P = p;
}

template <bool BWD>
static void run_main(const LossParams& P, const float* gout, double* sums) {
    std::vector<float> arena(ring_floats(P, BWD));
    const int LWp = (P.LW + 31) & ~31;
    const float gd = BWD ? gout[0] : 0.f, ge = BWD ? gout[1] : 0.f;
    for (int b = 0; b < P.B; ++b)
        for (int ya = 0; ya < P.h; ya += P.R)
            for (int xa = 0; xa < P.w; xa += P.TW) {
                // poison the rings: nothing may be read before it is written
                for (auto& v : arena) v = NAN;
                Tile T;
                T.b = b; T.xa = xa; T.xb = xa + P.TW < P.w ? xa + P.TW : P.w;
                T.ya = ya; T.yb = ya + P.R < P.h ? ya + P.R : P.h;
                T.cbeg = xa - HALO_L; T.cta = 0;
                const Rings S = carve(P, arena.data(), BWD);
                float acc[NUM_ACC] = {0, 0, 0, 0, 0, 0};
                if (BWD) phase_init_bwd(P, T, S, 0, 1);
                for (int r = first_step(T); r <= last_step(T); ++r) {
                    phase_A(P, T, S, r, 0, 1);
                    phase_B<BWD>(P, T, S, r, 0, 1, LWp, acc, gd, ge);
                    phase_C<BWD>(P, T, S, r, 0, 1, LWp, gd);
                    phase_D<BWD>(P, T, S, r, 0, 1, LWp, acc, gd, ge);
                }
                if (!BWD)
                    for (int k = 0; k < NUM_ACC; ++k) sums[k] += acc[k];
            }
}

extern "C" int emu_loss_fwd(const UslLossConfig* cfg, const UslLossScale* s,
                            int TW, int R, double* sums) {
    LossParams P;
    to_params(cfg, s, TW, R, &P);
    for (int k = 0; k < NUM_ACC; ++k) sums[k] = 0.0;
    run_main<false>(P, nullptr, sums);
    return 0;
}

static void run_scatter(const LossParams& P, int consR, const float* gout,
                        int accumulate = 0) {
    ConsParams C = {};
    C.B = P.B; C.h = P.h; C.w = P.w;
    C.disp = P.disp; C.d_bs = P.d_bs; C.d_cs = P.d_cs;
    C.unc = P.unc; C.u_bs = P.u_bs; C.u_cs = P.u_cs;
    C.gout_d = gout; C.gout_e = gout + 1;
    C.grad_disp = P.grad_disp; C.gd_bs = P.gd_bs; C.gd_cs = P.gd_cs;
    C.terms = P.terms & (TERM_CONS_D | TERM_CONS_U);
    C.coef_dd = P.coef[ACC_CONS_D]; C.coef_ud = P.coef[ACC_CONS_U];
    C.R = consR > C.h ? C.h : consR;
    C.accumulate = accumulate;
    std::vector<float> arena(cons_ring_floats(C.w));
    for (int b = 0; b < C.B; ++b)
        for (int ya = 0; ya < C.h; ya += C.R) {
            for (auto& v : arena) v = NAN;
            ConsTile T;
            T.b = b; T.ya = ya; T.yb = ya + C.R < C.h ? ya + C.R : C.h;
            const ConsRings S = cons_carve(C.w, arena.data());
            for (int r = cons_first_step(T); r <= cons_last_step(T); ++r) {
                cons_phase_A(C, T, S, r, 0, 1);
                cons_phase_B(C, T, S, r, 0, 1, gout[0], gout[1]);
                cons_phase_C_host(C, S, r);
                cons_phase_D(C, T, S, r, 0, 1);
            }
        }
}

// The transposed warp of the consistency terms alone (pure store into
// grad_disp); the marching emulation adds the rest on top of it.
extern "C" int emu_cons_scatter(const UslLossConfig* cfg, const UslLossScale* s,
                                int consR, const float* gout) {
    LossParams P;
    to_params(cfg, s, 16, 16, &P);
    if (P.terms & (TERM_CONS_D | TERM_CONS_U)) run_scatter(P, consR, gout);
    return 0;
}

// ... added to what the column-marching emulation stored before.
extern "C" int emu_cons_scatter_add(const UslLossConfig* cfg, const UslLossScale* s,
                                    int consR, const float* gout) {
    LossParams P;
    to_params(cfg, s, 16, 16, &P);
    if (P.terms & (TERM_CONS_D | TERM_CONS_U)) run_scatter(P, consR, gout, 1);
    return 0;
}

extern "C" int emu_loss_bwd(const UslLossConfig* cfg, const UslLossScale* s,
                            int TW, int R, int consR, const float* gout) {
    LossParams P;
    to_params(cfg, s, TW, R, &P);
    P.gout_d = gout; P.gout_e = gout + 1;
    if (P.terms & (TERM_CONS_D | TERM_CONS_U)) {
        run_scatter(P, consR, gout);
        P.grad_disp_accumulate = 1;
    }
    run_main<true>(P, gout, nullptr);
    return 0;
}
