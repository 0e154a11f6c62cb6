// Deterministic, atomic-free transposed warp: the part of kernel (2) that
// scatters.
//
// ConsistencyLoss (loss.py:167-188) compares a_v with warp(b_opp; -/+ a_v); its
// gradient w.r.t. the *sampled* map b_opp is the transpose of a bilinear gather
// with data-dependent columns.  ATen's grid_sampler backward uses float
// atomics for this (non-deterministic); here:
//
//   * the vertical taps depend on the source row only, so each source row r
//     is first scattered horizontally into a private row H(r) in shared
//     memory, and destination row y' is then assembled from H(y'-1), H(y'),
//     H(y'+1) in fixed order with the row weights;
//   * the horizontal scatter of one row is done by ONE warp walking the row in
//     32-column chunks, in order.  Within a chunk lanes with the same
//     destination are grouped (runs of a monotone chunk, else match.any),
//     summed by the group leader in lane order, and the leaders -- whose
//     destinations are now unique -- do
//     plain shared-memory read-modify-writes (first tap, __syncwarp, second
//     tap).  No atomics; the summation order is a pure function of the data.
//
// Both consistency terms of a scale scatter into the same disparity planes:
//   term dd: a = disparity,   b = disparity  (loss.py:545)
//   term ud: a = uncertainty, b = disparity  (loss.py:430-431)
#pragma once

#include "loss_core.cuh"

namespace usl {

struct ConsParams {
    int B, h, w;
    const float* disp; long long d_bs, d_cs;
    const float* unc;  long long u_bs, u_cs;
    const float* gout_d;                     // device scalars: upstream grads
    const float* gout_e;                     // (NULL = gout_default)
    float gout_default;
    float* grad_disp; long long gd_bs, gd_cs;  // both planes
    int accumulate;                          // add to what is there (else store)
    unsigned terms;                          // TERM_CONS_D | TERM_CONS_U
    float coef_dd, coef_ud;
    int R;                                   // strip height
    const float* scat;                       // LossParams::scat of the column kernels:
                                             // [b][view][h*w] x {d, u, s_dd, s_ud}
};

struct ConsTile { int b, ya, yb; };

struct ConsRings {
    float* Vd;     // [2][w]   blended opposite disparity for source view v
    float* H;      // [4][2][w] slot, destination view, column
    int* dest;     // [4][w]   job = term*2 + view
    float* c0;     // [4][w]
    float* c1;     // [4][w]
};

USL_HD size_t cons_ring_floats(int w) { return (size_t)(2 + 8 + 12) * w; }

USL_HD ConsRings cons_carve(int w, float* base) {
    ConsRings S;
    S.Vd = base; base += 2 * (size_t)w;
    S.H = base; base += 8 * (size_t)w;
    S.dest = reinterpret_cast<int*>(base); base += 4 * (size_t)w;
    S.c0 = base; base += 4 * (size_t)w;
    S.c1 = base;
    return S;
}

USL_HD int cons_first_step(const ConsTile& T) { return T.ya - 1; }
USL_HD int cons_last_step(const ConsTile& T) { return T.yb; }

// A: blended opposite-disparity row + clear this row's H slot.
USL_HD void cons_phase_A(const ConsParams& P, const ConsTile& T,
                         const ConsRings& S, int r, int tid, int nt) {
    const int slot = mod4(r);   // 4 slots: D(r-1) may still read row r-3
    for (int it = tid; it < 2 * P.w; it += nt)
        S.H[(size_t)slot * 2 * P.w + it] = 0.0f;
    if (r < 0 || r >= P.h) return;
    const Tap2 ty = warp_row_taps(r, P.h);
    const bool ok0 = ty.i0 >= 0 && ty.i0 < P.h;
    const bool ok1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < P.h;
    const float w0 = ok0 ? ty.w0 : 0.0f, w1 = ok1 ? ty.w1 : 0.0f;
    const long long o0 = (long long)(ok0 ? ty.i0 : 0) * P.w;
    const long long o1 = (long long)(ok1 ? ty.i0 + 1 : 0) * P.w;
    for (int it = tid; it < 2 * P.w; it += nt) {
        const int v = it / P.w, x = it - v * P.w;
        const float* pd = plane(P.disp, P.d_bs, P.d_cs, T.b, 1 - v);
        S.Vd[it] = w0 * USL_LDG(pd + o0 + x) + w1 * USL_LDG(pd + o1 + x);
    }
}

// B: per source pixel, destination column and the two tap contributions.
USL_HD void cons_phase_B(const ConsParams& P, const ConsTile& T,
                         const ConsRings& S, int r, int tid, int nt,
                         float gd_up, float ge_up) {
    if (r < 0 || r >= P.h) return;
    for (int it = tid; it < 2 * P.w; it += nt) {
        const int v = it / P.w, x = it - v * P.w;
        const float sign = v ? 1.0f : -1.0f;
        const long long pix = (long long)r * P.w + x;
        const float* Vrow = S.Vd + (size_t)v * P.w;
        for (int term = 0; term < 2; ++term) {
            if (!(P.terms & (term ? TERM_CONS_U : TERM_CONS_D))) continue;
            const float a = term
                ? USL_LDG(plane(P.unc, P.u_bs, P.u_cs, T.b, v) + pix)
                : USL_LDG(plane(P.disp, P.d_bs, P.d_cs, T.b, v) + pix);
            const Tap2 tx = split_coord(warp_coord(x, P.w, sign * a));
            const float f0 = (tx.i0 >= 0 && tx.i0 < P.w) ? Vrow[tx.i0] : 0.0f;
            const float f1 = (tx.i0 + 1 >= 0 && tx.i0 + 1 < P.w) ? Vrow[tx.i0 + 1] : 0.0f;
            const float f = a - (tx.w0 * f0 + tx.w1 * f1);
            const float rr = (term ? ge_up * P.coef_ud : gd_up * P.coef_dd) * sgnf(f);
            const size_t j = (size_t)(term * 2 + v) * P.w + x;
            S.dest[j] = tx.i0;
            S.c0[j] = -rr * tx.w0;
            S.c1[j] = -rr * tx.w1;
        }
    }
}

// C (host form): sequential scatter in column order.  The device form lives
// in loss_kernels.cu (warp-serial, same destination set, fixed order).
inline void cons_phase_C_host(const ConsParams& P, const ConsRings& S, int r) {
    if (r < 0 || r >= P.h) return;
    for (int v = 0; v < 2; ++v) {
        float* Hrow = S.H + ((size_t)mod4(r) * 2 + (1 - v)) * P.w;
        for (int term = 0; term < 2; ++term) {
            if (!(P.terms & (term ? TERM_CONS_U : TERM_CONS_D))) continue;
            const size_t j = (size_t)(term * 2 + v) * P.w;
            for (int x = 0; x < P.w; ++x) {
                const int d = S.dest[j + x];
                if (d >= 0 && d < P.w) Hrow[d] += S.c0[j + x];
                if (d + 1 >= 0 && d + 1 < P.w) Hrow[d + 1] += S.c1[j + x];
            }
        }
    }
}

// D: assemble destination row y' = r - 1 from the three source rows around it.
USL_HD void cons_phase_D(const ConsParams& P, const ConsTile& T,
                         const ConsRings& S, int r, int tid, int nt) {
    const int yd = r - 1;
    if (yd < T.ya || yd >= T.yb) return;
    float wgt[3];
    for (int k = 0; k < 3; ++k) {
        const int rs = yd - 1 + k;
        wgt[k] = 0.0f;
        if (rs < 0 || rs >= P.h) continue;
        const Tap2 ty = warp_row_taps(rs, P.h);
        if (ty.i0 == yd) wgt[k] = ty.w0;
        else if (ty.i0 + 1 == yd) wgt[k] = ty.w1;
    }
    for (int it = tid; it < 2 * P.w; it += nt) {
        const int o = it / P.w, x = it - o * P.w;
        float total = 0.0f;
        for (int k = 0; k < 3; ++k) {
            const int rs = yd - 1 + k;
            if (rs < 0 || rs >= P.h) continue;
            total += wgt[k] * S.H[((size_t)mod4(rs) * 2 + o) * P.w + x];
        }
        float* out = P.grad_disp + (long long)T.b * P.gd_bs + o * P.gd_cs +
                     (long long)yd * P.w + x;
        *out = P.accumulate ? *out + total : total;
    }
}

}  // namespace usl
