// Register-marching form of the fused loss kernels -- the hot path of the
// training step (warp in-kernel, reprojection term on).  Same mathematics as
// loss_core.cuh (see the header there for the reference citations: utils.py:
// 65-135, loss.py:43-151,167-188,208-264,389-434,539-568); re-organised so that
// a pixel costs as few issued instructions as possible, because on B200 this
// path is bound by the instruction issue rate long before HBM:
//
//  * one CTA owns one (sample, row strip, column tile) of one pyramid scale,
//    BOTH views; a thread owns TWO adjacent columns of ONE view for the whole
//    strip and marches down it one image row per step;
//  * everything vertical lives in registers: the 3x3 SSIM window sums are
//    separable, a thread keeps the horizontal 3-sums of the two previous rows
//    of {x, y, x^2+y^2, xy} per channel (the backward: of the three dSSIM maps)
//    and adds the current row -- 2 adds per quantity instead of a 9-tap loop.
//    The row loop is unrolled by two so that the two history slots swap roles
//    by register renaming instead of moves;
//  * everything horizontal goes through ONE shared-memory row per quantity
//    (the right/left neighbour pair), 64-bit accesses, conflict free;
//  * inputs are read exactly once per CTA from HBM as coalesced 64-bit loads
//    straight into registers, one row ahead of their use; the opposite view's
//    vertically blended row V(r) -- {R,G,B,disparity} packed as one 16-byte
//    element per column -- is the only staged input (the gather is data
//    dependent);
//  * GRAD mode produces the loss sums AND the gradient in the same pass: the
//    backward needs every forward intermediate anyway, and the upstream
//    gradients enter linearly (device scalars, NULL = 1).
//
// Step r of a strip [ya, yb):      (row r enters; results trail by 2 rows)
//   pB  warp row r from V(r): recon y(r), d(recon)/d(shift), |x - y|;
//       consistency terms; smoothness edges (r-1,r) and the row's own
//   pC  horizontal 3-sums of row r, SSIM of window row q = r-2 -> dssim(q),
//       dSSIM/d{mean_y, E[y^2], E[xy]} maps G(q)
//   pD  3x3 box of G -> d(loss)/d(recon) of row r-2 -> through the warp;
//       error map row r-2 (up-sampled dssim + L1) -> reprojection and
//       uncertainty terms; gradient row r-2 written; V(r+1) produced
//
// Every function is host+device: tests/emu runs the phases on the CPU with one
// TState per emulated thread (same rings, same step order).
#pragma once

#include "loss_core.cuh"

namespace usl {
namespace mk {

#if defined(__CUDA_ARCH__)
typedef float2 P2;
#else
struct alignas(8) P2 { float x, y; };
#endif

USL_HD P2 p2(float a, float b) { P2 r; r.x = a; r.y = b; return r; }
USL_HD P2 operator+(P2 a, P2 b) { return p2(a.x + b.x, a.y + b.y); }
USL_HD P2 operator-(P2 a, P2 b) { return p2(a.x - b.x, a.y - b.y); }
USL_HD P2 operator*(P2 a, P2 b) { return p2(a.x * b.x, a.y * b.y); }
USL_HD P2 operator*(float s, P2 a) { return p2(s * a.x, s * a.y); }
USL_HD P2 fma2(P2 a, P2 b, P2 c) { return p2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
USL_HD P2 fma2(float s, P2 b, P2 c) { return p2(fmaf(s, b.x, c.x), fmaf(s, b.y, c.y)); }
USL_HD P2 fma2(float s, P2 b, float c) { return p2(fmaf(s, b.x, c), fmaf(s, b.y, c)); }
USL_HD P2 ld2(const float* p) { return *reinterpret_cast<const P2*>(p); }
USL_HD void st2(float* p, P2 v) { *reinterpret_cast<P2*>(p) = v; }
#if defined(__CUDA_ARCH__)
USL_HD P2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
#else
USL_HD P2 ldg2(const float* p) { return ld2(p); }
#endif

// s * sgn(v)   (torch: d|v|/dv = sign(v), 0 at 0)
USL_HD float sgn_mul(float v, float s) {
    return v > 0.0f ? s : (v < 0.0f ? -s : 0.0f);
}

constexpr int NH_GRAD = 8;   // hist planes: y[3], d(recon)/d(shift)[3], l1, u
constexpr int NH_FWD = 2;    // l1, u

// One row of a thread's own inputs: its two columns, the pair to the right.
struct RowIn {
    P2 x[3], xn[3];     // image channels of the own view at c0,c0+1 / c0+2,c0+3
    P2 d, u;            // disparity, uncertainty of the own view
    float dn, un;       // ... at column c0+2
};

struct TState {
    RowIn in[2];        // rows r (slot r&1) and r-1
    P2 y[3];            // recon of the current row (pB -> pC)
    P2 H[2][3][4];      // horizontal 3-sums of the two previous rows
    P2 HG[2][3][3];     // ... of the G maps of the two previous window rows
    P2 gd[3], gu[3];    // gradient accumulators of rows r, r-1, r-2
    float acc[NUM_ACC];
    // per-thread constants
    P2 xbase;           // linspace(0,1,w) at c0, c0+1
    P2 txw;             // transposed up-sample column weights (GRAD)
    P2 axw;             // up-sample column tap weight of i1
    int ax0, ax1;       // local ds columns (i0 | i1 << 16) of c0 / c0+1
    P2 own;             // 1 where the column belongs to the tile
    int v, k, c0;
    bool active;
};

struct Geo {                 // per-CTA constants
    int b, xa, xb, ya, yb, cbeg;
    int npairs;              // active pairs per view in this tile
    int np;                  // thread slots per view
    int qlo;                 // first dssim row this strip forms
    float sH, sW;            // (h-2 -> h), (w-2 -> w) align-corners scales
    float gd_up, ge_up;      // upstream gradients (GRAD)
};

struct MRings {
    F4* V;          // [2][w]
    float* yx;      // [6][LW]
    float* gx;      // [18][LW]            (GRAD)
    float* ds;      // [4][2][LW]
    float* hist;    // [3][NH][2][LW]      thread-private
    float* edge;    // [2][LW]  (s_d, s_u) of the pair's right edge   (GRAD)
    float* tyw;     // [R + 4]  transposed up-sample row weights      (GRAD)
    int LW;
};

USL_HD int first_row(const Geo& G) { return G.ya - 2; }
USL_HD int last_row(const Geo& G) { return G.yb + 1; }

USL_HD size_t march_floats(const LossParams& P, bool grad) {
    const size_t LW = P.LW;
    size_t n = (size_t)8 * P.w + 6 * LW + 8 * LW;
    n += (size_t)3 * (grad ? NH_GRAD : NH_FWD) * 2 * LW;
    if (grad) n += 18 * LW + 2 * LW + (size_t)(P.R + 4);
    return n;
}

USL_HD MRings march_carve(const LossParams& P, float* base, bool grad) {
    MRings S;
    S.LW = P.LW;
    const size_t LW = P.LW;
    S.V = reinterpret_cast<F4*>(base); base += (size_t)8 * P.w;
    S.yx = base; base += 6 * LW;
    S.ds = base; base += 8 * LW;
    S.hist = base; base += (size_t)3 * (grad ? NH_GRAD : NH_FWD) * 2 * LW;
    S.gx = S.edge = S.tyw = nullptr;
    if (grad) {
        S.gx = base; base += 18 * LW;
        S.edge = base; base += 2 * LW;
        S.tyw = base;
    }
    return S;
}

// ---- CTA prologue -----------------------------------------------------------
template <bool GRAD>
USL_HD void thread_init(const LossParams& P, const Geo& G, const MRings& S,
                        int tid, TState& T) {
    T.v = tid >= G.np ? 1 : 0;
    T.k = tid - T.v * G.np;
    T.c0 = G.cbeg + 2 * T.k;
    T.active = T.k < G.npairs && tid < 2 * G.np;
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 3; ++c) {
            for (int q = 0; q < 4; ++q) T.H[a][c][q] = p2(0.f, 0.f);
            for (int m = 0; m < 3; ++m) T.HG[a][c][m] = p2(0.f, 0.f);
        }
    for (int a = 0; a < 3; ++a) T.gd[a] = T.gu[a] = p2(0.f, 0.f);
    for (int c = 0; c < 3; ++c) T.y[c] = p2(0.f, 0.f);
    for (int k = 0; k < NUM_ACC; ++k) T.acc[k] = 0.f;
    for (int s = 0; s < 2; ++s) {
        RowIn& I = T.in[s];
        for (int c = 0; c < 3; ++c) I.x[c] = I.xn[c] = p2(0.f, 0.f);
        I.d = I.u = p2(0.f, 0.f); I.dn = I.un = 0.f;
    }
    const int w = P.w;
    T.xbase = p2(linspace01(T.c0, w), linspace01(T.c0 + 1, w));
    T.own = p2((T.c0 >= G.xa && T.c0 < G.xb) ? 1.f : 0.f,
               (T.c0 + 1 >= G.xa && T.c0 + 1 < G.xb) ? 1.f : 0.f);
    T.txw = p2(0.f, 0.f);
    T.axw = p2(0.f, 0.f);
    T.ax0 = T.ax1 = 0;
    if (!T.active) return;
    for (int j = 0; j < 2; ++j) {
        const int c = T.c0 + j;
        if (c >= w) continue;
        const TapAC ax = ac_taps(c, G.sW, w - 2);
        const int packed = (ax.i0 - G.cbeg) | ((ax.i1 - G.cbeg) << 16);
        if (j == 0) { T.ax0 = packed; T.axw.x = ax.w1; }
        else { T.ax1 = packed; T.axw.y = ax.w1; }
        if (GRAD && c <= w - 3) {
            const float t = upsample_transpose_weight(c, w - 2, w, G.sW);
            if (j == 0) T.txw.x = t; else T.txw.y = t;
        }
    }
}

// Row weights of the transposed up-sample for the window rows of the strip.
USL_HD void cta_init_tables(const LossParams& P, const Geo& G, const MRings& S,
                            int tid, int nt) {
    for (int i = tid; i < P.R + 4; i += nt) {
        const int q = G.ya - 2 + i;
        S.tyw[i] = (q >= 0 && q <= P.h - 3)
                       ? upsample_transpose_weight(q, P.h - 2, P.h, G.sH) : 0.0f;
    }
}

// ---- loads of a thread's own row --------------------------------------------
USL_HD void load_row(const LossParams& P, const Geo& G, const TState& T, int r,
                     RowIn& I) {
    if (!T.active || r < 0 || r >= P.h) return;
    const int w = P.w;
    const long long o = (long long)r * w + T.c0;
    const float* im = plane(P.img, P.img_bs, P.img_cs, G.b, T.v * 3);
    const bool nb = T.c0 + 2 < w;
    for (int c = 0; c < 3; ++c) {
        I.x[c] = ldg2(im + c * P.img_cs + o);
        I.xn[c] = nb ? ldg2(im + c * P.img_cs + o + 2) : p2(0.f, 0.f);
    }
    const float* pd = plane(P.disp, P.d_bs, P.d_cs, G.b, T.v) + o;
    const float* pu = plane(P.unc, P.u_bs, P.u_cs, G.b, T.v) + o;
    I.d = ldg2(pd);
    I.u = ldg2(pu);
    I.dn = nb ? USL_LDG(pd + 2) : 0.f;
    I.un = nb ? USL_LDG(pu + 2) : 0.f;
}

// ---- pV: V(r) for the whole row, both views ---------------------------------
USL_HD void pV(const LossParams& P, const Geo& G, const MRings& S, int r,
               int tid, int nt) {
    if (r < 0 || r >= P.h) return;
    const int w = P.w;
    const Tap2 ty = warp_row_taps(r, P.h);
    const bool ok0 = ty.i0 >= 0 && ty.i0 < P.h;
    const bool ok1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < P.h;
    const float w0 = ok0 ? ty.w0 : 0.0f, w1 = ok1 ? ty.w1 : 0.0f;
    const long long o0 = (long long)(ok0 ? ty.i0 : ty.i0 + 1) * w;
    const long long o1 = (long long)(ok1 ? ty.i0 + 1 : ty.i0) * w;
    const int half = w >> 1;
    for (int it = tid; it < w; it += nt) {
        const int v = it >= half ? 1 : 0;
        const int c = (it - v * half) << 1;
        const int opp = 1 - v;
        const float* im = plane(P.img, P.img_bs, P.img_cs, G.b, opp * 3) + c;
        const float* pd = plane(P.disp, P.d_bs, P.d_cs, G.b, opp) + c;
        const P2 a0 = ldg2(im + o0), b0 = ldg2(im + o1);
        const P2 a1 = ldg2(im + P.img_cs + o0), b1 = ldg2(im + P.img_cs + o1);
        const P2 a2 = ldg2(im + 2 * P.img_cs + o0), b2 = ldg2(im + 2 * P.img_cs + o1);
        const P2 da = ldg2(pd + o0), db = ldg2(pd + o1);
        F4 e0, e1;
        e0.x = w0 * a0.x + w1 * b0.x; e1.x = w0 * a0.y + w1 * b0.y;
        e0.y = w0 * a1.x + w1 * b1.x; e1.y = w0 * a1.y + w1 * b1.y;
        e0.z = w0 * a2.x + w1 * b2.x; e1.z = w0 * a2.y + w1 * b2.y;
        e0.w = w0 * da.x + w1 * db.x; e1.w = w0 * da.y + w1 * db.y;
        S.V[v * w + c] = e0;
        S.V[v * w + c + 1] = e1;
    }
}

USL_HD float edge_w(P2 a0, P2 a1, P2 a2, P2 b0, P2 b1, P2 b2, int j) {
    const float g = j == 0
        ? fabsf(a0.x - b0.x) + fabsf(a1.x - b1.x) + fabsf(a2.x - b2.x)
        : fabsf(a0.y - b0.y) + fabsf(a1.y - b1.y) + fabsf(a2.y - b2.y);
    return USL_EXP(g * (-1.0f / 3.0f));
}

USL_HD float* hist_at(const MRings& S, int nh, int r, int q, int v, int l0) {
    return S.hist + ((size_t)((mod3(r) * nh + q) * 2 + v)) * S.LW + l0;
}

// ---- pB: warp of row r, row-local terms --------------------------------------
// PAR = r & 1 (compile time): T.in[PAR] holds row r, T.in[PAR ^ 1] row r-1.
template <bool GRAD, int PAR>
USL_HD void pB(const LossParams& P, const Geo& G, const MRings& S, int r,
               TState& T) {
    // rotate the gradient accumulators (row r enters at index 0)
    if (GRAD) {
        T.gd[2] = T.gd[1]; T.gd[1] = T.gd[0]; T.gd[0] = p2(0.f, 0.f);
        T.gu[2] = T.gu[1]; T.gu[1] = T.gu[0]; T.gu[0] = p2(0.f, 0.f);
    }
    if (!T.active || r < 0 || r >= P.h) return;
    const int w = P.w, v = T.v, l0 = 2 * T.k;
    const int nh = GRAD ? NH_GRAD : NH_FWD;
    const float sign = v ? 1.0f : -1.0f;
    const float fw = (float)w;
    const RowIn& I = T.in[PAR];
    const RowIn& Ip = T.in[PAR ^ 1];
    const F4* Vrow = S.V + v * w;
    const float dj[2] = {I.d.x, I.d.y}, uj[2] = {I.u.x, I.u.y};
    const bool c1ok = T.c0 + 1 < w;
    float yj[3][2], dIj[3][2], wd[2], dwd[2];
    for (int j = 0; j < 2; ++j) {
        const float xb = j ? T.xbase.y : T.xbase.x;
        const float g = fmaf(2.0f, xb + sign * dj[j], -1.0f);
        const Tap2 tx = split_coord(fmaf(g + 1.0f, 0.5f * fw, -0.5f));
        const TapPair t = gather2(Vrow, tx.i0, w);
        yj[0][j] = tx.w0 * t.f0.x + tx.w1 * t.f1.x;
        yj[1][j] = tx.w0 * t.f0.y + tx.w1 * t.f1.y;
        yj[2][j] = tx.w0 * t.f0.z + tx.w1 * t.f1.z;
        wd[j] = tx.w0 * t.f0.w + tx.w1 * t.f1.w;
        if (GRAD) {
            dIj[0][j] = fw * (t.f1.x - t.f0.x);
            dIj[1][j] = fw * (t.f1.y - t.f0.y);
            dIj[2][j] = fw * (t.f1.z - t.f0.z);
            dwd[j] = fw * (t.f1.w - t.f0.w);
        }
    }
    if (!c1ok) { yj[0][1] = yj[1][1] = yj[2][1] = 0.f; }
    P2 l1 = p2(0.f, 0.f);
    for (int c = 0; c < 3; ++c) {
        T.y[c] = p2(yj[c][0], yj[c][1]);
        l1.x += fabsf(I.x[c].x - yj[c][0]);
        l1.y += fabsf(I.x[c].y - yj[c][1]);
        st2(S.yx + (v * 3 + c) * S.LW + l0, T.y[c]);
        if (GRAD) {
            st2(hist_at(S, nh, r, c, v, l0), T.y[c]);
            st2(hist_at(S, nh, r, 3 + c, v, l0), p2(dIj[c][0], dIj[c][1]));
        }
    }
    st2(hist_at(S, nh, r, nh - 2, v, l0), l1);
    st2(hist_at(S, nh, r, nh - 1, v, l0), I.u);
    const bool own_row = r >= G.ya && r < G.yb;
    if (!GRAD && P.recon_out && own_row) {
        const long long hw = (long long)P.h * w;
        for (int c = 0; c < 3; ++c) {
            float* o = P.recon_out + ((long long)G.b * 6 + v * 3 + c) * hw +
                       (long long)r * w + T.c0;
            if (T.own.x != 0.f) o[0] = yj[c][0];
            if (T.own.y != 0.f && c1ok) o[1] = yj[c][1];
        }
    }
    const float ownj[2] = {T.own.x, c1ok ? T.own.y : 0.f};
    const unsigned terms = P.terms;

    // ---- consistency terms of the row (own pixels) ----
    if (own_row) {
        float gdj[2] = {0.f, 0.f}, guj[2] = {0.f, 0.f};
        for (int j = 0; j < 2; ++j) {
            if (terms & TERM_CONS_D) {
                const float f = dj[j] - wd[j];
                T.acc[ACC_CONS_D] += ownj[j] * fabsf(f);
                if (GRAD)
                    gdj[j] = sgn_mul(f, G.gd_up * P.coef[ACC_CONS_D]) *
                             (1.0f - sign * dwd[j]);
            }
            if (terms & TERM_CONS_U) {
                const float xb = j ? T.xbase.y : T.xbase.x;
                const float g = fmaf(2.0f, xb + sign * uj[j], -1.0f);
                const Tap2 tu = split_coord(fmaf(g + 1.0f, 0.5f * fw, -0.5f));
                const float g0 = (tu.i0 >= 0 && tu.i0 < w) ? Vrow[tu.i0].w : 0.0f;
                const float g1 = (tu.i0 + 1 >= 0 && tu.i0 + 1 < w) ? Vrow[tu.i0 + 1].w : 0.0f;
                const float f = uj[j] - (tu.w0 * g0 + tu.w1 * g1);
                T.acc[ACC_CONS_U] += ownj[j] * fabsf(f);
                if (GRAD)
                    guj[j] = sgn_mul(f, G.ge_up * P.coef[ACC_CONS_U]) *
                             (1.0f - sign * fw * (g1 - g0));
            }
        }
        if (GRAD) { T.gd[0] = p2(gdj[0], gdj[1]); T.gu[0] = p2(guj[0], guj[1]); }
    }

    // ---- smoothness: horizontal edges of row r, vertical edges (r-1, r) ----
    if (terms & (TERM_SMOOTH_D | TERM_SMOOTH_U)) {
        const bool prev_own = r - 1 >= G.ya && r - 1 < G.yb;   // implies r >= 1
        const bool sd = (terms & TERM_SMOOTH_D) != 0, su = (terms & TERM_SMOOTH_U) != 0;
        const float kd = G.gd_up * P.coef[ACC_SMOOTH_D];
        const float ku = G.ge_up * P.coef[ACC_SMOOTH_U];
        if (own_row) {
            // edge (c0, c1) and edge (c1, c1 + 1)
            const float wx0 = c1ok ? edge_w(I.x[0], I.x[1], I.x[2],
                                            p2(I.x[0].y, 0.f), p2(I.x[1].y, 0.f),
                                            p2(I.x[2].y, 0.f), 0) : 0.f;
            const bool e1ok = T.c0 + 2 < w;
            const float wx1 = e1ok ? edge_w(p2(I.x[0].y, 0.f), p2(I.x[1].y, 0.f),
                                            p2(I.x[2].y, 0.f), I.xn[0], I.xn[1],
                                            I.xn[2], 0) : 0.f;
            float ed = 0.f, eu = 0.f;
            if (sd) {
                const float g0 = c1ok ? I.d.x - I.d.y : 0.f;
                const float g1 = e1ok ? I.d.y - I.dn : 0.f;
                T.acc[ACC_SMOOTH_D] += ownj[0] * fabsf(g0) * wx0 + ownj[1] * fabsf(g1) * wx1;
                if (GRAD) {
                    const float s0 = sgn_mul(g0, kd * wx0), s1 = sgn_mul(g1, kd * wx1);
                    T.gd[0].x += s0; T.gd[0].y += s1 - s0; ed = s1;
                }
            }
            if (su) {
                const float g0 = c1ok ? I.u.x - I.u.y : 0.f;
                const float g1 = e1ok ? I.u.y - I.un : 0.f;
                T.acc[ACC_SMOOTH_U] += ownj[0] * fabsf(g0) * wx0 + ownj[1] * fabsf(g1) * wx1;
                if (GRAD) {
                    const float s0 = sgn_mul(g0, ku * wx0), s1 = sgn_mul(g1, ku * wx1);
                    T.gu[0].x += s0; T.gu[0].y += s1 - s0; eu = s1;
                }
            }
            if (GRAD) st2(S.edge + (size_t)v * S.LW + l0, p2(ed, eu));
        }
        if ((own_row || prev_own) && r >= 1) {
            const float wy0 = edge_w(Ip.x[0], Ip.x[1], Ip.x[2], I.x[0], I.x[1], I.x[2], 0);
            const float wy1 = edge_w(Ip.x[0], Ip.x[1], Ip.x[2], I.x[0], I.x[1], I.x[2], 1);
            const float po = prev_own ? 1.f : 0.f, co = own_row ? 1.f : 0.f;
            if (sd) {
                const float g0 = Ip.d.x - I.d.x, g1 = Ip.d.y - I.d.y;
                T.acc[ACC_SMOOTH_D] += po * (ownj[0] * fabsf(g0) * wy0 + ownj[1] * fabsf(g1) * wy1);
                if (GRAD) {
                    const float s0 = sgn_mul(g0, kd * wy0), s1 = sgn_mul(g1, kd * wy1);
                    T.gd[1].x += po * s0; T.gd[1].y += po * s1;
                    T.gd[0].x -= co * s0; T.gd[0].y -= co * s1;
                }
            }
            if (su) {
                const float g0 = Ip.u.x - I.u.x, g1 = Ip.u.y - I.u.y;
                T.acc[ACC_SMOOTH_U] += po * (ownj[0] * fabsf(g0) * wy0 + ownj[1] * fabsf(g1) * wy1);
                if (GRAD) {
                    const float s0 = sgn_mul(g0, ku * wy0), s1 = sgn_mul(g1, ku * wy1);
                    T.gu[1].x += po * s0; T.gu[1].y += po * s1;
                    T.gu[0].x -= co * s0; T.gu[0].y -= co * s1;
                }
            }
        }
    }
}

// ---- pC: SSIM of window row q = r - 2 ---------------------------------------
template <bool GRAD, int PAR>
USL_HD void pC(const LossParams& P, const Geo& G, const MRings& S, int r,
               TState& T) {
    if (!T.active) return;
    const int w = P.w, v = T.v, l0 = 2 * T.k, q = r - 2;
    const bool row_ok = r >= 0 && r < P.h;
    const bool own_row = r >= G.ya && r < G.yb;
    // left neighbour's right edge lands on column c0 (smoothness backward)
    if (GRAD && own_row && T.k > 0 &&
        (P.terms & (TERM_SMOOTH_D | TERM_SMOOTH_U))) {
        const P2 e = ld2(S.edge + (size_t)v * S.LW + l0 - 2);
        T.gd[0].x -= e.x;
        T.gu[0].x -= e.y;
    }
    const RowIn& I = T.in[PAR];
    P2 H[3][4];
    if (row_ok) {
        const bool nb = T.k + 1 < G.npairs;
        for (int c = 0; c < 3; ++c) {
            const P2 xa = I.x[c], xb = I.xn[c], ya = T.y[c];
            const P2 yb = nb ? ld2(S.yx + (v * 3 + c) * S.LW + l0 + 2) : p2(0.f, 0.f);
            float t;
            t = xa.y + xb.x;  H[c][0] = p2(xa.x + t, t + xb.y);
            t = ya.y + yb.x;  H[c][1] = p2(ya.x + t, t + yb.y);
            const float q0 = fmaf(xa.x, xa.x, ya.x * ya.x), q1 = fmaf(xa.y, xa.y, ya.y * ya.y);
            const float q2 = fmaf(xb.x, xb.x, yb.x * yb.x), q3 = fmaf(xb.y, xb.y, yb.y * yb.y);
            t = q1 + q2;      H[c][2] = p2(q0 + t, t + q3);
            t = fmaf(xa.y, ya.y, xb.x * yb.x);
            H[c][3] = p2(fmaf(xa.x, ya.x, t), fmaf(xb.y, yb.y, t));
        }
    } else {
        for (int c = 0; c < 3; ++c)
            for (int m = 0; m < 4; ++m) H[c][m] = p2(0.f, 0.f);
    }
    const bool q_ok = q >= G.qlo && q <= P.h - 3;
    if (q_ok) {
        const float inv9 = 1.0f / 9.0f;
        const bool p0ok = T.c0 <= w - 3, p1ok = T.c0 + 1 <= w - 3;
        float kk = 0.f;
        if (GRAD)
            kk = -0.5f * G.gd_up * P.coef[ACC_REPROJ] * P.alpha * (1.0f / 3.0f) *
                 S.tyw[q - (G.ya - 2)];
        P2 dsum = p2(0.f, 0.f);
        for (int c = 0; c < 3; ++c) {
            const P2 sx = T.H[0][c][0] + T.H[1][c][0] + H[c][0];
            const P2 sy = T.H[0][c][1] + T.H[1][c][1] + H[c][1];
            const P2 sq = T.H[0][c][2] + T.H[1][c][2] + H[c][2];
            const P2 sxy = T.H[0][c][3] + T.H[1][c][3] + H[c][3];
            const P2 mx = inv9 * sx, my = inv9 * sy;
            const P2 mm = mx * my;
            const P2 m2 = fma2(mx, mx, my * my);
            const P2 n1 = fma2(2.0f, mm, P.c1);
            const P2 d1 = p2(m2.x + P.c1, m2.y + P.c1);
            const P2 sig = fma2(inv9, sq, p2(-m2.x, -m2.y));      // var_x + var_y
            const P2 d2 = p2(sig.x + P.c2, sig.y + P.c2);
            const P2 vxy = fma2(inv9, sxy, p2(-mm.x, -mm.y));
            const P2 n2 = fma2(2.0f, vxy, P.c2);
            const P2 i1 = p2(USL_DIV(1.0f, d1.x), USL_DIV(1.0f, d1.y));
            const P2 i2 = p2(USL_DIV(1.0f, d2.x), USL_DIV(1.0f, d2.y));
            const P2 inv = i1 * i2;
            const P2 nn = n1 * n2;
            const P2 ssim = nn * inv;
            const P2 raw = fma2(-0.5f, ssim, 0.5f);
            dsum.x += p0ok ? fminf(fmaxf(raw.x, 0.0f), 1.0f) : 0.f;
            dsum.y += p1ok ? fminf(fmaxf(raw.y, 0.0f), 1.0f) : 0.f;
            if (GRAD) {
                // d reproj / d dssim(q) = coef * alpha/3 * T(q); the clamp
                // passes the gradient on the closed interval
                const float gb0 = (p0ok && raw.x >= 0.0f && raw.x <= 1.0f) ? kk * T.txw.x : 0.0f;
                const float gb1 = (p1ok && raw.y >= 0.0f && raw.y <= 1.0f) ? kk * T.txw.y : 0.0f;
                const P2 gb = p2(gb0, gb1);
                // dssim/dA = 2 mx (n2 - n1) inv - 2 my nn (d2 - d1) inv^2
                const P2 t1 = (2.0f * mx) * (n2 - n1);
                const P2 t2 = (2.0f * my) * (ssim * (d2 - d1));
                const P2 gA = gb * ((t1 - t2) * inv);
                const P2 gQ = p2(-gb.x, -gb.y) * (ssim * i2);
                const P2 gC = gb * (2.0f * (n1 * inv));
                st2(S.gx + ((v * 9 + c * 3 + 0) * (size_t)S.LW) + l0, gA);
                st2(S.gx + ((v * 9 + c * 3 + 1) * (size_t)S.LW) + l0, gQ);
                st2(S.gx + ((v * 9 + c * 3 + 2) * (size_t)S.LW) + l0, gC);
            }
        }
        st2(S.ds + (size_t)(mod4(q) * 2 + v) * S.LW + l0, dsum);
    } else if (GRAD) {
        for (int m = 0; m < 9; ++m)
            st2(S.gx + ((v * 9 + m) * (size_t)S.LW) + l0, p2(0.f, 0.f));
    }
    // the older history slot takes the new row (roles swap with PAR)
    for (int c = 0; c < 3; ++c)
        for (int m = 0; m < 4; ++m) T.H[PAR][c][m] = H[c][m];
}

USL_HD float unc_loss2(int type, float u, float e) { return unc_loss(type, u, e); }

// ---- pD: everything that needs G(q) / the error map; rows r-2 ---------------
template <bool GRAD, int PAR>
USL_HD void pD(const LossParams& P, const Geo& G, const MRings& S, int r,
               TState& T) {
    if (!T.active) return;
    const int w = P.w, v = T.v, l0 = 2 * T.k;
    const int nh = GRAD ? NH_GRAD : NH_FWD;
    const int y = r - 2;
    const bool own_row = y >= G.ya && y < G.yb;
    const bool c1ok = T.c0 + 1 < w;
    const float sign = v ? 1.0f : -1.0f;
    const long long pix = (long long)y * w + T.c0;
    const long long hw = (long long)P.h * w;
    if (GRAD) {
        // box-row sums of G(q = r-2) over p in [c-2, c]; zeros stored by pC
        // when q is not a window row
        P2 HGq[3][3];
        for (int c = 0; c < 3; ++c)
            for (int m = 0; m < 3; ++m) {
                const float* g = S.gx + ((v * 9 + c * 3 + m) * (size_t)S.LW) + l0;
                const P2 own = ld2(g);
                const P2 lf = T.k > 0 ? ld2(g - 2) : p2(0.f, 0.f);
                const float t = lf.y + own.x;
                HGq[c][m] = p2(lf.x + t, t + own.y);
            }
        if (own_row && (T.own.x != 0.f || T.own.y != 0.f)) {
            const float* im = plane(P.img, P.img_bs, P.img_cs, G.b, v * 3) + pix;
            const float kl1 = G.gd_up * P.coef[ACC_REPROJ] * (1.0f - P.alpha) * (1.0f / 3.0f);
            P2 gs = p2(0.f, 0.f);
            for (int c = 0; c < 3; ++c) {
                const P2 rc = ld2(hist_at(S, nh, y, c, v, l0));
                const P2 di = ld2(hist_at(S, nh, y, 3 + c, v, l0));
                const P2 x = ldg2(im + c * P.img_cs);
                const P2 sA = T.HG[0][c][0] + T.HG[1][c][0] + HGq[c][0];
                const P2 sQ = T.HG[0][c][1] + T.HG[1][c][1] + HGq[c][1];
                const P2 sC = T.HG[0][c][2] + T.HG[1][c][2] + HGq[c][2];
                P2 g = (1.0f / 9.0f) * fma2(x, sC, fma2(2.0f * rc, sQ, sA));
                g.x -= sgn_mul(x.x - rc.x, kl1);
                g.y -= sgn_mul(x.y - rc.y, kl1);
                if (P.grad_recon_in) {
                    const float* gi = P.grad_recon_in + ((long long)G.b * 6 + v * 3 + c) * hw + pix;
                    g.x += USL_LDG(gi);
                    if (c1ok) g.y += USL_LDG(gi + 1);
                }
                gs = fma2(g, di, gs);
            }
            T.gd[2].x += sign * gs.x;
            T.gd[2].y += sign * gs.y;
        }
        for (int c = 0; c < 3; ++c)
            for (int m = 0; m < 3; ++m) T.HG[PAR][c][m] = HGq[c][m];
    }
    if (!own_row) return;
    const float ownj[2] = {T.own.x, c1ok ? T.own.y : 0.f};
    if (ownj[0] == 0.f && ownj[1] == 0.f) return;
    // error map row y: bilinear (h-2, w-2) -> (h, w) of dssim + L1
    const TapAC ay = ac_taps(y, G.sH, P.h - 2);
    const float* d0 = S.ds + (size_t)(mod4(ay.i0) * 2 + v) * S.LW;
    const float* d1 = S.ds + (size_t)(mod4(ay.i1) * 2 + v) * S.LW;
    const P2 l1 = ld2(hist_at(S, nh, y, nh - 2, v, l0));
    const P2 uu = ld2(hist_at(S, nh, y, nh - 1, v, l0));
    const float l1j[2] = {l1.x, l1.y}, uj[2] = {uu.x, uu.y};
    float ej[2] = {0.f, 0.f}, guj[2] = {0.f, 0.f};
    for (int j = 0; j < 2; ++j) {
        if (ownj[j] == 0.f) continue;
        const int pk = j ? T.ax1 : T.ax0;
        const int i0 = pk & 0xffff, i1 = pk >> 16;
        const float w1 = j ? T.axw.y : T.axw.x, w0 = 1.0f - w1;
        // a zero-weight tap may point at a row/column that was never formed
        const float a00 = d0[i0], a01 = w1 != 0.f ? d0[i1] : 0.f;
        float up = w0 * a00 + w1 * a01;
        if (ay.w1 != 0.f) {
            const float b00 = d1[i0], b01 = w1 != 0.f ? d1[i1] : 0.f;
            up = ay.w0 * up + ay.w1 * (w0 * b00 + w1 * b01);
        } else {
            up = ay.w0 * up;
        }
        const float e = (P.alpha * up + (1.0f - P.alpha) * l1j[j]) * (1.0f / 3.0f);
        ej[j] = e;
        T.acc[ACC_REPROJ] += e;
        if (P.terms & TERM_UNC) {
            T.acc[ACC_UNC] += unc_loss(P.loss_type, uj[j], e);
            if (GRAD)
                guj[j] = G.ge_up * P.coef[ACC_UNC] * unc_loss_grad(P.loss_type, uj[j], e);
        }
    }
    if (P.err_out) {
        float* o = P.err_out + ((long long)G.b * 2 + v) * hw + pix;
        if (ownj[0] != 0.f) o[0] = ej[0];
        if (ownj[1] != 0.f) o[1] = ej[1];
    }
    if (GRAD) {
        float* od = P.grad_disp + (long long)G.b * P.gd_bs + v * P.gd_cs + pix;
        float* ou = P.grad_unc + (long long)G.b * P.gu_bs + v * P.gu_cs + pix;
        P2 a = T.gd[2];
        const P2 bu = p2(T.gu[2].x + guj[0], T.gu[2].y + guj[1]);
        if (ownj[0] != 0.f && ownj[1] != 0.f) {
            if (P.grad_disp_accumulate) a = a + ld2(od);
            st2(od, a);
            st2(ou, bu);
        } else if (ownj[0] != 0.f) {
            od[0] = P.grad_disp_accumulate ? od[0] + a.x : a.x;
            ou[0] = bu.x;
        } else {
            od[1] = P.grad_disp_accumulate ? od[1] + a.y : a.y;
            ou[1] = bu.y;
        }
    }
}

}  // namespace mk
}  // namespace usl
