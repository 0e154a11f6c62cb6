// Core of kernels (1) and (2): the fused per-scale stereo loss, forward and
// backward, as a strip-marching stencil.
//
// Reference being replaced (all under /root/reference/train):
//   utils.py:65-135   reconstruct / reconstruct_pyramid (the disparity warp)
//   loss.py:43-151    WeightedSSIMLoss (3x3 SSIM + L1, bilinear up-sample)
//   loss.py:167-188   ConsistencyLoss
//   loss.py:208-264   SmoothnessLoss
//   loss.py:389-434   ReprojectionErrorLoss (l1 / bayesian / log_bayesian)
//   loss.py:539-568   TukraUncertaintyLoss (per-scale sum)
//
// Work decomposition.  One CTA owns one (sample, row strip [ya,yb), column
// tile [xa,xb)) of one pyramid scale and marches down the strip one image row
// per step.  Because the warp's vertical taps depend on the row only
// (usl_math.cuh), the four-tap warp of row r factors into
//   (A) a vertical blend of two source rows into a full-width row V(r) held in
//       shared memory -- {R,G,B,disparity} of the opposite view packed as one
//       16-byte element per column -- followed by
//   (B) a two-tap horizontal gather from V(r) at the disparity-shifted column.
// The 3x3 SSIM windows and the (h-2,w-2)->(h,w) align-corners up-sample are
// evaluated from small row rings in shared memory that trail the marching row:
//   step r:  A  V(r)
//            B  recon(r), |I-recon|(r), cons/smoothness terms of row r
//            C  dssim(q=r-2)   [bwd: + dSSIM/d{mean, E[y^2], E[xy]} maps G(q)]
//            D  E(y=r-3) -> reprojection + uncertainty terms
//               [bwd: SSIM/L1 gradient of row r-2 pushed through the warp,
//                uncertainty gradient of row r-3, row r-3 flushed to HBM]
// Nothing but the inputs is read from HBM and nothing but the results
// (per-CTA partial sums; gradient rows) is written.
//
// Every function here is host+device: tests/emu compiles this header with g++
// and runs the phases sequentially to check the kernel logic on the CPU.
#pragma once

#include "usl_math.cuh"

#if defined(__CUDA_ARCH__)
#define USL_LDG(p) __ldg(p)
#define USL_EXP(x) __expf(x)
#define USL_LOG(x) __logf(x)
#define USL_DIV(a, b) __fdividef(a, b)
#else
#define USL_LDG(p) (*(p))
#define USL_EXP(x) expf(x)
#define USL_LOG(x) logf(x)
#define USL_DIV(a, b) ((a) / (b))
#endif

namespace usl {

enum : unsigned {
    TERM_REPROJ = 1u,     // WeightedSSIMLoss           (loss.py:544)
    TERM_CONS_D = 2u,     // ConsistencyLoss(disp)      (loss.py:545)
    TERM_SMOOTH_D = 4u,   // SmoothnessLoss(disp, img)  (loss.py:546)
    TERM_UNC = 8u,        // l1/bayesian/log_bayesian   (loss.py:426)
    TERM_SMOOTH_U = 16u,  // SmoothnessLoss(unc, img)   (loss.py:428-429)
    TERM_CONS_U = 32u,    // ConsistencyLoss(unc, disp) (loss.py:430-431)
};
enum { ACC_REPROJ = 0, ACC_CONS_D, ACC_SMOOTH_D, ACC_UNC, ACC_SMOOTH_U,
       ACC_CONS_U, NUM_ACC };
enum { LOSS_L1 = 0, LOSS_BAYESIAN = 1, LOSS_LOG_BAYESIAN = 2 };

struct alignas(16) F4 { float x, y, z, w; };

constexpr int HALO_L = 2;   // recon columns needed left of the tile
constexpr int HALO_R = 3;   // ... and right of it
constexpr int HALO_T = 2;   // recon rows needed above the strip
constexpr int HALO_B = 2;   // ... and below it
constexpr int LAG_E = 3;    // E(y) is evaluated at step y + LAG_E

struct LossParams {
    int B, h, w;
    // ---- inputs (planes are h*w contiguous; batch/channel strides in floats)
    const float* img;  long long img_bs, img_cs;    // 6 planes  L_rgb, R_rgb
    const float* disp; long long d_bs, d_cs;        // 2 planes  d_L, d_R
    const float* unc;  long long u_bs, u_cs;        // 2 planes  u_L, u_R
    const float* recon_in; long long ri_bs, ri_cs;  // optional given recon
    const float* err_in;   long long ei_bs, ei_cs;  // optional given error
    // ---- forward outputs
    float* recon_out;   // optional (B,6,h,w) contiguous
    float* err_out;     // optional (B,2,h,w) contiguous
    float* partials;    // [num CTAs][NUM_ACC]
    // ---- backward
    const float* gout_d;         // device scalars: upstream gradients of the
    const float* gout_e;         // disparity / error loss (NULL = 0)
    const float* grad_recon_in;  // optional (B,6,h,w) contiguous, added to dL/drecon
    float* grad_disp; long long gd_bs, gd_cs;   // 2 planes
    float* grad_unc;  long long gu_bs, gu_cs;   // 2 planes
    float* grad_recon_out;       // (B,6,h,w) contiguous when recon is given
    int grad_disp_accumulate;    // add to what the scatter kernel stored
    // input of the lane-per-row transposed warp of the consistency terms
    // (column kernels -> cons_rows_kernel), or NULL: [b][view][h*w] x
    // {d, u, s of term dd, s of term ud}, s = coefficient * sign(a - warp(b))
    float* scat;
    // ---- configuration
    unsigned terms;
    int loss_type;
    int recon_given, err_given;
    float alpha, c1, c2;
    // per-term constants: loss = sum_k coef[k] * S_k; backward multiplies them
    // by the upstream gradient of the output each term belongs to.
    float coef[NUM_ACC];
    // ---- tiling
    int TW, R;          // tile width, strip height
    int LW;             // local ring width  = TW + HALO_L + HALO_R
};

struct Tile {
    int b, xa, xb, ya, yb, cbeg;   // cbeg = xa - HALO_L  (local col 0)
    int cta;                       // linear CTA index -> partials row
};

// Shared-memory arena (all pointers into one float buffer).
struct Rings {
    F4* V;          // [2][w]
    float* recon;   // [3][6][LW]
    float* l1s;     // [4][2][LW]
    float* ds;      // [4][2][LW]
    float* tx;      // [LW]  column factor of the transposed up-sample (bwd)
    float* dIds;    // [3][6][LW]            (bwd)
    float* G;       // [3][2][3][3][LW]      (bwd)  slot, view, ch, {A,Q,C}
    float* grad;    // [4][4][LW]            (bwd)  slot, {dL,dR,uL,uR}
};

USL_HD bool needs_V(const LossParams& P) {
    return ((P.terms & TERM_REPROJ) && !P.recon_given) ||
           (P.terms & (TERM_CONS_D | TERM_CONS_U));
}

USL_HD size_t ring_floats(const LossParams& P, bool bwd) {
    size_t n = 0;
    n += (size_t)2 * P.w * 4;
    n += (size_t)3 * 6 * P.LW + (size_t)4 * 2 * P.LW * 2;
    if (bwd) n += (size_t)P.LW + (size_t)3 * 6 * P.LW +
                  (size_t)3 * 18 * P.LW + (size_t)4 * 4 * P.LW;
    return n;
}

USL_HD Rings carve(const LossParams& P, float* base, bool bwd) {
    Rings S;
    S.V = reinterpret_cast<F4*>(base); base += (size_t)2 * P.w * 4;
    S.recon = base; base += (size_t)3 * 6 * P.LW;
    S.l1s = base; base += (size_t)4 * 2 * P.LW;
    S.ds = base; base += (size_t)4 * 2 * P.LW;
    S.tx = S.dIds = S.G = S.grad = nullptr;
    if (bwd) {
        S.tx = base; base += P.LW;
        S.dIds = base; base += (size_t)3 * 6 * P.LW;
        S.G = base; base += (size_t)3 * 18 * P.LW;
        S.grad = base;
    }
    return S;
}

USL_HD int mod3(int r) { return ((r % 3) + 3) % 3; }
USL_HD int mod4(int r) { return r & 3; }

USL_HD const float* plane(const float* base, long long bs, long long cs, int b,
                          int c) {
    return base + (long long)b * bs + (long long)c * cs;
}

// Column factor of the transposed (w-2 -> w) up-sample: sum over destination
// columns of the weight they put on source column p.
USL_HD float upsample_transpose_weight(int p, int in, int out, float scale) {
    float t = 0.0f;
    for (int d = p; d <= p + 2 && d < out; ++d) {
        const TapAC a = ac_taps(d, scale, in);
        if (a.i0 == p) t += a.w0;
        if (a.i1 == p) t += a.w1;   // i1 == i0 at the clamped end: both count
    }
    return t;
}

// -------------------------------------------------------------------------
// Phase A: V(r)[v][x] = vertical blend of the opposite view's {rgb, disp}.
// -------------------------------------------------------------------------
USL_HD void phase_A(const LossParams& P, const Tile& T, const Rings& S, int r,
                    int tid, int nt) {
    if (r < 0 || r >= P.h || !needs_V(P)) return;
    const Tap2 ty = warp_row_taps(r, P.h);
    const bool ok0 = ty.i0 >= 0 && ty.i0 < P.h;
    const bool ok1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < P.h;
    const float w0 = ok0 ? ty.w0 : 0.0f, w1 = ok1 ? ty.w1 : 0.0f;
    const long long o0 = (long long)(ok0 ? ty.i0 : 0) * P.w;
    const long long o1 = (long long)(ok1 ? ty.i0 + 1 : 0) * P.w;
    const bool need_img = (P.terms & TERM_REPROJ) && !P.recon_given;
    for (int it = tid; it < 2 * P.w; it += nt) {
        const int v = it / P.w, x = it - v * P.w;
        const int opp = 1 - v;
        F4 o;
        o.x = o.y = o.z = 0.0f;
        if (need_img) {
            const float* p0 = plane(P.img, P.img_bs, P.img_cs, T.b, opp * 3);
            const float* p1 = p0 + P.img_cs;
            const float* p2 = p1 + P.img_cs;
            o.x = w0 * USL_LDG(p0 + o0 + x) + w1 * USL_LDG(p0 + o1 + x);
            o.y = w0 * USL_LDG(p1 + o0 + x) + w1 * USL_LDG(p1 + o1 + x);
            o.z = w0 * USL_LDG(p2 + o0 + x) + w1 * USL_LDG(p2 + o1 + x);
        }
        const float* pd = plane(P.disp, P.d_bs, P.d_cs, T.b, opp);
        o.w = w0 * USL_LDG(pd + o0 + x) + w1 * USL_LDG(pd + o1 + x);
        S.V[(size_t)v * P.w + x] = o;
    }
}

struct TapPair { F4 f0, f1; };

USL_HD TapPair gather2(const F4* Vrow, int x0, int w) {
    TapPair t;
    const F4 z = {0.0f, 0.0f, 0.0f, 0.0f};
    t.f0 = (x0 >= 0 && x0 < w) ? Vrow[x0] : z;
    t.f1 = (x0 + 1 >= 0 && x0 + 1 < w) ? Vrow[x0 + 1] : z;
    return t;
}

// edge-aware weight exp(-mean_c |I(a) - I(b)|)  (loss.py:220-222)
USL_HD float edge_weight(const float* im0, long long cs, long long pa,
                         long long pb) {
    const float g = fabsf(USL_LDG(im0 + pa) - USL_LDG(im0 + pb)) +
                    fabsf(USL_LDG(im0 + cs + pa) - USL_LDG(im0 + cs + pb)) +
                    fabsf(USL_LDG(im0 + 2 * cs + pa) - USL_LDG(im0 + 2 * cs + pb));
    return USL_EXP(-g * (1.0f / 3.0f));
}

// Smoothness of one map `m` (disparity or uncertainty) at pixel (r, c).
//   forward : |dx m * wx| + |dy m * wy|
//   backward: d/dm(r,c) of the sum over all pixels of the above
template <bool BWD>
USL_HD float smooth_at(const float* m, const float* im0, long long cs, int r,
                       int c, int h, int w) {
    const long long p = (long long)r * w + c;
    const float mc = USL_LDG(m + p);
    float out = 0.0f;
    if (c + 1 < w) {
        const float wx = edge_weight(im0, cs, p, p + 1);
        const float gx = mc - USL_LDG(m + p + 1);
        out += BWD ? sgnf(gx) * wx : fabsf(gx * wx);
    }
    if (r + 1 < h) {
        const float wy = edge_weight(im0, cs, p, p + w);
        const float gy = mc - USL_LDG(m + p + w);
        out += BWD ? sgnf(gy) * wy : fabsf(gy * wy);
    }
    if (BWD) {
        if (c >= 1) {
            const float wx = edge_weight(im0, cs, p - 1, p);
            out -= sgnf(USL_LDG(m + p - 1) - mc) * wx;
        }
        if (r >= 1) {
            const float wy = edge_weight(im0, cs, p - w, p);
            out -= sgnf(USL_LDG(m + p - w) - mc) * wy;
        }
    }
    return out;
}

// -------------------------------------------------------------------------
// Phase B: recon(r) and the row-local terms.
// -------------------------------------------------------------------------
template <bool BWD>
USL_HD void phase_B(const LossParams& P, const Tile& T, const Rings& S, int r,
                    int tid, int nt, int LWp, float* acc, float gd_up,
                    float ge_up) {
    if (r < 0 || r >= P.h) return;
    const bool own_row = r >= T.ya && r < T.yb;
    const bool reproj = (P.terms & TERM_REPROJ) != 0;
    const int s3 = mod3(r), s4 = mod4(r);
    for (int it = tid; it < 2 * LWp; it += nt) {
        const int v = it / LWp, lc = it - v * LWp;
        const int c = T.cbeg + lc;
        if (lc >= P.LW || c < 0 || c >= P.w) continue;
        const float sign = v ? 1.0f : -1.0f;
        const long long pix = (long long)r * P.w + c;
        const F4* Vrow = S.V + (size_t)v * P.w;
        const bool own = own_row && c >= T.xa && c < T.xb;
        float d = 0.0f;
        TapPair t;
        Tap2 tx;
        const bool have_d = needs_V(P);
        if (have_d) {
            d = USL_LDG(plane(P.disp, P.d_bs, P.d_cs, T.b, v) + pix);
            tx = split_coord(warp_coord(c, P.w, sign * d));
            t = gather2(Vrow, tx.i0, P.w);
        }
        if (reproj && !P.err_given) {
            float rc[3];
            if (P.recon_given) {
                for (int ch = 0; ch < 3; ++ch)
                    rc[ch] = USL_LDG(plane(P.recon_in, P.ri_bs, P.ri_cs, T.b,
                                           v * 3 + ch) + pix);
            } else {
                rc[0] = tx.w0 * t.f0.x + tx.w1 * t.f1.x;
                rc[1] = tx.w0 * t.f0.y + tx.w1 * t.f1.y;
                rc[2] = tx.w0 * t.f0.z + tx.w1 * t.f1.z;
                if (BWD) {
                    const float fw = (float)P.w;
                    S.dIds[((size_t)s3 * 6 + v * 3 + 0) * P.LW + lc] = fw * (t.f1.x - t.f0.x);
                    S.dIds[((size_t)s3 * 6 + v * 3 + 1) * P.LW + lc] = fw * (t.f1.y - t.f0.y);
                    S.dIds[((size_t)s3 * 6 + v * 3 + 2) * P.LW + lc] = fw * (t.f1.z - t.f0.z);
                }
            }
            float l1 = 0.0f;
            for (int ch = 0; ch < 3; ++ch) {
                const float im = USL_LDG(plane(P.img, P.img_bs, P.img_cs, T.b,
                                               v * 3 + ch) + pix);
                l1 += fabsf(im - rc[ch]);
                S.recon[((size_t)s3 * 6 + v * 3 + ch) * P.LW + lc] = rc[ch];
                if (!BWD && own && P.recon_out)
                    P.recon_out[((long long)T.b * 6 + v * 3 + ch) * P.h * P.w + pix] = rc[ch];
            }
            S.l1s[((size_t)s4 * 2 + v) * P.LW + lc] = l1;
        }
        if (!own) continue;
        float gd = 0.0f, gu = 0.0f;
        if (P.terms & TERM_CONS_D) {
            const float wd = tx.w0 * t.f0.w + tx.w1 * t.f1.w;
            const float f = d - wd;
            if (BWD)
                gd += gd_up * P.coef[ACC_CONS_D] * sgnf(f) *
                      (1.0f - sign * (float)P.w * (t.f1.w - t.f0.w));
            else
                acc[ACC_CONS_D] += fabsf(f);
        }
        if (P.terms & TERM_CONS_U) {
            const float u = USL_LDG(plane(P.unc, P.u_bs, P.u_cs, T.b, v) + pix);
            const Tap2 tu = split_coord(warp_coord(c, P.w, sign * u));
            const TapPair g = gather2(Vrow, tu.i0, P.w);
            const float wu = tu.w0 * g.f0.w + tu.w1 * g.f1.w;
            const float f = u - wu;
            if (BWD)
                gu += ge_up * P.coef[ACC_CONS_U] * sgnf(f) *
                      (1.0f - sign * (float)P.w * (g.f1.w - g.f0.w));
            else
                acc[ACC_CONS_U] += fabsf(f);
        }
        if (P.terms & (TERM_SMOOTH_D | TERM_SMOOTH_U)) {
            const float* im0 = plane(P.img, P.img_bs, P.img_cs, T.b, v * 3);
            if (P.terms & TERM_SMOOTH_D) {
                const float s = smooth_at<BWD>(
                    plane(P.disp, P.d_bs, P.d_cs, T.b, v), im0, P.img_cs, r, c,
                    P.h, P.w);
                if (BWD) gd += gd_up * P.coef[ACC_SMOOTH_D] * s;
                else acc[ACC_SMOOTH_D] += s;
            }
            if (P.terms & TERM_SMOOTH_U) {
                const float s = smooth_at<BWD>(
                    plane(P.unc, P.u_bs, P.u_cs, T.b, v), im0, P.img_cs, r, c,
                    P.h, P.w);
                if (BWD) gu += ge_up * P.coef[ACC_SMOOTH_U] * s;
                else acc[ACC_SMOOTH_U] += s;
            }
        }
        if (BWD) {
            S.grad[((size_t)s4 * 4 + v) * P.LW + lc] = gd;
            S.grad[((size_t)s4 * 4 + 2 + v) * P.LW + lc] = gu;
        }
    }
}

// -------------------------------------------------------------------------
// Phase C: dssim(q) summed over the 3 channels of each view; backward also
// stores dLoss/d{P(y), P(y^2), P(xy)} at q for every channel.
// -------------------------------------------------------------------------
template <bool BWD>
USL_HD void phase_C(const LossParams& P, const Tile& T, const Rings& S, int r,
                    int tid, int nt, int LWp, float gd_up) {
    const int q = r - 2;
    if (!(P.terms & TERM_REPROJ) || P.err_given) return;
    int qlo = T.ya - HALO_T; if (qlo < 0) qlo = 0;
    if (q < qlo || q > P.h - 3) return;
    float ty = 0.0f;
    if (BWD) ty = upsample_transpose_weight(q, P.h - 2, P.h,
                                            ac_scale(P.h - 2, P.h));
    const float inv9 = 1.0f / 9.0f;
    for (int it = tid; it < 2 * LWp; it += nt) {
        const int v = it / LWp, lc = it - v * LWp;
        const int p = T.cbeg + lc;
        if (lc + 2 >= P.LW || p < 0 || p > P.w - 3) continue;
        float dsum = 0.0f;
        for (int ch = 0; ch < 3; ++ch) {
            const float* im = plane(P.img, P.img_bs, P.img_cs, T.b, v * 3 + ch);
            float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
            for (int dy = 0; dy < 3; ++dy) {
                const float* rrow = S.recon +
                    ((size_t)mod3(q + dy) * 6 + v * 3 + ch) * P.LW + lc;
                const float* irow = im + (long long)(q + dy) * P.w + p;
                for (int dx = 0; dx < 3; ++dx) {
                    const float x = USL_LDG(irow + dx), y = rrow[dx];
                    sx += x; sy += y;
                    sxx = fmaf(x, x, sxx); syy = fmaf(y, y, syy);
                    sxy = fmaf(x, y, sxy);
                }
            }
            const float mx = sx * inv9, my = sy * inv9;
            const float vx = sxx * inv9 - mx * mx, vy = syy * inv9 - my * my;
            const float vxy = sxy * inv9 - mx * my;
            const float n1 = 2.0f * mx * my + P.c1, n2 = 2.0f * vxy + P.c2;
            const float d1 = mx * mx + my * my + P.c1, d2 = vx + vy + P.c2;
            const float inv = USL_DIV(1.0f, d1 * d2);
            const float ssim = n1 * n2 * inv;
            const float raw = (1.0f - ssim) * 0.5f;
            dsum += fminf(fmaxf(raw, 0.0f), 1.0f);
            if (BWD) {
                // d reproj / d dssim(q) = coef * alpha/3 * T(q); clamp passes
                // the gradient on the closed interval; d dssim/d ssim = -1/2.
                const float pass = (raw >= 0.0f && raw <= 1.0f) ? 1.0f : 0.0f;
                const float gb = -0.5f * pass * gd_up * P.coef[ACC_REPROJ] *
                                 P.alpha * (1.0f / 3.0f) * ty * S.tx[lc];
                const float gA = 2.0f * mx * (n2 - n1) * inv -
                                 2.0f * my * n1 * n2 * (d2 - d1) * inv * inv;
                const float gQ = -USL_DIV(n1 * n2 * inv, d2);
                const float gC = 2.0f * n1 * inv;
                float* G = S.G + (((size_t)mod3(q) * 2 + v) * 3 + ch) * 3 * P.LW + lc;
                G[0] = gb * gA;
                G[P.LW] = gb * gQ;
                G[2 * P.LW] = gb * gC;
            }
        }
        S.ds[((size_t)mod4(q) * 2 + v) * P.LW + lc] = dsum;
    }
}

USL_HD float unc_loss(int type, float u, float e) {
    if (type == LOSS_L1) return fabsf(u - e);
    if (type == LOSS_BAYESIAN) return USL_DIV(e, u) + USL_LOG(u);
    return e * USL_EXP(u) + u;            // the /2 is folded into coef
}

USL_HD float unc_loss_grad(int type, float u, float e) {
    if (type == LOSS_L1) return sgnf(u - e);
    if (type == LOSS_BAYESIAN) { const float iu = USL_DIV(1.0f, u); return iu - e * iu * iu; }
    return e * USL_EXP(u) + 1.0f;
}

// -------------------------------------------------------------------------
// Phase D: per-pixel error E and everything that depends on it.
// -------------------------------------------------------------------------
template <bool BWD>
USL_HD void phase_D(const LossParams& P, const Tile& T, const Rings& S, int r,
                    int tid, int nt, int LWp, float* acc, float gd_up,
                    float ge_up) {
    const bool reproj = (P.terms & TERM_REPROJ) != 0;
    const int y2 = r - 2;            // bwd: SSIM/L1 gradient row
    const int y3 = r - LAG_E;        // E row
    const bool do2 = BWD && reproj && !P.err_given && y2 >= T.ya && y2 < T.yb;
    const bool do3 = y3 >= T.ya && y3 < T.yb;
    if (!do2 && !do3) return;
    const float sH = ac_scale(P.h - 2, P.h), sW = ac_scale(P.w - 2, P.w);
    TapAC ay; ay.i0 = ay.i1 = 0; ay.w0 = ay.w1 = 0.f;
    if (do3) ay = ac_taps(y3, sH, P.h - 2);
    const long long hw = (long long)P.h * P.w;
    for (int it = tid; it < 2 * LWp; it += nt) {
        const int v = it / LWp, lc = it - v * LWp;
        const int c = T.cbeg + lc;
        if (lc >= P.LW || c < T.xa || c >= T.xb) continue;
        const float sign = v ? 1.0f : -1.0f;
        if (do2) {
            const long long pix = (long long)y2 * P.w + c;
            float gs = 0.0f;
            for (int ch = 0; ch < 3; ++ch) {
                float sA = 0.f, sQ = 0.f, sC = 0.f;
                for (int qy = y2 - 2; qy <= y2; ++qy) {
                    if (qy < 0 || qy > P.h - 3) continue;
                    const float* G = S.G + (((size_t)mod3(qy) * 2 + v) * 3 + ch) * 3 * P.LW;
                    for (int px = c - 2; px <= c; ++px) {
                        if (px < 0 || px > P.w - 3) continue;
                        const int l = px - T.cbeg;
                        sA += G[l]; sQ += G[P.LW + l]; sC += G[2 * P.LW + l];
                    }
                }
                const float im = USL_LDG(plane(P.img, P.img_bs, P.img_cs, T.b,
                                               v * 3 + ch) + pix);
                const float rc = S.recon[((size_t)mod3(y2) * 6 + v * 3 + ch) * P.LW + lc];
                float g = (sA + 2.0f * rc * sQ + im * sC) * (1.0f / 9.0f);
                g -= gd_up * P.coef[ACC_REPROJ] * (1.0f - P.alpha) *
                     (1.0f / 3.0f) * sgnf(im - rc);
                const long long o = ((long long)T.b * 6 + v * 3 + ch) * hw + pix;
                if (P.grad_recon_in) g += USL_LDG(P.grad_recon_in + o);
                if (P.recon_given) P.grad_recon_out[o] = g;
                else gs += g * S.dIds[((size_t)mod3(y2) * 6 + v * 3 + ch) * P.LW + lc];
            }
            if (!P.recon_given)
                S.grad[((size_t)mod4(y2) * 4 + v) * P.LW + lc] += sign * gs;
        }
        if (!do3) continue;
        const long long pix = (long long)y3 * P.w + c;
        float e = 0.0f;
        if (P.err_given) {
            e = USL_LDG(plane(P.err_in, P.ei_bs, P.ei_cs, T.b, v) + pix);
        } else if (reproj) {
            const TapAC ax = ac_taps(c, sW, P.w - 2);
            const float* d0 = S.ds + ((size_t)mod4(ay.i0) * 2 + v) * P.LW - T.cbeg;
            const float* d1 = S.ds + ((size_t)mod4(ay.i1) * 2 + v) * P.LW - T.cbeg;
            const float up = ay.w0 * (ax.w0 * d0[ax.i0] + ax.w1 * d0[ax.i1]) +
                             ay.w1 * (ax.w0 * d1[ax.i0] + ax.w1 * d1[ax.i1]);
            e = (P.alpha * up + (1.0f - P.alpha) *
                 S.l1s[((size_t)mod4(y3) * 2 + v) * P.LW + lc]) * (1.0f / 3.0f);
        }
        if (!BWD && reproj) {
            acc[ACC_REPROJ] += e;
            if (P.err_out) P.err_out[((long long)T.b * 2 + v) * hw + pix] = e;
        }
        if (P.terms & TERM_UNC) {
            const float u = USL_LDG(plane(P.unc, P.u_bs, P.u_cs, T.b, v) + pix);
            if (BWD)
                S.grad[((size_t)mod4(y3) * 4 + 2 + v) * P.LW + lc] +=
                    ge_up * P.coef[ACC_UNC] * unc_loss_grad(P.loss_type, u, e);
            else
                acc[ACC_UNC] += unc_loss(P.loss_type, u, e);
        }
        if (BWD) {
            const float gdv = S.grad[((size_t)mod4(y3) * 4 + v) * P.LW + lc];
            const float guv = S.grad[((size_t)mod4(y3) * 4 + 2 + v) * P.LW + lc];
            if (P.grad_disp) {
                float* o = P.grad_disp + (long long)T.b * P.gd_bs + v * P.gd_cs + pix;
                *o = P.grad_disp_accumulate ? (*o + gdv) : gdv;
            }
            if (P.grad_unc)
                P.grad_unc[(long long)T.b * P.gu_bs + v * P.gu_cs + pix] = guv;
        }
    }
}

// Prologue of the backward kernel: column factors of the transposed up-sample.
USL_HD void phase_init_bwd(const LossParams& P, const Tile& T, const Rings& S,
                           int tid, int nt) {
    const float sW = ac_scale(P.w - 2, P.w);
    for (int lc = tid; lc < P.LW; lc += nt) {
        const int p = T.cbeg + lc;
        S.tx[lc] = (p >= 0 && p <= P.w - 3)
                       ? upsample_transpose_weight(p, P.w - 2, P.w, sW) : 0.0f;
    }
}

USL_HD int first_step(const Tile& T) { return T.ya - HALO_T; }
USL_HD int last_step(const Tile& T) { return T.yb - 1 + LAG_E; }

}  // namespace usl
