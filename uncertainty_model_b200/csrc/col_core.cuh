// Column-marching form of the fused loss kernels (forward + one-pass gradient):
// the hot path of the training step.  Same mathematics as loss_core.cuh (see
// its header for the reference citations: utils.py:65-135, loss.py:43-151,
// 167-188,208-264,389-434,539-568), organised around what bounds it on B200.
//
// The path is NOT HBM bound: a pixel of one view costs ~500 issued instructions
// (warp, 3x3 SSIM of three channels forward and backward, two consistency
// warps, smoothness, uncertainty term) against 56 bytes of compulsory traffic,
// and it needs ~120 floats of on-chip state per column (two rows of window-sum
// history forward, two backward, the rows in flight).  Registers + shared
// memory of one SM hold ~512 columns of state, so:
//
//  * one THREAD owns ONE column of ONE view for a whole row strip and marches
//    down it, one image row per step; 512 threads (16 warps) per SM;
//  * everything vertical lives in registers: the separable 3x3 window sums keep
//    the horizontal 3-sums of the two previous rows (forward: {x, y, x^2+y^2,
//    xy} per channel; backward: the three dSSIM maps per channel);
//  * everything horizontal goes through single shared-memory rows (right
//    neighbours of the reconstruction, left neighbours of the dSSIM maps),
//    conflict free.  All rows of the CTA share ONE compile-time row stride and
//    a thread sits at the same position in each of them, so every access is
//    [thread base + immediate];
//  * the opposite view's vertically blended row V(r) -- {R,G,B,disparity} as one
//    16-byte element per column, zero padded -- is the only gathered input;
//    coordinates are clamped into the padding instead of being range tested;
//  * the unit of work -- one CTA -- is (sample, row strip, view set, column
//    tile): a full-width row of one view (w = 512), of both views (w <= 256), or
//    a column tile with a 2-column halo on each side (w > 512).  Every pyramid
//    scale is its own launch (its own block size and row stride), so the
//    geometry and the configuration are CTA-uniform kernel parameters;
//
// Step r of a strip [ya, yb)  (row r enters; results trail by two rows):
//   P1  warp of row r from V(r): recon y(r), d(recon)/d(shift), |x - y| and its
//       gradient; both consistency terms; vertical smoothness edge (r-1, r)
//   -- barrier --
//   P2  horizontal smoothness edge (c, c+1); horizontal 3-sums of row r; SSIM of
//       window row q = r-2 -> dssim(q) and the maps
//       G(q) = dSSIM/d{mean_y, E[y^2], E[xy]}
//   -- barrier --
//   P3  3x3 box of G -> d(loss)/d(recon) of row r-2, through the warp; error
//       map row r-2 (up-sampled dssim + L1) -> reprojection and uncertainty
//       terms; gradient row r-2 written; V(r+1) produced
//   -- barrier --
//
// Every function is host+device: tests/emu runs the phases on the CPU with one
// CState per emulated thread (same rings, same step order).
#pragma once

#include <stdint.h>
#include <string.h>

#include "loss_core.cuh"

namespace usl {
namespace ck {

#if defined(__CUDA_ARCH__)
USL_HD unsigned f2u(float f) { return __float_as_uint(f); }
USL_HD float u2f(unsigned u) { return __uint_as_float(u); }
USL_HD float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
USL_HD float fast_exp2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
USL_HD float fast_log2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
USL_HD F4 ld_f4(const F4* p) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    F4 r; r.x = t.x; r.y = t.y; r.z = t.z; r.w = t.w;
    return r;
}
USL_HD void st_f4(F4* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
USL_HD void st_f2(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
#else
USL_HD unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
USL_HD float u2f(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
USL_HD float fast_rcp(float x) { return 1.0f / x; }
USL_HD float fast_exp2(float x) { return exp2f(x); }
USL_HD float fast_log2(float x) { return log2f(x); }
USL_HD F4 ld_f4(const F4* p) { return *p; }
USL_HD void st_f4(F4* p, float a, float b, float c, float d) {
    p->x = a; p->y = b; p->z = c; p->w = d;
}
USL_HD void st_f2(float* p, float a, float b) { p[0] = a; p[1] = b; }
#endif

// s * sgn(v)   (torch: d|v|/dv = sign(v), 0 at 0): sign-bit transfer + zero test
USL_HD float sgn_mul(float v, float s) {
    const float r = u2f(f2u(s) ^ (f2u(v) & 0x80000000u));
    return v == 0.0f ? 0.0f : r;
}

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr int VPAD = 2;         // zero columns on each side of a V row

// ---- async row ring (TMA bulk copies + mbarriers) ---------------------------
#if defined(__CUDA_ARCH__)
USL_HD uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
USL_HD void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
USL_HD void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders the generic-proxy writes of this thread before later async-proxy
// (bulk copy) accesses to shared memory
USL_HD void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
USL_HD void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
USL_HD void bulk_g2s(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
USL_HD void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "USL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra USL_DONE;\n"
        "bra USL_WAIT;\n"
        "USL_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
#else
USL_HD uint32_t smem_u32(const void*) { return 0; }
USL_HD void mbar_init(uint64_t*, int) {}
USL_HD void mbar_fence_init() {}
USL_HD void fence_proxy_async() {}
USL_HD void mbar_expect_tx(uint64_t*, uint32_t) {}
USL_HD void bulk_g2s(float* dst, const float* src, uint32_t bytes, uint64_t*) {
    memcpy(dst, src, bytes);
}
USL_HD void mbar_wait(uint32_t, uint32_t) {}
#endif

// ---- shared-memory layout -----------------------------------------------------
// One arena per CTA: `SROW` floats per row (compile time); a view's segment of
// a row is [4 pad | LW columns | 4 pad] (16-byte aligned for the bulk copies),
// the two views side by side.  A thread sits at the same position `sb` in every
// row, so each access is [sb + imm].
constexpr int SEG_PAD = 4;
constexpr int NHIST_GRAD = 11;  // dI[3], y*dI[3], x*dI[3], l1, u
constexpr int NHIST_FWD = 2;    // l1, u
constexpr int NSLOT = 4;        // input ring depth (rows r, r+1, r+2 | r+3 in flight)
constexpr int NPL = 5;          // planes per ring slot: x[3], d, u
constexpr int NPO = 4;          // ... of the opposite view's ring: x[3], d
enum {
    ROW_Y = 0,      // [3]  recon of the current row        (right neighbours read)
    ROW_DS = 3,     // [4]  ring of channel-summed dssim rows
    ROW_IN = 7,     // [NSLOT][NPL] own view(s): image, disparity, uncertainty rows
    ROW_OPP = 27,   // [NSLOT][NPO] opposite view (one-view units): image, disparity
    ROW_HS = 43,    // [3][NH] thread-private history of the rows in flight
};
USL_HD constexpr int row_gx(bool grad) {     // [9] G(q) of the current window row
    return ROW_HS + 3 * (grad ? NHIST_GRAD : NHIST_FWD);
}
USL_HD constexpr int row_ed(bool grad) { return row_gx(grad) + 9; }   // [2]
USL_HD constexpr int n_rows(bool grad) { return grad ? row_ed(true) + 2 : row_gx(false); }

enum { MODE_PLAIN = 0, MODE_MASKED = 1, MODE_TILED = 2 };
// MODE_PLAIN : full-row unit, every thread owns a column, inputs arrive by bulk
//              async copies (needs w % 4 == 0 and 16-byte aligned planes)
// MODE_MASKED: full-row unit of any width / alignment; the threads fill the ring
// MODE_TILED : column tile with halos; V(r) swept straight from global memory

struct CGeo {                // per-CTA constants
    int b;                   // sample
    int v0, nv;              // first view, number of views (1 or 2)
    int xa, xb;              // owned columns
    int cbeg, LW;            // first column incl. halo; columns incl. halo
    int ya, yb;              // owned rows
    int qlo;                 // first window row this strip forms
    float sH, sW;            // (h-2 -> h), (w-2 -> w) align-corners scales
    float gd_up, ge_up;      // upstream gradients (GRAD)
};

// per-step tables (entry i: step r = ya - 3 + i; entry 0 only serves the
// prologue's V(ya - 2)), see c_init_unit
struct alignas(16) RowT {
    float tyw;               // transposed up-sample row weight of q = r - 2
    float ay_w1;             // up-sample (h-2 -> h) weight of tap i1, y = r - 2
    int ay_o0, ay_o1;        // float offsets (from sb) of the dssim ring rows i0, i1
};
struct alignas(16) RowU {
    int in_o;                // float offset of the input ring slot of row r
    int hs_w, hs_r;          // ... of the history slots of rows r (write), r-2 (read)
    int ds_w;                // ... of the dssim ring row of q = r - 2
};
struct alignas(16) RowV {    // vertical taps of the warp for row r + 1
    int o0, o1;              // float offsets of the ring slots of the source rows
    float w0, w1;            // weights (zero for rows outside the image)
};
struct alignas(16) RowW {    // ring synchronisation of the step
    unsigned wait_cur;       // (unused: the blend touches every row first)
    unsigned wait_v0, wait_v1;   // wait words (c_wait_word) of the source rows of
                             // V(r + 1) that no earlier step waited for (0: none)
    int issue;               // row to request at the start of the step (-1: none)
};

struct CRings {
    float* rows;    // [n_rows][SROW]
    F4* V;          // [nv][w + 2*VPAD]
    RowT* RT;       // [R + 8] each
    RowU* RU;
    RowV* RV;
    RowW* RW;
    uint64_t* mbar; // [NSLOT]
    uint32_t mbar_a;      // shared-window address of mbar[0]
    const float** isrc;   // [MAX_ISSUE] row 0 of every plane the ring holds
    int* idst;            // [MAX_ISSUE] float offset of its row inside a slot
    int* istr;            // [MAX_ISSUE] floats from one slot to the next
};
constexpr int MAX_ISSUE = 10;

USL_HD int c_first_row(const CGeo& G) { return G.ya - 2; }
USL_HD int c_last_row(const CGeo& G) { return G.yb + 1; }
USL_HD int c_step_index(const CGeo& G, int r) { return r - (G.ya - 3); }
USL_HD int c_ring_slot(const CGeo& G, int row) { return (row - (G.ya - 3)) % NSLOT; }
USL_HD int c_ring_parity(const CGeo& G, int row) { return ((row - (G.ya - 3)) / NSLOT) & 1; }

// floats of shared memory for one CTA
USL_HD size_t c_floats(int srow, int w, int nv, int R, bool grad) {
    size_t n = (size_t)n_rows(grad) * srow;
    n += (size_t)nv * (w + 2 * VPAD) * 4;
    n += (size_t)(R + 8) * 16;
    n += 8 + 2 * MAX_ISSUE + 12 + 12;
    return (n + 3) & ~(size_t)3;
}

USL_HD CRings c_carve(float* base, int srow, int w, int nv, int R, bool grad) {
    CRings S;
    S.V = reinterpret_cast<F4*>(base); base += (size_t)nv * (w + 2 * VPAD) * 4;
    S.RT = reinterpret_cast<RowT*>(base); base += (size_t)(R + 8) * 4;
    S.RU = reinterpret_cast<RowU*>(base); base += (size_t)(R + 8) * 4;
    S.RV = reinterpret_cast<RowV*>(base); base += (size_t)(R + 8) * 4;
    S.RW = reinterpret_cast<RowW*>(base); base += (size_t)(R + 8) * 4;
    S.mbar = reinterpret_cast<uint64_t*>(base); base += 8;
    S.mbar_a = smem_u32(S.mbar);
    S.isrc = reinterpret_cast<const float**>(base); base += 2 * MAX_ISSUE;
    S.idst = reinterpret_cast<int*>(base); base += 12;
    S.istr = reinterpret_cast<int*>(base); base += 12;
    S.rows = base;
    return S;
}

struct CState {
    float x[3], d, u;        // own inputs, row r
    float xp[3], dp, up;     // row r-1
    float y[3];              // recon of row r                    (P1 -> P2)
    float g[9];              // own G(q = r-2)                    (P2 -> P3)
    float H[2][3][4];        // horizontal 3-sums of the two previous rows
    float HG[2][3][3];       // ... of G of the two previous window rows
    float gd[3], gu[3];      // gradient accumulators of rows r, r-1, r-2
    float acc[NUM_ACC];
    // per-thread constants
    float xbase;             // linspace(0,1,w) at c
    float txw;               // transposed up-sample column weight (GRAD)
    float axw;               // up-sample column weight of tap i1
    float own;               // 1 where the column belongs to the unit
    float sign;              // -1 left view, +1 right view
    float* sb;               // the thread's position in row 0 of the arena
    const float* ob;         // the same column of the opposite view's ring rows
    const F4* vrow0;         // column 0 of the own view's V row
    int ax0, ax1;            // column offsets (from the own column) of the up-sample taps
    unsigned o_img, o_d, o_u;   // element offsets of (b, own view, row 0, c)
    unsigned o_oi, o_od;        // ... of (b, opposite view, row 0, c): img, disp
    unsigned o_gd, o_gu;        // element offsets of the gradient outputs: (b, own view,
                                // next row to be finalised, c)
    unsigned o_sc;              // element offset of (b, own view, row 0, c) in P.scat
    int v, vi, c, lc;
    bool active, win_ok, has1;
};

USL_HD float c_warp_coord(float xbase, float shift, float half_n) {
    const float x = xbase + shift;
    const float g = fmaf(2.0f, x, -1.0f);
    return fmaf(g + 1.0f, half_n, -0.5f);
}

// ---- CTA prologue -----------------------------------------------------------
template <int SROW, bool GRAD>
USL_HD void c_thread_init(const LossParams& P, const CGeo& G, const CRings& S,
                          int tid, CState& T) {
    const int w = P.w;
    T.vi = tid / G.LW;
    T.lc = tid - T.vi * G.LW;
    T.active = tid < G.nv * G.LW;
    if (!T.active) { T.vi = 0; T.lc = 0; }
    T.v = G.v0 + T.vi;
    T.c = G.cbeg + T.lc;
    T.sign = T.v ? 1.0f : -1.0f;
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 3; ++c) {
            for (int q = 0; q < 4; ++q) T.H[a][c][q] = 0.f;
            for (int m = 0; m < 3; ++m) T.HG[a][c][m] = 0.f;
        }
    for (int a = 0; a < 3; ++a) T.gd[a] = T.gu[a] = 0.f;
    for (int k = 0; k < NUM_ACC; ++k) T.acc[k] = 0.f;
    for (int c = 0; c < 3; ++c) T.x[c] = T.xp[c] = T.y[c] = 0.f;
    for (int k = 0; k < 9; ++k) T.g[k] = 0.f;
    T.d = T.dp = 0.f;
    T.u = T.up = 1.f;
    T.xbase = linspace01(T.c, w);
    T.own = (T.active && T.c >= G.xa && T.c < G.xb) ? 1.f : 0.f;
    T.win_ok = T.c <= w - 3;
    T.has1 = T.c + 1 < w;
    const int seg = G.LW + 2 * SEG_PAD;
    T.sb = S.rows + T.vi * seg + SEG_PAD + T.lc;
    // two-view units: the opposite view is the other segment of the input ring
    T.ob = G.nv == 2 ? S.rows + (1 - T.vi) * seg + SEG_PAD + T.lc + ROW_IN * SROW
                     : T.sb + ROW_OPP * SROW;
    T.vrow0 = S.V + T.vi * (w + 2 * VPAD) + VPAD;
    T.o_img = (unsigned)((long long)G.b * P.img_bs + (long long)T.v * 3 * P.img_cs + T.c);
    T.o_oi = (unsigned)((long long)G.b * P.img_bs + (long long)(1 - T.v) * 3 * P.img_cs + T.c);
    T.o_d = (unsigned)((long long)G.b * P.d_bs + (long long)T.v * P.d_cs + T.c);
    T.o_od = (unsigned)((long long)G.b * P.d_bs + (long long)(1 - T.v) * P.d_cs + T.c);
    T.o_u = (unsigned)((long long)G.b * P.u_bs + (long long)T.v * P.u_cs + T.c);
    T.o_gd = (unsigned)((long long)G.b * P.gd_bs + (long long)T.v * P.gd_cs +
                        (long long)G.ya * w + T.c);
    T.o_gu = (unsigned)((long long)G.b * P.gu_bs + (long long)T.v * P.gu_cs +
                        (long long)G.ya * w + T.c);
    T.o_sc = (unsigned)(((long long)G.b * 2 + T.v) * ((long long)P.h * w) + T.c);
    T.txw = 0.f; T.axw = 0.f; T.ax0 = T.ax1 = 0;
    if (!T.active) return;
    const TapAC ax = ac_taps(T.c, G.sW, w - 2);
    // a zero-weight tap may lie outside what the unit forms: point it at i0
    const int i1 = ax.w1 != 0.f ? ax.i1 : ax.i0;
    T.ax0 = ax.i0 - T.c;
    T.ax1 = i1 - T.c;
    T.axw = ax.w1;
    if (GRAD && T.c <= w - 3)
        T.txw = upsample_transpose_weight(T.c, w - 2, w, G.sW);
}

// `w`: mbarrier shared-window address | parity << 31 (see c_wait_word), < 0 as
// int only when the parity bit is set, so "no wait" is the value 0.
template <bool SURE>
USL_HD void c_ring_wait(unsigned w) {
    if (SURE || w != 0u) mbar_wait(w & 0x7fffffffu, w >> 31);
}
USL_HD unsigned c_wait_word(const CRings& S, int slot, int parity) {
    return (S.mbar_a + 8u * (unsigned)slot) | ((unsigned)parity << 31) | 0u;
}

// last image row the unit ever reads (own rows to yb+1, warp sources one more)
USL_HD int c_last_ring_row(const LossParams& P, const CGeo& G) {
    const int r = G.yb + 2;
    return r < P.h - 1 ? r : P.h - 1;
}

// Zeroes the exchange rows (the pads must be zero, the rest is overwritten
// before it is read) and the V pads; fills the per-step tables; arms the ring.
template <int SROW, bool GRAD, int MODE>
USL_HD void c_init_unit(const LossParams& P, const CGeo& G, const CRings& S,
                        int tid, int nt) {
    const int R = P.R;
    const int nh = GRAD ? NHIST_GRAD : NHIST_FWD;
    const int last = c_last_ring_row(P, G);
    for (int i = tid; i < R + 8; i += nt) {
        const int r = G.ya - 3 + i;         // step
        const int y = r - 2;                // row finalised / window row
        RowT t;
        t.tyw = (GRAD && y >= 0 && y <= P.h - 3)
                    ? upsample_transpose_weight(y, P.h - 2, P.h, G.sH) : 0.0f;
        t.ay_w1 = 0.f; t.ay_o0 = t.ay_o1 = ROW_DS * SROW;
        if (y >= 0 && y < P.h) {
            const TapAC ay = ac_taps(y, G.sH, P.h - 2);
            const int i1 = ay.w1 != 0.f ? ay.i1 : ay.i0;
            t.ay_w1 = ay.w1;
            t.ay_o0 = (ROW_DS + (ay.i0 & 3)) * SROW;
            t.ay_o1 = (ROW_DS + (i1 & 3)) * SROW;
        }
        S.RT[i] = t;
        RowU u;
        u.in_o = (ROW_IN + c_ring_slot(G, r) * NPL) * SROW;
        u.hs_w = (ROW_HS + mod3(r) * nh) * SROW;
        u.hs_r = (ROW_HS + mod3(y) * nh) * SROW;
        u.ds_w = (ROW_DS + mod4(y)) * SROW;
        S.RU[i] = u;
        // V(r + 1)
        RowV v;
        RowW ww;
        v.o0 = v.o1 = 0; v.w0 = v.w1 = 0.f;
        ww.wait_cur = ww.wait_v0 = ww.wait_v1 = 0u; ww.issue = -1;
        const int rn = r + 1;
        if (rn >= 0 && rn < P.h) {
            const Tap2 ty = warp_row_taps(rn, P.h);
            const bool ok0 = ty.i0 >= 0 && ty.i0 < P.h;
            const bool ok1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < P.h;
            v.w0 = ok0 ? ty.w0 : 0.0f;
            v.w1 = ok1 ? ty.w1 : 0.0f;
            const int i0 = ok0 ? ty.i0 : ty.i0 + 1;
            const int i1 = ok1 ? ty.i0 + 1 : ty.i0;
            // (the opposite view: the other segment of the own ring, or its own ring)
            const int per_slot = (G.nv == 2 ? NPL : NPO) * SROW;
            v.o0 = c_ring_slot(G, i0) * per_slot;
            v.o1 = c_ring_slot(G, i1) * per_slot;
            // A row is waited for where a thread first touches it, and never
            // again (an mbarrier wait costs ~90 cycles even when the row is
            // there): rows are touched in increasing order, the blend V(r + 1)
            // touches row r + 1 a phase before P1(r + 1) does, so it is only
            // the rows V(r + 1) uses that V(r) -- entry i - 1 -- did not.
            int p0 = -1, p1 = -1;
            if (i >= 1 && rn - 1 >= 0) {
                const Tap2 tp = warp_row_taps(rn - 1, P.h);
                const bool pk0 = tp.i0 >= 0 && tp.i0 < P.h;
                const bool pk1 = tp.i0 + 1 >= 0 && tp.i0 + 1 < P.h;
                p0 = pk0 ? tp.i0 : tp.i0 + 1;
                p1 = pk1 ? tp.i0 + 1 : tp.i0;
            }
            if (i0 != p0 && i0 != p1)
                ww.wait_v0 = c_wait_word(S, c_ring_slot(G, i0), c_ring_parity(G, i0));
            if (i1 != i0 && i1 != p0 && i1 != p1)
                ww.wait_v1 = c_wait_word(S, c_ring_slot(G, i1), c_ring_parity(G, i1));
        }
        S.RV[i] = v;
        if (r + 2 >= G.ya && r + 2 <= last) ww.issue = r + 2;
        S.RW[i] = ww;
    }
    for (int i = tid; i < G.nv * 2 * VPAD; i += nt) {
        const int vi = i / (2 * VPAD), k = i - vi * 2 * VPAD;
        const int col = k < VPAD ? k : P.w + k;
        st_f4(S.V + (size_t)vi * (P.w + 2 * VPAD) + col, 0.f, 0.f, 0.f, 0.f);
    }
    // the pads of every row ([4 | LW | 4] per view segment) must read as zero;
    // everything between them is written before it is read
    const int npad = G.nv * 2 * SEG_PAD;
    for (int i = tid; i < n_rows(GRAD) * npad; i += nt) {
        const int row = i / npad, k = i - row * npad;
        const int vi = k / (2 * SEG_PAD), j = k - vi * 2 * SEG_PAD;
        const int col = vi * (G.LW + 2 * SEG_PAD) + (j < SEG_PAD ? j : G.LW + j);
        S.rows[(size_t)row * SROW + col] = 0.f;
    }
    // (generic stores next to the rows the bulk copies will fill)
    if (MODE == MODE_PLAIN) fence_proxy_async();
    if (MODE == MODE_PLAIN && tid == 0) {
        for (int s = 0; s < NSLOT; ++s) mbar_init(S.mbar + s, 1);
        mbar_fence_init();
        // the planes of one ring slot: where row 0 of each starts in global
        // memory, and where its row lands inside the slot
        const int seg = G.LW + 2 * SEG_PAD;
        int n = 0;
        for (int vi = 0; vi < G.nv; ++vi) {
            const int v = G.v0 + vi;
            const int dst = ROW_IN * SROW + vi * seg + SEG_PAD;
            for (int k = 0; k < 3; ++k) {
                S.isrc[n] = plane(P.img, P.img_bs, P.img_cs, G.b, v * 3 + k);
                S.istr[n] = NPL * SROW;
                S.idst[n++] = dst + k * SROW;
            }
            S.isrc[n] = plane(P.disp, P.d_bs, P.d_cs, G.b, v);
            S.istr[n] = NPL * SROW;
            S.idst[n++] = dst + 3 * SROW;
            if (P.unc) {
                S.isrc[n] = plane(P.unc, P.u_bs, P.u_cs, G.b, v);
                S.istr[n] = NPL * SROW;
                S.idst[n++] = dst + 4 * SROW;
            }
        }
        if (G.nv == 1) {
            const int v = 1 - G.v0;
            const int dst = ROW_OPP * SROW + SEG_PAD;
            for (int k = 0; k < 3; ++k) {
                S.isrc[n] = plane(P.img, P.img_bs, P.img_cs, G.b, v * 3 + k);
                S.istr[n] = NPO * SROW;
                S.idst[n++] = dst + k * SROW;
            }
            S.isrc[n] = plane(P.disp, P.d_bs, P.d_cs, G.b, v);
            S.istr[n] = NPO * SROW;
            S.idst[n++] = dst + 3 * SROW;
        }
        for (; n < MAX_ISSUE; ++n) { S.isrc[n] = nullptr; S.idst[n] = 0; S.istr[n] = 0; }
    }
}

// ---- the input ring -----------------------------------------------------------
// Row `row` of the unit's views -> ring slot: image planes, disparity,
// uncertainty of the own view(s); image planes and disparity of the opposite
// view for one-view units.  MODE_PLAIN: bulk async copies that complete on the
// slot's mbarrier.
USL_HD int c_ring_planes(const LossParams& P, const CGeo& G) {
    const int n = G.nv == 2 ? 2 * NPL : NPL + 4;
    return P.unc ? n : n - G.nv;
}

// Called by every lane of ONE warp: lane 0 arms the slot's mbarrier with the
// byte count, then lane k requests plane k -- one bulk copy per lane, so a
// whole row set costs the warp a dozen instructions.
USL_HD void c_ring_issue(const LossParams& P, const CGeo& G, const CRings& S,
                         int srow, int row, int lane) {
    if (row >= P.h) return;
    const int slot = c_ring_slot(G, row);
    uint64_t* bar = S.mbar + slot;
    const int n = c_ring_planes(P, G);
    const unsigned bytes = (unsigned)G.LW * 4u;
    // a row above the image is never waited for, but its phase must still
    // complete so that the slot's parity keeps counting uses
    if (lane == 0) mbar_expect_tx(bar, row < 0 ? 0u : bytes * (unsigned)n);
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
    if (row < 0 || lane >= n) return;
    bulk_g2s(S.rows + S.idst[lane] + (size_t)slot * S.istr[lane],
             S.isrc[lane] + (long long)row * P.w, bytes, bar);
}

// MODE_MASKED / MODE_TILED: every thread moves its own column -- with
// asynchronous 4-byte copies (cp.async: no register, nothing waits here); the
// step waits for them (c_fill_wait) right before the block barrier that
// precedes the first read of the row, one P1 later.
#if defined(__CUDA_ARCH__)
USL_HD void cp_f32(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src)
                 : "memory");
}
USL_HD void c_fill_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
USL_HD void cp_f32(float* dst, const float* src) { *dst = *src; }
USL_HD void c_fill_wait() {}
#endif

template <int SROW, int MODE>
USL_HD void c_ring_fill(const LossParams& P, const CGeo& G, const CState& T, int row) {
    if (!T.active || row < 0 || row >= P.h) return;
    const int slot = c_ring_slot(G, row);
    const unsigned ro = (unsigned)(row * P.w);
    float* dst = T.sb + (ROW_IN + slot * NPL) * SROW;
    const float* im = P.img + (T.o_img + ro);
    cp_f32(dst, im);
    cp_f32(dst + SROW, im + P.img_cs);
    cp_f32(dst + 2 * SROW, im + 2 * P.img_cs);
    cp_f32(dst + 3 * SROW, P.disp + (T.o_d + ro));
    if (P.unc) cp_f32(dst + 4 * SROW, P.unc + (T.o_u + ro));
    if (MODE != MODE_TILED && G.nv == 1) {
        float* od = T.sb + (ROW_OPP + slot * NPO) * SROW;
        const float* oi = P.img + (T.o_oi + ro);
        cp_f32(od, oi);
        cp_f32(od + SROW, oi + P.img_cs);
        cp_f32(od + 2 * SROW, oi + 2 * P.img_cs);
        cp_f32(od + 3 * SROW, P.disp + (T.o_od + ro));
    }
}

// `sure`: the row is known to exist (no test of the table entry)
// ---- V(r+1): vertical blend of the opposite view, whole row -------------------
// Full-row units: thread (vi, lc) blends its own column from the ring.
// Column-tiled units: the threads of the unit sweep the whole row from global.
template <int SROW, int MODE, bool STEADY>
USL_HD void c_pV(const LossParams& P, const CGeo& G, const CRings& S, int r,
                 int tid, int nt, const CState& T) {
    const int rn = r + 1;
    if (!STEADY && (rn < 0 || rn >= P.h)) return;
    const int i = c_step_index(G, r);
    const RowV t = S.RV[i];
    if (MODE != MODE_TILED) {
        if (MODE == MODE_PLAIN) {
            const RowW ww = S.RW[i];
            c_ring_wait<false>(ww.wait_v0);
            c_ring_wait<false>(ww.wait_v1);
        }
        if (MODE == MODE_MASKED && !T.active) return;
        const float* a = T.ob + t.o0;
        const float* b = T.ob + t.o1;
        st_f4(const_cast<F4*>(T.vrow0) + T.c,
              t.w0 * a[0] + t.w1 * b[0],
              t.w0 * a[SROW] + t.w1 * b[SROW],
              t.w0 * a[2 * SROW] + t.w1 * b[2 * SROW],
              t.w0 * a[3 * SROW] + t.w1 * b[3 * SROW]);
    } else {
        const Tap2 ty = warp_row_taps(rn, P.h);
        const bool ok0 = ty.i0 >= 0 && ty.i0 < P.h;
        const bool ok1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < P.h;
        const unsigned r0 = (unsigned)((ok0 ? ty.i0 : ty.i0 + 1) * P.w);
        const unsigned r1 = (unsigned)((ok1 ? ty.i0 + 1 : ty.i0) * P.w);
        // (column-tiled units hold one view: col_plan)
        const int v = G.v0;
        const float* im0 = P.img + (unsigned)((long long)G.b * P.img_bs +
                                              (long long)(1 - v) * 3 * P.img_cs);
        const float* pd0 = P.disp + (unsigned)((long long)G.b * P.d_bs +
                                               (long long)(1 - v) * P.d_cs);
        for (int x = tid; x < P.w; x += nt) {
            const float* im = im0 + x;
            const float* pd = pd0 + x;
            st_f4(S.V + VPAD + x,
                  t.w0 * USL_LDG(im + r0) + t.w1 * USL_LDG(im + r1),
                  t.w0 * USL_LDG(im + P.img_cs + r0) + t.w1 * USL_LDG(im + P.img_cs + r1),
                  t.w0 * USL_LDG(im + 2 * P.img_cs + r0) + t.w1 * USL_LDG(im + 2 * P.img_cs + r1),
                  t.w0 * USL_LDG(pd + r0) + t.w1 * USL_LDG(pd + r1));
        }
#if defined(__CUDA_ARCH__)
        // the row the next step adds to this blend: ask for it now (one
        // request per 32-byte sector), so that its loads hit L1 then
        if (ty.i0 + 2 < P.h) {
            const unsigned rn2 = r1 + (unsigned)P.w;
            for (int x = tid * 8; x < P.w; x += nt * 8) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(im0 + x + rn2));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(im0 + P.img_cs + x + rn2));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(im0 + 2 * P.img_cs + x + rn2));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pd0 + x + rn2));
            }
        }
#endif
    }
}

USL_HD float edge_w3(const float* a, const float* b) {
    const float g = fabsf(a[0] - b[0]) + fabsf(a[1] - b[1]) + fabsf(a[2] - b[2]);
    return fast_exp2(g * (-LOG2E / 3.0f));
}

// Compile-time knowledge of a step.  STEADY: an interior step of the strip --
// rows r-2 .. r+2 exist, r-2 .. r are owned, every window row involved is
// formed here -- so all row tests fold away.  TERMS >= 0: the term mask is a
// compile-time constant.
template <int SROW_, bool GRAD_, int MODE_, int TERMS_, bool STEADY_>
struct Cfg {
    static constexpr int SROW = SROW_;
    static constexpr bool GRAD = GRAD_;
    static constexpr int MODE = MODE_;
    static constexpr bool MASKED = MODE_ != MODE_PLAIN;
    static constexpr int TERMS = TERMS_;
    static constexpr bool STEADY = STEADY_;
};

// ---- P1: warp of row r, row-local terms ---------------------------------------
template <class C>
USL_HD void c_p1(const LossParams& P, const CGeo& G, const CRings& S, int r,
                 CState& T) {
    constexpr int SROW = C::SROW;
    constexpr bool GRAD = C::GRAD, MASKED = C::MASKED, STEADY = C::STEADY;
    if (GRAD) {
        T.gd[2] = T.gd[1]; T.gd[1] = T.gd[0]; T.gd[0] = 0.f;
        T.gu[2] = T.gu[1]; T.gu[1] = T.gu[0]; T.gu[0] = 0.f;
    }
    // the row before becomes "previous"
    for (int c = 0; c < 3; ++c) T.xp[c] = T.x[c];
    T.dp = T.d; T.up = T.u;
    if (!STEADY && (r < 0 || r >= P.h)) return;
    const int i = c_step_index(G, r);
    if (MASKED && !T.active) return;
    const RowU ru = S.RU[i];
    const int nh = GRAD ? NHIST_GRAD : NHIST_FWD;
    const float sign = T.sign, fw = (float)P.w, hw = 0.5f * fw;
    const unsigned terms = C::TERMS >= 0 ? (unsigned)C::TERMS : P.terms;
    const bool own_row = STEADY || (r >= G.ya && r < G.yb);
    const float own = MASKED ? T.own : 1.0f;
    float* hs = T.sb + ru.hs_w;
    {
        const float* in = T.sb + ru.in_o;
        T.x[0] = in[0]; T.x[1] = in[SROW]; T.x[2] = in[2 * SROW];
        T.d = in[3 * SROW];
        if (P.unc) T.u = in[4 * SROW];
    }

    // ---- reconstruction of the row: two taps of V at the shifted column ----
    float wd, dwd = 0.f, dI[3];
    {
        const float ix = c_warp_coord(T.xbase, sign * T.d, hw);
        const float f = floorf(ix);
        const float w1 = ix - f, w0 = (f + 1.0f) - ix;
        const int xi = (int)fminf(fmaxf(f, -2.0f), fw);
        const F4 t0 = ld_f4(T.vrow0 + xi), t1 = ld_f4(T.vrow0 + xi + 1);
        T.y[0] = w0 * t0.x + w1 * t1.x;
        T.y[1] = w0 * t0.y + w1 * t1.y;
        T.y[2] = w0 * t0.z + w1 * t1.z;
        wd = w0 * t0.w + w1 * t1.w;
        if (GRAD) {
            dI[0] = fw * (t1.x - t0.x);
            dI[1] = fw * (t1.y - t0.y);
            dI[2] = fw * (t1.z - t0.z);
            dwd = fw * (t1.w - t0.w);
        }
    }
    float l1 = 0.f, gl1 = 0.f;
    for (int c = 0; c < 3; ++c) {
        const float a = T.x[c] - T.y[c];
        l1 += fabsf(a);
        T.sb[(ROW_Y + c) * SROW] = T.y[c];
        if (GRAD) {
            gl1 += sgn_mul(a, dI[c]);
            hs[c * SROW] = dI[c];
            hs[(3 + c) * SROW] = T.y[c] * dI[c];
            hs[(6 + c) * SROW] = T.x[c] * dI[c];
        }
    }
    hs[(nh - 2) * SROW] = l1;
    hs[(nh - 1) * SROW] = T.u;
    // (optional output, generic variant only: the adversarial step hands the
    //  reconstructions to a discriminator)
    if (C::TERMS < 0 && P.recon_out && own_row && own != 0.f) {
        const long long hwp = (long long)P.h * P.w;
        for (int c = 0; c < 3; ++c)
            P.recon_out[((long long)G.b * 6 + T.v * 3 + c) * hwp +
                        (long long)r * P.w + T.c] = T.y[c];
    }

    if (own_row) {
        float gdr = 0.f, gur = 0.f;
        if (GRAD) {
            // L1 part of d reproj / d recon, through the warp
            const float kl1 = G.gd_up * P.coef[ACC_REPROJ] * (1.0f - P.alpha) *
                              (1.0f / 3.0f);
            gdr = (-sign * kl1) * gl1;
        }
        // ---- consistency terms ----
        // The transposed warp of both terms runs in cons_rows_kernel, lane per
        // row; what it cannot recompute is the sign of (a - warp(b)) -- that
        // needs the blended row -- so the signed coefficients s = coefficient *
        // sign(.) of the pixel are left behind, with the two shifts beside them
        // (both live in registers here: one coalesced 16-byte store per pixel).
        float s_dd = 0.f, s_ud = 0.f;
        if (terms & TERM_CONS_D) {
            const float f = T.d - wd;
            T.acc[ACC_CONS_D] += own * fabsf(f);
            if (GRAD) {
                s_dd = sgn_mul(f, G.gd_up * P.coef[ACC_CONS_D]);
                gdr += s_dd * (1.0f - sign * dwd);
            }
        }
        if (terms & TERM_CONS_U) {
            const float ixu = c_warp_coord(T.xbase, sign * T.u, hw);
            const float f = floorf(ixu);
            const float w1 = ixu - f, w0 = (f + 1.0f) - ixu;
            const int xi = (int)fminf(fmaxf(f, -2.0f), fw);
            const float g0 = T.vrow0[xi].w, g1 = T.vrow0[xi + 1].w;
            const float e = T.u - (w0 * g0 + w1 * g1);
            T.acc[ACC_CONS_U] += own * fabsf(e);
            if (GRAD) {
                s_ud = sgn_mul(e, G.ge_up * P.coef[ACC_CONS_U]);
                gur = s_ud * (1.0f - (sign * fw) * (g1 - g0));
            }
        }
        if (GRAD && P.scat && (!MASKED || own != 0.f))
            st_f4(reinterpret_cast<F4*>(P.scat) + (T.o_sc + (unsigned)(r * P.w)),
                  T.d, T.u, s_dd, s_ud);
        if (GRAD) { T.gd[0] = gdr; T.gu[0] = gur; }
    }

    // ---- smoothness: vertical edge (r-1, r) of column c ----
    if (terms & (TERM_SMOOTH_D | TERM_SMOOTH_U)) {
        const bool prev_own = STEADY || (r - 1 >= G.ya && r - 1 < G.yb);   // => r >= 1
        if (STEADY || ((own_row || prev_own) && r >= 1)) {
            const float wy = edge_w3(T.xp, T.x);
            const float po = prev_own ? own : 0.f;
            if (terms & TERM_SMOOTH_D) {
                const float g = T.dp - T.d;
                T.acc[ACC_SMOOTH_D] += po * (fabsf(g) * wy);
                if (GRAD) {
                    const float s = sgn_mul(g, (G.gd_up * P.coef[ACC_SMOOTH_D]) * wy);
                    if (prev_own) T.gd[1] += s;
                    if (own_row) T.gd[0] -= s;
                }
            }
            if (terms & TERM_SMOOTH_U) {
                const float g = T.up - T.u;
                T.acc[ACC_SMOOTH_U] += po * (fabsf(g) * wy);
                if (GRAD) {
                    const float s = sgn_mul(g, (G.ge_up * P.coef[ACC_SMOOTH_U]) * wy);
                    if (prev_own) T.gu[1] += s;
                    if (own_row) T.gu[0] -= s;
                }
            }
        }
    }
}

// ---- P2: horizontal edges of row r; SSIM of window row q = r - 2 --------------
// PAR = r & 1 (compile time): T.H[PAR] holds the older history row and takes
// the new one.
template <class C, int PAR>
USL_HD void c_p2(const LossParams& P, const CGeo& G, const CRings& S, int r,
                 CState& T) {
    constexpr int SROW = C::SROW;
    constexpr bool GRAD = C::GRAD, MASKED = C::MASKED, STEADY = C::STEADY;
    if (MASKED && !T.active) return;
    const int q = r - 2;
    const int i = c_step_index(G, r);
    const bool row_ok = STEADY || (r >= 0 && r < P.h);
    const bool own_row = STEADY || (r >= G.ya && r < G.yb);
    const float own = MASKED ? T.own : 1.0f;
    const unsigned terms = C::TERMS >= 0 ? (unsigned)C::TERMS : P.terms;
    const RowU ru = S.RU[i];
    const float* in = T.sb + ru.in_o;
    float x1[3], h[3][4];
    if (row_ok) {
        for (int c = 0; c < 3; ++c) {
            const float* yr = T.sb + (ROW_Y + c) * SROW;
            const float* xr = in + c * SROW;
            const float y0 = T.y[c], y1 = yr[1], y2 = yr[2];
            const float x0 = T.x[c], x2 = xr[2];
            x1[c] = xr[1];
            h[c][0] = x0 + (x1[c] + x2);
            h[c][1] = y0 + (y1 + y2);
            const float q0 = fmaf(x0, x0, y0 * y0), q1 = fmaf(x1[c], x1[c], y1 * y1);
            const float q2 = fmaf(x2, x2, y2 * y2);
            h[c][2] = q0 + (q1 + q2);
            h[c][3] = fmaf(x0, y0, fmaf(x1[c], y1, x2 * y2));
        }
    } else {
        for (int c = 0; c < 3; ++c) {
            x1[c] = 0.f;
            for (int m = 0; m < 4; ++m) h[c][m] = 0.f;
        }
    }
    // ---- smoothness: edge (c, c+1) of row r (zero in the last column) ----
    if (own_row && (terms & (TERM_SMOOTH_D | TERM_SMOOTH_U))) {
        const float wx = T.has1 ? edge_w3(T.x, x1) : 0.f;
        float ed = 0.f, eu = 0.f;
        if (terms & TERM_SMOOTH_D) {
            const float g = T.d - in[3 * SROW + 1];
            T.acc[ACC_SMOOTH_D] += own * (fabsf(g) * wx);
            if (GRAD) {
                ed = sgn_mul(g, (G.gd_up * P.coef[ACC_SMOOTH_D]) * wx);
                T.gd[0] += ed;
            }
        }
        if (terms & TERM_SMOOTH_U) {
            const float g = T.u - in[4 * SROW + 1];
            T.acc[ACC_SMOOTH_U] += own * (fabsf(g) * wx);
            if (GRAD) {
                eu = sgn_mul(g, (G.ge_up * P.coef[ACC_SMOOTH_U]) * wx);
                T.gu[0] += eu;
            }
        }
        if (GRAD) {
            if (terms & TERM_SMOOTH_D) T.sb[(row_ed(true) + 0) * SROW] = ed;
            if (terms & TERM_SMOOTH_U) T.sb[(row_ed(true) + 1) * SROW] = eu;
        }
    }
    const bool q_ok = STEADY || (q >= G.qlo && q <= P.h - 3);
    if (q_ok) {
        const float inv9 = 1.0f / 9.0f;
        float kt = 0.f;
        if (GRAD)      // d reproj / d dssim(q) = coef * alpha/3 * T(q), times -1/2
            kt = (-0.5f * G.gd_up * P.coef[ACC_REPROJ] * P.alpha * (1.0f / 3.0f) *
                  inv9 * S.RT[i].tyw) * T.txw;
        float dsum = 0.f;
        for (int c = 0; c < 3; ++c) {
            const float sx = T.H[0][c][0] + T.H[1][c][0] + h[c][0];
            const float sy = T.H[0][c][1] + T.H[1][c][1] + h[c][1];
            const float sq = T.H[0][c][2] + T.H[1][c][2] + h[c][2];
            const float sxy = T.H[0][c][3] + T.H[1][c][3] + h[c][3];
            const float mx = inv9 * sx, my = inv9 * sy;
            const float mm = mx * my;
            const float m2 = fmaf(mx, mx, my * my);
            const float n1 = fmaf(2.0f, mm, P.c1);
            const float d1 = m2 + P.c1;
            const float d2 = fmaf(inv9, sq, -m2) + P.c2;      // var_x + var_y + c2
            const float n2 = fmaf(2.0f, fmaf(inv9, sxy, -mm), P.c2);
            const float inv = fast_rcp(d1 * d2);
            const float ssim = (n1 * n2) * inv;
            const float raw = fmaf(-0.5f, ssim, 0.5f);
            const float cl = fminf(fmaxf(raw, 0.0f), 1.0f);
            dsum += T.win_ok ? cl : 0.f;
            if (GRAD) {
                // the clamp passes the gradient on the closed interval
                const float gb = (T.win_ok && raw >= 0.0f && raw <= 1.0f) ? kt : 0.0f;
                // dssim/dA = 2 mx (n2 - n1) inv - 2 my nn (d2 - d1) inv^2
                const float t1 = (2.0f * mx) * (n2 - n1);
                const float t2 = (2.0f * my) * (ssim * (d2 - d1));
                float* gx = T.sb + (row_gx(true) + c * 3) * SROW;
                T.g[c * 3 + 0] = gb * ((t1 - t2) * inv);
                T.g[c * 3 + 1] = (-2.0f * gb) * (ssim * (inv * d1));   // (2 y) dssim/dQ: the 2
                T.g[c * 3 + 2] = gb * (2.0f * (n1 * inv));
                gx[0] = T.g[c * 3 + 0];
                gx[SROW] = T.g[c * 3 + 1];
                gx[2 * SROW] = T.g[c * 3 + 2];
            }
        }
        T.sb[ru.ds_w] = dsum;
    }
    for (int c = 0; c < 3; ++c)
        for (int m = 0; m < 4; ++m) T.H[PAR][c][m] = h[c][m];
}

// ---- P3: everything that needs G(q) / the error map; row r - 2 ---------------
template <class C, int PAR>
USL_HD void c_p3(const LossParams& P, const CGeo& G, const CRings& S, int r,
                 CState& T) {
    constexpr int SROW = C::SROW;
    constexpr bool GRAD = C::GRAD, MASKED = C::MASKED, STEADY = C::STEADY;
    if (MASKED && !T.active) return;
    const int nh = GRAD ? NHIST_GRAD : NHIST_FWD;
    const int y = r - 2;
    const int i = c_step_index(G, r);
    const bool own_row = STEADY || (y >= G.ya && y < G.yb);
    const bool mine = own_row && (!MASKED || T.own != 0.f);
    const unsigned terms = C::TERMS >= 0 ? (unsigned)C::TERMS : P.terms;
    const RowU ru = S.RU[i];
    const float* hs = T.sb + ru.hs_r;
    const unsigned ro = (unsigned)(y * P.w);
    if (GRAD) {
        // the left neighbour's right edge of row r lands on this column
        if ((STEADY || (r >= G.ya && r < G.yb)) &&
            (terms & (TERM_SMOOTH_D | TERM_SMOOTH_U))) {
            if (terms & TERM_SMOOTH_D) T.gd[0] -= T.sb[(row_ed(true) + 0) * SROW - 1];
            if (terms & TERM_SMOOTH_U) T.gu[0] -= T.sb[(row_ed(true) + 1) * SROW - 1];
        }
        const bool q_ok = STEADY || (y >= G.qlo && y <= P.h - 3);
        float hg[9];
        if (q_ok) {
            for (int k = 0; k < 9; ++k) {
                const float* g = T.sb + (row_gx(true) + k) * SROW;
                hg[k] = (g[-2] + g[-1]) + T.g[k];        // windows c-2, c-1, c
            }
        } else {
            for (int k = 0; k < 9; ++k) hg[k] = 0.f;
        }
        if (mine) {
            float gs = 0.f;
            for (int c = 0; c < 3; ++c) {
                const float a = hs[c * SROW];
                const float sA = T.HG[0][c][0] + T.HG[1][c][0] + hg[c * 3 + 0];
                const float sQ = T.HG[0][c][1] + T.HG[1][c][1] + hg[c * 3 + 1];
                const float sC = T.HG[0][c][2] + T.HG[1][c][2] + hg[c * 3 + 2];
                gs = fmaf(a, sA, gs);
                gs = fmaf(hs[(3 + c) * SROW], sQ, gs);
                gs = fmaf(hs[(6 + c) * SROW], sC, gs);
                if (C::TERMS < 0 && P.grad_recon_in)   // (adversarial step: generic variant)
                    gs = fmaf(USL_LDG(P.grad_recon_in +
                                      ((long long)G.b * 6 + T.v * 3 + c) *
                                          ((long long)P.h * P.w) + ro + T.c), a, gs);
            }
            T.gd[2] += T.sign * gs;
        }
        for (int c = 0; c < 3; ++c)
            for (int m = 0; m < 3; ++m) T.HG[PAR][c][m] = hg[c * 3 + m];
    }
    if (!mine) return;
    // error map row y: bilinear (h-2, w-2) -> (h, w) of dssim, plus L1
    const RowT rt = S.RT[i];
    const float* d0 = T.sb + rt.ay_o0;
    const float* d1 = T.sb + rt.ay_o1;
    const float cw1 = T.axw, cw0 = 1.0f - cw1;
    const float up0 = cw0 * d0[T.ax0] + cw1 * d0[T.ax1];
    const float up1 = cw0 * d1[T.ax0] + cw1 * d1[T.ax1];
    const float up = (1.0f - rt.ay_w1) * up0 + rt.ay_w1 * up1;
    const float l1 = hs[(nh - 2) * SROW];
    const float e = (P.alpha * up + (1.0f - P.alpha) * l1) * (1.0f / 3.0f);
    T.acc[ACC_REPROJ] += e;
    float gur = 0.f;
    if (terms & TERM_UNC) {
        const float u = hs[(nh - 1) * SROW];
        if (P.loss_type == LOSS_L1) {
            const float f = u - e;
            T.acc[ACC_UNC] += fabsf(f);
            if (GRAD) gur = sgn_mul(f, G.ge_up * P.coef[ACC_UNC]);
        } else if (P.loss_type == LOSS_BAYESIAN) {
            const float iu = fast_rcp(u);
            T.acc[ACC_UNC] += fmaf(e, iu, LN2 * fast_log2(u));
            if (GRAD) gur = (G.ge_up * P.coef[ACC_UNC]) * (iu - e * iu * iu);
        } else {
            const float eu = e * fast_exp2(u * LOG2E);
            T.acc[ACC_UNC] += eu + u;
            if (GRAD) gur = (G.ge_up * P.coef[ACC_UNC]) * (eu + 1.0f);
        }
    }
    if (C::TERMS < 0 && P.err_out)     // (optional outputs: generic variant only)
        P.err_out[((long long)G.b * 2 + T.v) * ((long long)P.h * P.w) + ro + T.c] = e;
    if (GRAD) {
        // rows are finalised in order: the output offsets walk down with them.
        // Pure stores: the scatter kernel adds its part afterwards.
        P.grad_disp[T.o_gd] = T.gd[2];
        P.grad_unc[T.o_gu] = T.gu[2] + gur;
        T.o_gd += (unsigned)P.w;
        T.o_gu += (unsigned)P.w;
    }
}

}  // namespace ck
}  // namespace usl
