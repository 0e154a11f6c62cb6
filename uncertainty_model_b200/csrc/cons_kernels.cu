// The scattering half of kernel (2): the transposed warp of the consistency
// terms, deterministic and atomic free.  Arithmetic of cons_core.cuh (see its
// header for the citations: loss.py:167-188, 430-431); this file is the
// warp-per-row schedule of it.
//
// A CTA owns the destination rows [ya, yb) of one sample; they receive from the
// source rows [ya-1, yb].  Every (source row, source view) is a JOB done by one
// warp from start to end, into a row H of shared memory no other warp touches:
//   * the warp blends the opposite view's disparity row it samples (Vd);
//   * it walks the row in 32-pixel chunks, in order; the lanes of a chunk that
//     hit the same destination column are grouped (contiguous runs where the
//     destinations are monotone -- one shuffle, two votes -- else match.any)
//     and summed by the group leader in lane order over shuffles; the leaders
//     -- whose destinations are now distinct -- add into H with plain
//     shared-memory read-modify-writes (first tap, __syncwarp, second tap),
//     the other lanes into spare cells, so that nothing branches.  The order
//     of every sum is a pure function of the data: no atomics, bit-identical
//     run to run.
// After ONE block barrier the destination rows are assembled from the three
// source rows around each of them, with the vertical tap weights, and stored
// (or added to what the fused kernel wrote before: `accumulate`).
#include <stdlib.h>

#include "cons_core.cuh"
#include "cons_launch.cuh"
#include "usl_common.cuh"

namespace usl {

constexpr int CONS2_THREADS = 512;   // 16 warps: 2 * (14 + 2) jobs = 2 each
constexpr int CONS2_R = 14;          // destination rows per CTA
constexpr int HPAD = 2;              // absorbs taps that fall outside the row
constexpr int NCH = 8;               // chunks of 32 pixels per span (registers)

__host__ __device__ inline int xb_floats(int w) {
    return (w + 32 * NCH - 1) / (32 * NCH) * (32 * NCH);
}

constexpr int DEAD_LO = -2000;          // keys of lanes left of the row (+ lane)
constexpr int DEAD_HI = 0x40000000;     // ... right of it, or past its end

// The lanes of the warp whose key equals this lane's.  Destination columns are
// non-decreasing along a chunk almost everywhere (they decrease only where the
// disparity jumps by more than a pixel per pixel); equal keys are then
// contiguous runs, found with one shuffle and two votes.  match.any -- one per
// ~67 cycles and scheduler -- is left to the chunks that are not monotone.
__device__ __forceinline__ unsigned equal_key_lanes(int key, int lane) {
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool first = lane == 0;
    if (!__all_sync(0xffffffffu, first || key >= prev))
        return __match_any_sync(0xffffffffu, key);
    const unsigned heads = __ballot_sync(0xffffffffu, first || key != prev);
    const unsigned upto = 0xffffffffu >> (31 - lane);        // lanes <= this one
    const int start = 31 - __clz(heads & upto);              // head of this run
    const unsigned later = heads & ~upto;
    const int end = later ? __ffs(later) - 1 : 32;           // one past its last lane
    return (0xffffffffu >> (32 - end)) & (0xffffffffu << start);
}

// One chunk of 32 sources into H.  `d`: destination column of the first tap
// (dead lanes: a unique key below -1 or from DEAD_HI up; they write nothing).
// Lanes that share a destination form a group (`grp`, from equal_key_lanes --
// issued by the caller for a whole span at once: match.any is slow and its
// latency long).  Each group LEADER -- its lowest lane -- adds up the
// contributions of its group in lane order, fetched with shuffles in a loop
// every lane walks together (its trip count is the size of the largest group,
// minus one: no divergent branches); then the leaders, whose destinations are
// distinct, update H: first taps, __syncwarp, second taps.  The other lanes do
// the same read-modify-writes on two private cells of `spare` (34 floats per
// warp) instead of branching.
__device__ __forceinline__ void scatter_chunk(float* Hrow, float* spare, int d,
                                              unsigned grp, float a0, float a1,
                                              int lane) {
    const bool leader = d >= -1 && d < DEAD_HI && (grp & ((1u << lane) - 1u)) == 0u;
    // own contribution first (the leader is the lowest lane), then the rest of
    // the group in lane order; most groups are singletons
    unsigned m = leader ? grp & (grp - 1u) : 0u;
    float s0 = a0, s1 = a1;
    while (__any_sync(0xffffffffu, m != 0u)) {
        const int src = m ? __ffs(m) - 1 : lane;
        const float t0 = __shfl_sync(0xffffffffu, a0, src);
        const float t1 = __shfl_sync(0xffffffffu, a1, src);
        if (m) { s0 += t0; s1 += t1; }
        m &= m - 1u;
    }
    float* p = leader ? Hrow + d : spare + lane;
    p[0] += leader ? s0 : 0.0f;
    __syncwarp();
    p[1] += leader ? s1 : 0.0f;
    __syncwarp();
}

// ALIGNED: every row width of the launch is a multiple of 32 (a chunk is either
// wholly inside its row or wholly outside).
template <bool ALIGNED>
__global__ void __launch_bounds__(CONS2_THREADS, 2)
cons_scatter2_kernel(const __grid_constant__ MultiCons M) {
    extern __shared__ float4 smem_raw[];
    int s = 0;
    while (s + 1 < M.n && (int)blockIdx.x >= M.cta_start[s + 1]) ++s;
    const ConsParams& P = M.P[s];
    const float gd_up = P.gout_d ? __ldg(P.gout_d) : P.gout_default;
    const float ge_up = P.gout_e ? __ldg(P.gout_e) : P.gout_default;
    if (M.skip_if_unit && gd_up == 1.0f && ge_up == 1.0f) return;
    const int local = blockIdx.x - M.cta_start[s];
    // strip-major: the (short) last strips of all samples come last
    const int b = local % P.B;
    const int ya = (local / P.B) * P.R;
    const int yb = min(P.h, ya + P.R);
    const int w = P.w, h = P.h;
    const int HW = w + 2 * HPAD;
    const int nr = yb - ya + 2;                 // source rows ya-1 .. yb
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    float* H = reinterpret_cast<float*>(smem_raw);        // [nr][2][HW]
    float* Vall = H + (size_t)(P.R + 2) * 2 * HW;         // [nwarps][HW]
    float* Vd = Vall + (size_t)warp * HW + HPAD;
    // linspace(0,1,w), padded to whole spans (every lane of a span reads it)
    float* xb = Vall + (size_t)nwarps * HW;
    const int xbn = xb_floats(w);
    float* spare = xb + xbn + warp * 64;
    if (lane < 2) { spare[lane] = 0.0f; spare[32 + lane] = 0.0f; }
    spare[2 + lane] = 0.0f;
    for (int x = tid; x < xbn; x += blockDim.x) xb[x] = x < w ? linspace01(x, w) : 0.0f;
    __syncthreads();
    const float fw = (float)w;

    for (int job = warp; job < 2 * nr; job += nwarps) {
        const int rsi = job >> 1, v = job & 1, rs = ya - 1 + rsi;
        float* Hrow = H + ((size_t)rsi * 2 + (1 - v)) * HW + HPAD;
        for (int x = lane - HPAD; x < w + HPAD; x += 32) Hrow[x] = 0.0f;
        if (rs < 0 || rs >= h) continue;
        // blended opposite disparity row for source view v
        {
            const Tap2 ty = warp_row_taps(rs, h);
            const bool ok0 = ty.i0 >= 0 && ty.i0 < h;
            const bool ok1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < h;
            const float w0 = ok0 ? ty.w0 : 0.0f, w1 = ok1 ? ty.w1 : 0.0f;
            const float* pd = plane(P.disp, P.d_bs, P.d_cs, b, 1 - v);
            const float* p0 = pd + (long long)(ok0 ? ty.i0 : 0) * w;
            const float* p1 = pd + (long long)(ok1 ? ty.i0 + 1 : 0) * w;
            for (int x = lane; x < w; x += 32)
                Vd[x] = w0 * __ldg(p0 + x) + w1 * __ldg(p1 + x);
            if (lane < HPAD) { Vd[-1 - lane] = 0.0f; Vd[w + lane] = 0.0f; }
        }
        __syncwarp();
        const float sign = v ? 1.0f : -1.0f;
        for (int term = 0; term < 2; ++term) {
            if (!(P.terms & (term ? TERM_CONS_U : TERM_CONS_D))) continue;
            const float* pa = (term ? plane(P.unc, P.u_bs, P.u_cs, b, v)
                                    : plane(P.disp, P.d_bs, P.d_cs, b, v)) +
                              (long long)rs * w;
            const float k = term ? ge_up * P.coef_ud : gd_up * P.coef_dd;
            // The row in spans of NCH chunks.  Pass 1 is latency tolerant: all
            // the loads of a span are in flight together, then destinations
            // and tap contributions of every pixel go to registers.  Pass 2 is
            // the ordered part: chunk after chunk into H.
            for (int span = 0; span < w; span += 32 * NCH) {
                float c0[NCH], c1[NCH];
                int dst[NCH];
                unsigned grp[NCH];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int x = span + j * 32 + lane;
                    c0[j] = (ALIGNED ? span + j * 32 < w : x < w) ? __ldg(pa + x) : 0.0f;
                }
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int x = span + j * 32 + lane;
                    const bool valid = ALIGNED ? span + j * 32 < w : x < w;
                    const float a = c0[j];
                    // warp_coord(x, w, sign * a) with the base grid from the table
                    const float g = fmaf(2.0f, xb[x] + sign * a, -1.0f);
                    const Tap2 tx = split_coord(fmaf(g + 1.0f, 0.5f * fw, -0.5f));
                    const int xi = min(max(tx.i0, -HPAD), w);
                    const float f0 = Vd[xi], f1 = Vd[xi + 1];
                    const float f = a - (tx.w0 * f0 + tx.w1 * f1);
                    // -k * sign(f), as two selects
                    float m = f > 0.0f ? -k : 0.0f;
                    m = f < 0.0f ? k : m;
                    // taps -1 and w land in the pads of the row; dead lanes get
                    // unique keys below / above every column, in lane order
                    const bool live = valid && (unsigned)(xi + 1) <= (unsigned)w;
                    dst[j] = live ? xi : (xi < 0 ? DEAD_LO : DEAD_HI) + lane;
                    c0[j] = m * tx.w0;
                    c1[j] = m * tx.w1;
                    grp[j] = equal_key_lanes(dst[j], lane);
                }
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    if (span + j * 32 >= w) break;
                    scatter_chunk(Hrow, spare, dst[j], grp[j], c0[j], c1[j], lane);
                }
            }
        }
    }
    __syncthreads();
    // destination rows: H(y'-1), H(y'), H(y'+1) with the vertical tap weights
    // (row by row, so that no thread divides by the width)
    for (int yi = warp; yi < yb - ya; yi += nwarps) {
        const int yd = ya + yi;
        float wgt[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const int rs = yd - 1 + kk;
            wgt[kk] = 0.0f;
            if (rs < 0 || rs >= h) continue;
            const Tap2 ty = warp_row_taps(rs, h);
            if (ty.i0 == yd) wgt[kk] = ty.w0;
            else if (ty.i0 + 1 == yd) wgt[kk] = ty.w1;
        }
        for (int o = 0; o < 2; ++o) {
            const float* h0 = H + ((size_t)yi * 2 + o) * HW + HPAD;
            float* out = P.grad_disp + (long long)b * P.gd_bs + o * P.gd_cs + (long long)yd * w;
            for (int x0 = lane; x0 < w; x0 += 128) {
                float old[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + 32 * u;
                    old[u] = (P.accumulate && x < w) ? out[x] : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int x = x0 + 32 * u;
                    if (x < w)
                        out[x] = old[u] + (wgt[0] * h0[x] + wgt[1] * h0[2 * HW + x] +
                                           wgt[2] * h0[4 * HW + x]);
                }
            }
        }
    }
}

int cons_scatter2_launch(MultiCons* C, cudaStream_t st) {
    // strips of 14 destination rows: 2 * 16 jobs over 16 warps, two CTAs per
    // SM -- or of 6 rows (one job per warp, half as long a CTA) when even those
    // all run at once (small batches: the step is as long as one CTA)
    size_t smem = 0;
    int R = CONS2_R;
    {
        long long ctas = 0;
        for (int k = 0; k < C->n; ++k) ctas += (long long)((C->P[k].h + 5) / 6) * C->P[k].B;
        if (ctas <= 2 * num_sms()) R = 6;
    }
    C->cta_start[0] = 0;
    for (int k = 0; k < C->n; ++k) {
        ConsParams& c = C->P[k];
        c.R = R;
        if (c.R > c.h) c.R = c.h;
        C->strips[k] = (c.h + c.R - 1) / c.R;
        C->cta_start[k + 1] = C->cta_start[k] + C->strips[k] * c.B;
        const size_t bytes = (((size_t)(c.R + 2) * 2 + CONS2_THREADS / 32) *
                                  (c.w + 2 * HPAD) + xb_floats(c.w) +
                              (size_t)CONS2_THREADS * 2) * sizeof(float);
        if (bytes > smem) smem = bytes;
    }
    if (smem > 220 * 1024) return USL_ERR_UNSUPPORTED;
    bool aligned = true;
    for (int k = 0; k < C->n; ++k) aligned = aligned && (C->P[k].w % 32 == 0);
    if (aligned) {
        if (cudaFuncSetAttribute(cons_scatter2_kernel<true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return USL_ERR_CUDA;
        cons_scatter2_kernel<true><<<C->cta_start[C->n], CONS2_THREADS, smem, st>>>(*C);
    } else {
        if (cudaFuncSetAttribute(cons_scatter2_kernel<false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return USL_ERR_CUDA;
        cons_scatter2_kernel<false><<<C->cta_start[C->n], CONS2_THREADS, smem, st>>>(*C);
    }
    return check_launch();
}

}  // namespace usl
