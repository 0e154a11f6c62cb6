// Host side of the column-marching fused loss kernels (col_core.cuh /
// col_kernel_impl.cuh): planning of the units of every scale and the launches.
//
// One CTA = one unit (sample, row strip, view set, column tile).  Every pyramid
// scale is its own launch, and the launches of a step run concurrently
// (col_launch forks them over side streams), largest scale first, so the short
// CTAs of the small scales fill the tail of the large one.
#include <stdlib.h>

#include "col_core.cuh"
#include "col_launch.cuh"

namespace usl {

// ------------------------------------------------------------------ host ---
static int pos_or(int v, int dflt) { return v > 0 ? v : dflt; }

bool col_eligible(const UslLossConfig* cfgs, const UslLossScale* scales, int n) {
    if (knobs().no_col) return false;
    for (int i = 0; i < n; ++i) {
        const UslLossScale& s = scales[i];
        if (s.flags & USL_SCALE_GENERAL_KERNELS) return false;
        const unsigned t = cfgs[i].terms;
        if (!(t & TERM_REPROJ) || s.recon_in || s.err_in) return false;
        if (!s.images || !s.disp) return false;
        if ((t & (TERM_UNC | TERM_SMOOTH_U | TERM_CONS_U)) && !s.unc) return false;
        if (s.w < 4 || s.h < 4) return false;
        // 32-bit element offsets inside every tensor
        const long long lim = 0x7fffffffLL;
        if ((long long)s.B * s.img_bs > lim || (long long)s.B * s.disp_bs > lim ||
            (long long)s.B * s.unc_bs > lim || (long long)s.B * s.gd_bs > lim ||
            (long long)s.B * s.gu_bs > lim)
            return false;
    }
    return true;
}

// The plan with a given minimum strip height.
static int col_plan_floor(ColPlan* M, bool grad, int floorR) {
    const int maxT = COL_MAX_THREADS;
    // widest two-view unit (in threads): wider rows get one unit per view
    const int maxT2 = pos_or(knobs().col_maxt2, COL_MAX_THREADS);
    // Strip heights.  The CTAs of the largest scale are long and hold a whole
    // SM each (registers); the other scales run beside them only on the SMs
    // they leave free.  Measured best (profiles/, config 2): the largest scale
    // on ~2/3 of the SMs in ONE wave, 32-row strips below.
    const int R0 = pos_or(knobs().col_r0, 0);
    int R0_used = 0;
    long long rows = 0;
    for (int i = 0; i < M->n; ++i) {
        LossParams& p = M->P[i];
        int nv, tiles = 1, TW = p.w, LW = p.w;
        if (2 * p.w <= maxT2) nv = 2;
        else if (p.w <= maxT) nv = 1;
        else {
            nv = 1;
            tiles = (p.w + (maxT - 4) - 1) / (maxT - 4);
            TW = (p.w + tiles - 1) / tiles;
            tiles = (p.w + TW - 1) / TW;
            LW = TW + 4;
        }
        // strip height: long strips for the big scale (less halo work), short
        // ones for the small scales (they fill the tail of the step)
        // Below the largest scale: strips as tall as they can be (each costs 8
        // more steps of halo than it has rows) while one still takes clearly
        // less time than a strip of the largest scale -- those are the critical
        // path and the small scales' transposed warps have to fit in behind
        // (config 2: 64 rows instead of 32 = 0.266 instead of 0.280 ms a step).
        int wantR = R0;
        if (i > 0) {
            wantR = pos_or(knobs().col_r, 0);
            if (!wantR) {
                wantR = (int)(0.85f * (float)(R0_used + 8)) - 8;
                if (wantR < floorR) wantR = floorR;
            }
        }
        if (i == 0 && R0 == 0) {
            const int per_strip = tiles * p.B * (2 / nv);
            int want_strips = (7 * num_sms() / 10) / per_strip;
            if (want_strips < 1) want_strips = 1;
            wantR = (p.h + want_strips - 1) / want_strips;
            if (wantR < floorR) wantR = floorR;
            if (wantR > 128) {
                // one wave on 70 % of the SMs would need taller strips than the
                // per-step tables in shared memory allow: several waves on all
                // SMs, as few strip steps (rows + 8 of halo) in total as possible
                int best = 128, best_cost = 0x7fffffff;
                for (int s = (p.h + 127) / 128; s <= (p.h + 31) / 32; ++s) {
                    const int R = (p.h + s - 1) / s;
                    const int waves = (per_strip * s + num_sms() - 1) / num_sms();
                    const int cost = waves * (R + 8);
                    if (cost < best_cost) { best_cost = cost; best = R; }
                }
                wantR = best;
            }
        }
        int strips = (p.h + wantR - 1) / wantR;
        int R = (((p.h + strips - 1) / strips) + 1) & ~1;
        strips = (p.h + R - 1) / R;
        if (i == 0) R0_used = R;
        p.TW = TW; p.R = R; p.LW = LW;
        M->nv[i] = nv; M->tiles_x[i] = tiles; M->strips[i] = strips;
        M->units[i] = tiles * strips * p.B * (2 / nv);
        const int nt = nv * LW;
        int cls = 64;
        while (cls < nt) cls *= 2;
        if (cls > maxT) return USL_ERR_UNSUPPORTED;
        M->cls[i] = cls;
        M->threads[i] = (nt + 31) & ~31;
        // bulk async row copies: 16-byte aligned rows of every plane
        const bool al = (p.w % 4 == 0) &&
            ((((uintptr_t)p.img | (uintptr_t)p.disp | (uintptr_t)p.unc) & 15) == 0) &&
            (((p.img_bs | p.img_cs | p.d_bs | p.d_cs | p.u_bs | p.u_cs) & 3) == 0);
        M->mode[i] = tiles > 1 ? ck::MODE_TILED
                               : (((nt & 31) || !al || knobs().col_no_tma)
                                      ? ck::MODE_MASKED : ck::MODE_PLAIN);
        M->smem[i] = ck::c_floats(col_class_srow(cls), p.w, nv, R, grad) * sizeof(float);
        if (M->smem[i] > 220 * 1024) return USL_ERR_UNSUPPORTED;
        M->row_start[i] = (int)rows;
        rows += M->units[i];
    }
    M->row_start[M->n] = (int)rows;
    return USL_OK;
}

int col_plan(ColPlan* M, bool grad) {
    // Strips of at least 32 rows (8 more rows of halo work each) -- unless the
    // whole step then has so few units that most SMs stay idle (small batches):
    // halve the minimum while all the units still fit on the SMs at once, the
    // step is as long as its longest unit.
    int rc = col_plan_floor(M, grad, 32);
    for (int floorR = 16; rc == USL_OK && floorR >= 8; floorR /= 2) {
        if (2 * M->row_start[M->n] > num_sms()) break;
        rc = col_plan_floor(M, grad, floorR);
    }
    return rc;
}

template <int CLS>
int col_launch_class(const ColPlan* M, int i, bool grad, int skip_if_unit,
                     cudaStream_t st);
extern template int col_launch_class<512>(const ColPlan*, int, bool, int, cudaStream_t);
extern template int col_launch_class<256>(const ColPlan*, int, bool, int, cudaStream_t);
extern template int col_launch_class<128>(const ColPlan*, int, bool, int, cudaStream_t);
extern template int col_launch_class<64>(const ColPlan*, int, bool, int, cudaStream_t);

int col_launch_scale(const ColPlan* M, int i, bool grad, int skip_if_unit,
                     cudaStream_t st) {
    // (profiling aid: USL_COL_ONLY = bit mask of the scales to launch)
    if (!((knobs().col_only >> i) & 1)) return USL_OK;
    switch (M->cls[i]) {
        case 512: return col_launch_class<512>(M, i, grad, skip_if_unit, st);
        case 256: return col_launch_class<256>(M, i, grad, skip_if_unit, st);
        case 128: return col_launch_class<128>(M, i, grad, skip_if_unit, st);
        case 64: return col_launch_class<64>(M, i, grad, skip_if_unit, st);
    }
    return USL_ERR_UNSUPPORTED;
}

int col_launch(const ColPlan* M, bool grad, int skip_if_unit, cudaStream_t st,
               ColAfter after, void* after_ctx) {
    StreamPool* pool = (M->n > 1 && !knobs().col_serial) ? stream_pool() : nullptr;
    if (!pool) {
        for (int i = 0; i < M->n; ++i) {
            int rc = col_launch_scale(M, i, grad, skip_if_unit, st);
            mark(2 + i, st);
            if (rc == USL_OK && after) { rc = after(after_ctx, i, st); mark(6 + i, st); }
            if (rc != USL_OK) return rc;
        }
        return USL_OK;
    }
    // fork: what precedes on `st` is ordered before every scale.  Scale 0 goes
    // to the high-priority stream: when all four launches become ready at the
    // same moment (graph replay) its CTAs -- a whole SM each, the longest of
    // the step -- must be placed first, not behind the small scales' CTAs.
    if (cudaEventRecord(pool->fork, st) != cudaSuccess) return USL_ERR_CUDA;
    const bool hi = !knobs().col_no_priority;
    int rc = USL_OK;
    if (hi) {
        if (cudaStreamWaitEvent(pool->first, pool->fork, 0) != cudaSuccess) return USL_ERR_CUDA;
        rc = col_launch_scale(M, 0, grad, skip_if_unit, pool->first);
        mark(2, pool->first);
        if (rc == USL_OK && after) { rc = after(after_ctx, 0, pool->first); mark(6, pool->first); }
        if (cudaEventRecord(pool->join_first, pool->first) != cudaSuccess ||
            cudaStreamWaitEvent(st, pool->join_first, 0) != cudaSuccess)
            rc = rc == USL_OK ? USL_ERR_CUDA : rc;
    } else {
        rc = col_launch_scale(M, 0, grad, skip_if_unit, st);
        mark(2, st);
        if (rc == USL_OK && after) { rc = after(after_ctx, 0, st); mark(6, st); }
    }
    for (int i = 1; i < M->n && rc == USL_OK; ++i) {
        cudaStream_t s = pool->side[i - 1];
        if (cudaStreamWaitEvent(s, pool->fork, 0) != cudaSuccess) { rc = USL_ERR_CUDA; break; }
        rc = col_launch_scale(M, i, grad, skip_if_unit, s);
        mark(2 + i, s);
        if (rc == USL_OK && after) { rc = after(after_ctx, i, s); mark(6 + i, s); }
        // join even after a failed launch so that `st` stays well ordered
        if (cudaEventRecord(pool->join[i - 1], s) != cudaSuccess ||
            cudaStreamWaitEvent(st, pool->join[i - 1], 0) != cudaSuccess)
            rc = rc == USL_OK ? USL_ERR_CUDA : rc;
    }
    return rc;
}

}  // namespace usl
