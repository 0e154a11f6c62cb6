// Lane-per-row form of the deterministic transposed warp of the consistency
// terms (cons_core.cuh has the mathematics and the citations: loss.py:167-188,
// 430-431; cons_kernels.cu the warp-per-row form), for the training step.
//
// What makes the warp-per-row kernel expensive is not the arithmetic but
// keeping 32 lanes that may hit the same column from losing updates: on
// white-noise disparities every chunk needs a MATCH.ANY (about 8 cycles per
// distinct key on the match unit: ~250 cycles a chunk, 40 % of that kernel)
// plus leader sums over shuffles.  Here a lane owns a whole (source row, view):
// it walks its row from left to right and adds each pixel's two taps -- of both
// terms -- into a row H of shared memory no other lane touches.  No grouping,
// no shuffles, no __syncwarp: every cell is written by one thread in program
// order, so the result is deterministic by construction (order of every sum:
// columns ascending, term dd before term ud).  All lanes of a warp are at the
// same column at the same time, so the base grid value is warp uniform.
//
// The price is one private row per lane (w + 5 floats): 64 lanes per SM at
// w = 512, i.e. two walking warps whose pace is one dependent
// load-add-store per column.  So everything that is not that chain is taken out
// of them.  The column kernels leave {d, u, s_dd, s_ud} per pixel behind
// (ConsParams::scat: s = coef * upstream * sign(a - warp(b)) is the one thing
// that cannot be recomputed without the blended row).  The warps of the CTA that
// do not walk turn them, a tile of 16 columns ahead, into what the chain
// consumes -- the tap contributions s * w0, s * w1 of both terms and the row
// addresses of their first taps -- and hand them over through a 4-stage ring in
// shared memory (mbarriers; a 336-byte block per row and tile, odd in 16-byte
// units so that the column walk is bank-conflict free).  The records reach the
// ring by asynchronous 16-byte copies issued three tiles ahead and are decoded
// in place.  The walkers interleave the read-modify-writes of 8 columns with the
// ring loads of the next 8, and run the two terms of a column as ONE round trip
// (four loads, four adds, four stores; where taps of the two terms coincide the
// later store carries both contributions, summed off the dependent chain).
//
// CTA = destination rows [ya, yb) of one sample; lane t = rsi * 2 + v walks
// source row rs0 + rsi of view v.  At the end all warps assemble the destination
// rows from the three source rows around each, with the vertical tap weights,
// and add them to the gradient the column kernels stored.
#include "cons_core.cuh"
#include "cons_launch.cuh"
#include "usl_common.cuh"

namespace usl {

constexpr int R4_THREADS = 320;
constexpr int R4_TILE = 16;                 // columns per staged tile
constexpr int R4_PITCH = R4_TILE + 4 + 1;   // 16-byte elements per staged (row, tile) block:
                                            // 16 contribution quadruples, 16 column pairs (odd: the
                                            // column walk is bank-conflict free)
constexpr int R4_STAGES = 4;
constexpr int R4_PAD = 2;                   // row = [2 | w | 2 (+1)]: taps outside the
                                            // image land in the pads
constexpr int R4_BATCH = 8;
constexpr int R4_SMEM_LIMIT = 224 * 1024;    // of the 227 KB a CTA may have
constexpr int R4_MAXE = 4;                  // elements of a tile per preparing thread

// floats per private row: odd, so that lanes at the same column (smooth
// disparities) sit in 32 different banks
__host__ __device__ inline int r4_pitch(int w) { return (w + 2 * R4_PAD) | 1; }

struct R4Geo { int nl, R; size_t smem; };

// ring bytes (reused after the walk as the transposition scratch of the
// assembly: a 32 x 33 float tile per warp)
__host__ __device__ inline size_t r4_ring_bytes(int nl) {
    const size_t ring = (size_t)R4_STAGES * nl * R4_PITCH * 16;
    const size_t scratch = (size_t)(R4_THREADS / 32) * 32 * 33 * 4;
    return ring > scratch ? ring : scratch;
}
__host__ __device__ inline size_t r4_smem(int nl, int w) {
    return (size_t)nl * r4_pitch(w) * 4 + r4_ring_bytes(nl) + 64;
}

// walking lanes (32 .. 128) and strip height of a scale; nl = 0 if even 32
// private rows do not fit
static R4Geo r4_geometry(int h, int w, size_t budget) {
    R4Geo g; g.nl = 0; g.R = 0; g.smem = 0;
    for (int nl = 128; nl >= 32; nl -= 32) {
        if (r4_smem(nl, w) > budget) continue;
        // a preparing thread takes at most R4_MAXE of the nl * 16 elements of a tile
        if (nl * R4_TILE > R4_MAXE * (R4_THREADS - nl)) continue;
        // no point in more rows than the image has
        if (nl > 32 && (nl - 32) / 2 >= h) continue;
        g.nl = nl; g.smem = r4_smem(nl, w);
        break;
    }
    if (!g.nl) return g;
    const int RS = g.nl / 2;                // source rows per strip
    g.R = RS >= h ? h : RS - 2;
    return g;
}

__device__ __forceinline__ uint32_t s_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mb_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "R4_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra R4_DONE;\n"
        "bra R4_WAIT;\n"
        "R4_DONE:\n"
        "}" ::"r"(s_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ int mb_test(uint64_t* bar, uint32_t parity) {
    int ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}" : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
    return ok;
}
// c_warp_coord of col_core.cuh (same operations in the same order: the sampling
// column comes out bit-identical to the one the column kernels used)
__device__ __forceinline__ float r4_coord(float xbase, float shift, float half_n) {
    const float x = xbase + shift;
    const float g = fmaf(2.0f, x, -1.0f);
    return fmaf(g + 1.0f, half_n, -0.5f);
}

// first tap column + 2 (columns -2 .. w+1 exist in a row: taps outside the image
// land in pads nobody reads) and the two tap contributions of a pixel
__device__ __forceinline__ void r4_decode(float xb, float shift, float s, float hwf,
                                          float fw, unsigned& col, float& c0, float& c1) {
    const float ix = r4_coord(xb, shift, hwf);
    const float f = floorf(ix);
    const float w1 = ix - f, w0 = (f + 1.0f) - ix;
    col = (unsigned)((int)fminf(fmaxf(f, -2.0f), fw) + 2);
    c0 = s * w0;
    c1 = s * w1;
}

struct R4Batch {            // 8 columns: tap contributions and row addresses
    float4 c[R4_BATCH];     // {dd tap 0, dd tap 1, ud tap 0, ud tap 1}
    uint32_t pd[R4_BATCH], pu[R4_BATCH];   // shared-window addresses of the first taps
};

__global__ void __launch_bounds__(R4_THREADS, 1)
cons_rows_kernel(const __grid_constant__ MultiCons M) {
    extern __shared__ float4 smem_raw[];
    int s = 0;
    while (s + 1 < M.n && (int)blockIdx.x >= M.cta_start[s + 1]) ++s;
    const ConsParams& P = M.P[s];
    if (M.skip_if_unit) {
        const float gd_up = P.gout_d ? __ldg(P.gout_d) : P.gout_default;
        const float ge_up = P.gout_e ? __ldg(P.gout_e) : P.gout_default;
        if (gd_up == 1.0f && ge_up == 1.0f) return;
    }
    const int local = blockIdx.x - M.cta_start[s];
    const int b = local % P.B;                  // strip-major
    const int ya = (local / P.B) * P.R;
    const int yb = min(P.h, ya + P.R);
    const int w = P.w, h = P.h;
    const int HW = r4_pitch(w);
    const int nl = M.lanes[s];                  // walking lanes (multiple of 32)
    const int rs0 = max(ya - 1, 0);             // first source row of the strip
    const int rs1 = min(yb, h - 1);             // last one
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4* stage = smem_raw;                                       // [STAGES][nl][PITCH]
    float* H = reinterpret_cast<float*>(reinterpret_cast<char*>(stage) + r4_ring_bytes(nl));  // [nl / 32][HW][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(H + (size_t)nl * HW);  // [STAGES]
    uint64_t* empty = full + R4_STAGES;                                 // [STAGES]
    const long long hw = (long long)h * w;
    const int nwalk = nl >> 5;                  // walking warps: 0 .. nwalk-1
    const int ntile = (w + R4_TILE - 1) / R4_TILE;

    // H (nl * HW floats, nl a multiple of 32: whole 16-byte words) and the ring:
    // rows past the strip are never copied, their lanes must read zeros
    // (contributions 0 into column -2, a pad)
    for (int i = tid; i < (int)(r4_ring_bytes(nl) / 16) + nl * HW / 4; i += R4_THREADS)
        stage[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (tid == 0) {
        // (every preparing warp arrives once per tile, every walking warp releases it)
        for (int k = 0; k < R4_STAGES; ++k) {
            mb_init(full + k, R4_THREADS / 32 - nwalk);
            mb_init(empty + k, nwalk);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= nwalk) {
        // ---- the preparing warps: {d, u, s_dd, s_ud} of a tile (16 columns of
        // every row) -> tap contributions + tap columns in the ring.  The records
        // come in by 16-byte asynchronous copies STRAIGHT INTO the ring slots
        // they are decoded in, R4_STAGES - 1 tiles ahead: no registers or
        // scoreboard entries are held while they fly, so the memory latency of
        // three tiles overlaps the decoding of one.  (Two groups of warps on
        // alternate tiles with plain loads waited two thirds of the time; a
        // register prefetch of the next tile does not overlap anything -- the
        // wait for this tile's loads drains the scoreboard the new loads were
        // just put on.)  A thread decodes what it copied itself: its own
        // wait_group is all the synchronisation the copies need.  Consecutive
        // threads take consecutive columns of a row: 256 contiguous bytes.
        const int npw = R4_THREADS / 32 - nwalk;
        const int gt = (warp - nwalk) * 32 + lane, gn = npw * 32;
        const int nelem = (rs1 - rs0 + 1) * 2 * R4_TILE;      // per tile
        const float fw = (float)w, hwf = 0.5f * fw;
        const float4* src0 = reinterpret_cast<const float4*>(P.scat) + (long long)b * 2 * hw;
        // (what a thread has no element for goes to the pad slot of row 0's
        //  block: straight-line code, the four decodes of a thread interleave)
        const float step = w > 1 ? 1.0f / (float)(w - 1) : 0.0f;
        auto issue = [&](int k) {
            const int st = k % R4_STAGES;
#pragma unroll
            for (int i = 0; i < R4_MAXE; ++i) {
                const int e = gt + i * gn;
                const bool have = e < nelem;
                const int t = have ? e / R4_TILE : 0, j = have ? e % R4_TILE : R4_PITCH - 1;
                const int x = k * R4_TILE + (e % R4_TILE);
                const bool in = have && x < w;         // (past the row: zero fill)
                const float4* src = src0 + (long long)(t & 1) * hw +
                                    (long long)(rs0 + (t >> 1)) * w + (in ? x : 0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::
                             "r"(s_u32(stage + ((size_t)st * nl + t) * R4_PITCH + j)),
                             "l"(src), "r"(in ? 16 : 0) : "memory");
            }
        };
        // One copy group per tile, in tile order.  A stage is refilled as soon as
        // the walkers have given it back -- tested without blocking before every
        // tile, so that decoding runs up to R4_STAGES - 1 tiles ahead of the walk
        // instead of waiting for the walkers after each tile; the only blocking
        // wait is for the stage of the tile that is to be decoded next.
        int issued = 0;
        auto refill = [&](bool must) -> bool {
            // tile `issued` goes into the stage tile `issued - R4_STAGES` had
            if (issued >= R4_STAGES) {
                uint64_t* bar = empty + issued % R4_STAGES;
                const uint32_t parity = ((issued / R4_STAGES) - 1) & 1;
                if (must) {
                    mb_wait(bar, parity);
                } else {
                    int ok = 0;
                    if (lane == 0) ok = mb_test(bar, parity);
                    if (!__shfl_sync(0xffffffffu, ok, 0)) return false;
                }
            }
            issue(issued);
            asm volatile("cp.async.commit_group;" ::: "memory");
            ++issued;
            return true;
        };
        for (int k = 0; k < ntile; ++k) {
            const int st = k % R4_STAGES;
            if (issued <= k) refill(true);
            while (issued < ntile && issued < k + R4_STAGES && refill(false)) {}
            // tile k's copies (this thread's) have landed when only the groups
            // committed after it are still pending
            switch (issued - 1 - k) {
                case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
                case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
                case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
                default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
            }
#pragma unroll
            for (int i = 0; i < R4_MAXE; ++i) {
                const int e = gt + i * gn;
                const bool have = e < nelem;
                const int t = have ? e / R4_TILE : 0, j = e % R4_TILE;
                float4* blk = stage + ((size_t)st * nl + t) * R4_PITCH;
                const float4 q = blk[have ? j : R4_PITCH - 1];
                const float sign = (t & 1) ? 1.0f : -1.0f;
                // linspace01(x, w) (usl_math.cuh), the division hoisted
                const int x = k * R4_TILE + j;
                const float xb = (x < w / 2) ? step * (float)x
                                             : fmaf(-step, (float)(w - 1 - x), 1.0f);
                unsigned cd, cu;
                float4 c;
                float cz, cw;
                r4_decode(xb, sign * q.x, q.z, hwf, fw, cd, c.x, c.y);
                r4_decode(xb, sign * q.y, q.w, hwf, fw, cu, cz, cw);
                // Where a tap of term ud falls on a cell term dd updates in the
                // same column, the later store must carry both parts: the loaded
                // value of that cell is the same for both, so the walker
                // subtracts (dd's part + ud's part) from it.
                const float k0 = cu == cd ? c.x : (cu == cd + 1 ? c.y : 0.0f);
                const float k1 = cu == cd ? c.y : (cu + 1 == cd ? c.x : 0.0f);
                c.z = k0 + cz;
                c.w = k1 + cw;
                blk[have ? j : R4_PITCH - 1] = c;
                reinterpret_cast<unsigned*>(blk + R4_TILE)[have ? j : 4 * (R4_PITCH - 1 - R4_TILE)] =
                    cd | (cu << 16);
            }
            __syncwarp();
            if (lane == 0) mb_arrive(full + st);
        }
    } else if (warp < nwalk) {
        // ---- the walk ---------------------------------------------------------
        // Software pipeline over batches of 8 columns: the read-modify-writes
        // of batch i are interleaved with the staged loads of batch i + 1.
        // the rows of a walking warp lie TRANSPOSED: cell (column x, lane t) at
        // x * 32 + t, i.e. lane t only ever touches bank t -- conflict free
        // wherever the 32 lanes' taps fall
        const uint32_t row_a = s_u32(H + (size_t)warp * HW * 32 + lane);   // column -2 of the row
        constexpr int BPT = R4_TILE / R4_BATCH;               // batches per tile
        const int nbatch = (w + R4_BATCH - 1) / R4_BATCH;

        // staged loads of batch `bi`: tap contributions, tap columns -> addresses
        auto load = [&](int bi, R4Batch& D) {
            const int k = bi / BPT, st = k % R4_STAGES;
            const int j = (bi - k * BPT) * R4_BATCH;
            const float4* blk = stage + ((size_t)st * nl + tid) * R4_PITCH;
            const float4* a = blk + j;
            const uint4* o = reinterpret_cast<const uint4*>(blk + R4_TILE) + j / 4;
            const uint4 o0 = o[0], o1 = o[1];
            const uint32_t ow[R4_BATCH] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
            for (int u = 0; u < R4_BATCH; ++u) {
                D.c[u] = a[u];
                D.pd[u] = row_a + 128u * (ow[u] & 0xffffu);
                D.pu[u] = row_a + 128u * (ow[u] >> 16);
            }
        };
        auto acquire = [&](int bi) {
            if (bi % BPT == 0) {
                const int k = bi / BPT;
                mb_wait(full + k % R4_STAGES, (k / R4_STAGES) & 1);
            }
        };
        auto release = [&](int bi) {
            if (bi % BPT == BPT - 1 || bi == nbatch - 1) {
                __syncwarp();
                if (lane == 0) mb_arrive(empty + (bi / BPT) % R4_STAGES);
            }
        };
        // Both terms of a column as ONE round trip: four loads, four adds, four
        // stores, in that order (the preparing warps have folded term dd's part
        // into term ud's where their taps coincide: the later store carries
        // both).  Explicit shared-memory instructions: the four loads must all
        // be issued before the arithmetic -- left to itself the compiler sinks
        // the second pair behind a branch, two round trips a column.
        auto rmw = [&](const R4Batch& D, int u) {
            const uint32_t pd = D.pd[u], pu = D.pu[u];
            const float4 c = D.c[u];
            float d0, d1, u0, u1;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d0) : "r"(pd));
            asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(d1) : "r"(pd));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(u0) : "r"(pu));
            asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(u1) : "r"(pu));
            d0 -= c.x; d1 -= c.y;
            u0 -= c.z; u1 -= c.w;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pd), "f"(d0));
            asm volatile("st.shared.f32 [%0+128], %1;" ::"r"(pd), "f"(d1));
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pu), "f"(u0));
            asm volatile("st.shared.f32 [%0+128], %1;" ::"r"(pu), "f"(u1));
        };
        // (columns past the end of a row were filled with zero contributions)
        R4Batch A, B;
        acquire(0);
        load(0, A);
        release(0);
#pragma unroll 1
        for (int bi = 0; bi < nbatch; bi += 2) {
            if (bi + 1 < nbatch) {
                acquire(bi + 1);
                load(bi + 1, B);
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) rmw(A, u);
                release(bi + 1);
            } else {
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) rmw(A, u);
                break;
            }
            if (bi + 2 < nbatch) {
                acquire(bi + 2);
                load(bi + 2, A);
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) rmw(B, u);
                release(bi + 2);
            } else {
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) rmw(B, u);
            }
        }
    }
    __syncthreads();

    // ---- destination rows: H(y'-1), H(y'), H(y'+1) with the vertical weights ----
    // A warp takes a block of 32 columns x 32 destination (row, view)s: lane =
    // destination while it reads the transposed rows (three distinct banks per
    // lane), lane = column while it writes to global memory; in between the
    // block goes through a 32 x 33 tile of the ring's memory, which is free now.
    {
        float* tile = reinterpret_cast<float*>(stage) + (size_t)warp * (32 * 33);
        const int ndest = (yb - ya) * 2;
        const int nxb = (w + 31) / 32, nlb = (ndest + 31) / 32;
        // rows of the gradient 16-byte aligned: whole quads of columns
        const bool vec_ok = (w % 4 == 0) && (((uintptr_t)P.grad_disp & 15) == 0) &&
                            (((P.gd_bs | P.gd_cs) & 3) == 0);
        for (int blk = warp; blk < nxb * nlb; blk += R4_THREADS / 32) {
            const int xb = (blk % nxb) * 32, lb = (blk / nxb) * 32;
            const int L = lb + lane;
            const int yd = ya + (L >> 1), o = L & 1;
            float wgt[3];
            const float* src[3];
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
                const int rs = yd - 1 + kk;
                wgt[kk] = 0.0f;
                int T = 0;
                if (L < ndest && rs >= 0 && rs < h) {
                    // (source view 1 - o scatters into destination view o)
                    T = (rs - rs0) * 2 + (1 - o);
                    const Tap2 ty = warp_row_taps(rs, h);
                    if (ty.i0 == yd) wgt[kk] = ty.w0;
                    else if (ty.i0 + 1 == yd) wgt[kk] = ty.w1;
                }
                src[kk] = H + (size_t)(T >> 5) * HW * 32 + (T & 31) + (size_t)(R4_PAD + xb) * 32;
            }
            const int nx = min(32, w - xb);
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                float v = 0.0f;
                if (i < nx)
                    v = wgt[0] * src[0][i * 32] + wgt[1] * src[1][i * 32] + wgt[2] * src[2][i * 32];
                tile[lane * 33 + i] = v;
            }
            __syncwarp();
            // One add per destination element, by the only thread that owns it in
            // this launch, onto what an EARLIER launch stored: a reduction
            // instruction (no return value, the add happens in L2) gives the bits
            // of load-add-store without a round trip per element.
            const int nd = min(32, ndest - lb);
            if (vec_ok) {
                // four columns a lane, four destinations a warp instruction:
                // 128-bit reductions (REDG.ADD.F32x4), a quarter of the lane
                // operations the L2 reduction path is limited by
                const int jj = lane >> 3, x4 = 4 * (lane & 7);
                if (x4 < nx) {
#pragma unroll 2
                    for (int j0 = 0; j0 < nd; j0 += 4) {
                        const int j = j0 + jj;
                        if (j >= nd) break;
                        const int Lj = lb + j;
                        float* out = P.grad_disp + (long long)b * P.gd_bs + (Lj & 1) * P.gd_cs +
                                     (long long)(ya + (Lj >> 1)) * w + xb + x4;
                        const float* tp = tile + j * 33 + x4;
                        const float v0 = tp[0], v1 = tp[1], v2 = tp[2], v3 = tp[3];
                        if (P.accumulate)
                            asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(out),
                                         "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
                        else
                            *reinterpret_cast<float4*>(out) = make_float4(v0, v1, v2, v3);
                    }
                }
            } else if (lane < nx) {
#pragma unroll 4
                for (int j = 0; j < nd; ++j) {
                    const int Lj = lb + j;
                    float* out = P.grad_disp + (long long)b * P.gd_bs + (Lj & 1) * P.gd_cs +
                                 (long long)(ya + (Lj >> 1)) * w + xb + lane;
                    const float v = tile[j * 33 + lane];
                    if (P.accumulate)
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(out), "f"(v) : "memory");
                    else
                        *out = v;
                }
            }
            __syncwarp();
        }
    }
}

// Every scale must carry `scat`; USL_ERR_UNSUPPORTED when a row is too wide
// for 32 private rows (the caller falls back to the warp-per-row kernel).
int cons_rows_launch(MultiCons* C, cudaStream_t st) {
    size_t smem = 0;
    C->cta_start[0] = 0;
    for (int k = 0; k < C->n; ++k) {
        ConsParams& c = C->P[k];
        if (!c.scat || ((uintptr_t)c.scat & 15)) return USL_ERR_ARG;
        const R4Geo g = r4_geometry(c.h, c.w, R4_SMEM_LIMIT);
        if (!g.nl) return USL_ERR_UNSUPPORTED;
        c.R = g.R;
        C->lanes[k] = g.nl;
        C->strips[k] = (c.h + c.R - 1) / c.R;
        C->cta_start[k + 1] = C->cta_start[k] + C->strips[k] * c.B;
        if (g.smem > smem) smem = g.smem;
    }
    // (the limit, not this launch's need: launches of several scales are in
    //  flight on different streams)
    if (cudaFuncSetAttribute(cons_rows_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             R4_SMEM_LIMIT) != cudaSuccess)
        return USL_ERR_CUDA;
    cons_rows_kernel<<<C->cta_start[C->n], R4_THREADS, smem, st>>>(*C);
    return check_launch();
}

}  // namespace usl
