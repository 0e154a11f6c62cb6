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
// load-add-store per step.  So everything that is not that chain is taken out
// of it: the column kernels leave one 16-byte element {ix_dd, ix_ud, s_dd, s_ud} per
// pixel behind (ConsParams::scat: the sampling columns of the two terms and
// s = coef * upstream * sign(a - warp(b)), the one thing that cannot be
// recomputed without the blended row) and the rows of
// a 16-column tile arrive as bulk async copies (one per row, UBLKCP) on a
// 4-stage mbarrier ring fed by the warps that do not walk; per 8 columns the
// sampling columns, weights and addresses of both terms are computed first
// (16 independent chains), then the 8 read-modify-writes run back to back, the
// two terms of a column as ONE chain (loads, forwarding of the first term's
// sums where the taps coincide, stores in order).
//
// CTA = destination rows [ya, yb) of one sample; lane t = rsi * 2 + v walks
// source row rs0 + rsi of view v.  At the end all warps assemble the destination
// rows from the three source rows around each, with the vertical tap weights,
// and add them to the gradient the column kernels stored.
#include "cons_core.cuh"
#include "cons_launch.cuh"
#include "usl_common.cuh"

namespace usl {

constexpr int R4_THREADS = 256;
constexpr int R4_TILE = 16;                 // columns per staged tile
constexpr int R4_PITCH = R4_TILE + 1;       // 16-byte elements per staged row (odd: the
                                            // column walk is bank-conflict free)
constexpr int R4_STAGES = 4;
constexpr int R4_PAD = 2;                   // row = [2 | w | 2 (+1)]: taps outside the
                                            // image land in the pads
constexpr int R4_BATCH = 8;

// floats per private row: odd, so that lanes at the same column (smooth
// disparities) sit in 32 different banks
__host__ __device__ inline int r4_pitch(int w) { return (w + 2 * R4_PAD) | 1; }

struct R4Geo { int nl, R; size_t smem; };

__host__ __device__ inline size_t r4_smem(int nl, int w) {
    return (size_t)nl * r4_pitch(w) * 4 +
           (size_t)R4_STAGES * nl * R4_PITCH * 16 + 64;
}

// walking lanes (32 .. 128) and strip height of a scale; nl = 0 if even 32
// private rows do not fit
static R4Geo r4_geometry(int h, int w, size_t budget) {
    R4Geo g; g.nl = 0; g.R = 0; g.smem = 0;
    for (int nl = 128; nl >= 32; nl -= 32) {
        if (r4_smem(nl, w) > budget) continue;
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
__device__ __forceinline__ void mb_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(s_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_row(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

// floor() and the float -> int conversion with additions only (F2I / FRND run
// at a quarter of the rate, and a walking warp is alone on its scheduler):
// adding 1.5 * 2^23 rounds to the nearest integer and leaves it in the low
// mantissa bits; |x| < 2^22 here (columns).
constexpr float R4_MAGIC = 12582912.0f;
__device__ __forceinline__ float r4_floor(float x) {
    const float r = (x + R4_MAGIC) - R4_MAGIC;      // nearest integer
    return r > x ? r - 1.0f : r;
}
__device__ __forceinline__ int r4_int(float integral) {
    return __float_as_int(integral + R4_MAGIC) - 0x4B400000;
}

// column offset (from column 0 of the row) and the two tap contributions of a
// pixel with sampling column ix and signed coefficient s
__device__ __forceinline__ void r4_decode(float ix, float s, float fw, int& off,
                                          float& c0, float& c1) {
    const float f = r4_floor(ix);
    const float w1 = ix - f, w0 = (f + 1.0f) - ix;
    // columns -2 .. w+1 exist: taps outside the image land in pads nobody reads
    off = r4_int(fminf(fmaxf(f, -2.0f), fw));
    c0 = s * w0;
    c1 = s * w1;
}

struct R4Batch {
    int od[R4_BATCH], ou[R4_BATCH];
    float cd0[R4_BATCH], cd1[R4_BATCH], cu0[R4_BATCH], cu1[R4_BATCH];
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
    float* H = reinterpret_cast<float*>(stage + (size_t)R4_STAGES * nl * R4_PITCH);  // [nl][HW]
    uint64_t* full = reinterpret_cast<uint64_t*>(H + (size_t)nl * HW);  // [STAGES]
    uint64_t* empty = full + R4_STAGES;                                 // [STAGES]
    const long long hw = (long long)h * w;
    const int nwalk = nl >> 5;                  // walking warps: 0 .. nwalk-1
    const int ntile = (w + R4_TILE - 1) / R4_TILE;

    for (int i = tid; i < nl * HW; i += R4_THREADS) H[i] = 0.0f;
    // rows past the strip are never copied: their lanes must read zeros
    for (int i = tid; i < R4_STAGES * nl * R4_PITCH; i += R4_THREADS)
        stage[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        // (every feeding warp that holds rows arrives once per tile, with its bytes)
        const int nrows0 = (rs1 - rs0 + 1) * 2;
        const int feeders = min((nrows0 + 31) / 32, R4_THREADS / 32 - nwalk);
        for (int k = 0; k < R4_STAGES; ++k) { mb_init(full + k, feeders); mb_init(empty + k, nwalk); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp >= nwalk) {
        // ---- the feeding warps: one bulk copy per (row, view) and tile; a copy
        // is issued by one lane at a time (UBLKCP takes uniform operands), so the
        // rows are dealt out over all the warps that do not walk, one per lane --
        const int nrows = (rs1 - rs0 + 1) * 2;
        const int t = tid - nl;                           // this lane's row, if any
        const bool mine = t < nrows;
        const unsigned active = __ballot_sync(0xffffffffu, mine);
        if (active) {
            const int rs = rs0 + (t >> 1), v = t & 1;
            const float4* src = reinterpret_cast<const float4*>(P.scat) +
                                ((long long)b * 2 + v) * hw + (long long)rs * w;
            const int mycount = __popc(active);
            for (int k = 0; k < ntile; ++k) {
                const int st = k % R4_STAGES;
                if (k >= R4_STAGES) mb_wait(empty + st, ((k / R4_STAGES) - 1) & 1);
                const int c0 = k * R4_TILE;
                const uint32_t bytes = (uint32_t)min(R4_TILE, w - c0) * 16u;
                if (lane == 0) mb_expect_tx(full + st, bytes * (uint32_t)mycount);
                __syncwarp();
                if (mine)
                    bulk_row(stage + ((size_t)st * nl + t) * R4_PITCH, src + c0, bytes,
                             full + st);
            }
        }
    } else if (warp < nwalk) {
        // ---- the walk ---------------------------------------------------------
        // Software pipeline over batches of 8 columns: the read-modify-writes
        // of batch i are interleaved, instruction by instruction, with the
        // staged loads and the decoding of batch i + 1 (independent work that
        // fills the load-add-store bubbles of the one dependent chain).
        float* Hrow = H + (size_t)tid * HW + R4_PAD;
        const float fw = (float)w;
        constexpr int BPT = R4_TILE / R4_BATCH;              // batches per tile
        const int nbatch = (w + R4_BATCH - 1) / R4_BATCH;

        // staged loads + decode of batch `bi` (waits for its tile first;
        // releases the tile after its last batch)
        // (a term that is off has s = 0)
        auto decode = [&](int bi, R4Batch& D, int u) {
            const int k = bi / BPT, st = k % R4_STAGES;
            const int j = (bi - k * BPT) * R4_BATCH + u;
            const float4 q = stage[((size_t)st * nl + tid) * R4_PITCH + j];
            r4_decode(q.x, q.z, fw, D.od[u], D.cd0[u], D.cd1[u]);
            r4_decode(q.y, q.w, fw, D.ou[u], D.cu0[u], D.cu1[u]);
        };
        auto acquire = [&](int bi) {
            if (bi < nbatch && bi % BPT == 0) {
                const int k = bi / BPT;
                mb_wait(full + k % R4_STAGES, (k / R4_STAGES) & 1);
            }
        };
        auto release = [&](int bi) {
            if (bi < nbatch && (bi % BPT == BPT - 1 || bi == nbatch - 1)) {
                __syncwarp();
                if (lane == 0) mb_arrive(empty + (bi / BPT) % R4_STAGES);
            }
        };
        // both terms of a column as ONE chain: all loads, the sums of term dd
        // forwarded where the taps of term ud coincide with them, stores in order
        // (explicit shared-memory instructions: the four loads must all be
        //  issued before the arithmetic -- left to itself the compiler sinks
        //  the second pair behind a branch on `diff`, two round trips a column)
        const uint32_t hrow_a = s_u32(Hrow);
        auto rmw = [&](const R4Batch& D, int u) {
            const uint32_t pd = hrow_a + 4u * (uint32_t)D.od[u];
            const uint32_t pu = hrow_a + 4u * (uint32_t)D.ou[u];
            float d0, d1, u0, u1;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d0) : "r"(pd));
            asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(d1) : "r"(pd));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(u0) : "r"(pu));
            asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(u1) : "r"(pu));
            const int diff = D.ou[u] - D.od[u];
            d0 -= D.cd0[u]; d1 -= D.cd1[u];
            u0 = diff == 0 ? d0 : u0;
            u0 = diff == 1 ? d1 : u0;
            u1 = diff == 0 ? d1 : u1;
            u1 = diff == -1 ? d0 : u1;
            u0 -= D.cu0[u]; u1 -= D.cu1[u];
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pd), "f"(d0));
            asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(pd), "f"(d1));
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pu), "f"(u0));
            asm volatile("st.shared.f32 [%0+4], %1;" ::"r"(pu), "f"(u1));
        };
        // full batches run in the pipeline; a last partial one (row width not a
        // multiple of 8: its slot holds stale elements past the end of the row)
        // is decoded with a guard and scattered on its own
        const int nfull = w / R4_BATCH;
        R4Batch A, B;
        if (nfull > 0) {
            acquire(0);
#pragma unroll
            for (int u = 0; u < R4_BATCH; ++u) decode(0, A, u);
            release(0);
        }
#pragma unroll 1
        for (int bi = 0; bi < nfull; bi += 2) {
            if (bi + 1 < nfull) {
                acquire(bi + 1);
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) { rmw(A, u); decode(bi + 1, B, u); }
                release(bi + 1);
            } else {
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) rmw(A, u);
                break;
            }
            if (bi + 2 < nfull) {
                acquire(bi + 2);
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) { rmw(B, u); decode(bi + 2, A, u); }
                release(bi + 2);
            } else {
#pragma unroll
                for (int u = 0; u < R4_BATCH; ++u) rmw(B, u);
            }
        }
        if (nfull < nbatch) {
            acquire(nfull);
#pragma unroll 1
            for (int u = 0; u < w - nfull * R4_BATCH; ++u) {
                const int k = nfull / BPT, st = k % R4_STAGES;
                const int j = (nfull - k * BPT) * R4_BATCH + u;
                const float4 q = stage[((size_t)st * nl + tid) * R4_PITCH + j];
                r4_decode(q.x, q.z, fw, A.od[0], A.cd0[0], A.cd1[0]);
                r4_decode(q.y, q.w, fw, A.ou[0], A.cu0[0], A.cu1[0]);
                rmw(A, 0);
            }
            release(nfull);
        }
    }
    __syncthreads();

    // ---- destination rows: H(y'-1), H(y'), H(y'+1) with the vertical weights ----
    for (int yi = warp; yi < yb - ya; yi += R4_THREADS / 32) {
        const int yd = ya + yi;
        float wgt[3];
        int row[3];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const int rs = yd - 1 + kk;
            wgt[kk] = 0.0f;
            row[kk] = 0;
            if (rs < 0 || rs >= h) continue;
            row[kk] = rs - rs0;
            const Tap2 ty = warp_row_taps(rs, h);
            if (ty.i0 == yd) wgt[kk] = ty.w0;
            else if (ty.i0 + 1 == yd) wgt[kk] = ty.w1;
        }
        for (int o = 0; o < 2; ++o) {
            // (source view 1 - o scatters into destination view o)
            const float* h0 = H + ((size_t)row[0] * 2 + (1 - o)) * HW + R4_PAD;
            const float* h1 = H + ((size_t)row[1] * 2 + (1 - o)) * HW + R4_PAD;
            const float* h2 = H + ((size_t)row[2] * 2 + (1 - o)) * HW + R4_PAD;
            float* out = P.grad_disp + (long long)b * P.gd_bs + o * P.gd_cs + (long long)yd * w;
            for (int x0 = lane; x0 < w; x0 += 256) {
                float old[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u;
                    old[u] = (P.accumulate && x < w) ? out[x] : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u;
                    if (x < w)
                        out[x] = old[u] + (wgt[0] * h0[x] + wgt[1] * h1[x] + wgt[2] * h2[x]);
                }
            }
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
        const R4Geo g = r4_geometry(c.h, c.w, 210 * 1024);
        if (!g.nl || 2 * g.nl > R4_THREADS) return USL_ERR_UNSUPPORTED;   // one feeding lane per row
        c.R = g.R;
        C->lanes[k] = g.nl;
        C->strips[k] = (c.h + c.R - 1) / c.R;
        C->cta_start[k + 1] = C->cta_start[k] + C->strips[k] * c.B;
        if (g.smem > smem) smem = g.smem;
    }
    if (cudaFuncSetAttribute(cons_rows_kernel,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return USL_ERR_CUDA;
    cons_rows_kernel<<<C->cta_start[C->n], R4_THREADS, smem, st>>>(*C);
    return check_launch();
}

}  // namespace usl
