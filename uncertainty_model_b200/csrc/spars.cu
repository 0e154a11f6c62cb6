// Kernel family (4): sparsification curves / AUSE
// (reference: train/sparsification.py:8-61).
//
//   curve(oracle, predicted): avg-pool k x k both maps, sort every row of the
//   pooled predicted error in descending order, gather the pooled oracle error
//   in that order, and report for 100 cut points the mean of the remaining
//   tail, normalised by the row mean and averaged over rows.
//
// Pipeline per call (rows = frames * 2 views, n = (H-k+1)*(W-k+1) per row):
//   pool      bit-exact restatement of ATen avg_pool2d: fp32 window sum in
//             row-major order, one true division by k*k.  The predicted map is
//             emitted as a 32-bit radix key whose ascending order is the
//             descending float order (-0.0 canonicalised to +0.0).
//   sort      segmented LSD radix sort, 4 passes of 8 bits, stable: equal keys
//             keep ascending index order, i.e. exactly
//             argsort(descending=True, stable=True).  Per pass: per-tile digit
//             histogram -> per-row exclusive scan -> ranked scatter.  Ranking
//             inside a tile is warp match.any based (atomic-free, order
//             preserving).  Payload: the pooled oracle value (and, on request,
//             the element index, which is the argsort itself).
//   tail      canonical fp64 segment sums between consecutive cut points
//             (256 lanes, lane-strided, stride-halving tree), suffix scan,
//             per-row normalisation, fixed-order accumulation over rows.
// The summation orders are the ones spelled out in oracle/spars_port.py, so
// keys, permutation, curve and AUSE are compared bit for bit.
//
// HBM-bound: algorithmic bytes are 16 B per pooled element (read two maps,
// write two orderings); the radix passes really move ~20 B/element/pass.
#include "usl_common.cuh"

namespace usl {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 4096
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int WARP_SPAN = SORT_TILE / SORT_WARPS;      // 512
constexpr int RADIX = 256;
constexpr int SEG_LANES = 256;

__device__ __forceinline__ unsigned desc_key(float v) {
    v = __fadd_rn(v, 0.0f);                     // -0.0 -> +0.0
    unsigned u = __float_as_uint(v);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // ascending float order
    return ~u;                                   // ... reversed
}

// the float a key stands for (desc_key is a bijection on non-NaN floats but for
// -0.0, which it canonicalises to +0.0 -- as a sum term the same thing)
__device__ __forceinline__ float key_value(unsigned k) {
    unsigned u = ~k;
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}

// ---- pool ---------------------------------------------------------------
// grid (column groups, blocks of PY output rows, maps).  Each thread produces a
// PX x PY patch of outputs, so a window row is loaded once -- as 128-bit loads
// where the row allows it (the loads of a warp then cover one contiguous span:
// this kernel is bound by L1 wavefronts otherwise) -- and feeds the PX
// accumulators of every output row whose window holds it; every accumulator
// still adds its k*k values one by one in row-major order (input rows arrive in
// increasing order, so each output row sees its window rows in order).
constexpr int PX = 8;
constexpr int PY = 4;
constexpr int KMAX = 15;
constexpr int POOL_THREADS = 128;
constexpr int PV = (KMAX + PX - 1 + 3) / 4 * 4;     // values held per window row

// K: the window size when it is known at compile time (11, the reference's
// default: no predicated-off adds), 0 = the run-time `kk` (any size to KMAX).
template <int K>
__global__ void __launch_bounds__(POOL_THREADS)
pool_kernel(const float* __restrict__ in, int H, int W, int kk, int vec_ok,
            float* __restrict__ out_val, unsigned* __restrict__ out_key) {
    const int k = K ? K : kk;
    const int oh = H - k + 1, ow = W - k + 1;
    const int ox = (blockIdx.x * POOL_THREADS + threadIdx.x) * PX;
    if (ox >= ow) return;
    const int oy0 = blockIdx.y * PY, row = blockIdx.z;
    const int ny = min(PY, oh - oy0);               // output rows of this patch
    const float div = (float)(k * k);
    const float* p = in + ((long long)row * H + oy0) * W + ox;
    const int need = k + PX - 1;
    const bool vec = vec_ok && ox + ((need + 3) & ~3) <= W;
    constexpr int KU = K ? K : KMAX;                // unrolled window columns
    float acc[PY][PX];
#pragma unroll
    for (int r = 0; r < PY; ++r)
#pragma unroll
        for (int j = 0; j < PX; ++j) acc[r][j] = 0.0f;
    for (int y = 0; y < ny + k - 1; ++y) {          // input row oy0 + y
        float v[PV];
        if (vec) {
#pragma unroll
            for (int j = 0; j < PV; j += 4) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < need) t = __ldg(reinterpret_cast<const float4*>(p + j));
                v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < PV; ++j)
                v[j] = (j < need && ox + j < W) ? __ldg(p + j) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < PY; ++r) {
            if (y - r < 0 || y - r >= k || r >= ny) continue;   // (uniform)
#pragma unroll
            for (int dx = 0; dx < KU; ++dx)
                if (K || dx < k) {
#pragma unroll
                    for (int j = 0; j < PX; ++j)
                        acc[r][j] = __fadd_rn(acc[r][j], v[dx + j]);
                }
        }
        p += W;
    }
#pragma unroll
    for (int r = 0; r < PY; ++r) {
        if (r >= ny) break;
        const long long o = ((long long)row * oh + oy0 + r) * ow + ox;
#pragma unroll
        for (int j = 0; j < PX; ++j)
            if (ox + j < ow) {
                const float m = __fdiv_rn(acc[r][j], div);
                if (out_val) out_val[o + j] = m;
                if (out_key) out_key[o + j] = desc_key(m);
            }
    }
}

__global__ void __launch_bounds__(256)
iota_kernel(int* idx, long long rows, int n) {
    const long long total = rows * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < total; i += stride)
        idx[i] = (int)(i % n);
}

// ---- sort ---------------------------------------------------------------
// counts[row][digit][tile]
__global__ void __launch_bounds__(SORT_THREADS)
hist_kernel(const unsigned* __restrict__ keys, int n, int ntiles, int shift,
            unsigned* __restrict__ counts) {
    __shared__ unsigned h[RADIX];
    const int row = blockIdx.y, tile = blockIdx.x;
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned* k = keys + (long long)row * n;
    const int base = tile * SORT_TILE;
    if (base + SORT_TILE <= n && (((uintptr_t)(k + base)) & 15) == 0) {
        // whole, 16-byte aligned tile: four 128-bit loads per thread (the order
        // of the items does not matter to a histogram)
        const uint4* k4 = reinterpret_cast<const uint4*>(k + base);
        uint4 v[SORT_ITEMS / 4];
#pragma unroll
        for (int j = 0; j < SORT_ITEMS / 4; ++j) v[j] = __ldg(k4 + j * SORT_THREADS + threadIdx.x);
#pragma unroll
        for (int j = 0; j < SORT_ITEMS / 4; ++j) {
            atomicAdd(&h[(v[j].x >> shift) & 255u], 1u);   // integer: exact
            atomicAdd(&h[(v[j].y >> shift) & 255u], 1u);
            atomicAdd(&h[(v[j].z >> shift) & 255u], 1u);
            atomicAdd(&h[(v[j].w >> shift) & 255u], 1u);
        }
    } else {
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; ++j) {
            const int i = base + j * SORT_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&h[(k[i] >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    counts[((long long)row * RADIX + threadIdx.x) * ntiles + tile] = h[threadIdx.x];
}

// Exclusive scan of the (digit-major, tile-minor) counts of every row, in two
// levels so that the whole GPU works on it: scan_tiles scans the tiles of one
// (row, digit) in place and leaves the digit's total; scan_digits turns the 256
// totals of a row into exclusive digit bases.  scatter_kernel adds the two.
__global__ void __launch_bounds__(256)
scan_tiles_kernel(unsigned* __restrict__ counts, int ntiles,
                  unsigned* __restrict__ totals) {
    __shared__ unsigned wsum[8];
    const int digit = blockIdx.x, row = blockIdx.y;
    unsigned* c = counts + ((long long)row * RADIX + digit) * ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned carry = 0;
    for (int base = 0; base < ntiles; base += 256) {
        const int i = base + threadIdx.x;
        const unsigned v = i < ntiles ? c[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned t = wsum[w];
            if (w < warp) wbase += t;
            total += t;
        }
        if (i < ntiles) c[i] = carry + wbase + incl - v;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[row * RADIX + digit] = carry;
}

__global__ void __launch_bounds__(RADIX)
scan_digits_kernel(unsigned* __restrict__ totals) {
    __shared__ unsigned wsum[RADIX / 32];
    unsigned* t = totals + (long long)blockIdx.x * RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned v = t[threadIdx.x];
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned wbase = 0;
#pragma unroll
    for (int w = 0; w < RADIX / 32; ++w)
        if (w < warp) wbase += wsum[w];
    t[threadIdx.x] = wbase + incl - v;
}

// The lanes of the warp whose 8-bit digit equals this lane's, from eight votes
// (match.any issues once per ~67 cycles and scheduler: slower than the votes).
__device__ __forceinline__ unsigned same_digit_lanes(unsigned d, bool valid) {
    unsigned m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned v = __ballot_sync(0xffffffffu, bit);
        m &= bit ? v : ~v;
    }
    return m;
}

// One pass of the LSD sort for one tile: rank every item among the items of
// its digit (warp by warp, in item order: the pass is stable), put the tile in
// digit order in shared memory, and copy it out -- consecutive threads then
// write consecutive words of a digit's run (a scattered 4-byte store per lane
// costs a whole sector of load/store-unit time each).
// PAYLOAD: 0 = keys only (the oracle curve ranks the pooled map by itself: the
// value travels inside its key), 1 = values, 2 = values and element indices.
// (five CTAs per SM: the pass waits on global loads most of the time, 40 warps
//  hide more of that than the 24 its natural 80 registers allow -- 346 -> ~290
//  us per pass in spite of ~150 bytes of spills)
template <int PAYLOAD>
__global__ void __launch_bounds__(SORT_THREADS, 5)
scatter_kernel(const unsigned* __restrict__ keys_in,
               const float* __restrict__ vals_in, const int* __restrict__ idx_in,
               unsigned* __restrict__ keys_out, float* __restrict__ vals_out,
               int* __restrict__ idx_out, const unsigned* __restrict__ offsets,
               const unsigned* __restrict__ digit_base, int n, int ntiles,
               int shift) {
    __shared__ unsigned whist[SORT_WARPS][RADIX];
    __shared__ unsigned skeys[SORT_TILE];
    __shared__ unsigned sdata[SORT_TILE];      // payload words (values, then indices)
    __shared__ unsigned delta[RADIX];          // global position - position in the tile
    __shared__ unsigned wsum[SORT_WARPS];
    const int row = blockIdx.y, tile = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS)
        (&whist[0][0])[i] = 0;
    __syncthreads();
    const long long rbase = (long long)row * n;
    const int tbase = tile * SORT_TILE;
    const int wbase = tbase + warp * WARP_SPAN;
    const int cnt = min(SORT_TILE, n - tbase);
    unsigned key[SORT_ITEMS];
    unsigned short rank[SORT_ITEMS];
#pragma unroll
    for (int c = 0; c < SORT_ITEMS; ++c) {
        const int i = wbase + c * 32 + lane;
        key[c] = i < n ? keys_in[rbase + i] : 0u;
    }
    if (PAYLOAD) {
        // the payload is read after the ranking: have it in L2 by then (a line a lane)
        const int i = wbase + lane * 32;
        if (lane < SORT_ITEMS && i < n)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vals_in + rbase + i));
        if (PAYLOAD == 2 && lane < SORT_ITEMS && i < n)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(idx_in + rbase + i));
    }
#pragma unroll
    for (int c = 0; c < SORT_ITEMS; ++c) {
        const bool valid = wbase + c * 32 + lane < n;
        const unsigned d = (key[c] >> shift) & 255u;
        const unsigned grp = same_digit_lanes(d, valid);
        const unsigned prev = whist[warp][d];
        __syncwarp();
        if (valid && (__ffs(grp) - 1) == lane) whist[warp][d] = prev + __popc(grp);
        __syncwarp();
        rank[c] = (unsigned short)(prev + __popc(grp & lt));
    }
    __syncthreads();
    {   // digit threadIdx.x: exclusive bases over the warps of this tile, then
        // over the digits (block scan); delta = where the digit's run goes
        const int d = threadIdx.x;
        unsigned acc = 0;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            const unsigned t = whist[w][d];
            whist[w][d] = acc;
            acc += t;
        }
        unsigned incl = acc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned start = incl - acc;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w)
            if (w < warp) start += wsum[w];
        delta[d] = offsets[((long long)row * RADIX + d) * ntiles + tile] +
                   digit_base[row * RADIX + d] - start;
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) whist[w][d] += start;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < SORT_ITEMS; ++c) {      // rank -> position in the tile
        const int i = wbase + c * 32 + lane;
        const unsigned d = (key[c] >> shift) & 255u;
        rank[c] = (unsigned short)(whist[warp][d] + rank[c]);
        if (i < n) {
            skeys[rank[c]] = key[c];
            if (PAYLOAD) sdata[rank[c]] = __float_as_uint(vals_in[rbase + i]);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += SORT_THREADS) {
        const unsigned k = skeys[j];
        const long long pos = rbase + delta[(k >> shift) & 255u] + j;
        keys_out[pos] = k;
        if (PAYLOAD) vals_out[pos] = __uint_as_float(sdata[j]);
    }
    if (PAYLOAD == 2) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < SORT_ITEMS; ++c) {
            const int i = wbase + c * 32 + lane;
            if (i < n) sdata[rank[c]] = (unsigned)idx_in[rbase + i];
        }
        __syncthreads();
        for (int j = threadIdx.x; j < cnt; j += SORT_THREADS) {
            const long long pos = rbase + delta[(skeys[j] >> shift) & 255u] + j;
            idx_out[pos] = (int)sdata[j];
        }
    }
}

// ---- tail ---------------------------------------------------------------
struct Cuts { int v[USL_MAX_STEPS + 1]; };

// canonical fp64 sum of sorted values with rank in [cut_k, cut_{k+1})
// (KEYS: the sorted values are read back out of the sorted keys)
template <bool KEYS>
__global__ void __launch_bounds__(SEG_LANES)
segment_kernel(const float* __restrict__ vals, const unsigned* __restrict__ keys, int n,
               const Cuts cuts, int steps, double* __restrict__ seg) {
    __shared__ double lanes[SEG_LANES];
    const int k = blockIdx.x, row = blockIdx.y;
    const float* v = vals + (long long)row * n;
    const unsigned* kk = keys + (long long)row * n;
    double a = 0.0;
    for (int i = cuts.v[k] + threadIdx.x; i < cuts.v[k + 1]; i += SEG_LANES)
        a = __dadd_rn(a, (double)(KEYS ? key_value(kk[i]) : v[i]));
    lanes[threadIdx.x] = a;
    __syncthreads();
    for (int s = SEG_LANES / 2; s >= 1; s >>= 1) {
        if (threadIdx.x < s)
            lanes[threadIdx.x] = __dadd_rn(lanes[threadIdx.x], lanes[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) seg[(long long)row * steps + k] = lanes[0];
}

// seg[row][*] -> normalised tail means, in place
__global__ void row_norm_kernel(double* seg, int rows, int n, const Cuts cuts,
                                int steps) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    double* s = seg + (long long)row * steps;
    double run = 0.0;
    for (int k = steps - 1; k >= 0; --k) {
        run = (k == steps - 1) ? s[k] : __dadd_rn(s[k], run);
        s[k] = run;
    }
    const double mean = __ddiv_rn(s[0], (double)n);
    for (int k = 0; k < steps; ++k)
        s[k] = __ddiv_rn(__ddiv_rn(s[k], (double)(n - cuts.v[k])), mean);
}

// row_norm_sum[k] += sum over rows, in row order
__global__ void accumulate_kernel(const double* norm, int rows, int steps,
                                  double* row_norm_sum) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= steps) return;
    double a = row_norm_sum[k];
    for (int r = 0; r < rows; ++r) a = __dadd_rn(a, norm[(long long)r * steps + k]);
    row_norm_sum[k] = a;
}

__global__ void finish_kernel(const double* row_norm_sum, int steps,
                              double total_rows, float* curve) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < steps) curve[k] = (float)__ddiv_rn(row_norm_sum[k], total_rows);
}

__global__ void ause_kernel(const float* oracle, const float* pred, int steps,
                            float* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a = 0.0;
    for (int k = 0; k < steps; ++k)
        a = __dadd_rn(a, (double)__fsub_rn(pred[k], oracle[k]));
    *out = (float)__ddiv_rn(a, (double)steps);
}

// ---- host ----------------------------------------------------------------
struct SparsLayout {
    size_t keys[2], vals[2], idx[2], counts, seg, total;
    int n, ntiles;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static SparsLayout layout(int rows, int H, int W, int k, bool with_idx) {
    SparsLayout L;
    L.n = (H - k + 1) * (W - k + 1);
    L.ntiles = (L.n + SORT_TILE - 1) / SORT_TILE;
    const size_t e = (size_t)rows * L.n;
    size_t off = 0;
    for (int i = 0; i < 2; ++i) { L.keys[i] = off; off = align_up(off + e * 4); }
    for (int i = 0; i < 2; ++i) { L.vals[i] = off; off = align_up(off + e * 4); }
    for (int i = 0; i < 2; ++i) {
        L.idx[i] = off;
        if (with_idx) off = align_up(off + e * 4);
    }
    L.counts = off; off = align_up(off + (size_t)rows * RADIX * (L.ntiles + 1) * 4);
    L.seg = off; off = align_up(off + (size_t)rows * USL_MAX_STEPS * 8);
    L.total = off;
    return L;
}

static unsigned flat_grid(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace usl

using namespace usl;

extern "C" size_t usl_spars_workspace_bytes(int rows, int H, int W, int k,
                                            int with_order) {
    if (rows <= 0 || k < 1 || H < k || W < k) return 0;
    return layout(rows, H, W, k, with_order != 0).total;
}

extern "C" int usl_spars_curve(const float* oracle, const float* predicted,
                               int rows, int H, int W, int k, const int* cuts,
                               int steps, double* row_norm_sum,
                               int32_t* order_out, float* pooled_oracle_out,
                               float* pooled_pred_out, void* workspace,
                               size_t workspace_bytes, void* stream_) {
    if (!oracle || !predicted || !cuts || !row_norm_sum || !workspace ||
        rows <= 0 || steps < 1)
        return USL_ERR_ARG;
    if (k < 1 || k > KMAX || H < k || W < k || steps > USL_MAX_STEPS)
        return USL_ERR_UNSUPPORTED;
    DeviceGuard guard(oracle);
    if ((long long)rows * (H - k + 1) * (W - k + 1) >= (1ll << 31))
        return USL_ERR_UNSUPPORTED;
    const bool with_idx = order_out != nullptr;
    const SparsLayout L = layout(rows, H, W, k, with_idx);
    if (workspace_bytes < L.total) return USL_ERR_WORKSPACE;
    Cuts c;
    for (int i = 0; i <= steps; ++i) {
        c.v[i] = cuts[i];
        if (cuts[i] < 0 || cuts[i] > L.n || (i && cuts[i] < cuts[i - 1]))
            return USL_ERR_ARG;
    }
    if (c.v[steps] != L.n) return USL_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    char* ws = (char*)workspace;
    unsigned* keys[2] = {(unsigned*)(ws + L.keys[0]), (unsigned*)(ws + L.keys[1])};
    float* vals[2] = {(float*)(ws + L.vals[0]), (float*)(ws + L.vals[1])};
    int* idx[2] = {(int*)(ws + L.idx[0]), (int*)(ws + L.idx[1])};
    unsigned* counts = (unsigned*)(ws + L.counts);
    double* seg = (double*)(ws + L.seg);
    const int n = L.n, ntiles = L.ntiles;

    unsigned* digit_base = counts + (size_t)rows * RADIX * ntiles;
    // the oracle curve: the payload of a key is the float it encodes (the one
    // difference, -0.0 -> +0.0, is no difference to a sum): keys travel alone,
    // unless a caller wants the pooled values / the permutation as well
    const bool key_only = predicted == oracle && !with_idx && !pooled_oracle_out &&
                          !pooled_pred_out;
    {
        const int oh = H - k + 1, ow = W - k + 1;
        if (oh > 65535 || rows > 65535) return USL_ERR_UNSUPPORTED;
        const dim3 pgrid(((ow + PX - 1) / PX + POOL_THREADS - 1) / POOL_THREADS,
                         (oh + PY - 1) / PY, rows);
        const int vo = (W % 4 == 0) && (((uintptr_t)oracle & 15) == 0);
        const int vp = (W % 4 == 0) && (((uintptr_t)predicted & 15) == 0);
        auto pool = k == 11 ? pool_kernel<11> : pool_kernel<0>;
        if (predicted == oracle) {
            // the oracle curve (evaluate.py:155: curve(error, error)) ranks the
            // map by itself: one pooling pass yields payload and keys
            pool<<<pgrid, POOL_THREADS, 0, stream>>>(oracle, H, W, k, vo,
                                                     key_only ? nullptr : vals[0], keys[0]);
            count_launches(1);
            if (pooled_pred_out &&
                cudaMemcpyAsync(pooled_pred_out, vals[0], (size_t)rows * n * 4,
                                cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
                return USL_ERR_CUDA;
        } else {
            pool<<<pgrid, POOL_THREADS, 0, stream>>>(oracle, H, W, k, vo, vals[0], nullptr);
            pool<<<pgrid, POOL_THREADS, 0, stream>>>(predicted, H, W, k, vp,
                                                     pooled_pred_out, keys[0]);
            count_launches(2);
        }
    }
    if (pooled_oracle_out)
        if (cudaMemcpyAsync(pooled_oracle_out, vals[0], (size_t)rows * n * 4,
                            cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
            return USL_ERR_CUDA;
    if (with_idx) {
        iota_kernel<<<flat_grid((long long)rows * n), 256, 0, stream>>>(
            idx[0], rows, n);
        count_launches(1);
    }
    const dim3 grid(ntiles, rows);
    int cur = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = pass * 8;
        hist_kernel<<<grid, SORT_THREADS, 0, stream>>>(keys[cur], n, ntiles,
                                                       shift, counts);
        scan_tiles_kernel<<<dim3(RADIX, rows), 256, 0, stream>>>(counts, ntiles, digit_base);
        scan_digits_kernel<<<rows, RADIX, 0, stream>>>(digit_base);
        if (with_idx)
            scatter_kernel<2><<<grid, SORT_THREADS, 0, stream>>>(
                keys[cur], vals[cur], idx[cur], keys[cur ^ 1], vals[cur ^ 1],
                idx[cur ^ 1], counts, digit_base, n, ntiles, shift);
        else if (key_only)
            scatter_kernel<0><<<grid, SORT_THREADS, 0, stream>>>(
                keys[cur], nullptr, nullptr, keys[cur ^ 1], nullptr,
                nullptr, counts, digit_base, n, ntiles, shift);
        else
            scatter_kernel<1><<<grid, SORT_THREADS, 0, stream>>>(
                keys[cur], vals[cur], nullptr, keys[cur ^ 1], vals[cur ^ 1],
                nullptr, counts, digit_base, n, ntiles, shift);
        cur ^= 1;
        count_launches(3);
    }
    if (with_idx)
        if (cudaMemcpyAsync(order_out, idx[cur], (size_t)rows * n * 4,
                            cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
            return USL_ERR_CUDA;
    if (key_only)
        segment_kernel<true><<<dim3(steps, rows), SEG_LANES, 0, stream>>>(
            vals[cur], keys[cur], n, c, steps, seg);
    else
        segment_kernel<false><<<dim3(steps, rows), SEG_LANES, 0, stream>>>(
            vals[cur], keys[cur], n, c, steps, seg);
    count_launches(2);
    row_norm_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(seg, rows, n, c, steps);
    accumulate_kernel<<<(steps + 127) / 128, 128, 0, stream>>>(seg, rows, steps,
                                                              row_norm_sum);
    return check_launch();
}

extern "C" int usl_spars_finish(const double* row_norm_sum, int steps,
                                long long total_rows, float* curve,
                                void* stream) {
    if (!row_norm_sum || !curve || steps < 1 || total_rows < 1) return USL_ERR_ARG;
    DeviceGuard guard(row_norm_sum);
    finish_kernel<<<(steps + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        row_norm_sum, steps, (double)total_rows, curve);
    return check_launch();
}

extern "C" int usl_spars_ause(const float* oracle_curve, const float* pred_curve,
                              int steps, float* out, void* stream) {
    if (!oracle_curve || !pred_curve || !out || steps < 1) return USL_ERR_ARG;
    DeviceGuard guard(out);
    ause_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(oracle_curve, pred_curve,
                                                    steps, out);
    return check_launch();
}
