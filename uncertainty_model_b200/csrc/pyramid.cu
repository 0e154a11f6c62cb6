// Kernel (3) of the path: input pyramid (reference: train/utils.py:27-50).
//
// Every level is resampled from the full-resolution stereo pair with
// align_corners=True bilinear taps; level 0 is the input itself and is
// aliased by the host, never copied.  All levels >= 1 are produced by one
// launch, largest level first: a CTA owns a tile of consecutive output rows of
// one (level, sample, channel) plane; a thread keeps one output column (its
// horizontal taps are computed once) and walks down the rows of the tile, so
// an output costs ~30 instructions and the four taps of a warp read two
// contiguous runs of the source rows.
// HBM-bound: 24 B read + 7.875 B written per full-resolution pixel.
#include "usl_common.cuh"

namespace usl {

constexpr int PYR_THREADS = 256;
constexpr int PYR_ROWS = 8;        // output rows per thread

struct PyramidParams {
    const float* src;
    long long src_bs, src_cs;
    int B, C, H, W;
    int levels;                 // number of produced levels (scales - 1)
    float* dst[USL_MAX_SCALES];
    int h[USL_MAX_SCALES], w[USL_MAX_SCALES];
    float sy[USL_MAX_SCALES], sx[USL_MAX_SCALES];
    int rows[USL_MAX_SCALES];      // output rows per CTA
    int tiles[USL_MAX_SCALES];     // CTAs per plane
    int cta_start[USL_MAX_SCALES + 1];
};

__global__ void __launch_bounds__(PYR_THREADS)
pyramid_kernel(const __grid_constant__ PyramidParams p) {
    int l = 0;
    while (l + 1 < p.levels && (int)blockIdx.x >= p.cta_start[l + 1]) ++l;
    const int local = blockIdx.x - p.cta_start[l];
    const int plane_i = local / p.tiles[l];
    const int tile = local - plane_i * p.tiles[l];
    const int b = plane_i / p.C, c = plane_i - b * p.C;
    const int w = p.w[l], h = p.h[l];
    const float* s = p.src + b * p.src_bs + c * p.src_cs;
    float* d = p.dst[l] + (long long)plane_i * h * w;
    const int y0 = tile * p.rows[l];
    const int y1 = min(h, y0 + p.rows[l]);
    const float sy = p.sy[l];
    if (w <= PYR_THREADS) {
        const int rstep = PYR_THREADS / w;          // rows the block covers at once
        const int ry = threadIdx.x / w, x = threadIdx.x - ry * w;
        if (ry >= rstep) return;
        const TapAC tx = ac_taps(x, p.sx[l], p.W);
        const float* s0 = s + tx.i0;
        const float* s1 = s + tx.i1;
        for (int y = y0 + ry; y < y1; y += rstep) {
            const TapAC ty = ac_taps(y, sy, p.H);
            const long long o0 = (long long)ty.i0 * p.W, o1 = (long long)ty.i1 * p.W;
            const float v00 = __ldg(s0 + o0), v01 = __ldg(s1 + o0);
            const float v10 = __ldg(s0 + o1), v11 = __ldg(s1 + o1);
            // ATen: w0y * (w0x * v00 + w1x * v01) + w1y * (w0x * v10 + w1x * v11)
            const float top = tx.w0 * v00 + tx.w1 * v01;
            const float bot = tx.w0 * v10 + tx.w1 * v11;
            d[(long long)y * w + x] = ty.w0 * top + ty.w1 * bot;
        }
    } else {
        for (int y = y0; y < y1; ++y) {
            const TapAC ty = ac_taps(y, sy, p.H);
            const float* r0 = s + (long long)ty.i0 * p.W;
            const float* r1 = s + (long long)ty.i1 * p.W;
            for (int x = threadIdx.x; x < w; x += PYR_THREADS) {
                const TapAC tx = ac_taps(x, p.sx[l], p.W);
                const float v00 = __ldg(r0 + tx.i0), v01 = __ldg(r0 + tx.i1);
                const float v10 = __ldg(r1 + tx.i0), v11 = __ldg(r1 + tx.i1);
                const float top = tx.w0 * v00 + tx.w1 * v01;
                const float bot = tx.w0 * v10 + tx.w1 * v11;
                d[(long long)y * w + x] = ty.w0 * top + ty.w1 * bot;
            }
        }
    }
}

}  // namespace usl

extern "C" int usl_pyramid(const float* src, int B, int C, int H, int W,
                           long long src_bs, long long src_cs, int scales,
                           float* const* dst, void* stream) {
    using namespace usl;
    if (!src || !dst || B <= 0 || C <= 0 || H <= 0 || W <= 0 || scales < 1 ||
        scales > USL_MAX_SCALES)
        return USL_ERR_ARG;
    if (scales == 1) return USL_OK;
    DeviceGuard guard(src);
    PyramidParams p;
    p.src = src; p.src_bs = src_bs; p.src_cs = src_cs;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.levels = scales - 1;
    p.cta_start[0] = 0;
    for (int l = 0; l < p.levels; ++l) {
        const int h = H >> (l + 1), w = W >> (l + 1);
        if (h < 1 || w < 1 || !dst[l + 1]) return USL_ERR_ARG;
        p.dst[l] = dst[l + 1];
        p.h[l] = h; p.w[l] = w;
        p.sy[l] = ac_scale(H, h);
        p.sx[l] = ac_scale(W, w);
        p.rows[l] = (w <= PYR_THREADS ? PYR_THREADS / w : 1) * PYR_ROWS;
        p.tiles[l] = (h + p.rows[l] - 1) / p.rows[l];
        const long long ctas = (long long)B * C * p.tiles[l];
        if (p.cta_start[l] + ctas > 0x7fffffffLL) return USL_ERR_UNSUPPORTED;
        p.cta_start[l + 1] = p.cta_start[l] + (int)ctas;
    }
    mark(0, (cudaStream_t)stream);
    pyramid_kernel<<<(unsigned)p.cta_start[p.levels], PYR_THREADS, 0,
                     (cudaStream_t)stream>>>(p);
    mark(1, (cudaStream_t)stream);
    return check_launch();
}
