// Kernel (3) of the path: input pyramid (reference: train/utils.py:27-50).
//
// Every level is resampled from the full-resolution stereo pair with
// align_corners=True bilinear taps; level 0 is the input itself and is
// aliased by the host, never copied.  All levels >= 1 are produced by one
// launch: the flat output index space of the levels is concatenated and a
// grid-stride loop walks it, so the full-resolution rows a CTA touches for
// level 1 are the same rows (hot in L1/L2) its neighbours touch for levels 2
// and 3.  HBM-bound: 24 B read + 7.875 B written per full-resolution pixel.
#include "usl_common.cuh"

namespace usl {

struct PyramidParams {
    const float* src;
    long long src_bs, src_cs;
    int B, C, H, W;
    int levels;                 // number of produced levels (scales - 1)
    float* dst[USL_MAX_SCALES];
    int h[USL_MAX_SCALES], w[USL_MAX_SCALES];
    float sy[USL_MAX_SCALES], sx[USL_MAX_SCALES];
    long long start[USL_MAX_SCALES + 1];   // prefix of per-level element counts
};

__global__ void __launch_bounds__(256)
pyramid_kernel(const PyramidParams p) {
    const long long total = p.start[p.levels];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < total; i += stride) {
        int l = 0;
        while (l + 1 < p.levels && i >= p.start[l + 1]) ++l;
        long long r = i - p.start[l];
        const int w = p.w[l], h = p.h[l];
        const int x = (int)(r % w); r /= w;
        const int y = (int)(r % h); r /= h;
        const int c = (int)(r % p.C);
        const int b = (int)(r / p.C);
        const TapAC ty = ac_taps(y, p.sy[l], p.H);
        const TapAC tx = ac_taps(x, p.sx[l], p.W);
        const float* s = p.src + b * p.src_bs + c * p.src_cs;
        const float* r0 = s + (long long)ty.i0 * p.W;
        const float* r1 = s + (long long)ty.i1 * p.W;
        const float v00 = __ldg(r0 + tx.i0), v01 = __ldg(r0 + tx.i1);
        const float v10 = __ldg(r1 + tx.i0), v11 = __ldg(r1 + tx.i1);
        // ATen: w0y * (w0x * v00 + w1x * v01) + w1y * (w0x * v10 + w1x * v11)
        const float top = tx.w0 * v00 + tx.w1 * v01;
        const float bot = tx.w0 * v10 + tx.w1 * v11;
        p.dst[l][i - p.start[l]] = ty.w0 * top + ty.w1 * bot;
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
    PyramidParams p;
    p.src = src; p.src_bs = src_bs; p.src_cs = src_cs;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.levels = scales - 1;
    p.start[0] = 0;
    for (int l = 0; l < p.levels; ++l) {
        const int h = H >> (l + 1), w = W >> (l + 1);
        if (h < 1 || w < 1 || !dst[l + 1]) return USL_ERR_ARG;
        p.dst[l] = dst[l + 1];
        p.h[l] = h; p.w[l] = w;
        p.sy[l] = ac_scale(H, h);
        p.sx[l] = ac_scale(W, w);
        p.start[l + 1] = p.start[l] + (long long)B * C * h * w;
    }
    const long long total = p.start[p.levels];
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    pyramid_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}
