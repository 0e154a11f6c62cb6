// Coordinate conventions shared by every kernel on the path.
//
// These helpers pin down, in fp32 and in one place, the three resampling
// conventions the reference relies on (all citations: /root/reference):
//
//  * linspace01     torch.linspace(0, 1, n) as the reference builds its base
//                   grid on the CPU (utils.py:80-87): fp32 step, lower half
//                   step*i, upper half 1 - step*(n-1-i) with one rounding.
//  * warp taps      F.grid_sample(bilinear, zeros, align_corners=False) fed
//                   with 2*(base+shift)-1 (utils.py:93-97):
//                   i = (g + 1) * size/2 - 0.5, taps floor(i), floor(i)+1,
//                   out-of-range taps contribute zero.
//  * align-corners  F.interpolate(bilinear, align_corners=True)
//                   (utils.py:45-46, loss.py:120-121): src = dst*(in-1)/(out-1),
//                   i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0.
//
// Everything here is host+device so the CPU emulation harness under tests/emu
// executes exactly the arithmetic the kernels execute.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define USL_HD __host__ __device__ __forceinline__
#else
#define USL_HD inline
#endif

namespace usl {

USL_HD float linspace01(int i, int n) {
    if (n <= 1) return 0.0f;
    const float step = 1.0f / (float)(n - 1);
    return (i < n / 2) ? step * (float)i
                       : fmaf(-step, (float)(n - 1 - i), 1.0f);
}

// Unnormalised sampling coordinate for base position `i` of `n`, shifted by
// `shift` (in units of the normalised [0, 1] axis).
USL_HD float warp_coord(int i, int n, float shift) {
    const float x = linspace01(i, n) + shift;
    const float g = fmaf(2.0f, x, -1.0f);
    return fmaf(g + 1.0f, 0.5f * (float)n, -0.5f);
}

struct Tap2 {          // two neighbouring source indices and their weights
    int i0;            // floor(coord); the second tap is i0 + 1
    float w0, w1;      // weights of i0 and i0 + 1 (before range masking)
};

USL_HD Tap2 split_coord(float c) {
    const float f = floorf(c);
    Tap2 t;
    t.i0 = (int)f;
    t.w1 = c - f;
    t.w0 = (f + 1.0f) - c;
    return t;
}

// Vertical taps of the warp for output row y: they depend on the row only.
USL_HD Tap2 warp_row_taps(int y, int h) {
    return split_coord(warp_coord(y, h, 0.0f));
}

// align_corners=True bilinear source taps for destination index `d`.
struct TapAC {
    int i0, i1;
    float w0, w1;
};

USL_HD float ac_scale(int in, int out) {
    return (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.0f;
}

USL_HD TapAC ac_taps(int d, float scale, int in) {
    const float src = scale * (float)d;
    TapAC t;
    t.i0 = (int)src;
    if (t.i0 > in - 1) t.i0 = in - 1;
    t.i1 = t.i0 + ((t.i0 < in - 1) ? 1 : 0);
    float l1 = src - (float)t.i0;
    l1 = fminf(fmaxf(l1, 0.0f), 1.0f);
    t.w1 = l1;
    t.w0 = 1.0f - l1;
    return t;
}

USL_HD float sgnf(float v) {          // torch: d|v|/dv = sign(v), 0 at 0
    return (float)((v > 0.0f) - (v < 0.0f));
}

}  // namespace usl
