// The callers and data formats either side of the loss path:
// streaming kernels, one launch for all pyramid levels each, bounded by HBM.
//
//   n2  decoder disparity head      model/layers/decoder.py:239-246
//         pred = scale * sigmoid(logits), and its backward
//   n3  discriminator input glue    train/utils.py:53-62, 138-140, 248-273
//         [image pyramid ; reconstruction pyramid] along the batch axis,
//         the reconstruction half warped on the fly (never materialised) or
//         copied from a materialised pyramid -- one launch instead of a
//         clone + a cat per level
//   n4  evaluation post-processing  train/utils.py:177-245
//         combine_disparity (Monodepth2 blind-spot blend, fp64 like the numpy
//         original) and to_heatmap (colour-map lookup)
#include "usl_common.cuh"

namespace usl {

constexpr int GL_THREADS = 256;

struct Levels {
    const float* a[USL_MAX_SCALES];      // first input of level i
    const float* b[USL_MAX_SCALES];      // second input
    float* out[USL_MAX_SCALES];
    long long a_bs[USL_MAX_SCALES], a_cs[USL_MAX_SCALES];   // strides (elements)
    long long b_bs[USL_MAX_SCALES], b_cs[USL_MAX_SCALES];
    long long n[USL_MAX_SCALES];         // work items of level i
    int cta_start[USL_MAX_SCALES + 1];
    int B[USL_MAX_SCALES], h[USL_MAX_SCALES], w[USL_MAX_SCALES];
    int levels;
    float scale;
};

static int plan_ctas(Levels* L, int items_per_cta) {
    L->cta_start[0] = 0;
    for (int i = 0; i < L->levels; ++i) {
        long long c = (L->n[i] + items_per_cta - 1) / items_per_cta;
        const long long cap = (long long)num_sms() * 8;
        if (c > cap) c = cap;
        if (c < 1) c = 1;
        L->cta_start[i + 1] = L->cta_start[i] + (int)c;
    }
    return L->cta_start[L->levels];
}

__device__ __forceinline__ int find_level(const Levels& L, int cta) {
    int s = 0;
    while (s + 1 < L.levels && cta >= L.cta_start[s + 1]) ++s;
    return s;
}

// ---- n2: pred = scale * sigmoid(logits) -------------------------------------
// 128-bit accesses when the level is 16-byte aligned (contiguous tensors).
template <bool BWD>
__global__ void __launch_bounds__(GL_THREADS) head_kernel(const __grid_constant__ Levels L) {
    const int s = find_level(L, blockIdx.x);
    const long long n = L.n[s];
    const int nc = L.cta_start[s + 1] - L.cta_start[s];
    const long long stride = (long long)nc * GL_THREADS;
    const long long t0 = (long long)(blockIdx.x - L.cta_start[s]) * GL_THREADS + threadIdx.x;
    const float sc = L.scale, inv = 1.0f / L.scale;
    const float* a = L.a[s];      // fwd: logits          bwd: grad of pred
    const float* b = L.b[s];      //                      bwd: pred
    float* out = L.out[s];
    auto f = [&](float x, float p) {
        if (!BWD) return sc / (1.0f + __expf(-x));
        // d pred / d logit = scale * s * (1 - s) = pred * (1 - pred / scale)
        return x * (p * (1.0f - p * inv));
    };
    const bool vec = (n & 3) == 0 && (((uintptr_t)a | (uintptr_t)out |
                                       (BWD ? (uintptr_t)b : 0)) & 15) == 0;
    if (vec) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        float4* o4 = reinterpret_cast<float4*>(out);
        for (long long i = t0; i < n / 4; i += stride) {
            const float4 x = __ldg(a4 + i);
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (BWD) p = __ldg(b4 + i);
            o4[i] = make_float4(f(x.x, p.x), f(x.y, p.y), f(x.z, p.z), f(x.w, p.w));
        }
    } else {
        for (long long i = t0; i < n; i += stride)
            out[i] = f(__ldg(a + i), BWD ? __ldg(b + i) : 0.0f);
    }
}

// ---- n3: [images ; reconstructions] along the batch axis ----------------------
// out (2B,6,h,w): samples 0..B-1 = the image level, samples B..2B-1 = its
// reconstruction: warped here from the disparities (b = prediction, channels
// 0/1) when WARP, else copied from a materialised level (b = recon).
template <bool WARP>
__global__ void __launch_bounds__(GL_THREADS) disc_input_kernel(const __grid_constant__ Levels L) {
    const int s = find_level(L, blockIdx.x);
    const int B = L.B[s], h = L.h[s], w = L.w[s];
    const long long hw = (long long)h * w;
    const long long n = L.n[s];                 // B * h * w pixels
    const int nc = L.cta_start[s + 1] - L.cta_start[s];
    const long long stride = (long long)nc * GL_THREADS;
    float* out = L.out[s];
    for (long long i = (long long)(blockIdx.x - L.cta_start[s]) * GL_THREADS + threadIdx.x;
         i < n; i += stride) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const int b = (int)(i / hw);
        const long long pix = (long long)y * w + x;
        const float* img = L.a[s] + b * L.a_bs[s] + pix;
        float v[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            v[c] = __ldg(img + c * L.a_cs[s]);
            out[((long long)b * 6 + c) * hw + pix] = v[c];
        }
        float* ro = out + ((long long)(B + b) * 6) * hw + pix;
        if (!WARP) {
            const float* rc = L.b[s] + b * L.b_bs[s] + pix;
#pragma unroll
            for (int c = 0; c < 6; ++c) ro[c * hw] = __ldg(rc + c * L.b_cs[s]);
        } else {
            // utils.py:65-135: left view from the right image shifted by -d_L,
            // right view from the left image shifted by +d_R
            const Tap2 ty = warp_row_taps(y, h);
            const float* img0 = L.a[s] + b * L.a_bs[s];
#pragma unroll
            for (int view = 0; view < 2; ++view) {
                const float d = __ldg(L.b[s] + b * L.b_bs[s] + view * L.b_cs[s] + pix);
                const Tap2 tx = split_coord(warp_coord(x, w, view ? d : -d));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* pl = img0 + ((1 - view) * 3 + c) * L.a_cs[s];
                    auto tap = [&](int yy, int xx) {
                        return (yy >= 0 && yy < h && xx >= 0 && xx < w)
                                   ? __ldg(pl + (long long)yy * w + xx) : 0.0f;
                    };
                    ro[(view * 3 + c) * hw] =
                        tap(ty.i0, tx.i0) * (tx.w0 * ty.w0) +
                        tap(ty.i0, tx.i0 + 1) * (tx.w1 * ty.w0) +
                        tap(ty.i0 + 1, tx.i0) * (tx.w0 * ty.w1) +
                        tap(ty.i0 + 1, tx.i0 + 1) * (tx.w1 * ty.w1);
                }
            }
        }
    }
}

// ---- gradient arriving at the reconstructions -> disparities ---------------------
// grad_d_v[b,y,x] (+)= sign_v * w * sum_c g[b, 3v+c, y, x] *
//                      [(v01 - v00) * wy0 + (v11 - v10) * wy1]   (utils.py:65-135:
// the transpose of the warp w.r.t. its shift; a gather with the warp's own taps).
// a = images, b = prediction (channels 0/1 = d_L, d_R), out = gradient of the
// prediction (same strides as b: b_bs / b_cs), g in `extra`.
struct ReconBwd {
    Levels L;
    const float* g[USL_MAX_SCALES];     // contiguous (B,6,h,w)
    long long o_bs[USL_MAX_SCALES], o_cs[USL_MAX_SCALES];
    int accumulate;
};
__global__ void __launch_bounds__(GL_THREADS) recon_bwd_kernel(const __grid_constant__ ReconBwd A) {
    const Levels& L = A.L;
    const int s = find_level(L, blockIdx.x);
    const int h = L.h[s], w = L.w[s];
    const long long hw = (long long)h * w;
    const long long n = L.n[s];
    const int nc = L.cta_start[s + 1] - L.cta_start[s];
    const long long stride = (long long)nc * GL_THREADS;
    const float fw = (float)w;
    for (long long i = (long long)(blockIdx.x - L.cta_start[s]) * GL_THREADS + threadIdx.x;
         i < n; i += stride) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const int b = (int)(i / hw);
        const long long pix = (long long)y * w + x;
        const Tap2 ty = warp_row_taps(y, h);
        const float* img0 = L.a[s] + b * L.a_bs[s];
#pragma unroll
        for (int view = 0; view < 2; ++view) {
            const float d = __ldg(L.b[s] + b * L.b_bs[s] + view * L.b_cs[s] + pix);
            const float sign = view ? 1.0f : -1.0f;
            const Tap2 tx = split_coord(warp_coord(x, w, sign * d));
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* pl = img0 + ((1 - view) * 3 + c) * L.a_cs[s];
                auto tap = [&](int yy, int xx) {
                    return (yy >= 0 && yy < h && xx >= 0 && xx < w)
                               ? __ldg(pl + (long long)yy * w + xx) : 0.0f;
                };
                const float v00 = tap(ty.i0, tx.i0), v01 = tap(ty.i0, tx.i0 + 1);
                const float v10 = tap(ty.i0 + 1, tx.i0), v11 = tap(ty.i0 + 1, tx.i0 + 1);
                const float go = __ldg(A.g[s] + ((long long)b * 6 + view * 3 + c) * hw + pix);
                acc += go * ((v01 - v00) * ty.w0 + (v11 - v10) * ty.w1);
            }
            float* o = L.out[s] + b * A.o_bs[s] + view * A.o_cs[s] + pix;
            const float gd = acc * fw * sign;
            *o = A.accumulate ? *o + gd : gd;
        }
    }
}

// ---- n4 -----------------------------------------------------------------------
// utils.py:199-245 in the numpy original's fp64: x = linspace(0,1,W)[j];
// l = 1 - clip(alpha (x - beta), 0, 1); r = l mirrored; out = r * left +
// l * right + (1 - l - r) * (left + right) / 2.
__global__ void __launch_bounds__(GL_THREADS)
combine_disparity_kernel(const float* left, const float* right, int planes, int h, int w,
                         double alpha, double beta, double* out) {
    const long long n = (long long)planes * h * w;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double step = w > 1 ? 1.0 / (double)(w - 1) : 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int x = (int)(i % w);
        auto mask = [&](int j) {
            // numpy.linspace: start + j * step, the last sample exactly stop
            const double xv = (j == w - 1 && w > 1) ? 1.0 : (double)j * step;
            const double t = alpha * (xv - beta);
            return 1.0 - fmin(fmax(t, 0.0), 1.0);
        };
        const double l = mask(x), r = mask(w - 1 - x);
        const float a32 = __ldg(left + i), b32 = __ldg(right + i);
        // (the mean of the two fp32 maps is taken in fp32, utils.py:224)
        const double mean = (double)((a32 + b32) / 2.0f);
        const double a = (double)a32, b = (double)b32;
        out[i] = (r * a + l * b) + (1.0 - (l + r)) * mean;
    }
}

// utils.py:177-196: matplotlib's Colormap.__call__ on floats: index
// int(x * N) (x == 1 -> N - 1), below 0 -> first, above 1 -> last entry
// (the under / over colours of the listed maps), NaN -> black.  out (3,h,w).
__global__ void __launch_bounds__(GL_THREADS)
heatmap_kernel(const float* x, long long n, int inverse, const double* lut, int entries,
               double* out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = __ldg(x + i);
        if (inverse) v = 1.0f - v;
        double r = 0.0, g = 0.0, b = 0.0;
        if (v == v) {
            // (matplotlib works in fp64 on the fp32 pixel)
            double t = (double)v * (double)entries;
            int k = t < 0.0 ? 0 : (t >= (double)entries ? entries - 1 : (int)t);
            if (v == 1.0f) k = entries - 1;
            r = __ldg(lut + 3 * k); g = __ldg(lut + 3 * k + 1); b = __ldg(lut + 3 * k + 2);
        }
        out[i] = r; out[n + i] = g; out[2 * n + i] = b;
    }
}

static unsigned flat_blocks(long long n) {
    long long blocks = (n + GL_THREADS - 1) / GL_THREADS;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace usl

using namespace usl;

static int head_launch(const float* const* a, const float* const* b, float* const* out,
                       const long long* counts, int levels, float scale, bool bwd,
                       void* stream) {
    if (!a || !out || !counts || levels < 1 || levels > USL_MAX_SCALES || !(scale > 0.0f))
        return USL_ERR_ARG;
    Levels L = {};
    L.levels = levels; L.scale = scale;
    for (int i = 0; i < levels; ++i) {
        if (!a[i] || !out[i] || counts[i] < 0 || (bwd && (!b || !b[i]))) return USL_ERR_ARG;
        L.a[i] = a[i]; L.b[i] = bwd ? b[i] : nullptr; L.out[i] = out[i];
        L.n[i] = counts[i];
    }
    DeviceGuard guard(a[0]);
    // (a thread moves 4 elements per trip when the level is aligned)
    const int grid = plan_ctas(&L, GL_THREADS * 16);
    if (bwd) head_kernel<true><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(L);
    else head_kernel<false><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(L);
    return check_launch();
}

extern "C" int usl_head_fwd(const float* const* logits, float* const* pred,
                            const long long* counts, int levels, float scale,
                            void* stream) {
    return head_launch(logits, nullptr, pred, counts, levels, scale, false, stream);
}

extern "C" int usl_head_bwd(const float* const* grad_pred, const float* const* pred,
                            float* const* grad_logits, const long long* counts,
                            int levels, float scale, void* stream) {
    return head_launch(grad_pred, pred, grad_logits, counts, levels, scale, true, stream);
}

extern "C" int usl_disc_input(const UslDiscLevel* lv, int levels, void* stream) {
    if (!lv || levels < 1 || levels > USL_MAX_SCALES) return USL_ERR_ARG;
    Levels L = {};
    L.levels = levels;
    const bool warp = lv[0].pred != nullptr;
    for (int i = 0; i < levels; ++i) {
        const UslDiscLevel& v = lv[i];
        if (!v.images || !v.out || v.B < 1 || v.h < 1 || v.w < 1) return USL_ERR_ARG;
        if ((v.pred != nullptr) != warp || (!warp && !v.recon)) return USL_ERR_ARG;
        L.a[i] = v.images; L.a_bs[i] = v.img_bs; L.a_cs[i] = v.img_cs;
        if (warp) { L.b[i] = v.pred; L.b_bs[i] = v.pred_bs; L.b_cs[i] = v.pred_cs; }
        else { L.b[i] = v.recon; L.b_bs[i] = v.rec_bs; L.b_cs[i] = v.rec_cs; }
        L.out[i] = v.out;
        L.B[i] = v.B; L.h[i] = v.h; L.w[i] = v.w;
        L.n[i] = (long long)v.B * v.h * v.w;
    }
    DeviceGuard guard(lv[0].images);
    const int grid = plan_ctas(&L, GL_THREADS * 4);
    if (warp) disc_input_kernel<true><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(L);
    else disc_input_kernel<false><<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(L);
    return check_launch();
}

extern "C" int usl_recon_bwd(const UslDiscLevel* lv, const float* const* grad_recon,
                             float* const* grad_pred, const long long* gp_bs,
                             const long long* gp_cs, int levels, int accumulate,
                             void* stream) {
    if (!lv || !grad_recon || !grad_pred || !gp_bs || !gp_cs || levels < 1 ||
        levels > USL_MAX_SCALES)
        return USL_ERR_ARG;
    ReconBwd A = {};
    A.L.levels = levels;
    A.accumulate = accumulate;
    for (int i = 0; i < levels; ++i) {
        const UslDiscLevel& v = lv[i];
        if (!v.images || !v.pred || !grad_recon[i] || !grad_pred[i] || v.B < 1 ||
            v.h < 1 || v.w < 1)
            return USL_ERR_ARG;
        A.L.a[i] = v.images; A.L.a_bs[i] = v.img_bs; A.L.a_cs[i] = v.img_cs;
        A.L.b[i] = v.pred; A.L.b_bs[i] = v.pred_bs; A.L.b_cs[i] = v.pred_cs;
        A.L.out[i] = grad_pred[i]; A.o_bs[i] = gp_bs[i]; A.o_cs[i] = gp_cs[i];
        A.g[i] = grad_recon[i];
        A.L.B[i] = v.B; A.L.h[i] = v.h; A.L.w[i] = v.w;
        A.L.n[i] = (long long)v.B * v.h * v.w;
    }
    DeviceGuard guard(lv[0].images);
    const int grid = plan_ctas(&A.L, GL_THREADS * 4);
    recon_bwd_kernel<<<grid, GL_THREADS, 0, (cudaStream_t)stream>>>(A);
    return check_launch();
}

extern "C" int usl_combine_disparity(const float* left, const float* right, int planes,
                                     int h, int w, double alpha, double beta, double* out,
                                     void* stream) {
    if (!left || !right || !out || planes < 1 || h < 1 || w < 1) return USL_ERR_ARG;
    DeviceGuard guard(left);
    combine_disparity_kernel<<<flat_blocks((long long)planes * h * w), GL_THREADS, 0,
                               (cudaStream_t)stream>>>(left, right, planes, h, w, alpha,
                                                       beta, out);
    return check_launch();
}

extern "C" int usl_heatmap(const float* x, long long n, int inverse, const double* lut,
                           int entries, double* out, void* stream) {
    if (!x || !lut || !out || n < 1 || entries < 1) return USL_ERR_ARG;
    DeviceGuard guard(x);
    heatmap_kernel<<<flat_blocks(n), GL_THREADS, 0, (cudaStream_t)stream>>>(
        x, n, inverse, lut, entries, out);
    return check_launch();
}
