// Stand-alone disparity warp (reference: train/utils.py:65-109), used when a
// caller asks for a materialised reconstruction (evaluate.py:139-140, the
// adversarial branch of train.py:122-152).  The training hot path never
// launches this: it warps inside the fused loss kernels.
//
//   out[b,c,y,x] = sum over 4 taps of image[b,c,yt,xt] * wt
//   ix = warp_coord(x, w, sign * disp[b,y,x]),  iy = warp_coord(y, h, 0)
//
// Backward w.r.t. the disparity is a gather with the same taps:
//   d out / d shift = w * [ (v01 - v00) * wy0 + (v11 - v10) * wy1 ].
#include "usl_common.cuh"

namespace usl {

struct WarpParams {
    const float* disp; long long disp_bs;
    const float* image; long long img_bs, img_cs;
    const float* grad_out; long long go_bs, go_cs;
    float* out; long long out_bs, out_cs;
    float* grad_disp; long long gd_bs;
    float sign;
    int B, C, h, w;
};

__device__ __forceinline__ float tap(const float* plane, int y, int x, int h,
                                     int w) {
    return (y >= 0 && y < h && x >= 0 && x < w)
               ? __ldg(plane + (long long)y * w + x) : 0.0f;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(256) warp_kernel(const WarpParams p) {
    const long long total = (long long)p.B * p.h * p.w;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < total; i += stride) {
        const int x = (int)(i % p.w);
        const int y = (int)((i / p.w) % p.h);
        const int b = (int)(i / ((long long)p.w * p.h));
        const long long pix = (long long)y * p.w + x;
        const float d = __ldg(p.disp + b * p.disp_bs + pix);
        const Tap2 tx = split_coord(warp_coord(x, p.w, p.sign * d));
        const Tap2 ty = warp_row_taps(y, p.h);
        float gsum = 0.0f;
        for (int c = 0; c < p.C; ++c) {
            const float* plane = p.image + b * p.img_bs + c * p.img_cs;
            const float v00 = tap(plane, ty.i0, tx.i0, p.h, p.w);
            const float v01 = tap(plane, ty.i0, tx.i0 + 1, p.h, p.w);
            const float v10 = tap(plane, ty.i0 + 1, tx.i0, p.h, p.w);
            const float v11 = tap(plane, ty.i0 + 1, tx.i0 + 1, p.h, p.w);
            if (!BACKWARD) {
                p.out[b * p.out_bs + c * p.out_cs + pix] =
                    v00 * (tx.w0 * ty.w0) + v01 * (tx.w1 * ty.w0) +
                    v10 * (tx.w0 * ty.w1) + v11 * (tx.w1 * ty.w1);
            } else {
                const float go = __ldg(p.grad_out + b * p.go_bs + c * p.go_cs + pix);
                gsum += go * ((v01 - v00) * ty.w0 + (v11 - v10) * ty.w1);
            }
        }
        if (BACKWARD)
            p.grad_disp[b * p.gd_bs + pix] = gsum * (float)p.w * p.sign;
    }
}

static int launch_warp(const WarpParams& p, bool backward, void* stream) {
    if (p.B <= 0 || p.C <= 0 || p.h < 1 || p.w < 1) return USL_ERR_ARG;
    const long long total = (long long)p.B * p.h * p.w;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (backward)
        warp_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else
        warp_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch();
}

// ---- gradient w.r.t. the SAMPLED image (the transpose of the gather) --------
// ATen does it with float atomics.  Here, deterministic: one thread per (b, c,
// output row y) adds the horizontal taps of its row, in column order, into a
// private row of the workspace (columns -2 .. w+1: taps outside the image land
// in pads); a second launch blends the three rows around each image row with
// the vertical tap weights, in row order.  Not a hot path (the training step
// differentiates through a sampled map only inside ConsistencyLoss, which the
// fused kernels handle): plain global-memory read-modify-writes.
constexpr int WBI_PAD = 2;

__global__ void __launch_bounds__(128)
warp_bwd_image_rows_kernel(const WarpParams p, float* __restrict__ ws) {
    const long long rows = (long long)p.B * p.C * p.h;
    const int pitch = p.w + 2 * WBI_PAD;
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows;
         r += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(r % p.h);
        const int c = (int)((r / p.h) % p.C);
        const int b = (int)(r / ((long long)p.h * p.C));
        float* T = ws + r * pitch;
        for (int i = 0; i < pitch; ++i) T[i] = 0.0f;
        const float* d = p.disp + b * p.disp_bs + (long long)y * p.w;
        const float* go = p.grad_out + b * p.go_bs + c * p.go_cs + (long long)y * p.w;
        for (int x = 0; x < p.w; ++x) {
            const Tap2 tx = split_coord(warp_coord(x, p.w, p.sign * __ldg(d + x)));
            const float f = fminf(fmaxf((float)tx.i0, (float)-WBI_PAD), (float)p.w);
            const int col = (int)f + WBI_PAD;
            const float g = __ldg(go + x);
            T[col] += g * tx.w0;
            T[col + 1] += g * tx.w1;
        }
    }
}

__global__ void __launch_bounds__(256)
warp_bwd_image_blend_kernel(const WarpParams p, const float* __restrict__ ws,
                            float* __restrict__ gi, long long gi_bs, long long gi_cs) {
    const long long total = (long long)p.B * p.C * p.h * p.w;
    const int pitch = p.w + 2 * WBI_PAD;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % p.w);
        const int yy = (int)((i / p.w) % p.h);
        const long long plane = i / ((long long)p.w * p.h);     // b * C + c
        const int c = (int)(plane % p.C), b = (int)(plane / p.C);
        float acc = 0.0f;
        for (int y = yy - 1; y <= yy + 1; ++y) {
            if (y < 0 || y >= p.h) continue;
            const Tap2 ty = warp_row_taps(y, p.h);
            float wgt = 0.0f;
            if (ty.i0 == yy) wgt = ty.w0;
            else if (ty.i0 + 1 == yy) wgt = ty.w1;
            acc += wgt * ws[(plane * p.h + y) * pitch + x + WBI_PAD];
        }
        gi[b * gi_bs + c * gi_cs + (long long)yy * p.w + x] = acc;
    }
}

}  // namespace usl

extern "C" long long usl_warp_bwd_image_workspace_bytes(int B, int C, int h, int w) {
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return 0;
    return (long long)B * C * h * (w + 2 * usl::WBI_PAD) * (long long)sizeof(float);
}

extern "C" int usl_warp_bwd_image(const float* disp, long long disp_bs, float sign,
                                  const float* grad_out, long long go_bs,
                                  long long go_cs, int B, int C, int h, int w,
                                  void* workspace, float* grad_image,
                                  long long gi_bs, long long gi_cs, void* stream) {
    if (!disp || !grad_out || !workspace || !grad_image || B <= 0 || C <= 0 ||
        h < 1 || w < 1)
        return USL_ERR_ARG;
    usl::DeviceGuard guard(grad_out);
    usl::WarpParams p = {};
    p.disp = disp; p.disp_bs = disp_bs; p.sign = sign;
    p.grad_out = grad_out; p.go_bs = go_bs; p.go_cs = go_cs;
    p.B = B; p.C = C; p.h = h; p.w = w;
    const long long rows = (long long)B * C * h;
    long long blocks = (rows + 127) / 128;
    const long long cap = (long long)usl::num_sms() * 16;
    if (blocks > cap) blocks = cap;
    usl::warp_bwd_image_rows_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(
        p, static_cast<float*>(workspace));
    usl::count_launches(1);
    const long long total = rows * w;
    blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    usl::warp_bwd_image_blend_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        p, static_cast<const float*>(workspace), grad_image, gi_bs, gi_cs);
    return usl::check_launch();
}

extern "C" int usl_warp_fwd(const float* disp, long long disp_bs, float sign,
                            const float* image, long long img_bs,
                            long long img_cs, int B, int C, int h, int w,
                            float* out, long long out_bs, long long out_cs,
                            void* stream) {
    if (!disp || !image || !out) return USL_ERR_ARG;
    usl::DeviceGuard guard(image);
    usl::WarpParams p = {};
    p.disp = disp; p.disp_bs = disp_bs; p.sign = sign;
    p.image = image; p.img_bs = img_bs; p.img_cs = img_cs;
    p.out = out; p.out_bs = out_bs; p.out_cs = out_cs;
    p.B = B; p.C = C; p.h = h; p.w = w;
    return usl::launch_warp(p, false, stream);
}

extern "C" int usl_warp_bwd_disp(const float* disp, long long disp_bs,
                                 float sign, const float* image,
                                 long long img_bs, long long img_cs,
                                 const float* grad_out, long long go_bs,
                                 long long go_cs, int B, int C, int h, int w,
                                 float* grad_disp, long long gd_bs,
                                 void* stream) {
    if (!disp || !image || !grad_out || !grad_disp) return USL_ERR_ARG;
    usl::DeviceGuard guard(image);
    usl::WarpParams p = {};
    p.disp = disp; p.disp_bs = disp_bs; p.sign = sign;
    p.image = image; p.img_bs = img_bs; p.img_cs = img_cs;
    p.grad_out = grad_out; p.go_bs = go_bs; p.go_cs = go_cs;
    p.grad_disp = grad_disp; p.gd_bs = gd_bs;
    p.B = B; p.C = C; p.h = h; p.w = w;
    return usl::launch_warp(p, true, stream);
}
