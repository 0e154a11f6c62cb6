// Gaussian-window SSIM metric of the evaluation loop:
// torchmetrics.functional.structural_similarity_index_measure(preds, target,
// gaussian_kernel=True, sigma=1.5, kernel_size=11, reduction='sum',
// data_range=1.0) as evaluate.py:142-146 calls it.  torchmetrics is not
// vendored in the reference tree (requirements.txt lists it unpinned) and is
// absent from this image; the algorithm restated here is the published one of
// torchmetrics 1.x (functional/image/ssim.py, _ssim_update):
//   * k x k Gaussian window g (sigma), normalised, depth-wise over channels;
//   * reflect-pad by (k-1)/2, filter {p, t, p*p, t*t, p*t}, then crop the
//     padded border off again -- so only windows that lie inside the image
//     survive, and the padding never reaches the result: a VALID filter;
//   * ssim = (2 mu_p mu_t + c1)(2 s_pt + c2) / ((mu_p^2 + mu_t^2 + c1)
//     (s_pp + s_tt + c2)), s_pp = max(E[p^2] - mu_p^2, 0) (likewise s_tt),
//     s_pt = E[pt] - mu_p mu_t, c1 = (k1 R)^2, c2 = (k2 R)^2;
//   * per image: mean over channels and surviving pixels.
// The filter is applied separably (rows, then columns) from a shared-memory
// tile; HBM-bound: 8 B per pixel and channel.
#include "usl_common.cuh"

namespace usl {

constexpr int SS_TX = 32, SS_TY = 32, SS_THREADS = 256;
constexpr int SS_KMAX = 15;

struct SsimArgs {
    const float* p; long long p_bs, p_cs;
    const float* t; long long t_bs, t_cs;
    int B, C, H, W, k;
    float g[SS_KMAX];
    float c1, c2;
    float* partials;        // [B*C][tiles]
    int tiles_x, tiles_y;
};

__global__ void __launch_bounds__(SS_THREADS) ssim_kernel(const __grid_constant__ SsimArgs A) {
    extern __shared__ float sm[];
    const int k = A.k, halo = k - 1;
    const int IW = SS_TX + halo, IH = SS_TY + halo;
    float* sp = sm;                       // [IH][IW]
    float* st = sp + IH * IW;             // [IH][IW]
    float* hz = st + IH * IW;             // [5][IH][SS_TX] row-filtered quantities
    __shared__ float red[SS_THREADS / 32];
    const int plane = blockIdx.z, b = plane / A.C, c = plane % A.C;
    const int ox0 = blockIdx.x * SS_TX, oy0 = blockIdx.y * SS_TY;
    const int oh = A.H - halo, ow = A.W - halo;       // surviving windows
    const float* p = A.p + b * A.p_bs + c * A.p_cs;
    const float* t = A.t + b * A.t_bs + c * A.t_cs;
    const int tid = threadIdx.x;
    for (int i = tid; i < IH * IW; i += SS_THREADS) {
        const int y = oy0 + i / IW, x = ox0 + i % IW;
        const bool ok = y < A.H && x < A.W;
        sp[i] = ok ? __ldg(p + (long long)y * A.W + x) : 0.0f;
        st[i] = ok ? __ldg(t + (long long)y * A.W + x) : 0.0f;
    }
    __syncthreads();
    for (int i = tid; i < IH * SS_TX; i += SS_THREADS) {
        const int y = i / SS_TX, x = i % SS_TX;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
        for (int j = 0; j < k; ++j) {
            const float g = A.g[j];
            const float pv = sp[y * IW + x + j], tv = st[y * IW + x + j];
            a0 = fmaf(g, pv, a0); a1 = fmaf(g, tv, a1);
            a2 = fmaf(g, pv * pv, a2); a3 = fmaf(g, tv * tv, a3);
            a4 = fmaf(g, pv * tv, a4);
        }
        hz[(0 * IH + y) * SS_TX + x] = a0; hz[(1 * IH + y) * SS_TX + x] = a1;
        hz[(2 * IH + y) * SS_TX + x] = a2; hz[(3 * IH + y) * SS_TX + x] = a3;
        hz[(4 * IH + y) * SS_TX + x] = a4;
    }
    __syncthreads();
    float acc = 0.0f;
    for (int i = tid; i < SS_TY * SS_TX; i += SS_THREADS) {
        const int y = i / SS_TX, x = i % SS_TX;
        if (oy0 + y >= oh || ox0 + x >= ow) continue;
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < k; ++j) {
            const float g = A.g[j];
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] = fmaf(g, hz[(q * IH + y + j) * SS_TX + x], m[q]);
        }
        const float mpp = m[0] * m[0], mtt = m[1] * m[1], mpt = m[0] * m[1];
        const float spp = fmaxf(m[2] - mpp, 0.0f), stt = fmaxf(m[3] - mtt, 0.0f);
        const float spt = m[4] - mpt;
        const float upper = 2.0f * spt + A.c2, lower = spp + stt + A.c2;
        acc += ((2.0f * mpt + A.c1) * upper) / ((mpp + mtt + A.c1) * lower);
    }
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        for (int w = 0; w < SS_THREADS / 32; ++w) s += red[w];
        A.partials[(long long)plane * (A.tiles_x * A.tiles_y) +
                   blockIdx.y * A.tiles_x + blockIdx.x] = s;
    }
}

// per image: fixed-order fp64 sum of its planes' tile sums / count
__global__ void ssim_finish_kernel(const float* partials, int B, int per_image,
                                   double count, float* out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s = 0.0;
    for (int i = 0; i < per_image; ++i) s += (double)partials[(long long)b * per_image + i];
    out[b] = (float)(s / count);
}

}  // namespace usl

using namespace usl;

extern "C" size_t usl_ssim_workspace_bytes(int B, int C, int H, int W, int k) {
    if (B < 1 || C < 1 || k < 1 || k > SS_KMAX || !(k & 1) || H < k || W < k) return 0;
    const int tx = (W - k + 1 + SS_TX - 1) / SS_TX, ty = (H - k + 1 + SS_TY - 1) / SS_TY;
    return (size_t)B * C * tx * ty * sizeof(float);
}

extern "C" int usl_ssim_gauss(const float* pred, long long p_bs, long long p_cs,
                              const float* target, long long t_bs, long long t_cs,
                              int B, int C, int H, int W, int k, float sigma,
                              float data_range, float k1, float k2,
                              float* per_image, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!pred || !target || !per_image || !workspace || B < 1 || C < 1) return USL_ERR_ARG;
    if (k < 1 || k > SS_KMAX || !(k & 1) || H < k || W < k || !(sigma > 0.0f))
        return USL_ERR_UNSUPPORTED;
    const size_t need = usl_ssim_workspace_bytes(B, C, H, W, k);
    if (workspace_bytes < need) return USL_ERR_WORKSPACE;
    DeviceGuard guard(pred);
    SsimArgs A = {};
    A.p = pred; A.p_bs = p_bs; A.p_cs = p_cs;
    A.t = target; A.t_bs = t_bs; A.t_cs = t_cs;
    A.B = B; A.C = C; A.H = H; A.W = W; A.k = k;
    {   // torchmetrics _gaussian: exp(-(x / sigma)^2 / 2) on arange((1-k)/2, (1+k)/2),
        // normalised, in the inputs' dtype (fp32)
        float g[SS_KMAX], sum = 0.0f;
        for (int j = 0; j < k; ++j) {
            const float d = (float)((1 - k) / 2 + j) / sigma;
            g[j] = expf(-(d * d) / 2.0f);
            sum += g[j];
        }
        for (int j = 0; j < k; ++j) A.g[j] = g[j] / sum;
    }
    A.c1 = (k1 * data_range) * (k1 * data_range);
    A.c2 = (k2 * data_range) * (k2 * data_range);
    A.partials = (float*)workspace;
    A.tiles_x = (W - k + 1 + SS_TX - 1) / SS_TX;
    A.tiles_y = (H - k + 1 + SS_TY - 1) / SS_TY;
    if ((long long)B * C > 65535 || A.tiles_y > 65535) return USL_ERR_UNSUPPORTED;
    const int IW = SS_TX + k - 1, IH = SS_TY + k - 1;
    const size_t smem = ((size_t)2 * IH * IW + (size_t)5 * IH * SS_TX) * sizeof(float);
    if (cudaFuncSetAttribute(ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return USL_ERR_CUDA;
    ssim_kernel<<<dim3(A.tiles_x, A.tiles_y, B * C), SS_THREADS, smem,
                  (cudaStream_t)stream>>>(A);
    count_launches(1);
    const double count = (double)C * (H - k + 1) * (W - k + 1);
    ssim_finish_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        A.partials, B, C * A.tiles_x * A.tiles_y, count, per_image);
    return check_launch();
}
