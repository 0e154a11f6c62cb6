// Version / error strings and the 3x3 valid mean used by `pooling=True`
// (reference: train/loss.py:386-387, 420-422).
#include "usl_common.cuh"

namespace usl {

__global__ void __launch_bounds__(256)
pool3_fwd_kernel(const float* x, long long x_bs, long long x_cs, int B, int C,
                 int h, int w, float* out) {
    const int oh = h - 2, ow = w - 2;
    const long long total = (long long)B * C * oh * ow;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < total; i += stride) {
        const int ox = (int)(i % ow);
        const int oy = (int)((i / ow) % oh);
        const int c = (int)((i / ((long long)ow * oh)) % C);
        const int b = (int)(i / ((long long)ow * oh * C));
        const float* p = x + b * x_bs + c * x_cs + (long long)oy * w + ox;
        float s = 0.0f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) s += __ldg(p + dy * w + dx);
        out[i] = s / 9.0f;
    }
}

// grad_x[y,x] = (1/9) * sum of grad_out over the windows that contain (y,x)
__global__ void __launch_bounds__(256)
pool3_bwd_kernel(const float* go, int B, int C, int h, int w, float* gx) {
    const int oh = h - 2, ow = w - 2;
    const long long total = (long long)B * C * h * w;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < total; i += stride) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const long long bc = i / ((long long)w * h);
        const float* p = go + bc * oh * ow;
        float s = 0.0f;
        for (int qy = max(0, y - 2); qy <= min(oh - 1, y); ++qy)
            for (int qx = max(0, x - 2); qx <= min(ow - 1, x); ++qx)
                s += __ldg(p + (long long)qy * ow + qx);
        gx[i] = s / 9.0f;
    }
}

// buffers[i][0 .. counts[i]) *= *g unless *g == 1 (see usl_grad_rescale)
struct RescaleArgs {
    float* buf[USL_MAX_RESCALE];
    long long end[USL_MAX_RESCALE];    // running sum of the element counts
    int n;
};
__global__ void __launch_bounds__(256)
rescale_kernel(const float* g, const __grid_constant__ RescaleArgs A) {
    const float s = __ldg(g);
    if (s == 1.0f) return;
    const long long total = A.end[A.n - 1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    int k = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += stride) {
        while (i >= A.end[k]) ++k;
        A.buf[k][i - (k ? A.end[k - 1] : 0)] *= s;
    }
}

static unsigned grid_for(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace usl

namespace usl {
std::atomic<unsigned long long*>& timeline_buffer() {
    static std::atomic<unsigned long long*> p{nullptr};
    return p;
}
__global__ void stamp_kernel(unsigned long long* buf, int idx) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (idx == 0) {     // the previous step's start and its last stamp
        unsigned long long last = 0;
        for (int i = 1; i <= 12; ++i) last = buf[i] > last ? buf[i] : last;
        buf[13] = buf[0]; buf[14] = last;
    }
    buf[idx] = t;
}
void mark(int idx, cudaStream_t st) {
    unsigned long long* b = timeline_buffer().load(std::memory_order_relaxed);
    if (b && idx >= 0 && idx < USL_TIMELINE_SLOTS) stamp_kernel<<<1, 1, 0, st>>>(b, idx);
}
}  // namespace usl

extern "C" int usl_debug_timeline(void* slots) {
    usl::timeline_buffer().store(static_cast<unsigned long long*>(slots),
                                 std::memory_order_relaxed);
    return USL_OK;
}

extern "C" int usl_version(void) { return USL_VERSION; }

extern "C" long long usl_launch_count(void) {
    return usl::launch_counter().load(std::memory_order_relaxed);
}

extern "C" const char* usl_strerror(int rc) {
    switch (rc) {
        case USL_OK: return "ok";
        case USL_ERR_ARG: return "invalid argument";
        case USL_ERR_CUDA: return "CUDA launch failed";
        case USL_ERR_UNSUPPORTED: return "unsupported shape";
        case USL_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}

extern "C" int usl_pool3_fwd(const float* x, long long x_bs, long long x_cs,
                             int B, int C, int h, int w, float* out,
                             void* stream) {
    if (!x || !out || B <= 0 || C <= 0) return USL_ERR_ARG;
    usl::DeviceGuard guard(x);
    if (h < 3 || w < 3) return USL_ERR_UNSUPPORTED;
    const long long total = (long long)B * C * (h - 2) * (w - 2);
    usl::pool3_fwd_kernel<<<usl::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        x, x_bs, x_cs, B, C, h, w, out);
    return usl::check_launch();
}

extern "C" int usl_pool3_bwd(const float* grad_out, int B, int C, int h, int w,
                             float* grad_x, void* stream) {
    if (!grad_out || !grad_x || B <= 0 || C <= 0) return USL_ERR_ARG;
    usl::DeviceGuard guard(grad_out);
    if (h < 3 || w < 3) return USL_ERR_UNSUPPORTED;
    const long long total = (long long)B * C * h * w;
    usl::pool3_bwd_kernel<<<usl::grid_for(total), 256, 0, (cudaStream_t)stream>>>(
        grad_out, B, C, h, w, grad_x);
    return usl::check_launch();
}

extern "C" int usl_grad_rescale(const float* g, float* const* buffers,
                                const long long* counts, int n, void* stream) {
    if (!g || !buffers || !counts || n < 1 || n > USL_MAX_RESCALE) return USL_ERR_ARG;
    usl::DeviceGuard guard(g);
    usl::RescaleArgs A;
    long long total = 0;
    A.n = n;
    for (int i = 0; i < n; ++i) {
        if (!buffers[i] || counts[i] < 0) return USL_ERR_ARG;
        total += counts[i];
        A.buf[i] = buffers[i];
        A.end[i] = total;
    }
    if (total == 0) return USL_OK;
    usl::rescale_kernel<<<usl::grid_for(total), 256, 0, (cudaStream_t)stream>>>(g, A);
    usl::mark(12, (cudaStream_t)stream);
    return usl::check_launch();
}
