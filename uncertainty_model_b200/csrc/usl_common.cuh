// Launch/error plumbing shared by the translation units of libusl.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "usl_math.cuh"
#include "../../include/usl.h"

namespace usl {

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    return (e == cudaSuccess) ? USL_OK : USL_ERR_CUDA;
}

inline int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) !=
                cudaSuccess || n <= 0)
            return 148;
        cached = n;
    }
    return cached;
}

// Every entry point launches on the device that owns its tensors: the guard
// makes that device current for the call and restores the caller's afterwards
// (the reference's DDP launcher moves everything `.to(cuda:i)` without ever
// calling torch.cuda.set_device, parallel_main.py:152-160).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const void* p) {
        if (!p) return;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return; }
        if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return;
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) return;
        if (cur != a.device && cudaSetDevice(a.device) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// Side streams for launches that may run concurrently with what is on the
// caller's stream: one small pool per host thread and device, created on first
// use and never destroyed (immutable afterwards; nothing is shared between
// host threads).  side[0 .. USL_MAX_SCALES-2]: per-scale launches of the fused
// kernel; `first`: a stream of the highest priority for the launch whose CTAs
// must be placed before those of the others (the largest scale's: each takes
// a whole SM, and the step is as long as they are).
struct StreamPool {
    bool ready = false, failed = false;
    cudaStream_t side[USL_MAX_SCALES];
    cudaStream_t first;
    cudaEvent_t fork, fork2, join[USL_MAX_SCALES], join_first;
};
inline StreamPool* stream_pool() {
    constexpr int MAX_DEV = 64;
    thread_local StreamPool pools[MAX_DEV];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    StreamPool& p = pools[dev];
    if (p.failed) return nullptr;
    if (!p.ready) {
        int lo = 0, hi = 0;
        bool ok = cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess &&
                  cudaStreamCreateWithPriority(&p.first, cudaStreamNonBlocking, hi) == cudaSuccess &&
                  cudaEventCreateWithFlags(&p.join_first, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&p.fork2, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; ok && i < USL_MAX_SCALES; ++i)
            ok = cudaStreamCreateWithFlags(&p.side[i], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { p.failed = true; cudaGetLastError(); return nullptr; }
        p.ready = true;
    }
    return &p;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace usl
