// Launch/error plumbing shared by the translation units of libusl.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "usl_math.cuh"
#include "../../include/usl.h"

namespace usl {

// Kernel launches issued by the library so far (usl_launch_count: bench.py
// reports the launches of a step from it instead of claiming a constant).
inline std::atomic<long long>& launch_counter() {
    static std::atomic<long long> n{0};
    return n;
}

// (launches that are followed by more launches before the next check_launch)
inline void count_launches(int n) {
    launch_counter().fetch_add(n, std::memory_order_relaxed);
}

inline int check_launch() {
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return (e == cudaSuccess) ? USL_OK : USL_ERR_CUDA;
}

// SM count of the current device (cached per device; the cache is write-once
// per slot, so concurrent first calls agree).
inline int num_sms() {
    constexpr int MAX_DEV = 64;
    static std::atomic<int> cached[MAX_DEV];
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 148;
    n = cached[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) !=
            cudaSuccess || n <= 0)
        return 148;
    cached[dev].store(n, std::memory_order_relaxed);
    return n;
}

// Tuning / profiling knobs, read from the environment ONCE (first use; C++11
// guarantees the initialisation is thread safe) and immutable afterwards: the
// launch paths never call getenv.  Unset = the built-in heuristics.
struct Knobs {
    int no_col;          // USL_NO_COL: keep every call on the general strip kernels
    int col_maxt2;       // USL_COL_MAXT2: widest two-view unit, in threads
    int col_r0, col_r;   // USL_COL_R0 / USL_COL_R: strip heights (scale 0 / others)
    int col_no_tma;      // USL_COL_NO_TMA
    int col_only;        // USL_COL_ONLY: bit mask of the scales to launch
    int col_serial;      // USL_COL_SERIAL: all scales on the caller's stream
    int col_no_priority; // USL_COL_NO_PRIORITY
    int fwd_tw, fwd_r, bwd_tw, bwd_r;   // USL_{FWD,BWD}_{TW,R}: general kernels
    int cons_r;          // USL_CONS_R
    int scatter_v1;      // USL_SCATTER_V1
    int scatter_warp_per_row;  // USL_SCATTER_WARP_PER_ROW: round 1's transposed-warp kernel
    int scatter_after_all;     // USL_SCATTER_AFTER_ALL: one transposed-warp launch behind all column kernels
};
inline int knob_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}
inline const Knobs& knobs() {
    static const Knobs k = [] {
        Knobs x;
        x.no_col = knob_int("USL_NO_COL", 0);
        x.col_maxt2 = knob_int("USL_COL_MAXT2", 0);
        x.col_r0 = knob_int("USL_COL_R0", 0);
        x.col_r = knob_int("USL_COL_R", 0);
        x.col_no_tma = knob_int("USL_COL_NO_TMA", 0);
        x.col_only = knob_int("USL_COL_ONLY", 0xff);
        x.col_serial = knob_int("USL_COL_SERIAL", 0);
        x.col_no_priority = knob_int("USL_COL_NO_PRIORITY", 0);
        x.fwd_tw = knob_int("USL_FWD_TW", 0);
        x.fwd_r = knob_int("USL_FWD_R", 0);
        x.bwd_tw = knob_int("USL_BWD_TW", 0);
        x.bwd_r = knob_int("USL_BWD_R", 0);
        x.cons_r = knob_int("USL_CONS_R", 0);
        x.scatter_v1 = knob_int("USL_SCATTER_V1", 0);
        x.scatter_warp_per_row = knob_int("USL_SCATTER_WARP_PER_ROW", 0);
        x.scatter_after_all = knob_int("USL_SCATTER_AFTER_ALL", 0);
        return x;
    }();
    return k;
}

// Profiling aid (usl_debug_timeline): when a buffer is registered, `mark`
// enqueues a one-thread kernel that writes %globaltimer into slot `idx` -- a
// time line of the concurrent per-scale launches that also works inside a CUDA
// graph replay (ncu serialises the launches, events do not exist in a graph).
std::atomic<unsigned long long*>& timeline_buffer();
void mark(int idx, cudaStream_t st);

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
    cudaEvent_t col_done[USL_MAX_SCALES];   // a scale's column kernel (not its scatter) has finished
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
                 cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p.col_done[i], cudaEventDisableTiming) == cudaSuccess;
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
