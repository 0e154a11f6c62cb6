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

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace usl
