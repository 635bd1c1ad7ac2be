// Shared helpers for libwts (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/wts.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libwts is written for sm_100a (B200) only"
#endif

namespace wts {

void set_error(const char* fmt, ...);

#define WTS_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::wts::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,         \
                             cudaGetErrorString(_e));                                     \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

#define WTS_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::wts::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,     \
                             cudaGetErrorString(_e));                                     \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

}  // namespace wts
