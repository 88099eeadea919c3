// Shared helpers for the rgbd_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rgbd_b200.h"

extern "C" void rgbd_set_error(const char *fmt, ...);
extern "C" void rgbd_count_launch(int n);

#define RGBD_CHECK_ARG(cond, msg)                                  \
    do {                                                           \
        if (!(cond)) {                                             \
            rgbd_set_error("%s: invalid argument: %s", __func__, msg); \
            return RGBD_E_INVALID;                                 \
        }                                                          \
    } while (0)

#define RGBD_LAUNCH_CHECK()                                                        \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            rgbd_set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__)); \
            return RGBD_E_CUDA;                                                    \
        }                                                                          \
        rgbd_count_launch(1);                                                      \
    } while (0)

template <typename T> struct ElemIO;
template <> struct ElemIO<float> {
    static __device__ __forceinline__ float ld(const float *p) { return *p; }
    static __device__ __forceinline__ void st(float *p, float v) { *p = v; }
};
template <> struct ElemIO<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};

// value the consumer of an activation of type T will see after a store (fp32: identity)
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
}

static inline int rgbd_grid_for(int64_t work, int block, int max_blocks = 148 * 16) {
    int64_t g = (work + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}
