// Entropy-model elementwise kernels of the Bi-CEE context model:
// checkerboard squeeze/unsqueeze (reference utils/ckbd.py:6-125) fused with
// GaussianConditional.build_indexes / quantize / _likelihood
// (CompressAI/compressai/entropy_models/entropy_models.py:118-146, 534-568) and the
// factorised-prior bottleneck on z (entropy_models.py:369-446).
//
// Layout: latents and Gaussian parameters are NHWC (channel fastest); the coder's symbol
// order is the reference's [C, H, W/2] row-major per image (w' fastest).  Every kernel that
// converts between the two goes through a 32x32 shared-memory tile so that both the NHWC
// side (32 consecutive channels = 128 B) and the stream side (32 consecutive w') are
// coalesced.  All of these are HBM-bound: 24 B of algorithmic traffic per latent.
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int kTile = 32;
constexpr int kThreads = 256;

__device__ __forceinline__ int ckbd_col(int h, int wq, int parity) {
    // anchor (parity 0): even rows keep odd columns, odd rows keep even columns (ckbd.py:51-56);
    // non-anchor (parity 1): the complement (ckbd.py:59-64)
    return 2 * wq + ((h + parity + 1) & 1);
}

__device__ __forceinline__ float lower_bound_f(float v, float bound) {
    return v < bound ? bound : v;  // torch.max(x, bound): NaN stays NaN
}

// idx = (n-1) - #{i < n-1 : s <= table[i]}   (entropy_models.py:561-568)
__device__ __forceinline__ int scale_index(float s, const float *table, int n) {
    int lo = 0, hi = n - 1;  // first i in [0, n-1) with s <= table[i], else n-1
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s <= table[mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

struct TileCoord {
    int n, h, c0, w0;
};
__device__ __forceinline__ TileCoord tile_coord(int H) {
    TileCoord t;
    t.n = blockIdx.x / H;
    t.h = blockIdx.x % H;
    t.c0 = blockIdx.y * kTile;
    t.w0 = blockIdx.z * kTile;
    return t;
}

template <typename TY>
__global__ void __launch_bounds__(kThreads)
ckbd_quantize_index_kernel(const float *__restrict__ y, int y_cstride, int y_coff,
                           const float *__restrict__ params, const float *__restrict__ scale_table,
                           int n_scales, float scale_bound, int H, int W, int g, int parity,
                           int32_t *__restrict__ sym, uint8_t *__restrict__ idx, int64_t stream_stride,
                           int64_t chunk_off, TY *__restrict__ yhat, int yhat_cstride, int yhat_coff) {
    __shared__ int32_t s_sym[kTile][kTile + 1];
    __shared__ uint8_t s_idx[kTile][kTile + 4];
    __shared__ float s_table[256];
    for (int i = threadIdx.x; i < n_scales; i += kThreads) s_table[i] = scale_table[i];
    __syncthreads();
    const TileCoord t = tile_coord(H);
    const int Wq = W >> 1;
    for (int e = threadIdx.x; e < kTile * kTile; e += kThreads) {
        const int cl = e & 31, wl = e >> 5;
        const int c = t.c0 + cl, wq = t.w0 + wl;
        if (c < g && wq < Wq) {
            const int w = ckbd_col(t.h, wq, parity);
            const int64_t pix = ((int64_t)t.n * H + t.h) * W + w;
            const float scale = params[pix * (2 * g) + c];
            const float mean = params[pix * (2 * g) + g + c];
            const float yv = y[pix * y_cstride + y_coff + c];
            const float r = rintf(yv - mean);  // torch.round = half-to-even
            const int32_t s = (int32_t)r;
            s_sym[cl][wl] = s;
            s_idx[cl][wl] = (uint8_t)scale_index(lower_bound_f(scale, scale_bound), s_table, n_scales);
            ElemIO<TY>::st(yhat + pix * yhat_cstride + yhat_coff + c, (float)s + mean);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTile * kTile; e += kThreads) {
        const int wl = e & 31, cl = e >> 5;
        const int c = t.c0 + cl, wq = t.w0 + wl;
        if (c < g && wq < Wq) {
            const int64_t j = (int64_t)t.n * stream_stride + chunk_off + ((int64_t)c * H + t.h) * Wq + wq;
            sym[j] = s_sym[cl][wl];
            idx[j] = s_idx[cl][wl];
        }
    }
}

__global__ void __launch_bounds__(kThreads)
ckbd_index_kernel(const float *__restrict__ params, const float *__restrict__ scale_table, int n_scales,
                  float scale_bound, int H, int W, int g, int parity, uint8_t *__restrict__ idx,
                  int64_t stream_stride, int64_t chunk_off) {
    __shared__ uint8_t s_idx[kTile][kTile + 4];
    __shared__ float s_table[256];
    for (int i = threadIdx.x; i < n_scales; i += kThreads) s_table[i] = scale_table[i];
    __syncthreads();
    const TileCoord t = tile_coord(H);
    const int Wq = W >> 1;
    for (int e = threadIdx.x; e < kTile * kTile; e += kThreads) {
        const int cl = e & 31, wl = e >> 5;
        const int c = t.c0 + cl, wq = t.w0 + wl;
        if (c < g && wq < Wq) {
            const int w = ckbd_col(t.h, wq, parity);
            const int64_t pix = ((int64_t)t.n * H + t.h) * W + w;
            const float scale = params[pix * (2 * g) + c];
            s_idx[cl][wl] = (uint8_t)scale_index(lower_bound_f(scale, scale_bound), s_table, n_scales);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTile * kTile; e += kThreads) {
        const int wl = e & 31, cl = e >> 5;
        const int c = t.c0 + cl, wq = t.w0 + wl;
        if (c < g && wq < Wq)
            idx[(int64_t)t.n * stream_stride + chunk_off + ((int64_t)c * H + t.h) * Wq + wq] = s_idx[cl][wl];
    }
}

template <typename TY>
__global__ void __launch_bounds__(kThreads)
ckbd_dequant_scatter_kernel(const int32_t *__restrict__ sym, int64_t stream_stride, int64_t chunk_off,
                            const float *__restrict__ params, int H, int W, int g, int parity,
                            TY *__restrict__ yhat, int yhat_cstride, int yhat_coff) {
    __shared__ int32_t s_sym[kTile][kTile + 1];
    const TileCoord t = tile_coord(H);
    const int Wq = W >> 1;
    for (int e = threadIdx.x; e < kTile * kTile; e += kThreads) {
        const int wl = e & 31, cl = e >> 5;
        const int c = t.c0 + cl, wq = t.w0 + wl;
        if (c < g && wq < Wq)
            s_sym[cl][wl] = sym[(int64_t)t.n * stream_stride + chunk_off + ((int64_t)c * H + t.h) * Wq + wq];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTile * kTile; e += kThreads) {
        const int cl = e & 31, wl = e >> 5;
        const int c = t.c0 + cl, wq = t.w0 + wl;
        if (c < g && wq < Wq) {
            const int w = ckbd_col(t.h, wq, parity);
            const int64_t pix = ((int64_t)t.n * H + t.h) * W + w;
            const float mean = params[pix * (2 * g) + g + c];
            ElemIO<TY>::st(yhat + pix * yhat_cstride + yhat_coff + c, (float)s_sym[cl][wl] + mean);
        }
    }
}

__device__ __forceinline__ float std_cumulative(float v) {
    // GaussianConditional._standardized_cumulative (entropy_models.py:489-494)
    return 0.5f * erfcf(-0.70710678118654752440f * v);
}

template <typename TY>
__global__ void __launch_bounds__(kThreads)
ckbd_ste_likelihood_kernel(const float *__restrict__ y, int y_cstride, int y_coff,
                           const float *__restrict__ params, float scale_bound, float lik_bound, int N,
                           int H, int W, int g, int parity, TY *__restrict__ yhat, int yhat_cstride,
                           int yhat_coff, float *__restrict__ lik, int lik_C, int lik_coff) {
    const int Wq = W >> 1;
    const int64_t total = (int64_t)N * H * Wq * g;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % g);
        int64_t r = e / g;
        const int wq = (int)(r % Wq);
        r /= Wq;
        const int h = (int)(r % H);
        const int n = (int)(r / H);
        const int w = ckbd_col(h, wq, parity);
        const int64_t pix = ((int64_t)n * H + h) * W + w;
        const float scale = lower_bound_f(params[pix * (2 * g) + c], scale_bound);
        const float mean = params[pix * (2 * g) + g + c];
        const float yv = y[pix * y_cstride + y_coff + c];
        // ste_round(x) = round(x) - x + x   (compressai/ops/ops.py:32), then + mean
        const float d = yv - mean;
        const float ste = (rintf(d) - d) + d;
        ElemIO<TY>::st(yhat + pix * yhat_cstride + yhat_coff + c, ste + mean);
        // GaussianConditional.forward: dequantise, then likelihood of (deq - mean)
        const float deq = rintf(d) + mean;
        const float v = fabsf(deq - mean);
        const float upper = std_cumulative((0.5f - v) / scale);
        const float lower = std_cumulative((-0.5f - v) / scale);
        float l = upper - lower;
        l = l < lik_bound ? lik_bound : l;
        lik[(((int64_t)n * lik_C + lik_coff + c) * H + h) * W + w] = l;
    }
}

// ------------------------------- factorised prior (z) ------------------------------------
template <typename TZ>
__global__ void eb_quantize_kernel(const float *__restrict__ z, int z_cstride, int N, int HW, int C,
                                   const float *__restrict__ medians, int32_t *__restrict__ sym,
                                   uint8_t *__restrict__ idx, TZ *__restrict__ zhat, int zhat_cstride,
                                   int zhat_coff) {
    const int64_t total = (int64_t)N * C * HW;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(e % HW);
        const int c = (int)((e / HW) % C);
        const int n = (int)(e / ((int64_t)HW * C));
        const float med = medians[c];
        const float r = rintf(z[((int64_t)n * HW + p) * z_cstride + c] - med);
        const int32_t s = (int32_t)r;
        sym[e] = s;
        idx[e] = (uint8_t)c;
        ElemIO<TZ>::st(zhat + ((int64_t)n * HW + p) * zhat_cstride + zhat_coff + c, (float)s + med);
    }
}

template <typename TZ>
__global__ void eb_dequantize_kernel(const int32_t *__restrict__ sym, int N, int HW, int C,
                                     const float *__restrict__ medians, TZ *__restrict__ zhat,
                                     int zhat_cstride, int zhat_coff) {
    const int64_t total = (int64_t)N * C * HW;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(e % HW);
        const int c = (int)((e / HW) % C);
        const int n = (int)(e / ((int64_t)HW * C));
        ElemIO<TZ>::st(zhat + ((int64_t)n * HW + p) * zhat_cstride + zhat_coff + c, (float)sym[e] + medians[c]);
    }
}

// per-channel packed parameters (58 floats): softplus(matrix_i), bias_i, tanh(factor_i)
__device__ __forceinline__ float eb_logits(const float *P, float v) {
    float a[3], b[3];
    // layer 0: 1 -> 3
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float t = P[r] * v + P[3 + r];
        a[r] = t + P[6 + r] * tanhf(t);
    }
    const float *Q = P + 9;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float t = Q[r * 3 + 0] * a[0];
            t += Q[r * 3 + 1] * a[1];
            t += Q[r * 3 + 2] * a[2];
            t += Q[9 + r];
            b[r] = t + Q[12 + r] * tanhf(t);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) a[r] = b[r];
        Q += 15;
    }
    float t = Q[0] * a[0];
    t += Q[1] * a[1];
    t += Q[2] * a[2];
    return t + Q[3];
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

template <typename TZ>
__global__ void eb_likelihood_kernel(const float *__restrict__ z, int z_cstride, int N, int HW, int C,
                                     const float *__restrict__ ebp, float lik_bound, TZ *__restrict__ zhat,
                                     int zhat_cstride, int zhat_coff, float *__restrict__ lik) {
    const int64_t total = (int64_t)N * C * HW;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(e % HW);
        const int c = (int)((e / HW) % C);
        const int n = (int)(e / ((int64_t)HW * C));
        const float *P = ebp + (int64_t)c * 59;
        const float med = P[58];
        const float zv = z[((int64_t)n * HW + p) * z_cstride + c];
        const float d = zv - med;
        const float deq = rintf(d) + med;  // quantize(..., "dequantize", medians)
        const float lower = eb_logits(P, deq - 0.5f);
        const float upper = eb_logits(P, deq + 0.5f);
        const float sum = lower + upper;
        const float sign = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
        float l = fabsf(sigmoidf_(sign * upper) - sigmoidf_(sign * lower));
        l = l < lik_bound ? lik_bound : l;
        lik[e] = l;  // NCHW
        const float ste = (rintf(d) - d) + d;  // forward(): ste_round(z - med) + med (elic_united.py:241-245)
        ElemIO<TZ>::st(zhat + ((int64_t)n * HW + p) * zhat_cstride + zhat_coff + c, ste + med);
    }
}

dim3 tile_grid(int N, int H, int W, int g) {
    return dim3((unsigned)(N * H), (unsigned)((g + kTile - 1) / kTile), (unsigned)(((W >> 1) + kTile - 1) / kTile));
}

}  // namespace

#define CKBD_ARGS_OK()                                                          \
    RGBD_CHECK_ARG(N > 0 && H > 0 && W > 0 && g > 0, "dims");                   \
    RGBD_CHECK_ARG((W & 1) == 0, "W must be even");                             \
    RGBD_CHECK_ARG(parity == 0 || parity == 1, "parity");                       \
    RGBD_CHECK_ARG((int64_t)N * H < 2147483647LL, "N*H too large")

extern "C" int rgbd_ckbd_quantize_index(const float *y, int32_t y_cstride, int32_t y_coff,
                                        const float *params, const float *scale_table, int32_t n_scales,
                                        float scale_bound, int32_t N, int32_t H, int32_t W, int32_t g,
                                        int32_t parity, int32_t *sym, uint8_t *idx, int64_t stream_stride,
                                        int64_t chunk_off, void *yhat, int32_t yhat_dtype,
                                        int32_t yhat_cstride, int32_t yhat_coff, void *stream) {
    RGBD_CHECK_ARG(y && params && scale_table && sym && idx && yhat, "null pointer");
    CKBD_ARGS_OK();
    RGBD_CHECK_ARG(n_scales >= 1 && n_scales <= 256, "n_scales must be in 1..256");
    const dim3 grid = tile_grid(N, H, W, g);
    cudaStream_t st = (cudaStream_t)stream;
    if (yhat_dtype == RGBD_DT_F32)
        ckbd_quantize_index_kernel<float><<<grid, kThreads, 0, st>>>(
            y, y_cstride, y_coff, params, scale_table, n_scales, scale_bound, H, W, g, parity, sym, idx,
            stream_stride, chunk_off, (float *)yhat, yhat_cstride, yhat_coff);
    else
        ckbd_quantize_index_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
            y, y_cstride, y_coff, params, scale_table, n_scales, scale_bound, H, W, g, parity, sym, idx,
            stream_stride, chunk_off, (__nv_bfloat16 *)yhat, yhat_cstride, yhat_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_ckbd_index(const float *params, const float *scale_table, int32_t n_scales,
                               float scale_bound, int32_t N, int32_t H, int32_t W, int32_t g, int32_t parity,
                               uint8_t *idx, int64_t stream_stride, int64_t chunk_off, void *stream) {
    RGBD_CHECK_ARG(params && scale_table && idx, "null pointer");
    CKBD_ARGS_OK();
    RGBD_CHECK_ARG(n_scales >= 1 && n_scales <= 256, "n_scales must be in 1..256");
    ckbd_index_kernel<<<tile_grid(N, H, W, g), kThreads, 0, (cudaStream_t)stream>>>(
        params, scale_table, n_scales, scale_bound, H, W, g, parity, idx, stream_stride, chunk_off);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_ckbd_dequant_scatter(const int32_t *sym, int64_t stream_stride, int64_t chunk_off,
                                         const float *params, int32_t N, int32_t H, int32_t W, int32_t g,
                                         int32_t parity, void *yhat, int32_t yhat_dtype,
                                         int32_t yhat_cstride, int32_t yhat_coff, void *stream) {
    RGBD_CHECK_ARG(sym && params && yhat, "null pointer");
    CKBD_ARGS_OK();
    const dim3 grid = tile_grid(N, H, W, g);
    cudaStream_t st = (cudaStream_t)stream;
    if (yhat_dtype == RGBD_DT_F32)
        ckbd_dequant_scatter_kernel<float><<<grid, kThreads, 0, st>>>(
            sym, stream_stride, chunk_off, params, H, W, g, parity, (float *)yhat, yhat_cstride, yhat_coff);
    else
        ckbd_dequant_scatter_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(
            sym, stream_stride, chunk_off, params, H, W, g, parity, (__nv_bfloat16 *)yhat, yhat_cstride,
            yhat_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_ckbd_ste_likelihood(const float *y, int32_t y_cstride, int32_t y_coff,
                                        const float *params, float scale_bound, float lik_bound, int32_t N,
                                        int32_t H, int32_t W, int32_t g, int32_t parity, void *yhat,
                                        int32_t yhat_dtype, int32_t yhat_cstride, int32_t yhat_coff,
                                        float *lik, int32_t lik_C, int32_t lik_coff, void *stream) {
    RGBD_CHECK_ARG(y && params && yhat && lik, "null pointer");
    CKBD_ARGS_OK();
    const int64_t total = (int64_t)N * H * (W / 2) * g;
    const int grid = rgbd_grid_for(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (yhat_dtype == RGBD_DT_F32)
        ckbd_ste_likelihood_kernel<float><<<grid, 256, 0, st>>>(
            y, y_cstride, y_coff, params, scale_bound, lik_bound, N, H, W, g, parity, (float *)yhat,
            yhat_cstride, yhat_coff, lik, lik_C, lik_coff);
    else
        ckbd_ste_likelihood_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            y, y_cstride, y_coff, params, scale_bound, lik_bound, N, H, W, g, parity, (__nv_bfloat16 *)yhat,
            yhat_cstride, yhat_coff, lik, lik_C, lik_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_eb_quantize(const float *z, int32_t z_cstride, int32_t N, int32_t HW, int32_t C,
                                const float *medians, int32_t *sym, uint8_t *idx, void *zhat,
                                int32_t zhat_dtype, int32_t zhat_cstride, int32_t zhat_coff, void *stream) {
    RGBD_CHECK_ARG(z && medians && sym && idx && zhat, "null pointer");
    RGBD_CHECK_ARG(N > 0 && HW > 0 && C > 0 && C <= 256, "dims (C <= 256: index is a byte)");
    const int grid = rgbd_grid_for((int64_t)N * HW * C, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (zhat_dtype == RGBD_DT_F32)
        eb_quantize_kernel<float><<<grid, 256, 0, st>>>(z, z_cstride, N, HW, C, medians, sym, idx, (float *)zhat,
                                                        zhat_cstride, zhat_coff);
    else
        eb_quantize_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(z, z_cstride, N, HW, C, medians, sym, idx,
                                                                (__nv_bfloat16 *)zhat, zhat_cstride, zhat_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_eb_dequantize(const int32_t *sym, int32_t N, int32_t HW, int32_t C, const float *medians,
                                  void *zhat, int32_t zhat_dtype, int32_t zhat_cstride, int32_t zhat_coff,
                                  void *stream) {
    RGBD_CHECK_ARG(sym && medians && zhat, "null pointer");
    RGBD_CHECK_ARG(N > 0 && HW > 0 && C > 0, "dims");
    const int grid = rgbd_grid_for((int64_t)N * HW * C, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (zhat_dtype == RGBD_DT_F32)
        eb_dequantize_kernel<float><<<grid, 256, 0, st>>>(sym, N, HW, C, medians, (float *)zhat, zhat_cstride,
                                                          zhat_coff);
    else
        eb_dequantize_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(sym, N, HW, C, medians, (__nv_bfloat16 *)zhat,
                                                                  zhat_cstride, zhat_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_eb_likelihood(const float *z, int32_t z_cstride, int32_t N, int32_t HW, int32_t C,
                                  const float *eb_params, float lik_bound, void *zhat, int32_t zhat_dtype,
                                  int32_t zhat_cstride, int32_t zhat_coff, float *lik, void *stream) {
    RGBD_CHECK_ARG(z && eb_params && zhat && lik, "null pointer");
    RGBD_CHECK_ARG(N > 0 && HW > 0 && C > 0, "dims");
    const int grid = rgbd_grid_for((int64_t)N * HW * C, 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (zhat_dtype == RGBD_DT_F32)
        eb_likelihood_kernel<float><<<grid, 128, 0, st>>>(z, z_cstride, N, HW, C, eb_params, lik_bound,
                                                          (float *)zhat, zhat_cstride, zhat_coff, lik);
    else
        eb_likelihood_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(z, z_cstride, N, HW, C, eb_params, lik_bound,
                                                                  (__nv_bfloat16 *)zhat, zhat_cstride, zhat_coff,
                                                                  lik);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
