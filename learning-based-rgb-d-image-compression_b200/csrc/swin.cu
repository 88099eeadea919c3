// Token-side kernels of SymmetricalTransFormerUnited (reference models/stf_united.py): LayerNorm, the PatchMerging gather +
// LayerNorm, PixelShuffle(2) of PatchSplit / the end convs, and shifted-window multi-head attention with the relative
// position bias.  Tokens are the pixels of our NHWC views, so "B, H*W, C" of the reference is the layout every other kernel
// already uses; the Linear layers run as 1x1 convs on the conv kernels.  All arithmetic in fp32 whatever the storage type;
// every reduction has a fixed order (one warp per token / per (window, head)).
#include "common.cuh"

namespace {

constexpr int kLnMaxPerLane = 24;     // LayerNorm width <= 768 channels (4 x 192 in the deepest PatchMerging)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y[p, :] = (x[p, :] - mean) * rsqrt(var + eps) * gamma + beta over C channels (nn.LayerNorm: biased variance).
// gather2x2 (PatchMerging, stf_united.py:236-244): the token is the concatenation [x(2i, 2j) | x(2i+1, 2j) | x(2i, 2j+1) |
// x(2i+1, 2j+1)] of four input pixels of Cq = C / 4 channels each; output pixel (i, j) of a [N, H/2, W/2] map.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
layernorm_kernel(const TI *__restrict__ x, TO *__restrict__ y, int64_t npix, int C, int xs, int xo, int ys, int yo,
                 const float *__restrict__ gamma, const float *__restrict__ beta, float eps, int gather2x2, int Hin, int Win) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int Cq = C >> 2, Ho = Hin >> 1, Wo = Win >> 1;
    for (int64_t p = warp; p < npix; p += nwarps) {
        const TI *src[4];
        if (gather2x2) {
            const int j = (int)(p % Wo), i = (int)((p / Wo) % Ho);
            const int64_t n = p / ((int64_t)Wo * Ho);
            const int64_t base = (n * Hin + 2 * i) * Win + 2 * j;
            src[0] = x + base * xs + xo;
            src[1] = x + (base + Win) * xs + xo;
            src[2] = x + (base + 1) * xs + xo;
            src[3] = x + (base + Win + 1) * xs + xo;
        } else {
            src[0] = src[1] = src[2] = src[3] = x + p * xs + xo;
        }
        float v[kLnMaxPerLane];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kLnMaxPerLane; ++k) {
            const int c = lane + 32 * k;
            v[k] = 0.f;
            if (c < C) {
                v[k] = gather2x2 ? ElemIO<TI>::ld(src[c / Cq] + (c % Cq)) : ElemIO<TI>::ld(src[0] + c);
                s += v[k];
            }
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < kLnMaxPerLane; ++k) {
            const int c = lane + 32 * k;
            if (c < C) {
                const float d = v[k] - mean;
                q += d * d;
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        TO *dst = y + p * ys + yo;
#pragma unroll
        for (int k = 0; k < kLnMaxPerLane; ++k) {
            const int c = lane + 32 * k;
            if (c < C) ElemIO<TO>::st(dst + c, (v[k] - mean) * rstd * gamma[c] + beta[c]);
        }
    }
}

// nn.PixelShuffle(2) on NHWC: y[n, 2h + i, 2w + j, c] = x[n, h, w, 4 c + 2 i + j]
template <typename T>
__global__ void pixel_shuffle2_kernel(const T *__restrict__ x, T *__restrict__ y, int N, int H, int W, int Co, int xs, int xo,
                                      int ys, int yo) {
    const int64_t total = (int64_t)N * 2 * H * 2 * W * Co;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % Co);
        const int64_t op = e / Co;
        const int ox = (int)(op % (2 * W)), oy = (int)((op / (2 * W)) % (2 * H));
        const int64_t n = op / ((int64_t)4 * W * H);
        const int64_t ip = (n * H + (oy >> 1)) * W + (ox >> 1);
        y[op * ys + yo + c] = x[ip * xs + xo + 4 * c + 2 * (oy & 1) + (ox & 1)];
    }
}

// Shifted-window attention (stf_united.py:83-115, 162-212, 334-352).  One warp per (image, window, head); lane l owns the
// query tokens l, l + 32.  K and V of the window sit in shared memory; softmax is computed online (running maximum), so
// no score matrix is stored.  The cyclic shift is index arithmetic: token (hs, ws) of the shifted map is pixel
// ((hs + shift) % H, (ws + shift) % W), both when reading q / k / v and when writing the result back.
constexpr int kAttnWarps = 4, kAttnMaxT = 64, kAttnMaxHd = 32;

template <typename T>
__global__ void __launch_bounds__(kAttnWarps * 32)
window_attention_kernel(const T *__restrict__ qkv, T *__restrict__ out, int N, int H, int W, int C, int heads, int ws, int shift,
                        const float *__restrict__ bias_table, float scale, int qs, int qo, int os, int oo) {
    __shared__ float Ks[kAttnWarps][kAttnMaxT * kAttnMaxHd / 2], Vs[kAttnWarps][kAttnMaxT * kAttnMaxHd / 2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = C / heads, NT = ws * ws;
    const int nwx = W / ws, nwy = H / ws;
    const int64_t units = (int64_t)N * nwy * nwx * heads;
    for (int64_t u = (int64_t)blockIdx.x * kAttnWarps + warp; u < units; u += (int64_t)gridDim.x * kAttnWarps) {
        const int head = (int)(u % heads);
        const int64_t w_ = u / heads;
        const int wx = (int)(w_ % nwx), wy = (int)((w_ / nwx) % nwy);
        const int64_t n = w_ / ((int64_t)nwx * nwy);
        auto pixel = [&](int t) -> int64_t {        // token t of this window -> pixel index in the unshifted map
            const int hs = wy * ws + t / ws, wsx = wx * ws + t % ws;
            const int h = (hs + shift) % H, w = (wsx + shift) % W;
            return (n * H + h) * W + w;
        };
        auto region = [&](int t) -> int {           // id of the mask region of token t (BasicLayer.forward's img_mask)
            const int hs = wy * ws + t / ws, wsx = wx * ws + t % ws;
            const int rh = hs < H - ws ? 0 : (hs < H - shift ? 1 : 2);
            const int rw = wsx < W - ws ? 0 : (wsx < W - shift ? 1 : 2);
            return 3 * rh + rw;
        };
        __syncwarp();
        for (int e = lane; e < NT * hd; e += 32) {
            const int t = e / hd, d = e - t * hd;
            const T *row = qkv + pixel(t) * qs + qo + head * hd + d;
            Ks[warp][e] = ElemIO<T>::ld(row + C);
            Vs[warp][e] = ElemIO<T>::ld(row + 2 * C);
        }
        __syncwarp();
        for (int i = lane; i < NT; i += 32) {
            float q[kAttnMaxHd], o[kAttnMaxHd];
            const int64_t pi = pixel(i);
            const T *qrow = qkv + pi * qs + qo + head * hd;
#pragma unroll
            for (int d = 0; d < kAttnMaxHd; ++d) {
                q[d] = d < hd ? ElemIO<T>::ld(qrow + d) * scale : 0.f;
                o[d] = 0.f;
            }
            const int ri = shift ? region(i) : 0;
            const int iy = i / ws, ix = i % ws;
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < NT; ++j) {
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < kAttnMaxHd; ++d)
                    if (d < hd) s = fmaf(q[d], Ks[warp][j * hd + d], s);
                const int rel = (iy - j / ws + ws - 1) * (2 * ws - 1) + (ix - j % ws + ws - 1);
                s += bias_table[rel * heads + head];
                if (shift && region(j) != ri) s += -100.0f;
                const float mn = fmaxf(m, s);
                const float corr = __expf(m - mn), pj = __expf(s - mn);
                l = l * corr + pj;
#pragma unroll
                for (int d = 0; d < kAttnMaxHd; ++d)
                    if (d < hd) o[d] = fmaf(pj, Vs[warp][j * hd + d], o[d] * corr);
                m = mn;
            }
            const float inv = 1.f / l;
            T *orow = out + pi * os + oo + head * hd;
#pragma unroll
            for (int d = 0; d < kAttnMaxHd; ++d)
                if (d < hd) ElemIO<T>::st(orow + d, o[d] * inv);
        }
    }
}

}  // namespace

extern "C" int rgbd_layernorm(const void *x, int32_t x_dtype, void *y, int32_t y_dtype, int64_t npix, int32_t C, int32_t x_cstride,
                              int32_t x_coff, int32_t y_cstride, int32_t y_coff, const float *gamma, const float *beta, float eps,
                              int32_t gather2x2, int32_t Hin, int32_t Win, void *stream) {
    RGBD_CHECK_ARG(x && y && gamma && beta, "null pointer");
    RGBD_CHECK_ARG(npix > 0 && C > 0 && C <= 32 * kLnMaxPerLane, "C must be at most 768");
    RGBD_CHECK_ARG(!gather2x2 || (C % 4 == 0 && Hin % 2 == 0 && Win % 2 == 0 && Hin > 0 && Win > 0), "gather2x2 needs C % 4 == 0 and even H, W");
    const int grid = rgbd_grid_for(npix * 32, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define RGBD_LN(TI, TO)                                                                                                        \
    layernorm_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI *)x, (TO *)y, npix, C, x_cstride, x_coff, y_cstride, y_coff, gamma, \
                                                   beta, eps, gather2x2, Hin, Win)
    if (x_dtype == RGBD_DT_F32 && y_dtype == RGBD_DT_F32) RGBD_LN(float, float);
    else if (x_dtype == RGBD_DT_F32) RGBD_LN(float, __nv_bfloat16);
    else if (y_dtype == RGBD_DT_F32) RGBD_LN(__nv_bfloat16, float);
    else RGBD_LN(__nv_bfloat16, __nv_bfloat16);
#undef RGBD_LN
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_pixel_shuffle2(const void *x, void *y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t Cout,
                                   int32_t x_cstride, int32_t x_coff, int32_t y_cstride, int32_t y_coff, void *stream) {
    RGBD_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && Cout > 0, "arguments");
    const int grid = rgbd_grid_for((int64_t)N * 4 * H * W * Cout, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        pixel_shuffle2_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, N, H, W, Cout, x_cstride, x_coff, y_cstride, y_coff);
    else
        pixel_shuffle2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, N, H, W, Cout, x_cstride,
                                                                   x_coff, y_cstride, y_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_window_attention(const void *qkv, void *out, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                                     int32_t heads, int32_t window, int32_t shift, const float *bias_table, float scale,
                                     int32_t qkv_cstride, int32_t qkv_coff, int32_t out_cstride, int32_t out_coff, void *stream) {
    RGBD_CHECK_ARG(qkv && out && bias_table, "null pointer");
    RGBD_CHECK_ARG(N > 0 && heads > 0 && C % heads == 0 && window > 0, "dims");
    RGBD_CHECK_ARG(H % window == 0 && W % window == 0, "H and W must be multiples of the window (inputs are multiples of 64)");
    RGBD_CHECK_ARG(0 <= shift && shift < window, "0 <= shift < window");
    const int hd = C / heads, T = window * window;
    RGBD_CHECK_ARG(T <= kAttnMaxT && hd <= kAttnMaxHd && T * hd <= kAttnMaxT * kAttnMaxHd / 2, "window / head size too large");
    const int64_t units = (int64_t)N * (H / window) * (W / window) * heads;
    const int grid = rgbd_grid_for(units * 32, kAttnWarps * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        window_attention_kernel<float><<<grid, kAttnWarps * 32, 0, st>>>((const float *)qkv, (float *)out, N, H, W, C, heads, window, shift,
                                                                         bias_table, scale, qkv_cstride, qkv_coff, out_cstride, out_coff);
    else
        window_attention_kernel<__nv_bfloat16><<<grid, kAttnWarps * 32, 0, st>>>((const __nv_bfloat16 *)qkv, (__nv_bfloat16 *)out, N, H, W, C,
                                                                                 heads, window, shift, bias_table, scale, qkv_cstride,
                                                                                 qkv_coff, out_cstride, out_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
