// Small spatial / channel kernels of the transforms: SE_Block channel attention
// (reference modules/transform/attention.py:52-67), ESA's max-pool (attention.py:88) and the
// NCHW <-> NHWC boundary conversions (+ the final clamp of elic_united.py:452).
#include "common.cuh"
#include <float.h>

namespace {

// partial[n][chunk][c] = sum over the chunk's pixels, fixed order; channels [c0, c1) of the view, rows of the partial
// table `pstride` floats apart (a caller that keeps the table can refresh only the channels that changed)
template <typename T>
__global__ void se_partial_kernel(const T *__restrict__ x, int HW, int c0, int c1, int cstride, int coff,
                                  int nchunk, float *__restrict__ partial, int pstride) {
    const int n = blockIdx.x / nchunk, chunk = blockIdx.x % nchunk;
    const int per = (HW + nchunk - 1) / nchunk;
    const int p0 = chunk * per;
    const int p1 = min(HW, p0 + per);
    for (int c = c0 + blockIdx.y * blockDim.x + threadIdx.x; c < c1; c += gridDim.y * blockDim.x) {
        const T *px = x + ((int64_t)n * HW + p0) * cstride + coff + c;
        float s = 0.f;
        for (int p = p0; p < p1; ++p, px += cstride) s += ElemIO<T>::ld(px);
        partial[((int64_t)n * nchunk + chunk) * pstride + c] = s;
    }
}

// SE squeeze-excite MLP, spread over many CTAs (the one-CTA-per-image form was latency bound):
//   se_mean_kernel   : mean[n][c]   = (sum of the partials in fixed order) / HW
//   se_hidden_kernel : hid[n][r]    = relu(W1[r,:] . mean[n,:])          one warp per (n, r)
//   se_gate_kernel   : scale[n][c]  = sigmoid(W2[c,:] . hid[n,:]) (+1)   one warp per (n, c)
__global__ void se_mean_kernel(const float *__restrict__ partial, int pstride, int nchunk, int HW, int N, int C,
                               float *__restrict__ mean) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)N * C) return;
    const int n = (int)(e / C), c = (int)(e % C);
    float s = 0.f;
    for (int k = 0; k < nchunk; ++k) s += partial[((int64_t)n * nchunk + k) * pstride + c];
    mean[e] = s / (float)HW;
}

__global__ void se_hidden_kernel(const float *__restrict__ mean, const float *__restrict__ w1, int N, int C, int Cr,
                                 float *__restrict__ hid) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (int64_t)N * Cr) return;
    const int n = (int)(wid / Cr), r = (int)(wid % Cr);
    const float *m = mean + (int64_t)n * C, *w = w1 + (int64_t)r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w[c], m[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) hid[wid] = s > 0.f ? s : 0.f;
}

__global__ void se_gate_kernel(const float *__restrict__ hid, const float *__restrict__ w2, int N, int C, int Cr,
                               int plus_one, float *__restrict__ scale) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (int64_t)N * C) return;
    const int n = (int)(wid / C), c = (int)(wid % C);
    const float *h = hid + (int64_t)n * Cr, *w = w2 + (int64_t)c * Cr;
    float s = 0.f;
    for (int r = lane; r < Cr; r += 32) s = fmaf(w[r], h[r], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const float g = 1.0f / (1.0f + expf(-s));
        scale[wid] = plus_one ? 1.0f + g : g;
    }
}

template <typename T>
__global__ void maxpool7s3_kernel(const T *__restrict__ x, T *__restrict__ y, int N, int H, int W, int C,
                                  int Ho, int Wo) {
    const int64_t total = (int64_t)N * Ho * Wo * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        int64_t r = e / C;
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho);
        const int n = (int)(r / Ho);
        float m = -FLT_MAX;
        for (int ky = 0; ky < 7; ++ky)
            for (int kx = 0; kx < 7; ++kx) {
                const float v = ElemIO<T>::ld(x + (((int64_t)n * H + oy * 3 + ky) * W + ox * 3 + kx) * C + c);
                m = v > m ? v : m;
            }
        ElemIO<T>::st(y + e, m);
    }
}

// split3: write [hi | lo | hi] with hi = T(x), lo = T(x - hi): a two-term bf16 expansion of the fp32
// image (exact for 16-bit depth), consumed by a first conv packed as [w_hi | w_hi | w_lo], i.e.
// x*w ~= hi*w_hi + lo*w_hi + hi*w_lo, all inside the one K=16 MMA a tap costs anyway.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float *__restrict__ x, T *__restrict__ y, int N, int C, int H, int W,
                                    int cstride, int coff, int split3) {
    const int64_t total = (int64_t)N * C * H * W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = e % ((int64_t)H * W);
        const int c = (int)((e / ((int64_t)H * W)) % C);
        const int n = (int)(e / ((int64_t)H * W * C));
        const float v = x[e];
        if (split3 == 2) {
            // split + space-to-depth: pixel (2Y + ry, 2X + rx) -> pixel (Y, X) of an H/2 x W/2 map, channel block
            // (2 ry + rx) of 3C channels [hi | lo | hi]
            const int yy = (int)(hw / W), xx = (int)(hw % W);
            T *dst = y + (((int64_t)n * (H / 2) + (yy >> 1)) * (W / 2) + (xx >> 1)) * cstride + coff +
                     (2 * (yy & 1) + (xx & 1)) * 3 * C + c;
            const float hi = round_to<T>(v);
            ElemIO<T>::st(dst, v);
            ElemIO<T>::st(dst + C, v - hi);
            ElemIO<T>::st(dst + 2 * C, v);
            continue;
        }
        T *dst = y + ((int64_t)n * H * W + hw) * cstride + coff + c;
        ElemIO<T>::st(dst, v);
        if (split3) {
            const float hi = round_to<T>(v);
            ElemIO<T>::st(dst + C, v - hi);
            ElemIO<T>::st(dst + 2 * C, v);
        }
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T *__restrict__ x, float *__restrict__ y, int N, int C, int H, int W,
                                    int cstride, int coff, int clamp01) {
    const int64_t total = (int64_t)N * C * H * W;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = e % ((int64_t)H * W);
        const int c = (int)((e / ((int64_t)H * W)) % C);
        const int n = (int)(e / ((int64_t)H * W * C));
        float v = ElemIO<T>::ld(x + ((int64_t)n * H * W + hw) * cstride + coff + c);
        if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
        y[e] = v;
    }
}

template <typename T>
__global__ void copy_view_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t npix, int C, int xs, int xo,
                                 int ys, int yo) {
    const int64_t total = npix * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / C;
        const int c = (int)(e % C);
        y[p * ys + yo + c] = x[p * xs + xo + c];
    }
}

__global__ void cast_view_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y, int64_t npix, int C,
                                      int xs, int xo, int ys, int yo) {
    const int64_t total = npix * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / C;
        const int c = (int)(e % C);
        y[p * ys + yo + c] = __float2bfloat16_rn(x[p * xs + xo + c]);
    }
}

template <typename T>
__global__ void scale_channels_kernel(const T *__restrict__ x, T *__restrict__ y, const float *__restrict__ scale,
                                      int64_t pix_per_img, int64_t npix, int C, int xs, int xo, int ys, int yo) {
    const int64_t total = npix * C;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = e / C;
        const int c = (int)(e % C);
        const int n = (int)(p / pix_per_img);
        ElemIO<T>::st(y + p * ys + yo + c, ElemIO<T>::ld(x + p * xs + xo + c) * scale[(int64_t)n * C + c]);
    }
}

// per-image filter copies with the input-channel gate folded in (8 input channels = 16 B per thread)
__global__ void scale_weights_kernel(const float *__restrict__ w, const float *__restrict__ scale,
                                     __nv_bfloat16 *__restrict__ out, int64_t per_image, int cin_pad, int Cin, int N) {
    const int64_t groups = per_image / 8;
    const int64_t total = groups * N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(e / groups);
        const int64_t g = e - (int64_t)n * groups;
        const int ci = (int)((g * 8) % cin_pad);
        const float4 a = reinterpret_cast<const float4 *>(w + g * 8)[0];
        const float4 b = reinterpret_cast<const float4 *>(w + g * 8)[1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t packed[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float s0 = ci + 2 * i < Cin ? scale[(int64_t)n * Cin + ci + 2 * i] : 0.f;
            const float s1 = ci + 2 * i + 1 < Cin ? scale[(int64_t)n * Cin + ci + 2 * i + 1] : 0.f;
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i] * s0, v[2 * i + 1] * s1);
            packed[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        reinterpret_cast<uint4 *>(out + (int64_t)n * per_image + g * 8)[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

}  // namespace

extern "C" int rgbd_scale_weights(const float *w, const float *scale, void *w_out, int32_t N, int32_t taps,
                                  int32_t cout_pad, int32_t cin_pad, int32_t Cin, void *stream) {
    RGBD_CHECK_ARG(w && scale && w_out, "null pointer");
    RGBD_CHECK_ARG(N > 0 && taps > 0 && cout_pad > 0 && cin_pad > 0 && (cin_pad & 7) == 0 && Cin > 0 && Cin <= cin_pad, "dims");
    RGBD_CHECK_ARG(((uintptr_t)w & 15) == 0 && ((uintptr_t)w_out & 15) == 0, "w / w_out must be 16-byte aligned");
    const int64_t per_image = (int64_t)taps * cout_pad * cin_pad;
    const int grid = rgbd_grid_for(per_image / 8 * N, 256);
    scale_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, scale, (__nv_bfloat16 *)w_out, per_image, cin_pad, Cin, N);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_scale_channels(const void *x, void *y, int32_t dtype, const float *scale, int32_t N, int64_t HW,
                                   int32_t C, int32_t x_cstride, int32_t x_coff, int32_t y_cstride, int32_t y_coff,
                                   void *stream) {
    RGBD_CHECK_ARG(x && y && scale, "null pointer");
    RGBD_CHECK_ARG(N > 0 && HW > 0 && C > 0, "dims");
    const int grid = rgbd_grid_for((int64_t)N * HW * C, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        scale_channels_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, scale, HW, (int64_t)N * HW, C,
                                                           x_cstride, x_coff, y_cstride, y_coff);
    else
        scale_channels_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, scale,
                                                                    HW, (int64_t)N * HW, C, x_cstride, x_coff,
                                                                    y_cstride, y_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_copy_view(const void *x, void *y, int32_t dtype, int64_t npix, int32_t C, int32_t x_cstride,
                              int32_t x_coff, int32_t y_cstride, int32_t y_coff, void *stream) {
    RGBD_CHECK_ARG(x && y, "null pointer");
    RGBD_CHECK_ARG(npix > 0 && C > 0, "dims");
    const int grid = rgbd_grid_for(npix * C, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        copy_view_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, npix, C, x_cstride, x_coff,
                                                      y_cstride, y_coff);
    else
        copy_view_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, npix, C,
                                                               x_cstride, x_coff, y_cstride, y_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_cast_view_bf16(const float *x, void *y, int64_t npix, int32_t C, int32_t x_cstride, int32_t x_coff,
                                   int32_t y_cstride, int32_t y_coff, void *stream) {
    RGBD_CHECK_ARG(x && y, "null pointer");
    RGBD_CHECK_ARG(npix > 0 && C > 0, "dims");
    cast_view_bf16_kernel<<<rgbd_grid_for(npix * C, 256), 256, 0, (cudaStream_t)stream>>>(
        x, (__nv_bfloat16 *)y, npix, C, x_cstride, x_coff, y_cstride, y_coff);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_zero(void *p, int64_t bytes, void *stream) {
    RGBD_CHECK_ARG(p && bytes >= 0, "pointer/size");
    cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        rgbd_set_error("rgbd_zero: %s", cudaGetErrorString(e));
        return RGBD_E_CUDA;
    }
    return RGBD_OK;
}

// SE_Block in two halves so that a caller can keep the table of partial sums across calls and refresh only the channels
// that were rewritten (the context buffer of the Bi-CEE chain: 1280 hyper-prior channels stay fixed over all 20 stages):
//   rgbd_se_partial : partial[n][chunk][c] for the channels [c0, c1) of the view
//   rgbd_se_gate    : scale[n][c] = sigmoid(W2 relu(W1 mean)) (+ 1) over the first C channels of the table
// The sums are per channel and in a fixed order, so refreshing a sub-range gives the same bits as recomputing everything.
extern "C" int rgbd_se_partial(const void *x, int32_t dtype, int32_t N, int32_t HW, int32_t cstride, int32_t coff,
                               int32_t c0, int32_t c1, int32_t nchunk, float *partial, int32_t pstride, void *stream) {
    RGBD_CHECK_ARG(x && partial, "null pointer");
    RGBD_CHECK_ARG(N > 0 && HW > 0 && nchunk > 0 && 0 <= c0 && c0 < c1 && c1 <= pstride, "dims");
    dim3 grid((unsigned)(N * nchunk), (unsigned)((c1 - c0 + 127) / 128));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        se_partial_kernel<float><<<grid, 128, 0, st>>>((const float *)x, HW, c0, c1, cstride, coff, nchunk, partial, pstride);
    else
        se_partial_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16 *)x, HW, c0, c1, cstride, coff, nchunk,
                                                                partial, pstride);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

// work: N * (C + Cr) floats (mean, hidden)
extern "C" int rgbd_se_gate(const float *partial, int32_t pstride, int32_t nchunk, int32_t N, int32_t HW, int32_t C,
                            const float *w1, const float *w2, int32_t Cr, int32_t plus_one, float *work, float *scale,
                            void *stream) {
    RGBD_CHECK_ARG(partial && w1 && w2 && work && scale, "null pointer");
    RGBD_CHECK_ARG(N > 0 && HW > 0 && C > 0 && Cr > 0 && nchunk > 0 && C <= pstride, "dims");
    cudaStream_t st = (cudaStream_t)stream;
    float *mean = work;
    float *hid = mean + (int64_t)N * C;
    se_mean_kernel<<<(unsigned)(((int64_t)N * C + 255) / 256), 256, 0, st>>>(partial, pstride, nchunk, HW, N, C, mean);
    RGBD_LAUNCH_CHECK();
    se_hidden_kernel<<<(unsigned)(((int64_t)N * Cr * 32 + 255) / 256), 256, 0, st>>>(mean, w1, N, C, Cr, hid);
    RGBD_LAUNCH_CHECK();
    se_gate_kernel<<<(unsigned)(((int64_t)N * C * 32 + 255) / 256), 256, 0, st>>>(hid, w2, N, C, Cr, plus_one, scale);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_se_scale(const void *x, int32_t dtype, int32_t N, int32_t HW, int32_t C, int32_t cstride,
                             int32_t coff, const float *w1, const float *w2, int32_t Cr, int32_t plus_one,
                             float *partial, int32_t nchunk, float *scale, void *stream) {
    RGBD_CHECK_ARG(C > 0, "dims");
    int rc = rgbd_se_partial(x, dtype, N, HW, cstride, coff, 0, C, nchunk, partial, C, stream);
    if (rc) return rc;
    // work buffers behind the partial sums: mean [N][C], hidden [N][Cr]
    return rgbd_se_gate(partial, C, nchunk, N, HW, C, w1, w2, Cr, plus_one, partial + (int64_t)N * nchunk * C, scale, stream);
}

extern "C" int rgbd_maxpool7s3(const void *x, void *y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                               void *stream) {
    RGBD_CHECK_ARG(x && y, "null pointer");
    RGBD_CHECK_ARG(N > 0 && H >= 7 && W >= 7 && C > 0, "dims (needs H, W >= 7)");
    const int Ho = (H - 7) / 3 + 1, Wo = (W - 7) / 3 + 1;
    const int grid = rgbd_grid_for((int64_t)N * Ho * Wo * C, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        maxpool7s3_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, N, H, W, C, Ho, Wo);
    else
        maxpool7s3_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, N, H,
                                                                W, C, Ho, Wo);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_nchw_to_nhwc(const float *x, void *y, int32_t dtype, int32_t N, int32_t C, int32_t H,
                                 int32_t W, int32_t y_cstride, int32_t y_coff, int32_t split3, void *stream) {
    RGBD_CHECK_ARG(x && y, "null pointer");
    RGBD_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "dims");
    const int grid = rgbd_grid_for((int64_t)N * C * H * W, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(x, (float *)y, N, C, H, W, y_cstride, y_coff, split3);
    else
        nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, (__nv_bfloat16 *)y, N, C, H, W, y_cstride,
                                                                  y_coff, split3);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_nhwc_to_nchw(const void *x, int32_t dtype, float *y, int32_t N, int32_t C, int32_t H,
                                 int32_t W, int32_t x_cstride, int32_t x_coff, int32_t clamp01, void *stream) {
    RGBD_CHECK_ARG(x && y, "null pointer");
    RGBD_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "dims");
    const int grid = rgbd_grid_for((int64_t)N * C * H * W, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == RGBD_DT_F32)
        nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float *)x, y, N, C, H, W, x_cstride, x_coff,
                                                         clamp01);
    else
        nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, y, N, C, H, W,
                                                                  x_cstride, x_coff, clamp01);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
