// Quality metrics and image export next to the codec path: what testing/tester_united.py:92-123 does to
// every reconstruction — utils/metrics.py:8-14 (PSNR and pytorch_msssim.ms_ssim on the GPU) and the 8-bit / 16-bit
// quantisation behind saveImg / cv2.imwrite.  All reductions are two-stage with a fixed order (no atomics): the same
// inputs give the same bits on every run.
#include "common.cuh"

namespace {

constexpr int kWin = 11;          // pytorch_msssim default window (win_size = 11, win_sigma = 1.5)
constexpr int kTileW = 32, kTileH = 8;

__device__ __forceinline__ float clamp01f(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// ---- squared error: partial[n][blk] = sum over the block's grid-stride elements of (a - b)^2 (double) ----
__global__ void __launch_bounds__(256)
sqerr_partial_kernel(const float *__restrict__ a, const float *__restrict__ b, int64_t n_per_image, int clamp01,
                     double *__restrict__ partial) {
    const int n = blockIdx.y;
    const float *pa = a + (int64_t)n * n_per_image, *pb = b + (int64_t)n * n_per_image;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_image; i += (int64_t)gridDim.x * blockDim.x) {
        float x = pa[i], y = pb[i];
        if (clamp01) {
            x = clamp01f(x);
            y = clamp01f(y);
        }
        const float d = x - y;
        acc += (double)d * (double)d;
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(int64_t)n * gridDim.x + blockIdx.x] = red[0];
}

// out[i * n_out + k] = sum_j partial[(i * n_out + k) * n_part + j], in index order
__global__ void reduce_rows_kernel(const double *__restrict__ partial, int n_rows, int n_part, double *__restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    double acc = 0.0;
    for (int j = 0; j < n_part; ++j) acc += partial[(int64_t)r * n_part + j];
    out[r] = acc;
}

// ---- one SSIM level (pytorch_msssim _ssim): separable 11-tap Gaussian, valid window ----
// grid: (tiles_x, tiles_y, planes); partial[plane][tile][2] = (sum ssim_map, sum cs_map) over the tile
__global__ void __launch_bounds__(kTileW * kTileH)
ssim_level_kernel(const float *__restrict__ X, const float *__restrict__ Y, int H, int W, float C1, float C2, int clamp01,
                  double *__restrict__ partial) {
    __shared__ float win[kWin];
    __shared__ float sx[kTileH + kWin - 1][kTileW + kWin - 1], sy[kTileH + kWin - 1][kTileW + kWin - 1];
    __shared__ float hq[5][kTileH + kWin - 1][kTileW];
    __shared__ double red[2][kTileW * kTileH];
    const int tid = threadIdx.y * kTileW + threadIdx.x;
    if (tid < kWin) {
        float s = 0.f, g[kWin];
        for (int i = 0; i < kWin; ++i) {
            const float c = (float)(i - kWin / 2);
            g[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f));
            s += g[i];
        }
        win[tid] = g[tid] / s;
    }
    const int Ho = H - kWin + 1, Wo = W - kWin + 1;
    const int plane = blockIdx.z;
    const float *px = X + (int64_t)plane * H * W, *py = Y + (int64_t)plane * H * W;
    const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH;
    for (int i = tid; i < (kTileH + kWin - 1) * (kTileW + kWin - 1); i += kTileW * kTileH) {
        const int r = i / (kTileW + kWin - 1), c = i - r * (kTileW + kWin - 1);
        const int iy = oy0 + r, ix = ox0 + c;
        float vx = 0.f, vy = 0.f;
        if (iy < H && ix < W) {
            vx = px[(int64_t)iy * W + ix];
            vy = py[(int64_t)iy * W + ix];
            if (clamp01) {
                vx = clamp01f(vx);
                vy = clamp01f(vy);
            }
        }
        sx[r][c] = vx;
        sy[r][c] = vy;
    }
    __syncthreads();
    // horizontal pass: the filter runs over W first in pytorch_msssim only when applied in dim order (H, then W); a
    // separable product of the same 1-D window, so the order changes rounding only
    for (int i = tid; i < (kTileH + kWin - 1) * kTileW; i += kTileW * kTileH) {
        const int r = i / kTileW, c = i - r * kTileW;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win[k], x = sx[r][c + k], y = sy[r][c + k];
            a0 += w * x;
            a1 += w * y;
            a2 += w * x * x;
            a3 += w * y * y;
            a4 += w * x * y;
        }
        hq[0][r][c] = a0; hq[1][r][c] = a1; hq[2][r][c] = a2; hq[3][r][c] = a3; hq[4][r][c] = a4;
    }
    __syncthreads();
    double ssim = 0.0, cs = 0.0;
    const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
    if (oy < Ho && ox < Wo) {
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float w = win[k];
            m1 += w * hq[0][threadIdx.y + k][threadIdx.x];
            m2 += w * hq[1][threadIdx.y + k][threadIdx.x];
            e11 += w * hq[2][threadIdx.y + k][threadIdx.x];
            e22 += w * hq[3][threadIdx.y + k][threadIdx.x];
            e12 += w * hq[4][threadIdx.y + k][threadIdx.x];
        }
        const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
        const float s11 = e11 - m11, s22 = e22 - m22, s12 = e12 - m12;
        const float csv = (2.f * s12 + C2) / (s11 + s22 + C2);
        const float sv = ((2.f * m12 + C1) / (m11 + m22 + C1)) * csv;
        ssim = (double)sv;
        cs = (double)csv;
    }
    red[0][tid] = ssim;
    red[1][tid] = cs;
    __syncthreads();
    for (int s = kTileW * kTileH / 2; s > 0; s >>= 1) {
        if (tid < s) {
            red[0][tid] += red[0][tid + s];
            red[1][tid] += red[1][tid + s];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int64_t tile = (int64_t)blockIdx.y * gridDim.x + blockIdx.x, ntile = (int64_t)gridDim.x * gridDim.y;
        partial[((int64_t)plane * ntile + tile) * 2] = red[0][0];
        partial[((int64_t)plane * ntile + tile) * 2 + 1] = red[1][0];
    }
}

// sums[plane][2] = ordered sum of the tile partials
__global__ void ssim_reduce_kernel(const double *__restrict__ partial, int planes, int ntile, double *__restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= planes * 2) return;
    const int plane = i >> 1, which = i & 1;
    double acc = 0.0;
    for (int t = 0; t < ntile; ++t) acc += partial[((int64_t)plane * ntile + t) * 2 + which];
    sums[i] = acc;
}

// F.avg_pool2d(kernel_size = 2, padding = (H % 2, W % 2)), count_include_pad = True (pytorch_msssim between levels)
__global__ void avgpool2_kernel(const float *__restrict__ x, float *__restrict__ y, int planes, int H, int W, int Ho, int Wo,
                                int clamp01) {
    const int64_t total = (int64_t)planes * Ho * Wo;
    const int ph = H & 1, pw = W & 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
        const int64_t pl = i / ((int64_t)Wo * Ho);
        float acc = 0.f;
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                const int iy = 2 * oy - ph + dy, ix = 2 * ox - pw + dx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                    float v = x[(pl * H + iy) * W + ix];
                    acc += clamp01 ? clamp01f(v) : v;
                }
            }
        y[i] = acc * 0.25f;
    }
}

// ToPILImage of a float tensor (torchvision: pic.mul(255).byte(), i.e. truncation) after clamp_(0, 1): NCHW fp32 -> NHWC u8
__global__ void quantize_u8_kernel(const float *__restrict__ x, uint8_t *__restrict__ y, int N, int C, int H, int W, int h, int w) {
    const int64_t total = (int64_t)N * h * w * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int ix = (int)((i / C) % w), iy = (int)((i / ((int64_t)C * w)) % h);
        const int64_t n = i / ((int64_t)C * w * h);
        const float v = clamp01f(x[((n * C + c) * H + iy) * (int64_t)W + ix]) * 255.f;
        y[i] = (uint8_t)(int)v;
    }
}

// (x * scale).astype("uint16") of testing/tester_united.py:101-105: truncation toward zero, wrap modulo 2^16
__global__ void quantize_u16_kernel(const float *__restrict__ x, uint16_t *__restrict__ y, int N, int H, int W, int h, int w, float scale) {
    const int64_t total = (int64_t)N * h * w;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ix = (int)(i % w), iy = (int)((i / w) % h);
        const int64_t n = i / ((int64_t)w * h);
        const float v = x[(n * H + iy) * (int64_t)W + ix] * scale;
        y[i] = (uint16_t)((long long)v & 0xFFFF);
    }
}

}  // namespace

extern "C" int rgbd_sq_error_sums(const float *a, const float *b, int32_t N, int64_t n_per_image, int32_t clamp01,
                                  double *work, int32_t n_part, double *out, void *stream) {
    RGBD_CHECK_ARG(a && b && work && out, "null pointer");
    RGBD_CHECK_ARG(N > 0 && n_per_image > 0 && n_part > 0 && n_part <= 4096, "sizes");
    cudaStream_t st = (cudaStream_t)stream;
    sqerr_partial_kernel<<<dim3((unsigned)n_part, (unsigned)N), 256, 0, st>>>(a, b, n_per_image, clamp01, work);
    RGBD_LAUNCH_CHECK();
    reduce_rows_kernel<<<(N + 127) / 128, 128, 0, st>>>(work, N, n_part, out);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_ssim_level(const float *x, const float *y, int32_t planes, int32_t H, int32_t W, float data_range,
                               int32_t clamp01, double *work, double *sums, void *stream) {
    RGBD_CHECK_ARG(x && y && work && sums, "null pointer");
    RGBD_CHECK_ARG(planes > 0 && H >= kWin && W >= kWin, "every side must be at least the 11-tap window");
    const int Ho = H - kWin + 1, Wo = W - kWin + 1;
    const dim3 grid((unsigned)((Wo + kTileW - 1) / kTileW), (unsigned)((Ho + kTileH - 1) / kTileH), (unsigned)planes);
    RGBD_CHECK_ARG(planes <= 65535, "too many planes");
    const float C1 = (0.01f * data_range) * (0.01f * data_range), C2 = (0.03f * data_range) * (0.03f * data_range);
    cudaStream_t st = (cudaStream_t)stream;
    ssim_level_kernel<<<grid, dim3(kTileW, kTileH), 0, st>>>(x, y, H, W, C1, C2, clamp01, work);
    RGBD_LAUNCH_CHECK();
    ssim_reduce_kernel<<<(planes * 2 + 127) / 128, 128, 0, st>>>(work, planes, (int)(grid.x * grid.y), sums);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int64_t rgbd_ssim_work_elems(int32_t planes, int32_t H, int32_t W) {
    if (H < kWin || W < kWin) return 0;
    const int64_t tx = (W - kWin + 1 + kTileW - 1) / kTileW, ty = (H - kWin + 1 + kTileH - 1) / kTileH;
    return 2 * (int64_t)planes * tx * ty;
}

extern "C" int rgbd_avgpool2(const float *x, float *y, int32_t planes, int32_t H, int32_t W, int32_t clamp01, void *stream) {
    RGBD_CHECK_ARG(x && y && planes > 0 && H > 0 && W > 0, "arguments");
    const int Ho = (H + 2 * (H & 1) - 2) / 2 + 1, Wo = (W + 2 * (W & 1) - 2) / 2 + 1;
    avgpool2_kernel<<<rgbd_grid_for((int64_t)planes * Ho * Wo, 256), 256, 0, (cudaStream_t)stream>>>(x, y, planes, H, W, Ho, Wo, clamp01);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_quantize_u8(const float *x, uint8_t *y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t crop_h,
                                int32_t crop_w, void *stream) {
    RGBD_CHECK_ARG(x && y && N > 0 && C > 0 && crop_h > 0 && crop_w > 0 && crop_h <= H && crop_w <= W, "arguments");
    quantize_u8_kernel<<<rgbd_grid_for((int64_t)N * C * crop_h * crop_w, 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, C, H, W, crop_h, crop_w);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_quantize_u16(const float *x, uint16_t *y, int32_t N, int32_t H, int32_t W, int32_t crop_h, int32_t crop_w,
                                 float scale, void *stream) {
    RGBD_CHECK_ARG(x && y && N > 0 && crop_h > 0 && crop_w > 0 && crop_h <= H && crop_w <= W, "arguments");
    quantize_u16_kernel<<<rgbd_grid_for((int64_t)N * crop_h * crop_w, 256), 256, 0, (cudaStream_t)stream>>>(x, y, N, H, W, crop_h, crop_w, scale);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
