// fp32-accumulate implicit-GEMM convolution on the CUDA cores (NHWC pixels x packed taps).
//
// This is the fp32 parity path of the conv family (and the path for the 3/1-channel
// image-side layers in bf16 mode): GEMM view  D[pixel, co] = sum_{tap, ci} A[pixel@tap, ci] *
// W[tap][ci][co], tile BM pixels x BN output channels per CTA, BK = 16 input channels per
// k-step, register-prefetched double buffering, TM x 4 register micro-tile per thread.
// The reduction order over (tap, ci) is fixed and independent of the batch size and of the
// tile a pixel falls in, so encoder and decoder reproduce the same bits.
//
// Replaces nn.Conv2d / nn.ConvTranspose2d of the reference (modules/layers/conv.py:7-34).
#include "common.cuh"

namespace {

constexpr int BK = 16;
constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ float4 load4(const T *p, bool vec, int nvalid);
template <> __device__ __forceinline__ float4 load4<float>(const float *p, bool vec, int nvalid) {
    if (vec && nvalid >= 4) return *reinterpret_cast<const float4 *>(p);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid > 0) v.x = p[0];
    if (nvalid > 1) v.y = p[1];
    if (nvalid > 2) v.z = p[2];
    if (nvalid > 3) v.w = p[3];
    return v;
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16 *p, bool vec, int nvalid) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec && nvalid >= 4) {
        const uint2 raw = *reinterpret_cast<const uint2 *>(p);
        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&raw.x);
        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162 *>(&raw.y);
        v.x = __low2float(lo); v.y = __high2float(lo); v.z = __low2float(hi); v.w = __high2float(hi);
        return v;
    }
    if (nvalid > 0) v.x = __bfloat162float(p[0]);
    if (nvalid > 1) v.y = __bfloat162float(p[1]);
    if (nvalid > 2) v.z = __bfloat162float(p[2]);
    if (nvalid > 3) v.w = __bfloat162float(p[3]);
    return v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == RGBD_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == RGBD_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
    if (act == RGBD_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    return v;
}

// F.interpolate(mode="bilinear", align_corners=False) source index (ATen
// area_pixel_compute_source_index): src = scale * (dst + 0.5) - 0.5, clamped at 0
__device__ __forceinline__ void bilerp_axis(int dst, int in_size, int out_size, int &i0, int &i1, float &l1) {
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

template <typename TIn, typename TOut, int BM, int BN, int TM>
__global__ void __launch_bounds__(kThreads)
conv_simt_kernel(const rgbd_conv_desc d) {
    constexpr int TN = 4;
    static_assert((BM / TM) * (BN / TN) == kThreads && TM % 4 == 0, "tile/thread mismatch");
    constexpr int A_VECS = BM * BK / 4;            // float4 per A tile
    constexpr int A_ITERS = (A_VECS + kThreads - 1) / kThreads;
    constexpr int B_VECS = BK * BN / 4;
    constexpr int B_ITERS = (B_VECS + kThreads - 1) / kThreads;
    constexpr int AS_LD = BM + 4;

    __shared__ __align__(16) float As[BK][AS_LD];
    __shared__ __align__(16) float Bs[BK][BN];

    const TIn *__restrict__ x = reinterpret_cast<const TIn *>(d.x);
    const float *__restrict__ w = reinterpret_cast<const float *>(d.w);

    const int tid = threadIdx.x;
    const int64_t M = (int64_t)d.N * d.Hs * d.Ws;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int co0 = blockIdx.y * BN;

    // per-thread A-load assignment: pixel rows and a 4-channel group inside the k-chunk
    int a_n[A_ITERS], a_oy[A_ITERS], a_ox[A_ITERS];
    bool a_ok[A_ITERS];
    const int a_kq = (tid & 3) * 4;
#pragma unroll
    for (int r = 0; r < A_ITERS; ++r) {
        const int ml = (tid >> 2) + r * (kThreads / 4);
        const int64_t m = m0 + ml;
        a_ok[r] = (ml < BM) && (m < M);
        const int64_t mm = a_ok[r] ? m : 0;
        a_n[r] = (int)(mm / ((int64_t)d.Hs * d.Ws));
        const int rem = (int)(mm % ((int64_t)d.Hs * d.Ws));
        a_oy[r] = rem / d.Ws;
        a_ox[r] = rem % d.Ws;
    }
    const bool x_vec = ((d.x_cstride | d.x_coff) & 3) == 0;

    const int cin_chunks = (d.Cin + BK - 1) / BK;
    const int n_kc = d.ntaps * cin_chunks;

    float4 a_reg[A_ITERS];
    float4 b_reg[B_ITERS];

    auto load_chunk = [&](int kc) {
        const int t = kc / cin_chunks;
        const int ci0 = (kc - t * cin_chunks) * BK;
        const int dy = d.dy[t], dx = d.dx[t];
        const int ci = ci0 + a_kq;
        const int nvalid = d.Cin - ci;
#pragma unroll
        for (int r = 0; r < A_ITERS; ++r) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int iy = a_oy[r] * d.i_step + dy;
            const int ix = a_ox[r] * d.i_step + dx;
            if (a_ok[r] && nvalid > 0 && iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) {
                const int64_t pix = ((int64_t)a_n[r] * d.H + iy) * d.W + ix;
                v = load4<TIn>(x + pix * d.x_cstride + d.x_coff + ci, x_vec, nvalid);
                if (d.in_scale) {
                    const float *sc = d.in_scale + (int64_t)a_n[r] * d.Cin + ci;
                    v.x *= sc[0];
                    if (nvalid > 1) v.y *= sc[1];
                    if (nvalid > 2) v.z *= sc[2];
                    if (nvalid > 3) v.w *= sc[3];
                }
            }
            a_reg[r] = v;
        }
        const float *wt = w + (int64_t)d.wtap[t] * d.Cin * d.cout_pad;
#pragma unroll
        for (int r = 0; r < B_ITERS; ++r) {
            const int e = tid + r * kThreads;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < B_VECS) {
                const int k = e / (BN / 4);
                const int nq = (e % (BN / 4)) * 4;
                const int cik = ci0 + k;
                if (cik < d.Cin && co0 + nq < d.cout_pad)
                    v = *reinterpret_cast<const float4 *>(wt + (int64_t)cik * d.cout_pad + co0 + nq);
            }
            b_reg[r] = v;
        }
    };
    auto store_chunk = [&]() {
#pragma unroll
        for (int r = 0; r < A_ITERS; ++r) {
            const int ml = (tid >> 2) + r * (kThreads / 4);
            if (ml < BM) {
                As[a_kq + 0][ml] = a_reg[r].x;
                As[a_kq + 1][ml] = a_reg[r].y;
                As[a_kq + 2][ml] = a_reg[r].z;
                As[a_kq + 3][ml] = a_reg[r].w;
            }
        }
#pragma unroll
        for (int r = 0; r < B_ITERS; ++r) {
            const int e = tid + r * kThreads;
            if (e < B_VECS) {
                const int k = e / (BN / 4);
                const int nq = (e % (BN / 4)) * 4;
                *reinterpret_cast<float4 *>(&Bs[k][nq]) = b_reg[r];
            }
        }
    };

    const int tx = tid % (BN / TN);  // along output channels
    const int ty = tid / (BN / TN);  // along pixels
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_chunk(0);
    for (int kc = 0; kc < n_kc; ++kc) {
        store_chunk();
        __syncthreads();
        if (kc + 1 < n_kc) load_chunk(kc + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&As[k][ty * TM + i]);
                a[i] = v.x;
                a[i + 1] = v.y;
                a[i + 2] = v.z;
                a[i + 3] = v.w;
            }
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * TN]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue ----
    TOut *__restrict__ y = reinterpret_cast<TOut *>(d.y);
    TOut *__restrict__ y2 = reinterpret_cast<TOut *>(d.y2);
    const TIn *__restrict__ res = reinterpret_cast<const TIn *>(d.res);
    const TIn *__restrict__ mul = reinterpret_cast<const TIn *>(d.mul);
    const int co = co0 + tx * TN;
    if (co >= d.Cout) return;
    float bias[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) bias[j] = (d.bias && co + j < d.Cout) ? d.bias[co + j] : 0.f;

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty * TM + i;
        if (m >= M) continue;
        const int n = (int)(m / ((int64_t)d.Hs * d.Ws));
        const int rem = (int)(m % ((int64_t)d.Hs * d.Ws));
        const int oy = (rem / d.Ws) * d.o_step + d.o_off_y;
        const int ox = (rem % d.Ws) * d.o_step + d.o_off_x;
        const int64_t opix = ((int64_t)n * d.Ho + oy) * d.Wo + ox;
        int by0 = 0, by1 = 0, bx0 = 0, bx1 = 0;
        float ly = 0.f, lx = 0.f;
        if (d.epi == RGBD_EPI_BILERP) {
            bilerp_axis(oy, d.res_H, d.Ho, by0, by1, ly);
            bilerp_axis(ox, d.res_W, d.Wo, bx0, bx1, lx);
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (co + j >= d.Cout) break;
            float v = acc[i][j] + bias[j];
            if (d.epi == RGBD_EPI_LINEAR) {
                if (res) v += ElemIO<TIn>::ld(res + opix * d.res_cstride + d.res_coff + co + j);
                v = apply_act(v, d.act);
            } else if (d.epi == RGBD_EPI_GATE) {
                const float gate = 1.0f / (1.0f + expf(-v));
                v = ElemIO<TIn>::ld(mul + opix * d.mul_cstride + d.mul_coff + co + j) * gate;
                if (res) v += ElemIO<TIn>::ld(res + opix * d.res_cstride + d.res_coff + co + j);
            } else {  // RGBD_EPI_BILERP
                const int64_t rb = (int64_t)n * d.res_H * d.res_W;
                const int c = d.res_coff + co + j;
                const float v00 = ElemIO<TIn>::ld(res + (rb + (int64_t)by0 * d.res_W + bx0) * d.res_cstride + c);
                const float v01 = ElemIO<TIn>::ld(res + (rb + (int64_t)by0 * d.res_W + bx1) * d.res_cstride + c);
                const float v10 = ElemIO<TIn>::ld(res + (rb + (int64_t)by1 * d.res_W + bx0) * d.res_cstride + c);
                const float v11 = ElemIO<TIn>::ld(res + (rb + (int64_t)by1 * d.res_W + bx1) * d.res_cstride + c);
                const float up = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
                v = apply_act(v + up, d.act);
            }
            ElemIO<TOut>::st(y + opix * d.y_cstride + d.y_coff + co + j, v);
            if (y2) ElemIO<TOut>::st(y2 + opix * d.y2_cstride + d.y2_coff + co + j, v);
        }
    }
}

template <typename TIn, typename TOut>
int launch_simt(const rgbd_conv_desc *d, cudaStream_t st) {
    const int64_t M = (int64_t)d->N * d->Hs * d->Ws;
    if (d->Cout <= 16) {
        constexpr int BM = 256, BN = 16, TM = 4;
        dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->Cout + BN - 1) / BN));
        conv_simt_kernel<TIn, TOut, BM, BN, TM><<<grid, kThreads, 0, st>>>(*d);
    } else {
        constexpr int BM = 128, BN = 64, TM = 8;
        dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->Cout + BN - 1) / BN));
        conv_simt_kernel<TIn, TOut, BM, BN, TM><<<grid, kThreads, 0, st>>>(*d);
    }
    return 0;
}

}  // namespace

extern "C" int rgbd_conv_validate(const rgbd_conv_desc *d) {
    RGBD_CHECK_ARG(d != nullptr, "null descriptor");
    RGBD_CHECK_ARG(d->x && d->y && d->w, "null tensor pointer");
    RGBD_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "dims");
    RGBD_CHECK_ARG(d->Hs > 0 && d->Ws > 0 && d->o_step > 0 && d->i_step > 0, "lattice");
    RGBD_CHECK_ARG((d->Hs - 1) * d->o_step + d->o_off_y < d->Ho && (d->Ws - 1) * d->o_step + d->o_off_x < d->Wo,
                   "output lattice exceeds Ho x Wo");
    RGBD_CHECK_ARG(d->ntaps > 0 && d->ntaps <= RGBD_MAX_TAPS, "ntaps");
    RGBD_CHECK_ARG(d->cout_pad >= d->Cout && (d->cout_pad & 15) == 0, "cout_pad must be a multiple of 16 >= Cout");
    RGBD_CHECK_ARG(d->x_coff + d->Cin <= d->x_cstride && d->y_coff + d->Cout <= d->y_cstride, "channel view");
    RGBD_CHECK_ARG(d->epi >= 0 && d->epi <= 3, "epi");
    RGBD_CHECK_ARG(d->epi != RGBD_EPI_SHUFFLE2 ||
                       (d->o_step == 2 && d->o_off_y == 0 && d->o_off_x == 0 && d->cout_pad == 16 && d->Cout <= 4 &&
                        (d->Hs - 1) * 2 + 1 < d->Ho && (d->Ws - 1) * 2 + 1 < d->Wo && !d->res && !d->mul && !d->y2),
                   "SHUFFLE2 epilogue: o_step 2, no offsets, cout_pad 16, Cout <= 4, no res / mul / y2");
    RGBD_CHECK_ARG(d->epi != RGBD_EPI_GATE || d->mul, "GATE epilogue needs mul");
    RGBD_CHECK_ARG(d->epi != RGBD_EPI_BILERP || (d->res && d->res_H > 0 && d->res_W > 0), "BILERP needs res map");
    RGBD_CHECK_ARG((d->x_dtype | 1) == 1 && (d->y_dtype | 1) == 1, "dtype");
    return RGBD_OK;
}

extern "C" int rgbd_conv_simt(const rgbd_conv_desc *d, void *stream) {
    int rc = rgbd_conv_validate(d);
    if (rc) return rc;
    RGBD_CHECK_ARG(d->epi != RGBD_EPI_SHUFFLE2, "SHUFFLE2 epilogue exists on the tensor-core path only");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->x_dtype == RGBD_DT_F32 && d->y_dtype == RGBD_DT_F32) launch_simt<float, float>(d, st);
    else if (d->x_dtype == RGBD_DT_BF16 && d->y_dtype == RGBD_DT_BF16) launch_simt<__nv_bfloat16, __nv_bfloat16>(d, st);
    else if (d->x_dtype == RGBD_DT_BF16 && d->y_dtype == RGBD_DT_F32) launch_simt<__nv_bfloat16, float>(d, st);
    else launch_simt<float, __nv_bfloat16>(d, st);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
