// Fused bottleneck block on tcgen05 tensor cores:
//
//     y = act( res + W3 * relu( W2 (*) relu( W1 * x + b1 ) + b2 ) + b3 )        1x1 -> 3x3 -> 1x1 + skip
//
// i.e. ResidualBottleneck (modules/layers/res_blk.py:7-27: res = x or the 1x1 skip conv of x) and
// AttentionBlock.ResidualUnit (CompressAI/compressai/layers/layers.py:178-197: res = x, act = ReLU), with the two
// C/2-channel intermediates kept ON CHIP: unfused, a block moves 960 channel-units per pixel through HBM
// (x, t1 out/in, t2 out/in, x again, y); fused it moves 384 + halo.
//
// One persistent CTA per SM walks output tiles of R rows x TW columns (positions laid out with row pitch P = TW + 2,
// R * P <= 256 = two M tiles).  Per tile, three GEMM phases share the tensor pipe:
//   P1  t1 = relu(W1 x + b1) on the tile PLUS its one-pixel halo ((R + 2) x P positions <= 384 = three M tiles).  The x halo
//       tile streams through a 2-stage TMA ring one 64-channel block at a time; accumulators D1 live in TMEM columns [0, 288).
//       The epilogue warps turn D1 into bf16 t1 in shared memory (two 128-byte-row planes: channels 0-63 and 64-95, written
//       with the 128-byte swizzle a TMA load would have produced), forcing positions outside the image to ZERO — the 3x3
//       conv of the reference pads t1 with zeros, not with relu(b1).
//   P2  t2 = relu(W2 (*) t1 + b2): implicit GEMM over the 9 taps; the A operand of tap (ky, kx) is the t1 plane read
//       through a UMMA descriptor whose start address is shifted by ky * P + kx rows (the halo trick of conv_halo.cu).
//       D2 in TMEM columns [288, 480).  t2 overwrites t1 in shared memory (every P2 MMA has completed by then).
//   P3  y = act(W3 t2 + b3 + res): four 48-channel output blocks, double buffered in the D2 region so that the epilogue of
//       one block (residual from global / L2, bf16 NHWC stores) overlaps the MMAs of the next, and P1 of the NEXT tile
//       (D1 columns are free again) overlaps the last epilogue.
// Weights stream from L2 through a 4-stage ring of 12 KB planes [96 rows x 64 K]; the K tails (channels 64-95) of two
// consecutive 3x3 taps share one plane (no zero padding is fetched).  Warp roles: 0 = x producer, 1 = weight producer,
// 2 / 3 = MMA issuers (one accumulator M tile each; warp 2 also issues the third halo M tile and owns TMEM), 4-11 = epilogue.
// K order is fixed (channel blocks, taps, 16-channel steps), so results do not depend on batch size or tile position.
#include "tc_common.cuh"
#include <new>
#include <cstdlib>

namespace {

constexpr int kRbThreads = 384;
constexpr int kXStages = 2, kWStages = 4;
constexpr int kCm = 96;                       // bottleneck width (C / 2)
constexpr int kWStageBytes = kCm * 128;       // 12 KB: 96 rows x 64 K (bf16)
constexpr int kNB3 = 48;                      // output-channel block of P3
constexpr uint32_t kRbTmemCols = 512;
constexpr int kD2Col = 3 * kCm;               // 288
constexpr int kMaxCout = 192;

struct RbParams {
    CUtensorMap xmap, w1map, w2map, w3map;
    const float *b1, *b2, *b3;
    const __nv_bfloat16 *res;
    __nv_bfloat16 *y;
    int32_t N, H, W, Cin, Cout;
    int32_t res_cstride, res_coff, y_cstride, y_coff;
    int32_t TW, R, P, tiles_x, tiles_y, total_tiles;
    int32_t kb1, nblk3, final_relu;
    int32_t x_bytes, x_tx, t_rows, t_bytes;
};

struct RbTile {
    int n, oy0, ox0;
};
__device__ __forceinline__ RbTile rb_tile(const RbParams &p, int tile) {
    RbTile t;
    const int per_img = p.tiles_x * p.tiles_y;
    t.n = tile / per_img;
    const int r = tile - t.n * per_img;
    t.oy0 = (r / p.tiles_x) * p.R;
    t.ox0 = (r % p.tiles_x) * p.TW;
    return t;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(kRbThreads, 1)
rb_fused_kernel(const __grid_constant__ RbParams p) {
    extern __shared__ uint8_t smem_raw[];
    using bf16 = __nv_bfloat16;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t x_base = base;
    const uint32_t t_base = x_base + (uint32_t)(kXStages * p.x_bytes);          // plane 0, then plane 1
    const uint32_t w_base = t_base + 2u * (uint32_t)p.t_bytes;
    const uint32_t bar_base = w_base + (uint32_t)(kWStages * kWStageBytes);
    auto x_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto x_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kXStages + s); };
    auto w_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kXStages + s); };
    auto w_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kXStages + kWStages + s); };
    const uint32_t misc = bar_base + 8u * (uint32_t)(2 * kXStages + 2 * kWStages);
    const uint32_t d1_full = misc, t1_ready = misc + 8, d2_full = misc + 16, t2_ready = misc + 24;
    auto d3_full = [&](int b) { return misc + 32u + 8u * (uint32_t)b; };
    auto d3_empty = [&](int b) { return misc + 48u + 8u * (uint32_t)b; };
    const uint32_t tmem_slot = misc + 64;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));
    float *bias_s = reinterpret_cast<float *>(smem_raw + (tmem_slot + 16u - raw));   // b1[96] | b2[96] | b3[Cout]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(x_full(s), 1);
            mbar_init(x_empty(s), 2);
        }
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), 2);
        }
        mbar_init(d1_full, 2);
        mbar_init(t1_ready, 8);
        mbar_init(d2_full, 2);
        mbar_init(t2_ready, 8);
        for (int b = 0; b < 2; ++b) {
            mbar_init(d3_full(b), 2);
            mbar_init(d3_empty(b), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, kRbTmemCols);
    if (warp == 3) {
        for (int i = lane; i < kCm; i += 32) {
            bias_s[i] = p.b1 ? p.b1[i] : 0.f;
            bias_s[kCm + i] = p.b2 ? p.b2[i] : 0.f;
        }
        for (int i = lane; i < p.Cout; i += 32) bias_s[2 * kCm + i] = p.b3 ? p.b3[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int first = (int)blockIdx.x, step = (int)gridDim.x;

    if (warp == 0) {
        // =========================== x producer: halo tile, one 64-channel block per stage ===========================
        if (lane == 0) {
            RingPos rx = {0, 0};
            griddep_wait();
            for (int tile = first; tile < p.total_tiles; tile += step) {
                const RbTile tc = rb_tile(p, tile);
                for (int kb = 0; kb < p.kb1; ++kb, rx.next(kXStages)) {
                    mbar_wait(x_empty(rx.s), rx.ph ^ 1u);
                    mbar_expect_tx(x_full(rx.s), (uint32_t)p.x_tx);
                    tma_load_4d(x_base + (uint32_t)(rx.s * p.x_bytes), &p.xmap, x_full(rx.s), kb * kBlockK, tc.ox0 - 1, tc.oy0 - 1,
                                tc.n);
                }
            }
            griddep_launch();
        }
    } else if (warp == 1) {
        // =========================== weight producer ===========================
        if (lane == 0) {
            RingPos rw = {0, 0};
            auto load = [&](const CUtensorMap *map, uint32_t bytes, int c0, int c1, int c2) {
                mbar_wait(w_empty(rw.s), rw.ph ^ 1u);
                mbar_expect_tx(w_full(rw.s), bytes);
                tma_load_3d(w_base + (uint32_t)(rw.s * kWStageBytes), map, w_full(rw.s), c0, c1, c2);
                rw.next(kWStages);
            };
            for (int tile = first; tile < p.total_tiles; tile += step) {
                for (int kb = 0; kb < p.kb1; ++kb) load(&p.w1map, kWStageBytes, kb * kBlockK, 0, 0);
                for (int s = 0; s < 14; ++s) load(&p.w2map, kWStageBytes, 0, 0, s);
                for (int blk = 0; blk < p.nblk3; ++blk)
                    for (int pl = 0; pl < 2; ++pl) load(&p.w3map, kNB3 * 128, pl * kBlockK, blk * kNB3, 0);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // =========================== MMA issuers ===========================
        const int wi = warp - 2;                 // accumulator M tile of P2 / P3; P1: warp 2 -> halo M tiles 0 and 2, warp 3 -> 1
        const uint64_t desc0 = make_smem_desc(0);
        const uint32_t idesc96 = make_idesc(128, kCm), idesc48 = make_idesc(128, kNB3);
        const uint32_t xb = x_base >> 4, wb = w_base >> 4, t0b = t_base >> 4, t1b = (t_base + (uint32_t)p.t_bytes) >> 4;
        const uint32_t mrow = (128u * 128u) >> 4;                    // one M tile of rows in descriptor units
        RingPos rx = {0, 0}, rw = {0, 0};
        uint32_t it = 0;
        for (int tile = first; tile < p.total_tiles; tile += step, ++it) {
            // ---------------- P1 ----------------
            for (int kb = 0; kb < p.kb1; ++kb) {
                mbar_wait(x_full(rx.s), rx.ph);
                mbar_wait(w_full(rw.s), rw.ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = xb + (uint32_t)((rx.s * p.x_bytes) >> 4), b_lo = wb + (uint32_t)((rw.s * kWStageBytes) >> 4);
                    for (int m = wi; m < 3; m += 2) {
                        const uint32_t dcol = tmem_base + (uint32_t)(m * kCm);
                        issue_mmas(dcol, desc0 + (uint64_t)(a_lo + (uint32_t)m * mrow), desc0 + (uint64_t)b_lo, idesc96, kb > 0 ? 1u : 0u, 4);
                    }
                    umma_commit(w_empty(rw.s));
                    umma_commit(x_empty(rx.s));
                }
                __syncwarp();
                rx.next(kXStages);
                rw.next(kWStages);
            }
            if (elect_one()) umma_commit(d1_full);
            __syncwarp();
            // the D2 columns double as the P3 output buffers of the previous tile: wait until its epilogue drained them
            mbar_wait(d3_empty(0), ((2u * it) & 1u) ^ 1u);
            mbar_wait(d3_empty(1), ((2u * it) & 1u) ^ 1u);
            mbar_wait(t1_ready, it & 1u);
            tc_fence_after();
            // ---------------- P2 ----------------
            {
                const uint32_t dcol = tmem_base + (uint32_t)(kD2Col + wi * kCm);
                const uint32_t jrow = (uint32_t)wi * mrow;
                for (int s = 0; s < 14; ++s) {
                    mbar_wait(w_full(rw.s), rw.ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b_lo = wb + (uint32_t)((rw.s * kWStageBytes) >> 4);
                        if (s < 9) {
                            const uint32_t sh = (uint32_t)(((s / 3) * p.P + (s % 3)) * 8);     // rows * 128 B >> 4
                            issue_mmas(dcol, desc0 + (uint64_t)(t0b + jrow + sh), desc0 + (uint64_t)b_lo, idesc96, s > 0 ? 1u : 0u, 4);
                        } else {
                            for (int h = 0; h < 2; ++h) {
                                const int t = 2 * (s - 9) + h;
                                if (t < 9) {
                                    const uint32_t sh = (uint32_t)(((t / 3) * p.P + (t % 3)) * 8);
                                    // plane 1 holds channels 64-95 (32 = two 16-channel steps); the weight plane holds tap t's tail
                                    // in its first 64 bytes when t is even, in the next 64 bytes when t is odd
                                    issue_mmas(dcol, desc0 + (uint64_t)(t1b + jrow + sh), desc0 + (uint64_t)(b_lo + (uint32_t)h * 4u), idesc96, 1u, 2);
                                }
                            }
                        }
                        umma_commit(w_empty(rw.s));
                    }
                    __syncwarp();
                    rw.next(kWStages);
                }
                if (elect_one()) umma_commit(d2_full);
                __syncwarp();
            }
            mbar_wait(t2_ready, it & 1u);
            tc_fence_after();
            // ---------------- P3 ----------------
            for (int blk = 0; blk < p.nblk3; ++blk) {
                const int b = blk & 1;
                const uint32_t use = 2u * it + (uint32_t)(blk >> 1);
                mbar_wait(d3_empty(b), (use & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t dcol = tmem_base + (uint32_t)(kD2Col + b * kCm + wi * kNB3);
                for (int pl = 0; pl < 2; ++pl) {
                    mbar_wait(w_full(rw.s), rw.ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b_lo = wb + (uint32_t)((rw.s * kWStageBytes) >> 4);
                        const uint32_t a_lo = (pl == 0 ? t0b : t1b) + (uint32_t)wi * mrow;
                        issue_mmas(dcol, desc0 + (uint64_t)a_lo, desc0 + (uint64_t)b_lo, idesc48, pl > 0 ? 1u : 0u, pl == 0 ? 4 : 2);
                        umma_commit(w_empty(rw.s));
                    }
                    __syncwarp();
                    rw.next(kWStages);
                }
                if (elect_one()) umma_commit(d3_full(b));
                __syncwarp();
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int ew = warp - 4;
        const int quad = warp & 3;               // TMEM lane quadrant of this warp
        const int half = ew >> 2;                // P1 / P2: column half [half * 48, +48); P3: M tile
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        const float *b1s = bias_s, *b2s = bias_s + kCm, *b3s = bias_s + 2 * kCm;
        const uint32_t plane[2] = {t_base, t_base + (uint32_t)p.t_bytes};
        griddep_wait();
        uint32_t it = 0;
        uint32_t r0[16];
        for (int tile = first; tile < p.total_tiles; tile += step, ++it) {
            const RbTile tc = rb_tile(p, tile);
            // ---------------- epilogue 1: D1 -> t1 (bf16, swizzled, zero outside the image) ----------------
            mbar_wait(d1_full, it & 1u);
            tc_fence_after();
            for (int m = 0; m < 3; ++m) {
                const int q = m * 128 + quad * 32 + lane;           // halo position
                const int hr = q / p.P, hc = q - hr * p.P;
                const int iy = tc.oy0 - 1 + hr, ix = tc.ox0 - 1 + hc;
                const bool inside = hr < p.R + 2 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const bool wr = q < p.t_rows;
                for (int c = 0; c < 3; ++c) {
                    const int ch = half * 48 + c * 16;
                    tmem_ld16_issue(tmem_base + tlane + (uint32_t)(m * kCm + ch), r0);
                    tmem_ld_wait(r0);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = inside ? fmaxf(__uint_as_float(r0[i]) + b1s[ch + i], 0.f) : 0.f;
                    if (wr) {
                        const uint32_t rowa = plane[ch >> 6] + (uint32_t)q * 128u;
                        const int k16 = (ch & 63) >> 3, sw = q & 7;       // 16-byte chunk index inside the 128-byte row
                        sts128(rowa + (uint32_t)(((k16) ^ sw) << 4), pack8(v));
                        sts128(rowa + (uint32_t)(((k16 + 1) ^ sw) << 4), pack8(v + 8));
                    }
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(t1_ready);
            // ---------------- epilogue 2: D2 -> t2 (over t1) ----------------
            mbar_wait(d2_full, it & 1u);
            tc_fence_after();
            for (int j = 0; j < 2; ++j) {
                const int q = j * 128 + quad * 32 + lane;
                for (int c = 0; c < 3; ++c) {
                    const int ch = half * 48 + c * 16;
                    tmem_ld16_issue(tmem_base + tlane + (uint32_t)(kD2Col + j * kCm + ch), r0);
                    tmem_ld_wait(r0);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(r0[i]) + b2s[ch + i], 0.f);
                    const uint32_t rowa = plane[ch >> 6] + (uint32_t)q * 128u;
                    const int k16 = (ch & 63) >> 3, sw = q & 7;
                    sts128(rowa + (uint32_t)(((k16) ^ sw) << 4), pack8(v));
                    sts128(rowa + (uint32_t)(((k16 + 1) ^ sw) << 4), pack8(v + 8));
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(t2_ready);
            // ---------------- epilogue 3: D3 blocks -> y = act(acc + b3 + res) ----------------
            {
                const int pos = half * 128 + quad * 32 + lane;
                const int ty = pos / p.P, tx = pos - ty * p.P;
                const int oy = tc.oy0 + ty, ox = tc.ox0 + tx;
                const bool valid = ty < p.R && tx < p.TW && oy < p.H && ox < p.W;
                const int64_t pix = ((int64_t)tc.n * p.H + oy) * p.W + ox;
                const bf16 *rp = p.res + pix * p.res_cstride + p.res_coff;
                bf16 *yp = p.y + pix * p.y_cstride + p.y_coff;
                for (int blk = 0; blk < p.nblk3; ++blk) {
                    const int b = blk & 1;
                    const uint32_t use = 2u * it + (uint32_t)(blk >> 1);
                    uint4 rr[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) rr[i] = make_uint4(0, 0, 0, 0);
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) rr[i] = reinterpret_cast<const uint4 *>(rp + blk * kNB3)[i];
                    }
                    mbar_wait(d3_full(b), use & 1u);
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        tmem_ld16_issue(tmem_base + tlane + (uint32_t)(kD2Col + b * kCm + half * kNB3 + c * 16), r0);
                        tmem_ld_wait(r0);
                        if (valid) {
                            float v[16], r[16];
                            unpack8(rr[2 * c], r);
                            unpack8(rr[2 * c + 1], r + 8);
                            const int co = blk * kNB3 + c * 16;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                v[i] = __uint_as_float(r0[i]) + b3s[co + i] + r[i];
                                if (p.final_relu) v[i] = fmaxf(v[i], 0.f);
                            }
                            reinterpret_cast<uint4 *>(yp + co)[0] = pack8(v);
                            reinterpret_cast<uint4 *>(yp + co)[1] = pack8(v + 8);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(d3_empty(b));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kRbTmemCols);
    }
}

}  // namespace

struct rgbd_rb_plan {
    RbParams p;
    dim3 grid;
    size_t smem;
};

extern "C" int rgbd_rb_plan_create(const rgbd_rb_desc *d, rgbd_rb_plan **out) {
    RGBD_CHECK_ARG(d && out, "null pointer");
    RGBD_CHECK_ARG(d->x && d->res && d->y && d->w1 && d->w2 && d->w3, "null tensor");
    RGBD_CHECK_ARG(d->Cmid == kCm, "bottleneck width must be 96");
    RGBD_CHECK_ARG(d->Cin > 0 && d->Cin % kBlockK == 0, "Cin must be a multiple of 64");
    RGBD_CHECK_ARG(d->Cout == kMaxCout, "Cout must be 192 (four 48-channel output blocks, two uses of each TMEM buffer per tile)");
    RGBD_CHECK_ARG(d->N > 0 && d->H >= 2 && d->W >= 2, "sizes");
    RGBD_CHECK_ARG(((d->x_cstride | d->x_coff | d->res_cstride | d->res_coff | d->y_cstride | d->y_coff) & 7) == 0,
                   "views must be 16-byte aligned (cstride, coff % 8)");
    RGBD_CHECK_ARG((((uintptr_t)d->x | (uintptr_t)d->res | (uintptr_t)d->y | (uintptr_t)d->w1 | (uintptr_t)d->w2 | (uintptr_t)d->w3) & 15) == 0,
                   "base pointers must be 16-byte aligned");
    RGBD_CHECK_ARG((int64_t)d->N * d->H * d->W < 2147483647LL, "too many pixels");
    rgbd_rb_plan *pl = new (std::nothrow) rgbd_rb_plan();
    if (!pl) {
        rgbd_set_error("rb: out of host memory");
        return RGBD_E_INVALID;
    }
    RbParams &p = pl->p;
    p.b1 = d->b1; p.b2 = d->b2; p.b3 = d->b3;
    p.res = reinterpret_cast<const __nv_bfloat16 *>(d->res);
    p.y = reinterpret_cast<__nv_bfloat16 *>(d->y);
    p.N = d->N; p.H = d->H; p.W = d->W; p.Cin = d->Cin; p.Cout = d->Cout;
    p.res_cstride = d->res_cstride; p.res_coff = d->res_coff; p.y_cstride = d->y_cstride; p.y_coff = d->y_coff;
    p.kb1 = d->Cin / kBlockK;
    p.nblk3 = d->Cout / kNB3;
    p.final_relu = d->final_relu;
    // tile geometry: R * P <= 256 output positions (two M tiles), (R + 2) * P + 2 <= 384 halo positions (three M tiles);
    // fewest tiles wins (every tile streams the same weights), ties go to the wider tile (longer contiguous rows)
    long best = -1;
    for (int TW = 2; TW <= d->W && TW + 2 <= 256; ++TW) {
        const int P = TW + 2;
        int R = 256 / P;
        while (R > 0 && (R + 2) * P + 2 > 384) --R;
        if (R > d->H) R = d->H;
        if (R < 1) continue;
        {
            const long xb = ((long)(R + 2) * P * 128 + 1023) / 1024 * 1024;
            long tr = (258 + 2 * P + 7) / 8 * 8;
            if (tr > 384) tr = 384;
            const long tb = (tr * 128 + 1023) / 1024 * 1024;
            if (kXStages * xb + 2 * tb + kWStages * kWStageBytes + 4096 > 227 * 1024) continue;
        }
        const long tiles = (long)((d->W + TW - 1) / TW) * (long)((d->H + R - 1) / R);
        if (best < 0 || tiles <= best) {
            best = tiles;
            p.TW = TW; p.R = R; p.P = P;
        }
    }
    if (best < 0) {
        rgbd_set_error("rb: no tile geometry for %d x %d", d->H, d->W);
        delete pl;
        return RGBD_E_INVALID;
    }
    p.tiles_x = (d->W + p.TW - 1) / p.TW;
    p.tiles_y = (d->H + p.R - 1) / p.R;
    const long total = (long)d->N * p.tiles_x * p.tiles_y;
    p.total_tiles = (int)total;
    p.x_tx = (p.R + 2) * p.P * 128;
    p.x_bytes = (p.x_tx + 1023) / 1024 * 1024;
    p.t_rows = (258 + 2 * p.P + 7) / 8 * 8;                   // rows P2 can address: 128 + 127 + (2 P + 2) + 1
    if (p.t_rows > 384) p.t_rows = 384;
    p.t_bytes = (p.t_rows * 128 + 1023) / 1024 * 1024;
    pl->smem = 1024 + (size_t)kXStages * p.x_bytes + 2 * (size_t)p.t_bytes + (size_t)kWStages * kWStageBytes +
               8 * (2 * kXStages + 2 * kWStages) + 64 + 16 + 4 * (2 * kCm + kMaxCout) + 64;
    // P1 reads three full M tiles (384 rows) from a stage: the last stage's over-read must stay inside the allocation
    RGBD_CHECK_ARG((size_t)(kXStages - 1) * p.x_bytes + 384 * 128 <= (size_t)kXStages * p.x_bytes + 2 * (size_t)p.t_bytes,
                   "internal: halo over-read leaves the allocation");
    if (pl->smem > 227 * 1024) {
        rgbd_set_error("rb: %zu bytes of shared memory needed", pl->smem);
        delete pl;
        return RGBD_E_INVALID;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    pl->grid = dim3((unsigned)(p.total_tiles < sms ? p.total_tiles : sms));

    int rc;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->x_cstride * 2, (cuuint64_t)d->W * d->x_cstride * 2,
                                 (cuuint64_t)d->H * d->W * d->x_cstride * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)p.P, (cuuint32_t)(p.R + 2), 1};
        rc = encode_map(&p.xmap, reinterpret_cast<const char *>(d->x) + (int64_t)d->x_coff * 2, 4, dims, strides, box);
    }
    if (!rc) {   // W1: bf16 [1][96][Cin]
        cuuint64_t dims[3] = {(cuuint64_t)d->Cin, (cuuint64_t)kCm, 1};
        cuuint64_t strides[2] = {(cuuint64_t)d->Cin * 2, (cuuint64_t)d->Cin * kCm * 2};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)kCm, 1};
        rc = encode_map(&p.w1map, d->w1, 3, dims, strides, box);
    }
    if (!rc) {   // W2: bf16 [14 planes][96][64]
        cuuint64_t dims[3] = {(cuuint64_t)kBlockK, (cuuint64_t)kCm, 14};
        cuuint64_t strides[2] = {(cuuint64_t)kBlockK * 2, (cuuint64_t)kBlockK * kCm * 2};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)kCm, 1};
        rc = encode_map(&p.w2map, d->w2, 3, dims, strides, box);
    }
    if (!rc) {   // W3: bf16 [1][Cout][128] (K padded 96 -> 128)
        cuuint64_t dims[3] = {128, (cuuint64_t)d->Cout, 1};
        cuuint64_t strides[2] = {128 * 2, (cuuint64_t)128 * d->Cout * 2};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)kNB3, 1};
        rc = encode_map(&p.w3map, d->w3, 3, dims, strides, box);
    }
    if (rc) {
        delete pl;
        return rc;
    }
    {
        static bool configured_on[64] = {};
        std::lock_guard<std::mutex> lock(g_init_mutex);
        const int di = (dev >= 0 && dev < 64) ? dev : 0;
        if (!configured_on[di]) {
            cudaFuncSetAttribute(rb_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            configured_on[di] = true;
        }
    }
    if (getenv("RGBD_TC_VERBOSE"))
        fprintf(stderr, "rb_fused: %dx%d Cin %d Cout %d | TW %d R %d P %d tiles %d x_bytes %d t_rows %d smem %zu\n", d->H, d->W, d->Cin,
                d->Cout, p.TW, p.R, p.P, p.total_tiles, p.x_bytes, p.t_rows, pl->smem);
    *out = pl;
    return RGBD_OK;
}

extern "C" int rgbd_rb_run(const rgbd_rb_plan *pl, void *stream) {
    RGBD_CHECK_ARG(pl != nullptr, "null plan");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = pl->grid;
    cfg.blockDim = dim3(kRbThreads);
    cfg.dynamicSmemBytes = pl->smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = getenv("RGBD_TC_NOPDL") == nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, rb_fused_kernel, pl->p);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" void rgbd_rb_plan_destroy(rgbd_rb_plan *pl) { delete pl; }
