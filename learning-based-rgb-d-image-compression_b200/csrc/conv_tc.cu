// tcgen05 tensor-core implicit-GEMM convolution (bf16 operands, fp32 accumulators in TMEM).
//
// GEMM view: D[pixel, co] = sum_{tap, ci} A[pixel @ tap, ci] * W[tap][co][ci].
//   * M tile = 128 output pixels = a TH x TW rectangle of one image's output lattice.
//   * A operand: for each (tap, 64-channel block) ONE 4-D TMA box {64 ch, TW, TH, 1} of the NHWC
//     input at the tap's spatial offset lands in shared memory as 128 rows x 128 B with the
//     128-byte swizzle = exactly the K-major SWIZZLE_128B UMMA operand layout.  Out-of-image
//     rows/columns are zero-filled by TMA (= the reference's zero padding), so there is no
//     im2col buffer and no boundary code.  Stride-2 convs address one of four parity
//     sub-lattices of the input through tensor maps with doubled strides.
//   * B operand: 3-D TMA box {64 ci, BN co, 1 tap} of the packed bf16 weights [tap][co][ci].
//   * One elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16) into a TMEM accumulator;
//     tcgen05.commit releases smem stages back to the TMA producer through mbarriers.
//   * Epilogue warps read the accumulator with tcgen05.ld (one pixel row per thread) and apply
//     bias / activation / residual / sigmoid-gate / bilinear-add before storing NHWC.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_id % 4).  The kernel is persistent: 2 CTAs per
// SM walk the tile list; residual / gate operands are prefetched before the accumulator is ready.
//
// The K reduction order (tap-major, then channel blocks, fixed MMA order) does not depend on the
// batch size or on the tile a pixel falls in: encoder and decoder reproduce the same bits.
#include "common.cuh"
#include <cuda.h>
#include <new>

namespace {

constexpr int kEpiWarps = 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;     // TMA warp + MMA warp + epilogue warps
constexpr int kTileM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 B = one swizzle row
constexpr int kABytes = kTileM * kBlockK * 2;     // 16 KB
// two persistent CTAs per SM (two TMA / MMA issuers and two epilogues in flight); the budget leaves
// room for one rANS decode block (~56 KB) on the same SM
constexpr int kSmemBudget = 84 * 1024;
constexpr int kMaxStages = 6;
constexpr uint32_t kTmemCols = 256;               // per CTA: two accumulator buffers when BN <= 128, else one
constexpr int kMaxChunks = 16;                    // 16-column chunks per tile (BN / 16)
constexpr int kPrefetchLinear = 4;                // residual chunks requested ahead of use
constexpr int kPrefetchGate = 4;                  // residual + gate chunks (twice the registers)

struct TcParams {
    CUtensorMap amap[4];
    CUtensorMap bmap;
    rgbd_conv_desc d;
    int32_t TW, TH, tiles_x, tiles_y;
    int32_t BN, kblocks, stages, n_ntiles, total_tiles, acc_bufs;
    int8_t tap_map[RGBD_MAX_TAPS], qy[RGBD_MAX_TAPS], qx[RGBD_MAX_TAPS];
    int8_t _pad[5];
};

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a mis-encoded tensor map would otherwise hang the GPU forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {  // ~2 s
            printf("rgbd conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1),
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups), [46,48) version = 1,
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format
// BF16 (bits 7, 10), K-major A and B, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == RGBD_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == RGBD_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
    return v;
}
__device__ __forceinline__ void bilerp_axis(int dst, int in_size, int out_size, int &i0, int &i1, float &l1) {
    const float scale = (float)in_size / (float)out_size;
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

// 16 consecutive channels of one pixel
template <typename T> __device__ __forceinline__ void load16(const T *p, bool vec, int nvalid, float *v);
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16 *p, bool vec, int nvalid, float *v) {
    if (vec && nvalid >= 16) {
        const uint4 a = reinterpret_cast<const uint4 *>(p)[0];
        const uint4 b = reinterpret_cast<const uint4 *>(p)[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
            v[2 * i] = __low2float(h);
            v[2 * i + 1] = __high2float(h);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = i < nvalid ? __bfloat162float(p[i]) : 0.f;
    }
}
template <typename T> __device__ __forceinline__ void store16(T *p, bool vec, int nvalid, const float *v);
template <> __device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16 *p, bool vec, int nvalid, const float *v) {
    if (vec && nvalid >= 16) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        reinterpret_cast<uint4 *>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4 *>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
    }
}
template <> __device__ __forceinline__ void store16<float>(float *p, bool vec, int nvalid, const float *v) {
    if (vec && nvalid >= 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nvalid) p[i] = v[i];
    }
}

__device__ __forceinline__ void unpack8(const uint4 &q, float *v) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
        v[2 * i] = __low2float(h);
        v[2 * i + 1] = __high2float(h);
    }
}

struct TileCoord {
    int n, oy0, ox0, co0, bn;
};
__device__ __forceinline__ TileCoord tile_coord(const TcParams &p, int tile) {
    // n-tile fastest: CTAs that run side by side share the same A tile through L2
    TileCoord t;
    const int nt = tile % p.n_ntiles;
    const int mt = tile / p.n_ntiles;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    t.n = mt / tiles_per_img;
    const int trem = mt - t.n * tiles_per_img;
    t.oy0 = (trem / p.tiles_x) * p.TH;
    t.ox0 = (trem % p.tiles_x) * p.TW;
    t.co0 = nt * p.BN;
    t.bn = min(p.BN, p.d.cout_pad - t.co0);   // multiple of 16
    return t;
}

// Persistent kernel: two CTAs per SM loop over output tiles (tile = blockIdx.x + i * gridDim.x).
// The smem ring and its mbarrier phases run continuously across tiles; when the N tile is <= 128
// columns the accumulator is double buffered in the CTA's 256 TMEM columns so that the epilogue of
// tile i overlaps the MMAs of tile i + 1 (wider tiles rely on the sibling CTA for overlap).
template <typename TOut, int kEpi>
__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ TcParams p) {
    constexpr bool kGate = kEpi == RGBD_EPI_GATE;
    constexpr int kPrefetch = kGate ? kPrefetchGate : kPrefetchLinear;
    extern __shared__ uint8_t smem_raw[];
    using bf16 = __nv_bfloat16;
    const rgbd_conv_desc &d = p.d;
    // 1024-byte aligned carve-up: [stage: A 16 KB | B BN*128 B] ... | barriers | tmem ptr
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.BN * 128u;
    const uint32_t stage_bytes = kABytes + b_bytes;          // multiple of 1024 (BN % 16 == 0 -> 2 KB steps)
    const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
    auto tmem_full_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + a); };
    auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * kMaxStages + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full_bar(a), 1);
            mbar_init(tmem_empty_bar(a), kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int num_it = d.ntaps * p.kblocks;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            uint32_t g = 0;   // running k-iteration count across tiles
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord(p, tile);
                for (int it = 0; it < num_it; ++it, ++g) {
                    const int t = it / p.kblocks, kb = it - t * p.kblocks;
                    const int s = (int)(g % (uint32_t)p.stages);
                    const uint32_t ph = (g / (uint32_t)p.stages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    const uint32_t sa = base + (uint32_t)s * stage_bytes;
                    mbar_expect_tx(full_bar(s), (uint32_t)kABytes + b_bytes);   // TMA delivers full boxes
                    tma_load_4d(sa, &p.amap[p.tap_map[t]], full_bar(s), kb * kBlockK, tc.ox0 + p.qx[t],
                                tc.oy0 + p.qy[t], tc.n);
                    tma_load_3d(sa + kABytes, &p.bmap, full_bar(s), kb * kBlockK, tc.co0, d.wtap[t]);
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t g = 0;
            int li = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
                const TileCoord tc = tile_coord(p, tile);
                const int acc = li % p.acc_bufs;
                const uint32_t use = (uint32_t)(li / p.acc_bufs);
                mbar_wait(tmem_empty_bar(acc), (use & 1u) ^ 1u);   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t idesc = make_idesc(kTileM, tc.bn);
                const uint32_t dcol = tmem_base + (uint32_t)acc * 128u;
                for (int it = 0; it < num_it; ++it, ++g) {
                    const int kb = it % p.kblocks;
                    const int s = (int)(g % (uint32_t)p.stages);
                    const uint32_t ph = (g / (uint32_t)p.stages) & 1u;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t sa = base + (uint32_t)s * stage_bytes;
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + kABytes);
                    // only the 16-channel groups that hold real input channels (tail block may be short)
                    const int kleft = d.Cin - kb * kBlockK;
                    const int nk = kleft >= kBlockK ? 4 : (kleft + 15) >> 4;
                    for (int k = 0; k < nk; ++k)
                        umma_bf16(dcol, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                  (it > 0 || k > 0) ? 1u : 0u);
                    umma_commit(empty_bar(s));   // frees the smem stage once these MMAs have read it
                }
                umma_commit(tmem_full_bar(acc));
            }
        }
    } else {
        // ============ epilogue: TMEM -> registers -> NHWC global (8 warps) ============
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;          // pixel index inside the tile
        const int th = row / p.TW, tw = row - th * p.TW;
        TOut *y = reinterpret_cast<TOut *>(d.y);
        TOut *y2 = reinterpret_cast<TOut *>(d.y2);
        const bf16 *res = reinterpret_cast<const bf16 *>(d.res);
        const bf16 *mul = reinterpret_cast<const bf16 *>(d.mul);
        constexpr int kVecOut = 16 / (int)sizeof(TOut);   // elements per 16 B
        const bool y_vec = ((d.y_cstride | d.y_coff) % kVecOut) == 0;
        const bool y2_vec = ((d.y2_cstride | d.y2_coff) % kVecOut) == 0;
        const bool res_vec = ((d.res_cstride | d.res_coff) & 7) == 0;
        const bool mul_vec = ((d.mul_cstride | d.mul_coff) & 7) == 0;
        const bool res_direct = res != nullptr && kEpi != RGBD_EPI_BILERP;
        int li = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++li) {
            const TileCoord tc = tile_coord(p, tile);
            const int acc = li % p.acc_bufs;
            const uint32_t use = (uint32_t)(li / p.acc_bufs);
            const int sy = tc.oy0 + th, sx = tc.ox0 + tw;    // site in the Hs x Ws output lattice
            const bool valid = sy < d.Hs && sx < d.Ws;
            const int oy = sy * d.o_step + d.o_off_y, ox = sx * d.o_step + d.o_off_x;
            const int64_t opix = ((int64_t)tc.n * d.Ho + oy) * d.Wo + ox;
            const int c_begin = 0;
            const int c_end = tc.bn >> 4;
            // residual / gate operands of this thread's pixel are requested BEFORE the accumulator is
            // ready, so their HBM latency hides behind the MMAs
            uint4 pr[2 * kPrefetch];
            uint4 pm[kGate ? 2 * kPrefetch : 1];
            auto prefetch = [&](int k, int slot) {
                const int c = c_begin + k;
                const int co = tc.co0 + c * 16;
                const bool on = valid && c < c_end && d.Cout - co >= 16;
                pr[2 * slot] = pr[2 * slot + 1] = make_uint4(0, 0, 0, 0);
                if (on && res_direct && res_vec) {
                    const uint4 *q = reinterpret_cast<const uint4 *>(res + opix * d.res_cstride + d.res_coff + co);
                    pr[2 * slot] = q[0];
                    pr[2 * slot + 1] = q[1];
                }
                if (kGate) {
                    pm[2 * slot] = pm[2 * slot + 1] = make_uint4(0, 0, 0, 0);
                    if (on && mul_vec) {
                        const uint4 *q = reinterpret_cast<const uint4 *>(mul + opix * d.mul_cstride + d.mul_coff + co);
                        pm[2 * slot] = q[0];
                        pm[2 * slot + 1] = q[1];
                    }
                }
            };
#pragma unroll
            for (int k = 0; k < kPrefetch; ++k) prefetch(k, k);
            int by0 = 0, by1 = 0, bx0 = 0, bx1 = 0;
            float ly = 0.f, lx = 0.f;
            if (kEpi == RGBD_EPI_BILERP && valid) {
                bilerp_axis(oy, d.res_H, d.Ho, by0, by1, ly);
                bilerp_axis(ox, d.res_W, d.Wo, bx0, bx1, lx);
            }
            mbar_wait(tmem_full_bar(acc), use & 1u);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kMaxChunks; ++k) {
                const int c = c_begin + k;
                if (c < c_end) {            // warp-uniform
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 128 + c * 16), v);
                const int co = tc.co0 + c * 16;
                const int nvalid = d.Cout - co;
                if (valid && nvalid > 0) {
                const bool full16 = nvalid >= 16;
                if (d.bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += (i < nvalid) ? __ldg(d.bias + co + i) : 0.f;
                }
                float r[16];
                if (res_direct) {
                    if (res_vec && full16) {
                        unpack8(pr[2 * (k % kPrefetch)], r);
                        unpack8(pr[2 * (k % kPrefetch) + 1], r + 8);
                    } else {
                        load16<bf16>(res + opix * d.res_cstride + d.res_coff + co, false, nvalid, r);
                    }
                }
                if (kGate) {
                    float m[16];
                    if (mul_vec && full16) {
                        unpack8(pm[2 * (k % kPrefetch)], m);
                        unpack8(pm[2 * (k % kPrefetch) + 1], m + 8);
                    } else {
                        load16<bf16>(mul + opix * d.mul_cstride + d.mul_coff + co, false, nvalid, m);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = m[i] * __fdividef(1.0f, 1.0f + __expf(-v[i]));
                    if (res_direct) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += r[i];
                    }
                } else if (kEpi == RGBD_EPI_LINEAR) {
                    if (res_direct) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += r[i];
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = act_fn(v[i], d.act);
                } else {  // RGBD_EPI_BILERP
                    const int64_t rb = (int64_t)tc.n * d.res_H * d.res_W;
                    const int cc = d.res_coff + co;
                    float a00[16], a01[16], a10[16], a11[16];
                    load16<bf16>(res + (rb + (int64_t)by0 * d.res_W + bx0) * d.res_cstride + cc, res_vec, nvalid, a00);
                    load16<bf16>(res + (rb + (int64_t)by0 * d.res_W + bx1) * d.res_cstride + cc, res_vec, nvalid, a01);
                    load16<bf16>(res + (rb + (int64_t)by1 * d.res_W + bx0) * d.res_cstride + cc, res_vec, nvalid, a10);
                    load16<bf16>(res + (rb + (int64_t)by1 * d.res_W + bx1) * d.res_cstride + cc, res_vec, nvalid, a11);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float up = (1.f - ly) * ((1.f - lx) * a00[i] + lx * a01[i]) +
                                         ly * ((1.f - lx) * a10[i] + lx * a11[i]);
                        v[i] = act_fn(v[i] + up, d.act);
                    }
                }
                store16<TOut>(y + opix * d.y_cstride + d.y_coff + co, y_vec, nvalid, v);
                if (y2) store16<TOut>(y2 + opix * d.y2_cstride + d.y2_coff + co, y2_vec, nvalid, v);
                }
                if (k + kPrefetch < kMaxChunks) prefetch(k + kPrefetch, k % kPrefetch);   // refill the slot
                }
            }
            // all of this warp's tcgen05.ld have completed (wait::ld inside tmem_ld16): hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_map(CUtensorMap *m, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
               const cuuint32_t *box) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        rgbd_set_error("conv_tc: cuTensorMapEncodeTiled unavailable");
        return RGBD_E_CUDA;
    }
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides_bytes, box, ones,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rgbd_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu %llu box %u %u %u", (int)r,
                       rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                       box[0], box[1], box[2]);
        return RGBD_E_CUDA;
    }
    return RGBD_OK;
}

inline int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

}  // namespace

struct rgbd_conv_tc_plan {
    TcParams p;
    dim3 grid;
    size_t smem;
    int out_f32;
    int epi;
};

extern "C" int rgbd_conv_validate(const rgbd_conv_desc *d);

extern "C" int rgbd_conv_tc_plan_create(const rgbd_conv_desc *d, int32_t cin_pad, rgbd_conv_tc_plan **out) {
    int rc = rgbd_conv_validate(d);
    if (rc) return rc;
    RGBD_CHECK_ARG(out != nullptr, "null out");
    RGBD_CHECK_ARG(d->x_dtype == RGBD_DT_BF16, "tensor-core path needs bf16 activations");
    RGBD_CHECK_ARG(d->in_scale == nullptr, "in_scale is not supported on the tensor-core path (pre-scale the input)");
    RGBD_CHECK_ARG((d->x_cstride & 7) == 0 && (d->x_coff & 7) == 0, "x view must be 16-byte aligned (cstride, coff % 8)");
    RGBD_CHECK_ARG(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->w & 15) == 0, "x / w base must be 16-byte aligned");
    RGBD_CHECK_ARG(cin_pad >= d->Cin && (cin_pad % kBlockK) == 0, "cin_pad must be a multiple of 64 >= Cin");
    RGBD_CHECK_ARG(d->i_step == 1 || d->i_step == 2, "i_step must be 1 or 2");

    rgbd_conv_tc_plan *pl = new (std::nothrow) rgbd_conv_tc_plan();
    if (!pl) {
        rgbd_set_error("conv_tc: out of host memory");
        return RGBD_E_INVALID;
    }
    TcParams &p = pl->p;
    p.d = *d;
    // tile rectangle: the TW x TH = 128 shape that wastes the fewest lattice sites
    const int cand[][2] = {{8, 16}, {16, 8}, {32, 4}, {4, 32}, {64, 2}, {2, 64}, {128, 1}, {1, 128}};
    long best = -1;
    for (auto &c : cand) {
        const long tx = (d->Ws + c[0] - 1) / c[0], ty = (d->Hs + c[1] - 1) / c[1];
        const long area = tx * ty;
        if (best < 0 || area < best) {
            best = area;
            p.TW = c[0];
            p.TH = c[1];
            p.tiles_x = (int)tx;
            p.tiles_y = (int)ty;
        }
    }
    p.kblocks = (d->Cin + kBlockK - 1) / kBlockK;
    // N tile: whole cout_pad if <= 256, else the fewest equal-ish tiles (multiples of 16)
    const int ntiles = (d->cout_pad + 255) / 256;
    p.BN = ((d->cout_pad + ntiles - 1) / ntiles + 15) / 16 * 16;
    const int stage_bytes = kABytes + p.BN * 128;
    p.stages = kSmemBudget / stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    if (p.stages < 2) p.stages = 2;
    p.acc_bufs = p.BN <= 128 ? 2 : 1;
    p.n_ntiles = (d->cout_pad + p.BN - 1) / p.BN;
    const long total = (long)d->N * p.tiles_x * p.tiles_y * p.n_ntiles;
    RGBD_CHECK_ARG(total < 2147483647L, "too many tiles");
    p.total_tiles = (int)total;
    pl->smem = (size_t)p.stages * stage_bytes + 1024 /*align*/ + 8 * (2 * kMaxStages + 6);
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0)
            num_sms = 148;
    }
    pl->grid = dim3((unsigned)(p.total_tiles < 2 * num_sms ? p.total_tiles : 2 * num_sms));
    pl->out_f32 = d->y_dtype == RGBD_DT_F32;
    pl->epi = d->epi;

    // A tensor maps: one per input parity class (i_step == 2) or a single one
    const int st = d->i_step;
    const int nmap = st * st;
    const char *xb = reinterpret_cast<const char *>(d->x);
    for (int ry = 0; ry < st; ++ry)
        for (int rx = 0; rx < st; ++rx) {
            const int Hsub = (d->H - ry + st - 1) / st, Wsub = (d->W - rx + st - 1) / st;
            if (Hsub <= 0 || Wsub <= 0) continue;
            cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)Wsub, (cuuint64_t)Hsub, (cuuint64_t)d->N};
            cuuint64_t strides[3] = {(cuuint64_t)st * d->x_cstride * 2, (cuuint64_t)st * d->W * d->x_cstride * 2,
                                     (cuuint64_t)d->H * d->W * d->x_cstride * 2};
            cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
            void *basep = (void *)(xb + ((int64_t)(ry * d->W + rx) * d->x_cstride + d->x_coff) * 2);
            rc = encode_map(&p.amap[ry * st + rx], basep, 4, dims, strides, box);
            if (rc) {
                delete pl;
                return rc;
            }
        }
    for (int i = nmap; i < 4; ++i) p.amap[i] = p.amap[0];
    for (int t = 0; t < d->ntaps; ++t) {
        if (st == 1) {
            p.tap_map[t] = 0;
            p.qy[t] = d->dy[t];
            p.qx[t] = d->dx[t];
        } else {
            const int qy = floordiv2(d->dy[t]), qx = floordiv2(d->dx[t]);
            const int ry = d->dy[t] - 2 * qy, rx = d->dx[t] - 2 * qx;
            p.tap_map[t] = (int8_t)(ry * 2 + rx);
            p.qy[t] = (int8_t)qy;
            p.qx[t] = (int8_t)qx;
        }
    }
    // B tensor map over the packed weights [taps_total][cout_pad][cin_pad] (taps_total >= max wtap + 1)
    int max_tap = 0;
    for (int t = 0; t < d->ntaps; ++t) max_tap = d->wtap[t] > max_tap ? d->wtap[t] : max_tap;
    {
        cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)d->cout_pad, (cuuint64_t)(max_tap + 1)};
        cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * d->cout_pad * 2};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)p.BN, 1};
        rc = encode_map(&p.bmap, const_cast<void *>(d->w), 3, dims, strides, box);
        if (rc) {
            delete pl;
            return rc;
        }
    }
    static bool configured = false;
    if (!configured) {
        const int cap = 100 * 1024;
        cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_tc_kernel<__nv_bfloat16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_tc_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_tc_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_tc_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        configured = true;
    }
    *out = pl;
    return RGBD_OK;
}

extern "C" int rgbd_conv_tc_run(const rgbd_conv_tc_plan *pl, void *stream) {
    RGBD_CHECK_ARG(pl != nullptr, "null plan");
    cudaStream_t st = (cudaStream_t)stream;
#define RGBD_TC_LAUNCH(T, E) conv_tc_kernel<T, E><<<pl->grid, kThreads, pl->smem, st>>>(pl->p)
    if (pl->out_f32) {
        if (pl->epi == RGBD_EPI_GATE) RGBD_TC_LAUNCH(float, 1);
        else if (pl->epi == RGBD_EPI_BILERP) RGBD_TC_LAUNCH(float, 2);
        else RGBD_TC_LAUNCH(float, 0);
    } else {
        if (pl->epi == RGBD_EPI_GATE) RGBD_TC_LAUNCH(__nv_bfloat16, 1);
        else if (pl->epi == RGBD_EPI_BILERP) RGBD_TC_LAUNCH(__nv_bfloat16, 2);
        else RGBD_TC_LAUNCH(__nv_bfloat16, 0);
    }
#undef RGBD_TC_LAUNCH
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" void rgbd_conv_tc_plan_destroy(rgbd_conv_tc_plan *pl) { delete pl; }
