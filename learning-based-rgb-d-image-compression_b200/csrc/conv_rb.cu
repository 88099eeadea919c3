// Fused bottleneck block on tcgen05 tensor cores:
//
//     y = act( res + W3 * relu( W2 (*) relu( W1 * x + b1 ) + b2 ) + b3 )        1x1 -> 3x3 -> 1x1 + skip
//
// i.e. ResidualBottleneck (modules/layers/res_blk.py:7-27: res = x or the 1x1 skip conv of x) and
// AttentionBlock.ResidualUnit (CompressAI/compressai/layers/layers.py:178-197: res = x, act = ReLU), with the two
// C/2-channel intermediates kept ON CHIP: unfused, a block moves 960 channel-units per pixel through HBM
// (x, t1 out/in, t2 out/in, x again, y); fused it moves 384 + halo.
//
// Geometry.  Positions are laid out with a row pitch of P = 32, so one M tile of 128 positions is rpm = 4 whole rows, one
// epilogue warp (32 TMEM lanes) is one row, and position <-> (row, column) is a shift and a mask.  A tile produces 2 M
// tiles of output (R = 8 rows x TW = 30 useful columns); its one-pixel halo — 10 rows x 32 columns starting one pixel up
// and left — fits 3 M tiles (322 of their 384 rows are ever read).  One persistent CTA per SM takes tiles from a global
// counter (dynamic scheduler, as in conv_halo.cu).  Per tile, three GEMM phases share the tensor pipe:
//   P1  t1 = relu(W1 x + b1) on the halo (3 M tiles: 0 and 1 together, one issuer warp each, then 2 — so that the epilogues
//       of the first two run under the MMAs of the third).  x arrives as 16 KB pieces {64 channels, P columns, rpm rows} =
//       one (channel block, M tile) each, through a 4-deep TMA ring; the NEXT tile's pieces are prefetched into L2 while
//       this tile computes (cp.async.bulk.prefetch.tensor), so the ring refills at L2 latency, not HBM latency.
//       Accumulators D1[m] in TMEM columns [0, 288).  The epilogue warps turn D1 into bf16 t1 in shared memory (two
//       128-byte-row planes: channels 0-63 and 64-95, written with the 128-byte swizzle a TMA load would have produced),
//       forcing positions outside the image to ZERO — the 3x3 conv of the reference pads t1 with zeros, not with relu(b1).
//   P2  t2 = relu(W2 (*) t1 + b2): implicit GEMM over the 9 taps; the A operand of tap (ky, kx) is the t1 plane read
//       through a UMMA descriptor whose start address is shifted by ky * P + kx rows (the halo trick of conv_halo.cu).
//       D2[j] in TMEM columns [288, 480).  t2 overwrites t1 in shared memory (every P2 MMA has completed by then).
//   P3  y = act(W3 t2 + b3 + res): four 48-channel output blocks per M tile, double buffered in the D2[j] region so that
//       the epilogue of one block overlaps the MMAs of the next, and P1 of the NEXT tile (D1 columns are free again) overlaps
//       the last epilogues.  The epilogue is bound by the load / store unit (a thread owns a pixel, so every global access
//       of a warp lands in its own 128-byte line): the residual comes in with 32-byte loads, and the result leaves through a
//       3 KB staging row per warp and ONE TMA store {48 channels, 30 columns, 1 row} per warp and block, which also clips
//       the image border.
// Weights stream from L2 through a 4-stage ring of 12 KB planes [96 rows x 64 K]; the K tails (channels 64-95) of two
// consecutive 3x3 taps share one plane (no zero padding is fetched); a W3 stage holds both K planes of one output block.
// Warp roles: 0 = x producer, 1 = weight producer, 2 / 3 = MMA issuers (output M tile j = warp - 2; halo M tiles 0 and 2
// / 1; warp 2 owns TMEM), 4-11 = epilogue (TMEM lane quadrant = warp % 4; warps 4-7 serve M tile 0, warps 8-11 M tile 1,
// both halves of the third halo tile).  Every hand-over is its own mbarrier (per halo M tile, per output M tile, per
// output-block buffer), the epilogues run one TMEM load ahead of the arithmetic, and no integer division is left in them.
// K order is fixed (channel blocks, taps, 16-channel steps), so results do not depend on batch size or tile position.
#include "tc_common.cuh"
#include <new>
#include <cstdlib>

namespace {

constexpr int kRbThreads = 384;
constexpr int kXStages = 4, kWStages = 4;
constexpr int kXStageBytes = 128 * 128;       // 16 KB: one M tile of positions x 64 channels (bf16)
constexpr int kCm = 96;                       // bottleneck width (C / 2)
constexpr int kWStageBytes = kCm * 128;       // 12 KB: 96 rows x 64 K (bf16)
constexpr int kNB3 = 48;                      // output-channel block of P3
constexpr int kP = 32, kLP = 5, kRpm = 4, kR = 8, kTW = 30;   // tile geometry (see above)
constexpr int kTRows = 328;                   // rows of a t plane: P2 reads rows [0, 128 + 127 + 2 P + 2]
constexpr int kTBytes = kTRows * 128;         // 41 KB (a multiple of 1024: plane 1 keeps the swizzle phase)
constexpr int kYStageBytes = 32 * kNB3 * 2;   // per epilogue warp: 32 positions x 48 channels (bf16), packed
constexpr uint32_t kRbTmemCols = 512;
constexpr int kD2Col = 3 * kCm;               // 288
constexpr int kMaxCout = 192;
constexpr int kTileQ = 4;                     // depth of the CTA's tile queue (dynamic scheduler)
constexpr int kRbBarriers = 2 * kXStages + 2 * kWStages + 3 + 3 + 2 + 2 + 4 + 4 + 2 * kTileQ;

struct RbParams {
    CUtensorMap xmap, w1map, w2map, w3map, ymap;
    const float *b1, *b2, *b3;
    const __nv_bfloat16 *res;
    __nv_bfloat16 *y;
    int32_t N, H, W, Cin, Cout;
    int32_t res_cstride, res_coff, y_cstride, y_coff;
    int32_t tiles_x, tiles_y, total_tiles;
    int32_t kb1, nblk3, final_relu, prefetch, static_tiles, wide_io;   // wide_io: res rows are 32-byte aligned
    int32_t *sched;   // tile counter (device memory, zero between launches)
    long long *dbg;   // optional cycle counters of CTA 0 (development builds: RGBD_TIMING_PROBES + RGBD_TC_TRACE), else NULL
};

struct RbTile {
    int n, oy0, ox0;
};
__device__ __forceinline__ RbTile rb_tile(const RbParams &p, int tile) {
    RbTile t;
    const int per_img = p.tiles_x * p.tiles_y;
    t.n = tile / per_img;
    const int r = tile - t.n * per_img;
    const int tyi = r / p.tiles_x;
    t.oy0 = tyi * kR;
    t.ox0 = (r - tyi * p.tiles_x) * kTW;
    return t;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"(c3)
                 : "memory");
}

__global__ void __launch_bounds__(kRbThreads, 1)
rb_fused_kernel(const __grid_constant__ RbParams p) {
    extern __shared__ uint8_t smem_raw[];
    using bf16 = __nv_bfloat16;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t x_base = base;
    const uint32_t t_base = x_base + (uint32_t)(kXStages * kXStageBytes);          // plane 0, then plane 1
    const uint32_t w_base = t_base + 2u * (uint32_t)kTBytes;
    const uint32_t bar_base = w_base + (uint32_t)(kWStages * kWStageBytes);
    auto x_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto x_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kXStages + s); };
    auto w_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kXStages + s); };
    auto w_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kXStages + kWStages + s); };
    const uint32_t misc = bar_base + 8u * (uint32_t)(2 * kXStages + 2 * kWStages);
    auto d1_full = [&](int m) { return misc + 8u * (uint32_t)m; };                  // P1 MMAs of halo M tile m complete
    auto t1_ready = [&](int m) { return misc + 24u + 8u * (uint32_t)m; };           // t1 rows of halo M tile m are in shared memory
    auto d2_full = [&](int j) { return misc + 48u + 8u * (uint32_t)j; };
    auto t2_ready = [&](int j) { return misc + 64u + 8u * (uint32_t)j; };
    auto d3_full = [&](int j, int b) { return misc + 80u + 8u * (uint32_t)(2 * j + b); };
    auto d3_empty = [&](int j, int b) { return misc + 112u + 8u * (uint32_t)(2 * j + b); };
    auto tq_full = [&](int q) { return misc + 144u + 8u * (uint32_t)q; };
    auto tq_empty = [&](int q) { return misc + 144u + 8u * (uint32_t)(kTileQ + q); };
    const uint32_t tmem_slot = misc + 144u + 16u * (uint32_t)kTileQ;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));
    volatile int *tq_s = reinterpret_cast<volatile int *>(smem_raw + (tmem_slot + 16u - raw));   // kTileQ tile indices
    float *bias_s = reinterpret_cast<float *>(smem_raw + (tmem_slot + 32u - raw));   // b1[96] | b2[96] | b3[Cout]
    const uint32_t ystage_base = (tmem_slot + 32u + 4u * (uint32_t)(2 * kCm + kMaxCout) + 127u) & ~127u;   // 8 x kYStageBytes

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(x_full(s), 1);
            mbar_init(x_empty(s), 2);      // a piece feeds one issuer warp, but BOTH pass every stage (see the P1 loop)
        }
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), 2);      // every weight stage feeds both issuer warps
        }
        for (int m = 0; m < 3; ++m) {
            mbar_init(d1_full(m), 1);
            mbar_init(t1_ready(m), m < 2 ? 4 : 8);
        }
        for (int j = 0; j < 2; ++j) {
            mbar_init(d2_full(j), 1);
            mbar_init(t2_ready(j), 4);
            for (int b = 0; b < 2; ++b) {
                mbar_init(d3_full(j, b), 1);
                mbar_init(d3_empty(j, b), 4);
            }
        }
        // tile queue: filled by the x producer, read by the weight producer, the two issuer warps and the epilogue warps
        for (int q = 0; q < kTileQ; ++q) {
            mbar_init(tq_full(q), 1);
            mbar_init(tq_empty(q), 1 + 2 + 8);
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
    // Dynamic tile scheduler (as in conv_halo.cu): tiles are handed out in index order from a global counter, so a CTA that
    // starts late — its SM was busy with rANS coder blocks of another pipeline slot — does not hold the grid back; the others
    // take its share.  The x producer draws the tiles (one ahead, for the L2 prefetch) and publishes them through a small
    // queue; every other role reads the k-th entry.  Which CTA computes a tile has no effect on the result.
    auto next_tile = [&](int k) -> int {      // consumers: all lanes of the calling warp
        const int q = k % kTileQ;
        mbar_wait(tq_full(q), (uint32_t)((k / kTileQ) & 1));
        const int t = tq_s[q];
        __syncwarp();
        if (lane == 0) mbar_arrive(tq_empty(q));
        return t;
    };

    if (warp == 0) {
        // =========================== x producer: one (channel block, halo M tile) piece per stage ===========================
        if (lane == 0) {
            RingPos rx = {0, 0};
            const bool tr = p.dbg != nullptr && blockIdx.x == 0;
            long long wx = 0;
            griddep_wait();
            int drawn = 0;
            auto draw = [&]() -> int {
                if (p.static_tiles) {      // A / B switch of development runs (RGBD_RB_STATIC=1): tile = blockIdx.x + k * gridDim.x
                    const int t = (int)blockIdx.x + drawn++ * (int)gridDim.x;
                    return t < p.total_tiles ? t : p.total_tiles;
                }
                int t = atomicAdd(p.sched, 1);
                if (t >= p.total_tiles) {
                    // every CTA draws exactly one value past the end; the last of them re-arms the counter for the next
                    // launch of this plan (stream order keeps that launch behind this one)
                    if (t == p.total_tiles + (int)gridDim.x - 1) atomicExch(p.sched, 0);
                    t = p.total_tiles;
                }
                return t;
            };
            int ahead = draw();
            for (int k = 0;; ++k) {
                const int tile = ahead;
                const int q = k % kTileQ;
                mbar_wait(tq_empty(q), (uint32_t)((k / kTileQ) & 1) ^ 1u);
                tq_s[q] = tile;
                mbar_arrive(tq_full(q));
                if (tile >= p.total_tiles) break;
                ahead = draw();
                const RbTile tc = rb_tile(p, tile);
                if (p.prefetch && ahead < p.total_tiles) {
                    // the next tile's pieces start their way HBM -> L2 now: the ring can hold only 4 of its 9+ pieces ahead of time
                    const RbTile tn = rb_tile(p, ahead);
                    for (int kb = 0; kb < p.kb1; ++kb)
                        for (int m = 0; m < 3; ++m) tma_prefetch_4d(&p.xmap, kb * kBlockK, tn.ox0 - 1, tn.oy0 - 1 + m * kRpm, tn.n);
                }
                // halo M tiles 0 and 1 together (channel block by channel block: one issuer warp each, sharing the W1 stages),
                // then M tile 2 alone: D1[0] and D1[1] complete two thirds of the way into P1, so their epilogues run under
                // the MMAs of M tile 2
                for (int part = 0; part < 2; ++part) {
                    for (int kb = 0; kb < p.kb1; ++kb) {
                        for (int m = part ? 2 : 0; m < (part ? 3 : 2); ++m, rx.next(kXStages)) {
                            mbar_wait_t(x_empty(rx.s), rx.ph ^ 1u, tr, wx);
                            mbar_expect_tx(x_full(rx.s), (uint32_t)kXStageBytes);
                            tma_load_4d(x_base + (uint32_t)(rx.s * kXStageBytes), &p.xmap, x_full(rx.s), kb * kBlockK, tc.ox0 - 1,
                                        tc.oy0 - 1 + m * kRpm, tc.n);
                        }
                    }
                }
            }
            griddep_launch();
            if (tr) p.dbg[0] = wx;
        }
    } else if (warp == 1) {
        // =========================== weight producer ===========================
        {
            RingPos rw = {0, 0};
            const bool tr = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
            long long ww = 0;
            for (int k = 0;; ++k) {
                if (next_tile(k) >= p.total_tiles) break;      // whole warp; lane 0 issues the loads
                if (lane != 0) continue;
                for (int mk = 0; mk < 2 * p.kb1; ++mk, rw.next(kWStages)) {      // W1 twice: for halo M tiles 0 + 1, then for M tile 2
                    mbar_wait_t(w_empty(rw.s), rw.ph ^ 1u, tr, ww);
                    mbar_expect_tx(w_full(rw.s), (uint32_t)kWStageBytes);
                    tma_load_3d(w_base + (uint32_t)(rw.s * kWStageBytes), &p.w1map, w_full(rw.s), (mk % p.kb1) * kBlockK, 0, 0);
                }
                for (int s = 0; s < 14; ++s, rw.next(kWStages)) {
                    mbar_wait_t(w_empty(rw.s), rw.ph ^ 1u, tr, ww);
                    mbar_expect_tx(w_full(rw.s), (uint32_t)kWStageBytes);
                    tma_load_3d(w_base + (uint32_t)(rw.s * kWStageBytes), &p.w2map, w_full(rw.s), 0, 0, s);
                }
                for (int blk = 0; blk < p.nblk3; ++blk, rw.next(kWStages)) {
                    // both K planes of one 48-channel output block in one stage: [48 rows x K 0-63] then [48 rows x K 64-127]
                    mbar_wait_t(w_empty(rw.s), rw.ph ^ 1u, tr, ww);
                    mbar_expect_tx(w_full(rw.s), (uint32_t)(2 * kNB3 * 128));
                    const uint32_t dst = w_base + (uint32_t)(rw.s * kWStageBytes);
                    tma_load_3d(dst, &p.w3map, w_full(rw.s), 0, blk * kNB3, 0);
                    tma_load_3d(dst + (uint32_t)(kNB3 * 128), &p.w3map, w_full(rw.s), kBlockK, blk * kNB3, 0);
                }
            }
            if (tr) p.dbg[1] = ww;
        }
    } else if (warp == 2 || warp == 3) {
        // =========================== MMA issuers ===========================
        const int wi = warp - 2;                 // output M tile of P2 / P3; P1: warp 2 -> halo M tiles 0 and 2, warp 3 -> 1
        const uint64_t desc0 = make_smem_desc(0);
        const uint32_t idesc96 = make_idesc(128, kCm), idesc48 = make_idesc(128, kNB3);
        const uint32_t xb = x_base >> 4, wb = w_base >> 4, t0b = t_base >> 4, t1b = (t_base + (uint32_t)kTBytes) >> 4;
        const uint32_t mrow = (128u * 128u) >> 4;                    // one M tile of rows in descriptor units
        RingPos rx = {0, 0}, rw = {0, 0};
        uint32_t it = 0;
        const bool tr = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
        long long m_w1 = 0, m_x = 0, m_pre2 = 0, m_w2 = 0, m_t2 = 0, m_e3 = 0, m_w3 = 0;
        const long long m_begin = tr ? clock64() : 0;
        for (;; ++it) {
            if (next_tile((int)it) >= p.total_tiles) break;
            // ---------------- P1 (D1[m] was drained by the epilogue of the previous tile: t1_ready waits below) ----------------
            // Both issuer warps wait for and release EVERY stage, also those whose MMAs the other warp issues: a parity wait is
            // only meaningful for a waiter that has seen the previous phase of the same barrier, so nobody may skip a stage
            // (and the producers must not run more than one ring round ahead of either warp).
            for (int part = 0; part < 2; ++part) {
                for (int kb = 0; kb < p.kb1; ++kb, rw.next(kWStages)) {
                    mbar_wait_t(w_full(rw.s), rw.ph, tr, m_w1);
                    const uint32_t b_lo = wb + (uint32_t)((rw.s * kWStageBytes) >> 4);
                    bool issued = false;
                    for (int m = part ? 2 : 0; m < (part ? 3 : 2); ++m, rx.next(kXStages)) {
                        const bool mine = part ? wi == 0 : m == wi;
                        mbar_wait_t(x_full(rx.s), rx.ph, tr, m_x);
                        tc_fence_after();
                        if (elect_one()) {
                            if (mine) {
                                const uint32_t a_lo = xb + (uint32_t)((rx.s * kXStageBytes) >> 4);
                                issue_mmas(tmem_base + (uint32_t)(m * kCm), desc0 + (uint64_t)a_lo, desc0 + (uint64_t)b_lo, idesc96, kb > 0 ? 1u : 0u, 4);
                                umma_commit(x_empty(rx.s));
                            } else {
                                mbar_arrive(x_empty(rx.s));
                            }
                        }
                        __syncwarp();
                        issued |= mine;
                    }
                    if (elect_one()) {
                        if (issued) umma_commit(w_empty(rw.s));
                        else mbar_arrive(w_empty(rw.s));
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    if (part == 0) umma_commit(d1_full(wi));
                    else if (wi == 0) umma_commit(d1_full(2));
                }
                __syncwarp();
            }
            // the D2[j] columns double as the P3 output buffers of the previous tile: wait until its epilogue drained them
            mbar_wait_t(d3_empty(wi, 0), ((2u * it) & 1u) ^ 1u, tr, m_pre2);
            mbar_wait_t(d3_empty(wi, 1), ((2u * it) & 1u) ^ 1u, tr, m_pre2);
            // output M tile j reads t1 rows [128 j, 128 j + 127 + 2 P + 2]: halo M tiles j and j + 1 (with P = 64, M tile 1 also
            // touches the first two rows of halo M tile 2 — only for the two padding columns, whose results are discarded)
            mbar_wait_t(t1_ready(wi), it & 1u, tr, m_pre2);
            mbar_wait_t(t1_ready(wi + 1), it & 1u, tr, m_pre2);
            tc_fence_after();
            // ---------------- P2 ----------------
            {
                const uint32_t dcol = tmem_base + (uint32_t)(kD2Col + wi * kCm);
                const uint32_t jrow = (uint32_t)wi * mrow;
                for (int s = 0; s < 14; ++s) {
                    mbar_wait_t(w_full(rw.s), rw.ph, tr, m_w2);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b_lo = wb + (uint32_t)((rw.s * kWStageBytes) >> 4);
                        if (s < 9) {
                            const uint32_t sh = (uint32_t)((((s / 3) << kLP) + (s % 3)) * 8);     // rows * 128 B >> 4
                            issue_mmas(dcol, desc0 + (uint64_t)(t0b + jrow + sh), desc0 + (uint64_t)b_lo, idesc96, s > 0 ? 1u : 0u, 4);
                        } else {
                            for (int h = 0; h < 2; ++h) {
                                const int t = 2 * (s - 9) + h;
                                if (t < 9) {
                                    const uint32_t sh = (uint32_t)((((t / 3) << kLP) + (t % 3)) * 8);
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
                if (elect_one()) umma_commit(d2_full(wi));
                __syncwarp();
            }
            // warp 2 issues the next tile's MMAs into D1[2]: its t1 must have left TMEM (warp 3 waited for it above)
            if (wi == 0) mbar_wait_t(t1_ready(2), it & 1u, tr, m_t2);
            mbar_wait_t(t2_ready(wi), it & 1u, tr, m_t2);
            tc_fence_after();
            // ---------------- P3 ----------------
            for (int blk = 0; blk < p.nblk3; ++blk) {
                const int b = blk & 1;
                const uint32_t use = 2u * it + (uint32_t)(blk >> 1);
                mbar_wait_t(d3_empty(wi, b), (use & 1u) ^ 1u, tr, m_e3);
                mbar_wait_t(w_full(rw.s), rw.ph, tr, m_w3);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t dcol = tmem_base + (uint32_t)(kD2Col + wi * kCm + b * kNB3);
                    const uint32_t b_lo = wb + (uint32_t)((rw.s * kWStageBytes) >> 4);
                    const uint32_t jrow = (uint32_t)wi * mrow;
                    issue_mmas(dcol, desc0 + (uint64_t)(t0b + jrow), desc0 + (uint64_t)b_lo, idesc48, 0u, 4);
                    issue_mmas(dcol, desc0 + (uint64_t)(t1b + jrow), desc0 + (uint64_t)(b_lo + (uint32_t)((kNB3 * 128) >> 4)), idesc48, 1u, 2);
                    umma_commit(w_empty(rw.s));
                    umma_commit(d3_full(wi, b));
                }
                __syncwarp();
                rw.next(kWStages);
            }
        }
        if (tr && wi == 0) {
            p.dbg[2] = m_w1; p.dbg[3] = m_x; p.dbg[4] = m_pre2; p.dbg[5] = m_w2; p.dbg[6] = m_t2; p.dbg[7] = m_e3; p.dbg[8] = m_w3;
            p.dbg[9] = clock64() - m_begin; p.dbg[10] = it;
        }
    } else {
        // =========================== epilogue warps ===========================
        const int ew = warp - 4;
        const int quad = warp & 3;               // TMEM lane quadrant of this warp
        const int set = ew >> 2;                 // M tile served: halo M tile `set` (all 96 columns) + half of halo M tile 2; output M tile `set`
        const uint32_t tlane = (uint32_t)(quad * 32) << 16;
        const float *b1s = bias_s, *b2s = bias_s + kCm, *b3s = bias_s + 2 * kCm;
        const uint32_t plane0 = t_base, plane1 = t_base + (uint32_t)kTBytes;
        const int LP = kLP, PM = kP - 1;
        const uint32_t ybuf = ystage_base + (uint32_t)(ew * kYStageBytes) + (uint32_t)(lane * kNB3 * 2);   // this thread's row of the warp's staging tile
        uint32_t ra[16], rb[16];
        // relu(acc + bias) (or zero) of 16 channels starting at ch -> bf16 -> this thread's row of the t planes
        auto to_plane = [&](const uint32_t *r, int ch, const float *bias, bool keep, int q) {
            float v[16];
            const float4 *bs = reinterpret_cast<const float4 *>(bias + ch);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 b4 = bs[i];
                v[4 * i] = keep ? fmaxf(__uint_as_float(r[4 * i]) + b4.x, 0.f) : 0.f;
                v[4 * i + 1] = keep ? fmaxf(__uint_as_float(r[4 * i + 1]) + b4.y, 0.f) : 0.f;
                v[4 * i + 2] = keep ? fmaxf(__uint_as_float(r[4 * i + 2]) + b4.z, 0.f) : 0.f;
                v[4 * i + 3] = keep ? fmaxf(__uint_as_float(r[4 * i + 3]) + b4.w, 0.f) : 0.f;
            }
            if (q >= kTRows) return;                          // halo rows 10 / 11 of M tile 2: never read
            const uint32_t rowa = (ch < kBlockK ? plane0 : plane1) + (uint32_t)q * 128u;
            const int k16 = (ch & 63) >> 3, sw = q & 7;       // 16-byte chunk index inside the 128-byte row
            sts128(rowa + (uint32_t)(((k16) ^ sw) << 4), pack8(v));
            sts128(rowa + (uint32_t)(((k16 + 1) ^ sw) << 4), pack8(v + 8));
        };
        // n chunks of 16 accumulator columns starting at TMEM column col0 (channel ch0), one TMEM load in flight ahead
        auto drain_to_plane = [&](uint32_t col0, int ch0, int n, const float *bias, bool keep, int q) {
            tmem_ld16_issue(col0, ra);
            for (int c = 0; c < n; c += 2) {
                tmem_ld_wait(ra);
                if (c + 1 < n) tmem_ld16_issue(col0 + (uint32_t)((c + 1) * 16), rb);
                to_plane(ra, ch0 + c * 16, bias, keep, q);
                if (c + 1 < n) {
                    tmem_ld_wait(rb);
                    if (c + 2 < n) tmem_ld16_issue(col0 + (uint32_t)((c + 2) * 16), ra);
                    to_plane(rb, ch0 + (c + 1) * 16, bias, keep, q);
                }
            }
        };
        griddep_wait();
        uint32_t it = 0;
        const bool tr = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
        long long e_d1 = 0, e_d2 = 0, e_d3 = 0, e_p1 = 0, e_p2 = 0, e_p3 = 0, e_res = 0, e_bw = 0, e_st = 0, e_top = 0;
        const long long e_begin = tr ? clock64() : 0;
        for (;; ++it) {
            const int tile = next_tile((int)it);
            if (tile >= p.total_tiles) break;
            const long long qt = tr ? clock64() : 0;
            const RbTile tc = rb_tile(p, tile);
            // this thread's output pixel (P3) and its residual row
            const int pos = set * 128 + quad * 32 + lane;
            const int ty = pos >> LP, tx = pos & PM;
            const int oy = tc.oy0 + ty, ox = tc.ox0 + tx;
            const bool valid = tx < kTW && oy < p.H && ox < p.W;
            const int64_t pix = ((int64_t)tc.n * p.H + oy) * p.W + ox;
            const bf16 *rp = p.res + pix * p.res_cstride + p.res_coff;
            if (tr) e_top += clock64() - qt;
            // ---------------- epilogue 1: D1 -> t1 (bf16, swizzled, zero outside the image) ----------------
            for (int pass = 0; pass < 2; ++pass) {
                const int mm = pass == 0 ? set : 2;
                const int q = mm * 128 + quad * 32 + lane;           // halo position
                const int iy = tc.oy0 - 1 + (q >> LP), ix = tc.ox0 - 1 + (q & PM);
                const bool inside = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const int ch0 = pass == 0 ? 0 : set * 48;
                mbar_wait_t(d1_full(mm), it & 1u, tr, e_d1);
                tc_fence_after();
                const long long c0 = tr ? clock64() : 0;
                drain_to_plane(tmem_base + tlane + (uint32_t)(mm * kCm + ch0), ch0, pass == 0 ? 6 : 3, b1s, inside, q);
                if (tr) e_p1 += clock64() - c0;
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(t1_ready(mm));
            }
            // the residual of the first output block travels while P2 runs
            uint4 rr[6];
            auto load_res = [&](const bf16 *src) {
                if (p.wide_io) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const U8 w = ldg256(src + 16 * i);
                        rr[2 * i] = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
                        rr[2 * i + 1] = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 6; ++i) rr[i] = reinterpret_cast<const uint4 *>(src)[i];
                }
            };
#pragma unroll
            for (int i = 0; i < 6; ++i) rr[i] = make_uint4(0, 0, 0, 0);
            if (valid) load_res(rp);
            // ---------------- epilogue 2: D2 -> t2 (over t1: every P2 MMA of BOTH output M tiles has completed) ----------------
            mbar_wait_t(d2_full(0), it & 1u, tr, e_d2);
            mbar_wait_t(d2_full(1), it & 1u, tr, e_d2);
            tc_fence_after();
            const long long c1 = tr ? clock64() : 0;
            drain_to_plane(tmem_base + tlane + (uint32_t)(kD2Col + set * kCm), 0, 6, b2s, true, pos);
            if (tr) e_p2 += clock64() - c1;
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(t2_ready(set));
            // ---------------- epilogue 3: D3 blocks -> y = act(acc + b3 + res) ----------------
            for (int blk = 0; blk < p.nblk3; ++blk) {
                const int b = blk & 1;
                const uint32_t use = 2u * it + (uint32_t)(blk >> 1);
                const long long q0 = tr ? clock64() : 0;
                uint4 cr[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) cr[i] = rr[i];
                if (valid && blk + 1 < p.nblk3) load_res(rp + (blk + 1) * kNB3);      // one block ahead of its use
                if (tr) e_res += clock64() - q0;
                mbar_wait_t(d3_full(set, b), use & 1u, tr, e_d3);
                tc_fence_after();
                const long long c2 = tr ? clock64() : 0;
                const uint32_t col0 = tmem_base + tlane + (uint32_t)(kD2Col + set * kCm + b * kNB3);
                auto finish = [&](const uint32_t *r, int c) {
                    float v[16], rs[16];
                    unpack8(cr[2 * c], rs);
                    unpack8(cr[2 * c + 1], rs + 8);
                    const int co = blk * kNB3 + c * 16;
                    const float4 *bs = reinterpret_cast<const float4 *>(b3s + co);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 b4 = bs[i];
                        v[4 * i] = __uint_as_float(r[4 * i]) + b4.x + rs[4 * i];
                        v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4.y + rs[4 * i + 1];
                        v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4.z + rs[4 * i + 2];
                        v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4.w + rs[4 * i + 3];
                    }
                    if (p.final_relu) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    sts128(ybuf + (uint32_t)(c * 32), pack8(v));
                    sts128(ybuf + (uint32_t)(c * 32 + 16), pack8(v + 8));
                };
                // the TMA store of the previous block has read the staging row
                const long long q1 = tr ? clock64() : 0;
                if (lane == 0) bulk_wait_read0();
                __syncwarp();
                if (tr) e_bw += clock64() - q1;
                tmem_ld16_issue(col0, ra);
                tmem_ld_wait(ra);
                tmem_ld16_issue(col0 + 16u, rb);
                finish(ra, 0);
                tmem_ld_wait(rb);
                tmem_ld16_issue(col0 + 32u, ra);
                finish(rb, 1);
                tmem_ld_wait(ra);
                // every TMEM read of this buffer has completed: the issuer may overwrite it while the last chunk is stored
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d3_empty(set, b));
                finish(ra, 2);
                const long long q2 = tr ? clock64() : 0;
                fence_async_smem();
                __syncwarp();
                if (tr) e_st += clock64() - q2;
                if (lane == 0) {      // this warp's row of the tile: columns beyond the image and rows below it are clipped by TMA
                    tma_store_4d(&p.ymap, ystage_base + (uint32_t)(ew * kYStageBytes), blk * kNB3, tc.ox0, tc.oy0 + (pos >> LP), tc.n);
                    bulk_commit();
                }
                if (tr) e_p3 += clock64() - c2;
            }
        }
        if (lane == 0) bulk_wait0();      // every TMA store of this warp has been written
        if (tr) {
            p.dbg[11] = e_d1; p.dbg[12] = e_p1; p.dbg[13] = e_d2; p.dbg[14] = e_p2; p.dbg[15] = e_d3; p.dbg[16] = e_p3;
            p.dbg[17] = clock64() - e_begin;
            p.dbg[18] = e_res; p.dbg[19] = e_bw; p.dbg[20] = e_st; p.dbg[21] = e_top;
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
    RGBD_CHECK_ARG(d->sched_ws != nullptr && ((uintptr_t)d->sched_ws & 3) == 0,
                   "sched_ws: the caller provides a zero-initialised int32 in device memory per plan (tile counter)");
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
    p.sched = reinterpret_cast<int32_t *>(d->sched_ws);
    p.prefetch = 1;
    p.static_tiles = 0;
#ifdef RGBD_TIMING_PROBES
    // A / B switches of development builds only (RGBD_BUILD_DEFINES=-DRGBD_TIMING_PROBES)
    p.prefetch = getenv("RGBD_RB_NOPREFETCH") == nullptr;
    p.static_tiles = getenv("RGBD_RB_STATIC") != nullptr;
#endif
    p.wide_io = ((d->res_cstride | d->res_coff) & 15) == 0 && ((uintptr_t)d->res & 31) == 0;
    p.dbg = nullptr;
#ifdef RGBD_TIMING_PROBES
    if (const char *e = getenv("RGBD_TC_TRACE")) p.dbg = (long long *)strtoull(e, nullptr, 0);
#endif
    // fixed tile geometry (kP, kR, kTW above): 8 rows x 30 useful columns of output per tile
    p.tiles_x = (d->W + kTW - 1) / kTW;
    p.tiles_y = (d->H + kR - 1) / kR;
    const long total = (long)d->N * p.tiles_x * p.tiles_y;
    RGBD_CHECK_ARG(total < 2147483647L, "too many tiles");
    p.total_tiles = (int)total;
    // [x ring | t plane 0 | t plane 1 | weight ring | barriers | TMEM slot | biases | output staging rows]
    pl->smem = 1024 + (size_t)kXStages * kXStageBytes + 2 * (size_t)kTBytes + (size_t)kWStages * kWStageBytes +
               8 * kRbBarriers + 16 + 4 * (2 * kCm + kMaxCout) + 128 + 8 * (size_t)kYStageBytes + 64;
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
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)kP, (cuuint32_t)kRpm, 1};
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
    if (!rc) {   // y: bf16 NHWC view, stored by TMA as packed boxes {48 channels, 30 columns, 1 row} (no swizzle)
        cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->y_cstride * 2, (cuuint64_t)d->W * d->y_cstride * 2,
                                 (cuuint64_t)d->H * d->W * d->y_cstride * 2};
        cuuint32_t box[4] = {(cuuint32_t)kNB3, (cuuint32_t)kTW, 1, 1};
        rc = encode_map(&p.ymap, reinterpret_cast<char *>(d->y) + (int64_t)d->y_coff * 2, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
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
        fprintf(stderr, "rb_fused: %dx%d Cin %d Cout %d | tiles %d smem %zu\n", d->H, d->W, d->Cin, d->Cout, p.total_tiles, pl->smem);
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
