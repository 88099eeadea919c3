// tcgen05 tensor-core implicit-GEMM convolution, halo-resident form (bf16 operands, fp32 TMEM
// accumulators).  One persistent CTA per SM.
//
// GEMM view: D[pos, co] = sum_{group, kblock, tap in group} A_group,kblock[pos + shift(tap), :] * W[tap][co][kblock]
//   * A "super-tile" is R rows x P columns of one image's output lattice laid out as R*P <= 256
//     consecutive *positions* (row pitch P = TW + ext_x - 1: TW useful columns plus the filter's
//     horizontal extent).  Positions 0..127 and 128..255 are the M = 128 rows of up to two
//     accumulators (MT = 1 or 2) that share every weight fetch.
//   * For each (tap group, 64-channel block) ONE 4-D TMA box {64 ch, P, R + ext_y - 1, 1} of the NHWC
//     input — the tile plus its halo — lands in shared memory as consecutive 128-byte rows with the
//     128-byte swizzle.  Because the swizzle is a function of the absolute shared-memory address, the
//     A operand of tap (qy, qx) is the SAME halo tile read through a UMMA descriptor whose start
//     address is shifted by (qy * P + qx) rows (verified on B200: profiles/tools/desc_test.cu).  A k x k
//     filter therefore fetches its input once instead of k*k times.  Out-of-image rows/columns are
//     zero-filled by TMA (= the reference's zero padding).  A tap group is the set of taps that read
//     the same input lattice: one group for stride 1, the four input parity classes for stride 2
//     (tensor maps with doubled strides).
//   * B operand: 3-D TMA box {64 ci, BN co, 1 tap} of the packed bf16 weights [tap][co][ci], in its
//     own, deeper ring; each B stage feeds the MMAs of both accumulators.
//   * Residual adds of RGBD_EPI_LINEAR (ResidualBottleneck / ResidualUnit skip connections) ride on
//     the tensor core: the residual tile is one more A operand multiplied by an identity B tile, so
//     its HBM latency is hidden by the same TMA pipeline as the activations and the epilogue issues
//     no global loads at all.
//   * Warp roles (384 threads): warp 0 = A producer, warp 1 = B producer, warps 2 / 3 = MMA issuers of
//     accumulator 0 / 1 (warp 2 also owns the TMEM allocation), warps 4..11 = epilogue (TMEM lane quadrant = warp % 4; warps
//     4..7 drain accumulator 0, warps 8..11 accumulator 1, or the odd column chunks when MT = 1).
//     TMEM: 512 columns = 2 buffers x 2 accumulators x 128 columns, so the epilogue of super-tile i
//     overlaps the MMAs of super-tile i + 1.
//
// The K reduction order (group, channel block, tap, 16-channel step; residual last) is fixed and does
// not depend on the batch size or on where a pixel falls inside a tile: encoder and decoder reproduce
// the same bits.
#include "tc_common.cuh"
#include <new>
#include <cstdlib>

namespace {

constexpr int kEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = 32 * (kFirstEpiWarp + kEpiWarps);   // 384
constexpr int kMaxA = 4, kMaxB = 8;                // ring depths
constexpr int kRingBudget = 160 * 1024;            // A ring + B ring; leaves room for a rANS block on the SM
constexpr int kStageBytes = kEpiWarps * 2048;      // per epilogue warp: 32 rows x 64 B output transpose tile
constexpr int kMaxAStage = 52 * 1024;              // halo tile (one channel block)
constexpr int kMinSmem = 120 * 1024;               // > half an SM: never two of these CTAs on one SM (512 TMEM columns each)
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxGroups = 4;
constexpr int kTileQ = 4;                          // depth of the CTA's tile queue (dynamic scheduler)
constexpr int kMaxBias = 2048;                   // widest layer: the Swin MLP of STF_united, 384 -> 1536

struct HParams {
    CUtensorMap amap[4];
    CUtensorMap bmap, rmap, emap, mmap;   // mmap: gate operand (RGBD_EPI_GATE `mul`) tiles
    rgbd_conv_desc d;
    int32_t TW, R, P, MT, box_rows;
    int32_t tiles_x, tiles_y, BN, kblocks, n_ntiles, total_tiles;
    int32_t nA, nB, a_bytes, a_tx, r_tx, b_bytes, tps;   // tps: taps per weight stage (stage = tps * b_bytes)
    int32_t mul_blocks, mul_bytes;   // gate operand staged by TMA: 64-channel blocks per N tile, bytes per block (0: epilogue loads it)
    int32_t ngroups, res_blocks, stage_epi;   // stage_epi: epilogue stores go through a shared-memory transpose
    int8_t g_map[kMaxGroups], g_qy[kMaxGroups], g_qx[kMaxGroups], g_first[kMaxGroups + 1];
    int16_t t_shift[RGBD_MAX_TAPS];
    int8_t t_w[RGBD_MAX_TAPS];
    long long *dbg;   // optional cycle counters of CTA 0 (RGBD_TC_TRACE = device pointer to 16 int64), else NULL
};

struct TileCoord {
    int n, oy0, ox0, co0, bn;
};
__device__ __forceinline__ TileCoord tile_coord(const HParams &p, int tile) {
    // n-tile fastest: CTAs that run side by side share the same A halo through L2
    TileCoord t;
    const int nt = tile % p.n_ntiles;
    const int mt = tile / p.n_ntiles;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    t.n = mt / tiles_per_img;
    const int trem = mt - t.n * tiles_per_img;
    t.oy0 = (trem / p.tiles_x) * p.R;
    t.ox0 = (trem % p.tiles_x) * p.TW;
    t.co0 = nt * p.BN;
    t.bn = min(p.BN, p.d.cout_pad - t.co0);   // multiple of 16
    return t;
}

// Register cap: 384 threads x 144 registers leave 10 K registers (and the ring budget leaves >= 60 KB of
// shared memory) on the SM, so that a rANS coder block (64 threads x 48 registers, 57 + 4 KB) can be
// co-resident — the serial rANS chains of the other pipeline slots must not fence SMs off from the convs.
template <typename TOut, int kEpi>
__global__ void __maxnreg__(144)
conv_halo_kernel(const __grid_constant__ HParams p) {
    constexpr bool kGate = kEpi == RGBD_EPI_GATE;
    extern __shared__ uint8_t smem_raw[];
    using bf16 = __nv_bfloat16;
    const rgbd_conv_desc &d = p.d;
    // 1024-byte aligned carve-up: [A ring | B ring | barriers | tmem ptr | bias]
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t a_base = base;
    const uint32_t b_base = a_base + (uint32_t)(p.nA * p.a_bytes);
    const uint32_t b_stage = (uint32_t)(p.tps * p.b_bytes);
    const uint32_t bar_base = b_base + (uint32_t)p.nB * b_stage;
    auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxA + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxA + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kMaxA + kMaxB + s); };
    auto tmem_full_bar = [&](int a, int j) { return bar_base + 8u * (uint32_t)(2 * kMaxA + 2 * kMaxB + 2 * a + j); };
    auto tmem_empty_bar = [&](int a, int j) { return bar_base + 8u * (uint32_t)(2 * kMaxA + 2 * kMaxB + 4 + 2 * a + j); };
    auto mul_full = [&](int b) { return bar_base + 8u * (uint32_t)(2 * kMaxA + 2 * kMaxB + 8 + b); };
    auto mul_empty = [&](int b) { return bar_base + 8u * (uint32_t)(2 * kMaxA + 2 * kMaxB + 10 + b); };
    auto tq_full = [&](int q) { return bar_base + 8u * (uint32_t)(2 * kMaxA + 2 * kMaxB + 12 + q); };
    auto tq_empty = [&](int q) { return bar_base + 8u * (uint32_t)(2 * kMaxA + 2 * kMaxB + 12 + kTileQ + q); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxA + 2 * kMaxB + 12 + 2 * kTileQ);
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));
    // per-tap A descriptor offsets (row shift * 128 B >> 4), then the bias
    uint32_t *shift_s = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot + 16u - raw));
    volatile int *tq_s = reinterpret_cast<volatile int *>(smem_raw + (tmem_slot + 16u + 104u - raw));   // kTileQ tile indices
    float *bias_s = reinterpret_cast<float *>(smem_raw + (tmem_slot + 16u + 128u - raw));
    const uint32_t stage_base = tmem_slot + 16u + 128u + 4u * (uint32_t)d.cout_pad;   // 16-byte aligned (cout_pad % 16 == 0)
    // gate operand tiles (1024-byte aligned: the TMA 128-byte swizzle is a function of the address)
    const uint32_t mul_base = (stage_base + (p.stage_epi ? (uint32_t)kStageBytes : 0u) + 1023u) & ~1023u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        // one MMA-issuing warp per accumulator: a stage is free once every issuer's MMAs have read it
        for (int s = 0; s < p.nA; ++s) {
            mbar_init(a_full(s), 1);
            mbar_init(a_empty(s), (uint32_t)p.MT);
        }
        for (int s = 0; s < p.nB; ++s) {
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), (uint32_t)p.MT);
        }
        for (int a = 0; a < 2; ++a)
            for (int j = 0; j < 2; ++j) {
                mbar_init(tmem_full_bar(a, j), 1);
                mbar_init(tmem_empty_bar(a, j), (uint32_t)(kEpiWarps / p.MT));
            }
        for (int b = 0; b < 2; ++b) {
            mbar_init(mul_full(b), 1);
            mbar_init(mul_empty(b), kEpiWarps);
        }
        // tile queue: filled by the A producer, read by the B producer, the MMA issuer warps and the epilogue warps
        for (int q = 0; q < kTileQ; ++q) {
            mbar_init(tq_full(q), 1);
            mbar_init(tq_empty(q), (uint32_t)(1 + p.MT + kEpiWarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
    if (warp == 3) {
        if (d.epi == RGBD_EPI_SHUFFLE2) {   // columns are (parity, channel): every parity gets the channel's bias
            if (lane < 16) bias_s[lane] = (d.bias && (lane & 3) < d.Cout) ? d.bias[lane & 3] : 0.f;
        } else {
            for (int i = lane; i < d.cout_pad; i += 32) bias_s[i] = (d.bias && i < d.Cout) ? d.bias[i] : 0.f;
        }
        if (lane < RGBD_MAX_TAPS) shift_s[lane] = (uint32_t)((int)p.t_shift[lane] * 8);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Dynamic tile scheduler: tiles are handed out in index order from a global counter (p.d.sched_ws), so a CTA that
    // starts late — its SM was busy with rANS coder blocks of another pipeline slot — does not hold the grid back: the
    // other CTAs take its share.  The A producer thread draws the tile and publishes it through a small queue; every
    // other role reads the k-th entry.  Which CTA computes a tile has no effect on the result.
    auto next_tile = [&](int k) -> int {      // consumers: all lanes of the calling warp
        const int q = k % kTileQ;
        mbar_wait(tq_full(q), (uint32_t)((k / kTileQ) & 1));
        const int t = tq_s[q];
        __syncwarp();
        if (lane == 0) mbar_arrive(tq_empty(q));
        return t;
    };

    if (warp == 0) {
        // =========================== A producer (halo tiles, residual tiles) ===========================
        if (lane == 0) {
            RingPos ra = {0, 0};
            int mli = 0;
            const bool tr = p.dbg != nullptr && blockIdx.x == 0;
            long long w_empty = 0;
            griddep_wait();
            int *counter = reinterpret_cast<int *>(d.sched_ws);
            for (int k = 0;; ++k) {
                // draw the next tile and publish it to the other roles
                const int q = k % kTileQ;
                mbar_wait(tq_empty(q), (uint32_t)((k / kTileQ) & 1) ^ 1u);
                int tile = atomicAdd(counter, 1);
                if (tile >= p.total_tiles) {
                    // every CTA draws exactly one value past the end; the last of them re-arms the counter for the
                    // next launch of this plan (stream order keeps that launch behind this one)
                    if (tile == p.total_tiles + (int)gridDim.x - 1) atomicExch(counter, 0);
                    tile = p.total_tiles;
                }
                tq_s[q] = tile;
                mbar_arrive(tq_full(q));
                if (tile >= p.total_tiles) break;
                const TileCoord tc = tile_coord(p, tile);
                for (int g = 0; g < p.ngroups; ++g) {
                    for (int kb = 0; kb < p.kblocks; ++kb, ra.next(p.nA)) {
                        const int s = ra.s;
                        mbar_wait_t(a_empty(s), ra.ph ^ 1u, tr, w_empty);
                        mbar_expect_tx(a_full(s), (uint32_t)p.a_tx);
                        tma_load_4d(a_base + (uint32_t)(s * p.a_bytes), &p.amap[p.g_map[g]], a_full(s), kb * kBlockK,
                                    tc.ox0 + p.g_qx[g], tc.oy0 + p.g_qy[g], tc.n);
                    }
                }
                const int rb = p.res_blocks ? (tc.bn + 63) >> 6 : 0;
                for (int jb = 0; jb < rb; ++jb, ra.next(p.nA)) {
                    const int s = ra.s;
                    mbar_wait(a_empty(s), ra.ph ^ 1u);
                    mbar_expect_tx(a_full(s), (uint32_t)p.r_tx);
                    tma_load_4d(a_base + (uint32_t)(s * p.a_bytes), &p.rmap, a_full(s), tc.co0 + jb * kBlockK, tc.ox0,
                                tc.oy0, tc.n);
                }
                if (p.mul_blocks) {
                    // gate operand of this tile: lands while the MMAs run; the epilogue reads it from shared memory
                    // (double buffered: tile i + 1's operand travels while the epilogue works on tile i)
                    const int nb = (tc.bn + 63) >> 6;
                    const int mb = mli & 1;
                    mbar_wait(mul_empty(mb), (uint32_t)((mli >> 1) & 1) ^ 1u);
                    mbar_expect_tx(mul_full(mb), (uint32_t)(nb * p.r_tx));
                    for (int jb = 0; jb < nb; ++jb)
                        tma_load_4d(mul_base + (uint32_t)((mb * p.mul_blocks + jb) * p.mul_bytes), &p.mmap, mul_full(mb),
                                    tc.co0 + jb * kBlockK, tc.ox0, tc.oy0, tc.n);
                    ++mli;
                }
            }
            // this CTA has requested its last input: the next kernel of the stream may start its prologue on the
            // SMs that drain first (it launches once every CTA of this grid got here or exited)
            griddep_launch();
            if (tr) p.dbg[0] = w_empty;
        }
    } else if (warp == 1) {
        // =========================== B producer (weights, identity tiles) ===========================
        {
            RingPos rbp = {0, 0};
            const bool tr = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
            long long w_empty = 0;
            griddep_wait();   // per-image filters (rgbd_scale_weights) are produced by the previous kernel
            for (int k = 0;; ++k) {
                const int tile = next_tile(k);      // whole warp; lane 0 issues the loads
                if (tile >= p.total_tiles) break;
                if (lane != 0) continue;
                const TileCoord tc = tile_coord(p, tile);
                for (int g = 0; g < p.ngroups; ++g) {
                    const int t0 = p.g_first[g], t1 = p.g_first[g + 1];
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        for (int t = t0; t < t1; t += p.tps, rbp.next(p.nB)) {
                            const int s = rbp.s;
                            const int n = min(p.tps, t1 - t);
                            mbar_wait_t(b_empty(s), rbp.ph ^ 1u, tr, w_empty);
                            mbar_expect_tx(b_full(s), (uint32_t)(n * p.b_bytes));
                            for (int tt = 0; tt < n; ++tt)
                                tma_load_3d(b_base + (uint32_t)s * b_stage + (uint32_t)(tt * p.b_bytes), &p.bmap, b_full(s),
                                            kb * kBlockK, tc.co0, p.t_w[t + tt] + tc.n * d.w_image_stride);
                        }
                    }
                }
                const int rb = p.res_blocks ? (tc.bn + 63) >> 6 : 0;
                for (int jb = 0; jb < rb; ++jb, rbp.next(p.nB)) {
                    const int s = rbp.s;
                    mbar_wait(b_empty(s), rbp.ph ^ 1u);
                    mbar_expect_tx(b_full(s), (uint32_t)p.b_bytes);
                    tma_load_3d(b_base + (uint32_t)s * b_stage, &p.emap, b_full(s), 0, 0, jb);
                }
            }
            if (tr) p.dbg[1] = w_empty;
        }
    } else if (warp == 2 || warp == 3) {
        // =========================== MMA issuers ===========================
        // Warp 2 owns accumulator (M tile) 0, warp 3 accumulator 1: one thread cannot issue
        // tcgen05.mma + barrier traffic fast enough to keep the tensor pipe busy at N <= 128, and two
        // issuers with one accumulator each keep every accumulator's K order fixed.  The whole warp
        // walks the loops and waits on the barriers; one elected lane issues.
        const int j = warp - 2;
        if (j < p.MT) {
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            int li = 0;
            const bool tr = p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && j == 0;
            long long w_tmem = 0, w_a = 0, w_b = 0;
            const long long t_begin = tr ? clock64() : 0;
            const uint64_t desc0 = make_smem_desc(0);
            const int nA = p.nA, nB = p.nB;
            const uint32_t a_step = (uint32_t)p.a_bytes >> 4, b_step = b_stage >> 4, b_tap = (uint32_t)p.b_bytes >> 4;
            const int tps = p.tps;
            const uint32_t a_lo0 = (a_base >> 4) + (uint32_t)(j * 1024), b_lo0 = b_base >> 4;
            uint32_t a_lo = a_lo0, b_lo = b_lo0;     // descriptor start-address fields of the current ring stages
            const int kblocks = p.kblocks, ngroups = p.ngroups;
            for (;; ++li) {
                const int tile = next_tile(li);
                if (tile >= p.total_tiles) break;
                const TileCoord tc = tile_coord(p, tile);
                const int acc = li & 1;
                const uint32_t use = (uint32_t)(li >> 1);
                mbar_wait_t(tmem_empty_bar(acc, j), (use & 1u) ^ 1u, tr, w_tmem);   // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t idesc = make_idesc(128, tc.bn);
                const uint32_t dcol = tmem_base + (uint32_t)(acc * 256 + j * 128);
                uint32_t started = 0;
                for (int g = 0; g < ngroups; ++g) {
                    const int t0 = p.g_first[g], t1 = p.g_first[g + 1];
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait_t(a_full(sa), pha, tr, w_a);
                        // only the 16-channel groups that hold real input channels (tail block may be short)
                        const int kleft = d.Cin - kb * kBlockK;
                        const int nk = kleft >= kBlockK ? 4 : (kleft + 15) >> 4;
                        for (int t = t0; t < t1; t += tps) {
                            const int n = min(tps, t1 - t);
                            mbar_wait_t(b_full(sb), phb, tr, w_b);
                            tc_fence_after();
                            if (elect_one()) {
                                issue_mmas(dcol, desc0 + (uint64_t)(a_lo + shift_s[t]), desc0 + (uint64_t)b_lo, idesc, started, nk);
                                for (int tt = 1; tt < n; ++tt)
                                    issue_mmas(dcol, desc0 + (uint64_t)(a_lo + shift_s[t + tt]), desc0 + (uint64_t)(b_lo + tt * b_tap),
                                               idesc, 1u, nk);
                                umma_commit(b_empty(sb));   // frees the weight stage once these MMAs have read it
                            }
                            __syncwarp();
                            started = 1;
                            b_lo += b_step;
                            if (++sb == nB) {
                                sb = 0;
                                phb ^= 1u;
                                b_lo = b_lo0;
                            }
                        }
                        if (elect_one()) umma_commit(a_empty(sa));
                        __syncwarp();
                        a_lo += a_step;
                        if (++sa == nA) {
                            sa = 0;
                            pha ^= 1u;
                            a_lo = a_lo0;
                        }
                    }
                }
                const int rb = p.res_blocks ? (tc.bn + 63) >> 6 : 0;
                for (int jb = 0; jb < rb; ++jb) {
                    mbar_wait_t(a_full(sa), pha, tr, w_a);
                    mbar_wait_t(b_full(sb), phb, tr, w_b);
                    tc_fence_after();
                    if (elect_one()) {
                        const int kvalid = min(kBlockK, tc.bn - jb * kBlockK);
                        issue_mmas(dcol, desc0 + (uint64_t)a_lo, desc0 + (uint64_t)b_lo, idesc, started, (kvalid + 15) >> 4);
                        umma_commit(b_empty(sb));
                        umma_commit(a_empty(sa));
                    }
                    __syncwarp();
                    started = 1;
                    b_lo += b_step;
                    if (++sb == nB) {
                        sb = 0;
                        phb ^= 1u;
                        b_lo = b_lo0;
                    }
                    a_lo += a_step;
                    if (++sa == nA) {
                        sa = 0;
                        pha ^= 1u;
                        a_lo = a_lo0;
                    }
                }
                if (elect_one()) umma_commit(tmem_full_bar(acc, j));
                __syncwarp();
            }
            if (tr) {
                p.dbg[2] = w_tmem;
                p.dbg[3] = w_a;
                p.dbg[4] = w_b;
                p.dbg[5] = clock64() - t_begin;
                p.dbg[6] = li;
            }
        }
    } else if (warp >= kFirstEpiWarp) {
        // ============ epilogue: TMEM -> registers -> NHWC global ============
        const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
        const int set = (warp - kFirstEpiWarp) >> 2;     // 0 / 1
        const int j = p.MT == 2 ? set : 0;               // accumulator (M tile) this warp drains
        const int cbfirst = p.MT == 2 ? 0 : set;         // MT == 1: the two warp sets split the 32-column blocks
        const int cbstep = p.MT == 2 ? 1 : 2;
        const int pos = j * 128 + quad * 32 + lane;      // position inside the super-tile
        const int ty = pos / p.P, tx = pos - ty * p.P;
        TOut *y = reinterpret_cast<TOut *>(d.y);
        TOut *y2 = reinterpret_cast<TOut *>(d.y2);
        const bf16 *res = reinterpret_cast<const bf16 *>(d.res);
        const bf16 *mul = reinterpret_cast<const bf16 *>(d.mul);
        constexpr int kVecOut = 16 / (int)sizeof(TOut);   // elements per 16 B
        const bool y_vec = ((d.y_cstride | d.y_coff) % kVecOut) == 0;
        const bool y2_vec = ((d.y2_cstride | d.y2_coff) % kVecOut) == 0;
        const bool res_vec = ((d.res_cstride | d.res_coff) & 7) == 0;
        const bool mul_vec = ((d.mul_cstride | d.mul_coff) & 7) == 0;
        // residual handled here only when it did not ride on the tensor core
        const bool res_direct = res != nullptr && kEpi != RGBD_EPI_BILERP && p.res_blocks == 0;
        const float slope = act_slope(d.act);
        const bool mul_staged = kGate && p.mul_blocks != 0;
        // Coalesced stores: a thread owns one pixel row of the accumulator, so a direct 16-byte store touches
        // 32 different 128-byte lines per warp instruction (one LSU wavefront each) — for the small-K layers
        // that is the bottleneck.  Instead each warp transposes 32 rows x 32 channels through its private
        // 2 KB tile (16-byte pieces XOR-swizzled by row pair) and stores 8 rows x 64 contiguous bytes per
        // instruction.
        const bool staged = p.stage_epi != 0 && sizeof(TOut) == 2 && kEpi != RGBD_EPI_BILERP && y_vec &&
                            (y2 == nullptr || y2_vec) && (d.Cout & 7) == 0;
        const uint32_t stg = stage_base + (uint32_t)(warp - kFirstEpiWarp) * 2048u;
        int li = 0;
        griddep_wait();   // the epilogue reads (residual / gate / bilinear operands) and writes activation memory
        const bool tr = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == kFirstEpiWarp * 32;
        long long w_full = 0, et[3] = {0, 0, 0};
        const long long t_begin = tr ? clock64() : 0;
        for (;; ++li) {
            const int tile = next_tile(li);
            if (tile >= p.total_tiles) break;
            const TileCoord tc = tile_coord(p, tile);
            const int acc = li & 1;
            const uint32_t use = (uint32_t)(li >> 1);
            const int sy = tc.oy0 + ty, sx = tc.ox0 + tx;    // site in the Hs x Ws output lattice
            const bool valid = ty < p.R && tx < p.TW && sy < d.Hs && sx < d.Ws;
            const int oy = sy * d.o_step + d.o_off_y, ox = sx * d.o_step + d.o_off_x;
            const int64_t opix = ((int64_t)tc.n * d.Ho + oy) * d.Wo + ox;
            const int nchunks = tc.bn >> 4;
            // gate / late-residual operands of this thread's pixel are requested BEFORE the accumulator
            // is ready, one 16-column chunk ahead of their use
            uint4 pr[2], pm[2];
            auto prefetch = [&](int c) {
                const int co = tc.co0 + c * 16;
                const bool on = valid && c < nchunks && d.Cout - co >= 16;
                pr[0] = pr[1] = pm[0] = pm[1] = make_uint4(0, 0, 0, 0);
                if (on && res_direct && res_vec) {
                    ldg_2x128(res + opix * d.res_cstride + d.res_coff + co, pr[0], pr[1]);
                }
                if (kGate && on && mul_vec && !mul_staged) {
                    ldg_2x128(mul + opix * d.mul_cstride + d.mul_coff + co, pm[0], pm[1]);
                }
            };
            const int nblk = (nchunks + 1) >> 1;
            // the chunk this warp handles after chunk c (its blocks are cbfirst, cbfirst + cbstep, ...)
            auto chunk_after = [&](int c) {
                if ((c & 1) == 0 && c + 1 < nchunks) return c + 1;
                const int nb = (c >> 1) + cbstep;
                return nb < nblk ? 2 * nb : nchunks;
            };
            const int pix_code = valid ? (int)opix : -1;
            if (kGate || res_direct) prefetch(2 * cbfirst);
            int by0 = 0, by1 = 0, bx0 = 0, bx1 = 0;
            float ly = 0.f, lx = 0.f;
            if (kEpi == RGBD_EPI_BILERP && valid) {
                bilerp_axis(oy, d.res_H, d.Ho, by0, by1, ly);
                bilerp_axis(ox, d.res_W, d.Wo, bx0, bx1, lx);
            }
            if (mul_staged) mbar_wait(mul_full(li & 1), (uint32_t)((li >> 1) & 1));
            mbar_wait_t(tmem_full_bar(acc, j), use & 1u, tr, w_full);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256 + j * 128);
            uint32_t r0[16], r1[16];
            auto process = [&](int c, const uint32_t *rv) {
                const int cnext = chunk_after(c);
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rv[i]);
                if constexpr (kEpi == RGBD_EPI_SHUFFLE2) {
                    // 16 columns = 4 output parities x 4 channels of a stride-2 transposed conv: pixel shuffle
                    if (valid) {
                        // bf16 rows padded to >= 4 channels: the 4 columns of a parity leave as ONE 8-byte store (the columns
                        // beyond Cout have zero weights and zero bias: the padding lanes receive act(0) = 0)
                        const bool pack4 = sizeof(TOut) == 2 && ((d.y_cstride | d.y_coff) & 3) == 0 &&
                                           (reinterpret_cast<uintptr_t>(y) & 7) == 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int64_t pix = ((int64_t)tc.n * d.Ho + 2 * sy + (q >> 1)) * d.Wo + 2 * sx + (q & 1);
                            TOut *dst = y + pix * d.y_cstride + d.y_coff;
                            float o[4];
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) o[ch] = act_fn(v[4 * q + ch] + bias_s[4 * q + ch], slope);
                            if (pack4) {
                                const __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
                                *reinterpret_cast<uint2 *>(dst) = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
                            } else {
#pragma unroll
                                for (int ch = 0; ch < 4; ++ch)
                                    if (ch < d.Cout) ElemIO<TOut>::st(dst + ch, o[ch]);
                            }
                        }
                    }
                    return;
                }
                const int co = tc.co0 + c * 16;
                const int nvalid = d.Cout - co;
                uint4 cr[2] = {pr[0], pr[1]}, cm[2] = {pm[0], pm[1]};
                if ((kGate || res_direct) && cnext < nchunks) prefetch(cnext);
                if (valid && nvalid > 0) {
                    const bool full16 = nvalid >= 16;
                    {
                        const float4 *bs = reinterpret_cast<const float4 *>(bias_s + co);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 b4 = bs[i];
                            v[4 * i] += b4.x;
                            v[4 * i + 1] += b4.y;
                            v[4 * i + 2] += b4.z;
                            v[4 * i + 3] += b4.w;
                        }
                    }
                    float r[16];
                    if (res_direct) {
                        if (res_vec && full16) {
                            unpack8(cr[0], r);
                            unpack8(cr[1], r + 8);
                        } else {
                            load16_bf16(res + opix * d.res_cstride + d.res_coff + co, false, nvalid, r);
                        }
                    }
                    if (kGate) {
                        // sigmoid(v) = 0.5 * tanh(0.5 v) + 0.5: one MUFU op per element instead of ex2 + rcp
                        uint4 q0, q1;
                        if (mul_staged) {
                            // this thread's pixel row of the TMA-staged tile: 16-byte chunk q lives at q ^ (row & 7)
                            const uint32_t rowa = mul_base + (uint32_t)(((li & 1) * p.mul_blocks + (c >> 2)) * p.mul_bytes + pos * 128);
                            const int qq = (c & 3) * 2, sw = pos & 7;
                            q0 = lds128(rowa + (uint32_t)(((qq) ^ sw) << 4));
                            q1 = lds128(rowa + (uint32_t)(((qq + 1) ^ sw) << 4));
                        } else if (mul_vec && full16) {
                            q0 = cm[0];
                            q1 = cm[1];
                        } else {
                            float m[16];
                            load16_bf16(mul + opix * d.mul_cstride + d.mul_coff + co, false, nvalid, m);
                            q0 = pack8(m);
                            q1 = pack8(m + 8);
                        }
                        float m[8];
                        unpack8(q0, m);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = m[i] * fmaf(tanh_approx(0.5f * v[i]), 0.5f, 0.5f);
                        unpack8(q1, m);
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[8 + i] = m[i] * fmaf(tanh_approx(0.5f * v[8 + i]), 0.5f, 0.5f);
                        if (res_direct) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] += r[i];
                        }
                    } else if (kEpi == RGBD_EPI_LINEAR) {
                        if (res_direct) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] += r[i];
                        }
                        if (d.act == RGBD_ACT_GELU) {      // exact GELU (the Swin MLPs): a uniform branch, off the hot layers' path
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752f));
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = act_fn(v[i], slope);
                        }
                    } else {  // RGBD_EPI_BILERP
                        const int64_t rbase = (int64_t)tc.n * d.res_H * d.res_W;
                        const int cc = d.res_coff + co;
                        float a00[16], a01[16], a10[16], a11[16];
                        load16_bf16(res + (rbase + (int64_t)by0 * d.res_W + bx0) * d.res_cstride + cc, res_vec, nvalid, a00);
                        load16_bf16(res + (rbase + (int64_t)by0 * d.res_W + bx1) * d.res_cstride + cc, res_vec, nvalid, a01);
                        load16_bf16(res + (rbase + (int64_t)by1 * d.res_W + bx0) * d.res_cstride + cc, res_vec, nvalid, a10);
                        load16_bf16(res + (rbase + (int64_t)by1 * d.res_W + bx1) * d.res_cstride + cc, res_vec, nvalid, a11);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float up = (1.f - ly) * ((1.f - lx) * a00[i] + lx * a01[i]) +
                                             ly * ((1.f - lx) * a10[i] + lx * a11[i]);
                            v[i] = act_fn(v[i] + up, slope);
                        }
                    }
                    if (staged) {
                        if constexpr (sizeof(TOut) == 2) {
                            const uint32_t rowaddr = stg + (uint32_t)(lane * 64);
                            const int f = (lane >> 1) & 3, h2 = (c & 1) * 2;
                            sts128(rowaddr + (uint32_t)(((h2) ^ f) << 4), pack8(v));
                            sts128(rowaddr + (uint32_t)(((h2 + 1) ^ f) << 4), pack8(v + 8));
                        }
                    } else {
                        store16<TOut>(y + opix * d.y_cstride + d.y_coff + co, y_vec, nvalid, v);
                        if (y2) store16<TOut>(y2 + opix * d.y2_cstride + d.y2_coff + co, y2_vec, nvalid, v);
                    }
                }
            };
            if constexpr (kGate || kEpi == RGBD_EPI_BILERP || kEpi == RGBD_EPI_SHUFFLE2) {
                // these epilogues keep their operands in registers too: one accumulator chunk in flight (no spills)
                for (int c = 2 * cbfirst; c < nchunks; c = chunk_after(c)) {
                    tmem_ld16_issue(trow + (uint32_t)(c * 16), r0);
                    tmem_ld_wait(r0);
                    process(c, r0);
                }
            } else {
                int cb = cbfirst;
                if (cb < nblk) {
                    tmem_ld16_issue(trow + (uint32_t)(cb * 32), r0);
                    if (2 * cb + 1 < nchunks) tmem_ld16_issue(trow + (uint32_t)(cb * 32 + 16), r1);
                }
                for (; cb < nblk; cb += cbstep) {
                    const bool has1 = 2 * cb + 1 < nchunks;
                    const long long e0 = tr ? clock64() : 0;
                    tmem_ld_wait(r0);
                    const long long e1 = tr ? clock64() : 0;
                    process(2 * cb, r0);
                    if (has1) {
                        tmem_ld_wait(r1);
                        process(2 * cb + 1, r1);
                    }
                    const long long e2 = tr ? clock64() : 0;
                    // the next block's accumulator columns travel during the store phase
                    const int nb = cb + cbstep;
                    if (nb < nblk) {
                        tmem_ld16_issue(trow + (uint32_t)(nb * 32), r0);
                        if (2 * nb + 1 < nchunks) tmem_ld16_issue(trow + (uint32_t)(nb * 32 + 16), r1);
                    }
                    if (staged) {
                        if constexpr (sizeof(TOut) == 2) {
                            __syncwarp();
                            const int q = lane & 3, rsub = lane >> 2;
                            const int chb = tc.co0 + cb * 32 + q * 8;
                            const bool pv = q * 8 < (has1 ? 32 : 16) && chb + 8 <= d.Cout;
    #pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = rsub + 8 * i;
                                const int pc = __shfl_sync(0xffffffffu, pix_code, r);
                                if (pv && pc >= 0) {
                                    const uint4 val = lds128(stg + (uint32_t)(r * 64 + ((q ^ ((r >> 1) & 3)) << 4)));
                                    *reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(y) + (int64_t)pc * d.y_cstride + d.y_coff + chb) = val;
                                    if (y2)
                                        *reinterpret_cast<uint4 *>(reinterpret_cast<bf16 *>(y2) + (int64_t)pc * d.y2_cstride + d.y2_coff + chb) = val;
                                }
                            }
                            __syncwarp();
                        }
                    }
                    if (tr) {
                        const long long e3 = clock64();
                        et[0] += e1 - e0;
                        et[1] += e2 - e1;
                        et[2] += e3 - e2;
                    }
                }
            }
            // all of this warp's tcgen05.ld have completed: hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(acc, j));
            if (mul_staged && lane == 0) mbar_arrive(mul_empty(li & 1));   // (__syncwarp above: every lane has read its row)
        }
        if (tr) {
            p.dbg[7] = w_full;
            p.dbg[8] = clock64() - t_begin;
            for (int i = 0; i < 3; ++i) p.dbg[9 + i] = et[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------- host side
inline int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

// identity tiles for the residual-through-MMA trick: E[jb][n][k] = (n == 64 * jb + k)
__device__ __nv_bfloat16 g_identity[2 * 128 * kBlockK];
__global__ void identity_init_kernel() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * 128 * kBlockK) {
        const int k = i % kBlockK, n = (i / kBlockK) % 128, jb = i / (kBlockK * 128);
        g_identity[i] = __float2bfloat16_rn(n == jb * kBlockK + k ? 1.f : 0.f);
    }
}
const void *identity_ptr() {
    static void *ptr[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_init_mutex);
    if (!ptr[dev]) {
        void *q = nullptr;
        if (cudaGetSymbolAddress(&q, g_identity) != cudaSuccess) return nullptr;
        identity_init_kernel<<<(2 * 128 * kBlockK + 255) / 256, 256>>>();
        if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
        ptr[dev] = q;
    }
    return ptr[dev];
}

}  // namespace

struct rgbd_conv_tc_plan {
    HParams p;
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
    RGBD_CHECK_ARG(d->cout_pad <= kMaxBias, "cout_pad too large");
    RGBD_CHECK_ARG(d->sched_ws != nullptr && ((uintptr_t)d->sched_ws & 3) == 0,
                   "sched_ws: the caller provides a zero-initialised int32 in device memory per plan (tile counter)");
    for (int t = 0; t < d->ntaps; ++t)
        RGBD_CHECK_ARG(d->w_image_stride == 0 || d->w_image_stride > d->wtap[t], "w_image_stride must cover every tap");
    RGBD_CHECK_ARG((int64_t)d->N * d->Ho * d->Wo < 2147483647LL, "too many output pixels");

    rgbd_conv_tc_plan *pl = new (std::nothrow) rgbd_conv_tc_plan();
    if (!pl) {
        rgbd_set_error("conv_tc: out of host memory");
        return RGBD_E_INVALID;
    }
    HParams &p = pl->p;
    p.d = *d;
    p.dbg = nullptr;
#ifdef RGBD_TIMING_PROBES
    // development builds only (RGBD_BUILD_DEFINES=-DRGBD_TIMING_PROBES): cycle counters of CTA 0 go to this device pointer
    if (const char *e = getenv("RGBD_TC_TRACE")) p.dbg = (long long *)strtoull(e, nullptr, 0);
#endif

    // ---- tap groups: taps that read the same input lattice (parity class when i_step == 2) ----
    const int st = d->i_step;
    int tq_y[RGBD_MAX_TAPS], tq_x[RGBD_MAX_TAPS], tg[RGBD_MAX_TAPS];
    for (int t = 0; t < d->ntaps; ++t) {
        if (st == 1) {
            tg[t] = 0;
            tq_y[t] = d->dy[t];
            tq_x[t] = d->dx[t];
        } else {
            const int qy = floordiv2(d->dy[t]), qx = floordiv2(d->dx[t]);
            tg[t] = (d->dy[t] - 2 * qy) * 2 + (d->dx[t] - 2 * qx);
            tq_y[t] = qy;
            tq_x[t] = qx;
        }
    }
    int order[RGBD_MAX_TAPS], nord = 0;
    int ext_x = 1, ext_y = 1;
    p.ngroups = 0;
    for (int g = 0; g < 4; ++g) {
        int qy0 = 127, qy1 = -127, qx0 = 127, qx1 = -127, cnt = 0;
        for (int t = 0; t < d->ntaps; ++t)
            if (tg[t] == g) {
                qy0 = tq_y[t] < qy0 ? tq_y[t] : qy0;
                qy1 = tq_y[t] > qy1 ? tq_y[t] : qy1;
                qx0 = tq_x[t] < qx0 ? tq_x[t] : qx0;
                qx1 = tq_x[t] > qx1 ? tq_x[t] : qx1;
                ++cnt;
            }
        if (!cnt) continue;
        const int gi = p.ngroups++;
        p.g_map[gi] = (int8_t)g;
        p.g_qy[gi] = (int8_t)qy0;
        p.g_qx[gi] = (int8_t)qx0;
        p.g_first[gi] = (int8_t)nord;
        for (int t = 0; t < d->ntaps; ++t)
            if (tg[t] == g) order[nord++] = t;
        ext_x = (qx1 - qx0 + 1) > ext_x ? (qx1 - qx0 + 1) : ext_x;
        ext_y = (qy1 - qy0 + 1) > ext_y ? (qy1 - qy0 + 1) : ext_y;
    }
    p.g_first[p.ngroups] = (int8_t)nord;
    for (int g = p.ngroups + 1; g <= kMaxGroups; ++g) p.g_first[g] = (int8_t)nord;

    p.kblocks = (d->Cin + kBlockK - 1) / kBlockK;
    // N tile <= 128: two accumulators (M tiles) x two buffers fit the 512 TMEM columns
    const int ntiles = (d->cout_pad + 127) / 128;
    p.BN = ((d->cout_pad + ntiles - 1) / ntiles + 15) / 16 * 16;
    p.n_ntiles = (d->cout_pad + p.BN - 1) / p.BN;
    p.b_bytes = p.BN * 128;

    // residual through the tensor core: plain residual adds on the full output lattice
    const bool res_mma = d->res != nullptr && d->epi == RGBD_EPI_LINEAR && d->o_step == 1 &&
                         ((d->res_cstride | d->res_coff) & 7) == 0 && ((uintptr_t)d->res & 15) == 0;
    p.res_blocks = res_mma ? (p.BN + 63) / 64 : 0;

    // ---- super-tile geometry: TW useful columns, R rows, pitch P = TW + ext_x - 1, R * P <= 256 ----
    // cost model (cycles per useful output site): MMA issue floor vs L2 -> shared-memory fetch; it does
    // not look at the batch size, so the geometry (and with it nothing numerically relevant) is the
    // same for every N
    double best = -1;
    const double cin_frac = (double)d->Cin / (double)(p.kblocks * kBlockK);
    for (int TW = 1; TW <= d->Ws && TW + ext_x - 1 <= 256; ++TW) {
        const int P = TW + ext_x - 1;
        for (int MT = 1; MT <= 2; ++MT) {
            int R = MT * 128 / P;
            if (R > d->Hs) R = d->Hs;
            if (R < 1) continue;
            if (MT == 2 && R * P <= 128) continue;
            const int box_rows = R + ext_y - 1;
            if (box_rows > 256) continue;
            long halo_rows = (long)box_rows * P + ext_x - 1;
            const long need_rows = (long)MT * 128 + (long)(ext_y - 1) * P + ext_x - 1;
            if (need_rows > halo_rows) halo_rows = need_rows;
            const long a_bytes = (halo_rows * 128 + 1023) / 1024 * 1024;
            if (a_bytes > kMaxAStage) continue;
            const double tiles = (double)((d->Ws + TW - 1) / TW) * (double)((d->Hs + R - 1) / R);
            const double mma = (double)MT * (d->ntaps * p.kblocks + p.res_blocks) * 4.0 * (p.BN / 2.0);
            const double l2 = ((double)p.ngroups * p.kblocks * box_rows * P * 128.0 * cin_frac +
                               (double)d->ntaps * p.kblocks * p.b_bytes + (double)p.res_blocks * (R * P * 128.0 + p.b_bytes)) /
                              42.0;
            // + a charge per image-row segment of the tile (TMA gathers and the epilogue's stores like long
            // contiguous runs; this also breaks the tie for 1x1 filters towards wide tiles)
            const double per_tile = (mma > l2 ? mma : l2) + 0.15 * (mma < l2 ? mma : l2) + 1500.0 + 40.0 * box_rows;
            const double cost = tiles * per_tile;
            if (best < 0 || cost < best) {
                best = cost;
                p.TW = TW;
                p.R = R;
                p.P = P;
                p.MT = MT;
                p.box_rows = box_rows;
            }
        }
    }
    if (best < 0) {
        rgbd_set_error("conv_tc: no tile geometry for Hs %d Ws %d ext %d x %d", d->Hs, d->Ws, ext_y, ext_x);
        delete pl;
        return RGBD_E_INVALID;
    }
    p.tiles_x = (d->Ws + p.TW - 1) / p.TW;
    p.tiles_y = (d->Hs + p.R - 1) / p.R;
    {
        // the stage must hold both the halo tile and the 128 * MT rows the MMAs address (+ tap shift)
        long rows = (long)p.box_rows * p.P + ext_x - 1;
        const long need = (long)p.MT * 128 + (long)(ext_y - 1) * p.P + ext_x - 1;
        if (need > rows) rows = need;
        p.a_bytes = (int)((rows * 128 + 1023) / 1024 * 1024);
    }
    p.a_tx = p.box_rows * p.P * 128;
    p.r_tx = p.R * p.P * 128;
    for (int i = 0; i < nord; ++i) {
        const int t = order[i];
        int gi = 0;
        while (p.g_map[gi] != tg[t]) ++gi;
        p.t_shift[i] = (int16_t)((tq_y[t] - p.g_qy[gi]) * p.P + (tq_x[t] - p.g_qx[gi]));
        p.t_w[i] = d->wtap[t];
    }
    // small-K layers: the epilogue (not the MMAs) paces the kernel -> coalesce its stores through shared memory
    // (measured on B200: once the activation is branch-free these layers are HBM-bound either way, so the
    // transpose path is opt-in: RGBD_TC_STAGE=1)
    static const bool env_stage = getenv("RGBD_TC_STAGE") != nullptr;
    p.stage_epi = (env_stage && d->ntaps * p.kblocks <= 8 && d->y_dtype == RGBD_DT_BF16) ? 1 : 0;
    // gate operand through TMA: only worth its shared memory when the layer is epilogue / memory bound
    p.mul_blocks = 0;
    p.mul_bytes = 0;
    const bool mul_tma = d->epi == RGBD_EPI_GATE && d->mul != nullptr && d->o_step == 1 && d->y_dtype == RGBD_DT_BF16 &&
                         ((d->mul_cstride | d->mul_coff) & 7) == 0 && ((uintptr_t)d->mul & 15) == 0 &&
                         (d->Cout & 15) == 0 && d->ntaps * p.kblocks <= 8;
    if (mul_tma) {
        p.mul_blocks = (p.BN + 63) / 64;
        p.mul_bytes = p.MT * 128 * 128;
        // needs room for two A and two weight stages beside the double-buffered tiles; else the epilogue loads it
        if (220 * 1024 - 1024 - 2 * p.mul_blocks * p.mul_bytes < 2 * p.a_bytes + 2 * p.b_bytes) p.mul_blocks = p.mul_bytes = 0;
    }
    // (these few small-K layers may use the whole SM: 220 KB in total, double-buffered gate tiles included)
    const int kBudget = p.mul_blocks ? 220 * 1024 - 1024 - 2 * p.mul_blocks * p.mul_bytes
                                     : kRingBudget - (p.stage_epi ? kStageBytes : 0);
    // Weight stages hold up to 3 consecutive taps (<= 36 KB): one barrier round trip + commit per stage
    // costs an issuing thread ~600 cycles, which 4 MMAs per accumulator do not cover.
    int max_group_taps = 1;
    for (int g = 0; g < p.ngroups; ++g) {
        const int n = p.g_first[g + 1] - p.g_first[g];
        max_group_taps = n > max_group_taps ? n : max_group_taps;
    }
    p.tps = 36 * 1024 / p.b_bytes;
    if (p.tps > 3) p.tps = 3;
    if (p.tps > max_group_taps) p.tps = max_group_taps;
    if (p.tps < 1) p.tps = 1;
    // ring depths: 2 A stages (more when affordable), at least 2 weight stages
    p.nA = 2;
    while (p.tps > 1 && (kBudget - p.nA * p.a_bytes) / (p.tps * p.b_bytes) < 2) --p.tps;
    const int b_stage = p.tps * p.b_bytes;
    p.nB = (kBudget - p.nA * p.a_bytes) / b_stage;
    if (p.nB > kMaxB) p.nB = kMaxB;
    if (p.nB < 2) {
        rgbd_set_error("conv_tc: ring budget too small (a_bytes %d b_bytes %d)", p.a_bytes, p.b_bytes);
        delete pl;
        return RGBD_E_INVALID;
    }
    const int a_iters = p.kblocks * p.ngroups + p.res_blocks;
    const int nb_keep = p.tps > 1 ? 2 : 5;
    while (p.nA < kMaxA && p.nA < a_iters + 1 &&
           (p.nA + 1) * p.a_bytes + (p.nB < nb_keep ? p.nB : nb_keep) * b_stage <= kBudget) {
        ++p.nA;
        const int nb = (kBudget - p.nA * p.a_bytes) / b_stage;
        if (nb < p.nB) p.nB = nb;
    }
    const long total = (long)d->N * p.tiles_x * p.tiles_y * p.n_ntiles;
    RGBD_CHECK_ARG(total < 2147483647L, "too many tiles");
    p.total_tiles = (int)total;
    pl->smem = (size_t)p.nA * p.a_bytes + (size_t)p.nB * p.tps * p.b_bytes + 1024 /*align*/ +
               8 * (2 * kMaxA + 2 * kMaxB + 12 + 2 * kTileQ) + 16 + 128 + 4 * (size_t)d->cout_pad + 64 + (p.stage_epi ? kStageBytes : 0) +
               (p.mul_blocks ? 1024 + 2 * (size_t)p.mul_blocks * p.mul_bytes : 0);
    if (pl->smem < (size_t)kMinSmem) pl->smem = kMinSmem;
    // per-device state (one process may hold nets on several GPUs; function attributes are per device)
    static int sms_of[64] = {};
    static bool configured_on[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    int num_sms;
    {
        std::lock_guard<std::mutex> lock(g_init_mutex);
        if (sms_of[dev] == 0) {
            if (cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms_of[dev] <= 0)
                sms_of[dev] = 148;
        }
        num_sms = sms_of[dev];
    }
    pl->grid = dim3((unsigned)(p.total_tiles < num_sms ? p.total_tiles : num_sms));
    pl->out_f32 = d->y_dtype == RGBD_DT_F32;
    pl->epi = d->epi;

    // A tensor maps: one per input parity class (i_step == 2) or a single one
    const char *xb = reinterpret_cast<const char *>(d->x);
    for (int ry = 0; ry < st; ++ry)
        for (int rx = 0; rx < st; ++rx) {
            const int Hsub = (d->H - ry + st - 1) / st, Wsub = (d->W - rx + st - 1) / st;
            if (Hsub <= 0 || Wsub <= 0) continue;
            cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)Wsub, (cuuint64_t)Hsub, (cuuint64_t)d->N};
            cuuint64_t strides[3] = {(cuuint64_t)st * d->x_cstride * 2, (cuuint64_t)st * d->W * d->x_cstride * 2,
                                     (cuuint64_t)d->H * d->W * d->x_cstride * 2};
            cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)p.P, (cuuint32_t)p.box_rows, 1};
            const void *basep = xb + ((int64_t)(ry * d->W + rx) * d->x_cstride + d->x_coff) * 2;
            rc = encode_map(&p.amap[ry * st + rx], basep, 4, dims, strides, box);
            if (rc) {
                delete pl;
                return rc;
            }
        }
    for (int i = st * st; i < 4; ++i) p.amap[i] = p.amap[0];
    // B tensor map over the packed weights [taps_total][cout_pad][cin_pad] (taps_total >= max wtap + 1)
    int max_tap = 0;
    for (int t = 0; t < d->ntaps; ++t) max_tap = d->wtap[t] > max_tap ? d->wtap[t] : max_tap;
    {
        // per-image weight sets (w_image_stride taps apart) are further slices of the same tap dimension
        cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)d->cout_pad,
                              (cuuint64_t)(max_tap + 1) + (cuuint64_t)d->w_image_stride * (cuuint64_t)(d->N - 1)};
        cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * d->cout_pad * 2};
        cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)p.BN, 1};
        rc = encode_map(&p.bmap, d->w, 3, dims, strides, box);
        if (rc) {
            delete pl;
            return rc;
        }
    }
    p.rmap = p.amap[0];
    p.emap = p.bmap;
    p.mmap = p.amap[0];
    if (p.mul_blocks) {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->mul_cstride * 2, (cuuint64_t)d->Wo * d->mul_cstride * 2,
                                 (cuuint64_t)d->Ho * d->Wo * d->mul_cstride * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)p.P, (cuuint32_t)p.R, 1};
        rc = encode_map(&p.mmap, reinterpret_cast<const char *>(d->mul) + (int64_t)d->mul_coff * 2, 4, dims, strides, box);
        if (rc) {
            delete pl;
            return rc;
        }
    }
    if (p.res_blocks) {
        cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)d->res_cstride * 2, (cuuint64_t)d->Wo * d->res_cstride * 2,
                                 (cuuint64_t)d->Ho * d->Wo * d->res_cstride * 2};
        cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)p.P, (cuuint32_t)p.R, 1};
        rc = encode_map(&p.rmap, reinterpret_cast<const char *>(d->res) + (int64_t)d->res_coff * 2, 4, dims, strides, box);
        if (!rc) {
            const void *idp = identity_ptr();
            if (!idp) {
                rgbd_set_error("conv_tc: identity tile unavailable");
                rc = RGBD_E_CUDA;
            } else {
                cuuint64_t edims[3] = {(cuuint64_t)kBlockK, 128, 2};
                cuuint64_t estr[2] = {(cuuint64_t)kBlockK * 2, (cuuint64_t)kBlockK * 128 * 2};
                cuuint32_t ebox[3] = {(cuuint32_t)kBlockK, (cuuint32_t)p.BN, 1};
                rc = encode_map(&p.emap, idp, 3, edims, estr, ebox);
            }
        }
        if (rc) {
            delete pl;
            return rc;
        }
    }
    std::unique_lock<std::mutex> init_lock(g_init_mutex);
    if (!configured_on[dev]) {
        const int cap = 227 * 1024;
        cudaFuncSetAttribute(conv_halo_kernel<__nv_bfloat16, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<__nv_bfloat16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<__nv_bfloat16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<__nv_bfloat16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        cudaFuncSetAttribute(conv_halo_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
        configured_on[dev] = true;
    }
    init_lock.unlock();
    if (getenv("RGBD_TC_VERBOSE"))
        fprintf(stderr, "conv_halo: %dx%d taps %d Cin %d Cout %d | TW %d R %d P %d MT %d box_rows %d groups %d kb %d BN %d nA %d nB %d tps %d a %d b %d res %d tiles %d smem %zu\n",
                d->Hs, d->Ws, d->ntaps, d->Cin, d->Cout, p.TW, p.R, p.P, p.MT, p.box_rows, p.ngroups, p.kblocks, p.BN, p.nA, p.nB, p.tps,
                p.a_bytes, p.b_bytes, p.res_blocks, p.total_tiles, pl->smem);
    *out = pl;
    return RGBD_OK;
}

extern "C" int rgbd_conv_tc_run(const rgbd_conv_tc_plan *pl, void *stream) {
    RGBD_CHECK_ARG(pl != nullptr, "null plan");
    cudaStream_t st = (cudaStream_t)stream;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = pl->grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = pl->smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = getenv("RGBD_TC_NOPDL") == nullptr;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
#define RGBD_TC_LAUNCH(T, E) cudaLaunchKernelEx(&cfg, conv_halo_kernel<T, E>, pl->p)
    if (pl->out_f32) {
        if (pl->epi == RGBD_EPI_GATE) RGBD_TC_LAUNCH(float, 1);
        else if (pl->epi == RGBD_EPI_BILERP) RGBD_TC_LAUNCH(float, 2);
        else if (pl->epi == RGBD_EPI_SHUFFLE2) RGBD_TC_LAUNCH(float, 3);
        else RGBD_TC_LAUNCH(float, 0);
    } else {
        if (pl->epi == RGBD_EPI_GATE) RGBD_TC_LAUNCH(__nv_bfloat16, 1);
        else if (pl->epi == RGBD_EPI_BILERP) RGBD_TC_LAUNCH(__nv_bfloat16, 2);
        else if (pl->epi == RGBD_EPI_SHUFFLE2) RGBD_TC_LAUNCH(__nv_bfloat16, 3);
        else RGBD_TC_LAUNCH(__nv_bfloat16, 0);
    }
#undef RGBD_TC_LAUNCH
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" void rgbd_conv_tc_plan_destroy(rgbd_conv_tc_plan *pl) { delete pl; }
