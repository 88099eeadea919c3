// GPU rANS coder, byte-identical to compressai.ans
// (reference: CompressAI/compressai/cpp_exts/rans/rans_interface.cpp:99-351 and
//  CompressAI/third_party/ryg_rans/rans64.h:59-142).
//
// A rANS stream is one serial dependency chain on the 64-bit state x, so the unit of parallelism
// is the stream: ONE WARP PER STREAM, all 32 lanes carrying x redundantly (uniform control flow).
// Everything that does not depend on x is taken off the chain:
//   encoder: per-symbol records (exact reciprocal of the range, bias, shift) come from a table
//            that is pre-computed once per CDF table set (rgbd_rans_tables.enc_rec); the lanes
//            fetch the records of the NEXT batch of 32 symbols from L2 while the chain runs, stage
//            them in shared memory, and the chain reads them back as prefetched broadcasts;
//   decoder: lane l tests the hypothesis "the symbol is l": it holds cdf[l], cdf[l+1] of the
//            symbol's table (prefetched one symbol ahead, the address does not depend on x),
//            pre-computes the state that hypothesis leads to, and a warp OR-reduction (REDUX)
//            selects the one lane that is right.  Tables with more than 33 entries first narrow
//            the window with a 32-ary ballot search.
// The compacted uint16 CDF tables (27 256 entries for the 64 Gaussian tables) live in shared
// memory for the whole decode kernel.
#include "common.cuh"
#include <cstdlib>
#include <mutex>

namespace {

constexpr int kWarpsPerBlock = 4;   // fewer, fatter blocks: each holds a 57 KB table copy that can fence a conv CTA off its SM
constexpr uint64_t kRansL = 1ull << 31;
constexpr int kMaxTables = 256;
constexpr uint32_t kFull = 0xffffffffu;

struct TableSmem {
    int32_t base[kMaxTables];
    int32_t length[kMaxTables];
    int32_t offset[kMaxTables];
};

__device__ __forceinline__ void load_meta(const rgbd_rans_tables &t, TableSmem &m) {
    for (int i = threadIdx.x; i < t.n_tables; i += blockDim.x) {
        m.base[i] = t.base[i];
        m.length[i] = t.length[i];
        m.offset[i] = t.offset[i];
    }
}

// ---------------------------------------------------------------------------------------
// Encoder
// ---------------------------------------------------------------------------------------
// One record per CDF bin, built on the host (entropy_models.py) with exact integer arithmetic:
//   range >= 2: rcp = ceil(2^(s+63) / range), s = ceil(log2 range), shift = s - 1, bias = start
//               => mulhi(x, rcp) >> shift == x / range for every x < 2^63   (rans64.h:211-247)
//   range == 1: rcp = 2^64 - 1, shift = 0, bias = start + 65535  (q = x - 1; rans64.h:191-210)
struct __align__(16) EncRec {
    uint64_t rcp;
    uint32_t bias;
    uint32_t packed;  // range [0,16) | shift [16,22)
};

struct EncState {
    uint64_t x;
    int64_t cur;  // next free word index (exclusive), counts down
    bool overflow;
};

__device__ __forceinline__ void enc_emit(EncState &st, uint32_t *out_base, int lane) {
    if (st.cur <= 0) {
        st.overflow = true;
    } else {
        st.cur -= 1;
        if (lane == 0) out_base[st.cur] = (uint32_t)st.x;
    }
    st.x >>= 32;
}

__device__ __forceinline__ void enc_put_bits4(EncState &st, uint32_t val, uint32_t *out_base, int lane) {
    // rans_interface.cpp:60-78 with nbits = 4: freq = 1 << 12, x_max = 2^15 * 2^32 * 2^12
    if (st.x >= (1ull << 59)) enc_emit(st, out_base, lane);
    st.x = (st.x << 4) | val;
}

__device__ __forceinline__ void enc_put(EncState &st, const EncRec &r, uint32_t *out_base, int lane) {
    const uint32_t range = r.packed & 0xFFFFu;
    // Rans64EncPut (rans64.h:77-93): x_max = ((L >> 16) << 32) * range
    if (st.x >= ((uint64_t)range << 47)) enc_emit(st, out_base, lane);
    const uint64_t q = __umul64hi(st.x, r.rcp) >> ((r.packed >> 16) & 63u);   // == x / range, exact
    st.x = st.x + r.bias + q * (uint64_t)(65536u - range);
}

// what a lane knows about "its" symbol of a batch before the record arrives
struct EncPrep {
    int32_t rec_index;  // bin index into enc_rec (>= 0), -1 for lanes beyond the stream start
    uint32_t raw;       // escape payload
    uint32_t esc;
};

__device__ __forceinline__ EncPrep enc_prepare(int32_t i, int32_t symv, int32_t ti, const TableSmem &meta) {
    EncPrep p;
    p.rec_index = -1;
    p.raw = 0;
    p.esc = 0;
    if (i >= 0) {
        const int32_t top = meta.length[ti] - 2;
        int32_t v = symv - meta.offset[ti];
        if (v < 0) {                       // rans_interface.cpp:125-131
            p.raw = (uint32_t)(-2 * v - 1);
            v = top;
        } else if (v >= top) {
            p.raw = (uint32_t)(2 * (v - top));
            v = top;
        }
        p.esc = (v == top) ? 1u : 0u;
        p.rec_index = meta.base[ti] + v;
    }
    return p;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rans_encode_kernel(const int32_t *__restrict__ sym, const uint8_t *__restrict__ idx,
                   int64_t stream_stride, int32_t n_sym, int32_t n_streams, rgbd_rans_tables t,
                   uint32_t *__restrict__ out, int64_t cap_words, int32_t *__restrict__ nwords) {
    __shared__ TableSmem meta;
    __shared__ EncRec s_rec[kWarpsPerBlock][32];
    __shared__ uint32_t s_raw[kWarpsPerBlock][32];
    load_meta(t, meta);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int s = blockIdx.x * kWarpsPerBlock + wib;
    if (s >= n_streams) return;

    const int32_t *my_sym = sym + (int64_t)s * stream_stride;
    const uint8_t *my_idx = idx + (int64_t)s * stream_stride;
    const EncRec *__restrict__ table = reinterpret_cast<const EncRec *>(t.enc_rec);
    uint32_t *out_base = out + (int64_t)s * cap_words;
    EncRec *rec = s_rec[wib];
    uint32_t *raws = s_raw[wib];

    EncState st;
    st.x = kRansL;
    st.cur = cap_words;
    st.overflow = false;

    // Three-deep software pipeline over batches of 32 symbols (processed from the END of the stream;
    // lane j owns symbol hi-1-j so that the chain consumes lanes 0, 1, 2, ... in order):
    //   stage A (2 batches ahead): symbol + index loads
    //   stage B (1 batch ahead):  escape mapping and the record fetch from the pre-computed table
    //   stage C (current):        records -> shared memory -> the serial chain
    auto load_A = [&](int32_t hi, int32_t &sv, int32_t &iv) {
        const int32_t i = hi - 1 - lane;
        sv = 0;
        iv = 0;
        if (hi > 0 && i >= 0) {
            sv = my_sym[i];
            iv = my_idx[i];
        }
    };
    int32_t a_sym, a_idx;          // stage A registers
    EncPrep b_prep;                // stage B registers
    EncRec b_rec;
    b_rec.rcp = ~0ull; b_rec.bias = 0; b_rec.packed = 1;

    load_A(n_sym, a_sym, a_idx);
    b_prep = enc_prepare(n_sym - 1 - lane, a_sym, a_idx, meta);
    if (b_prep.rec_index >= 0) b_rec = table[b_prep.rec_index];
    load_A(n_sym - 32, a_sym, a_idx);

    for (int32_t hi = n_sym; hi > 0; hi -= 32) {
        // stage C inputs
        const EncPrep c_prep = b_prep;
        const EncRec c_rec = b_rec;
        // advance stage B with the symbols loaded one iteration ago, then stage A
        b_prep = enc_prepare(hi - 32 > 0 ? hi - 33 - lane : -1, a_sym, a_idx, meta);
        if (b_prep.rec_index >= 0) b_rec = table[b_prep.rec_index];
        load_A(hi - 64, a_sym, a_idx);

        rec[lane] = c_rec;
        raws[lane] = c_prep.raw;
        const uint32_t esc_mask = __ballot_sync(kFull, c_prep.esc != 0);
        __syncwarp();
        const int nvalid = hi < 32 ? hi : 32;
        EncRec cur = rec[0];
        if (esc_mask == 0) {
            for (int j = 0; j < nvalid; ++j) {
                const EncRec nxt = rec[(j + 1) & 31];
                enc_put(st, cur, out_base, lane);
                cur = nxt;
            }
        } else {
            for (int j = 0; j < nvalid; ++j) {
                const EncRec nxt = rec[(j + 1) & 31];
                if ((esc_mask >> j) & 1u) {
                    // escape payload, emitted in REVERSE of the forward record order
                    // [bin][count: 15 * (n / 15), n % 15][nibbles LSB first]  (rans_interface.cpp:139-163)
                    const uint32_t raw = raws[j];
                    int nnib = 0;
                    while (nnib < 8 && (raw >> (nnib * 4)) != 0) ++nnib;
                    for (int k = nnib - 1; k >= 0; --k) enc_put_bits4(st, (raw >> (k * 4)) & 15u, out_base, lane);
                    enc_put_bits4(st, (uint32_t)(nnib % 15), out_base, lane);
                    for (int k = 0; k < nnib / 15; ++k) enc_put_bits4(st, 15u, out_base, lane);
                }
                enc_put(st, cur, out_base, lane);
                cur = nxt;
            }
        }
        __syncwarp();
    }
    // Rans64EncFlush (rans64.h:96-103)
    if (st.cur < 2) {
        st.overflow = true;
    } else {
        st.cur -= 2;
        if (lane == 0) {
            out_base[st.cur] = (uint32_t)st.x;
            out_base[st.cur + 1] = (uint32_t)(st.x >> 32);
        }
    }
    if (lane == 0) nwords[s] = st.overflow ? -1 : (int32_t)(cap_words - st.cur);
}

// ---------------------------------------------------------------------------------------
// Decoder
// ---------------------------------------------------------------------------------------
struct WordFeed {
    const uint32_t *words;
    int64_t len;     // words in this stream
    int64_t base;    // stream position of lane 0 of `cur`
    uint32_t cur, nxt;
    __device__ __forceinline__ uint32_t fetch(int64_t p) const { return p < len ? words[p] : 0u; }
    __device__ __forceinline__ void init(const uint32_t *w, int64_t l, int64_t pos, int lane) {
        words = w; len = l; base = pos;
        cur = fetch(base + lane);
        nxt = fetch(base + 32 + lane);
    }
    // word at absolute position p (monotonically increasing, p < base + 64)
    __device__ __forceinline__ uint32_t take(int64_t p, int lane) {
        if (p - base >= 32) {
            cur = nxt;
            base += 32;
            nxt = fetch(base + 32 + lane);
        }
        return __shfl_sync(kFull, cur, (int)(p - base));
    }
};

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
    return (uint32_t)v;
}

struct DecState {
    uint64_t x;
    int64_t pos;
    uint32_t wnext;  // the word the next renormalisation consumes (pre-fetched)
};

__device__ __forceinline__ void dec_renorm(DecState &d, WordFeed &feed, int lane) {
    if (d.x < kRansL) {
        d.x = (d.x << 32) | d.wnext;
        d.pos += 1;
        d.wnext = feed.take(d.pos, lane);
    }
}

// bypass nibbles of an escape symbol (rans_interface.cpp:80-96, 323-344)
__device__ __forceinline__ int32_t dec_escape(DecState &d, WordFeed &feed, int lane, int32_t top) {
    auto get4 = [&]() -> int32_t {
        const uint32_t v = (uint32_t)(d.x & 15u);
        d.x >>= 4;
        dec_renorm(d, feed, lane);
        return (int32_t)v;
    };
    int32_t dgt = get4();
    int32_t nnib = dgt;
    // valid streams carry at most 8 payload nibbles, so the unary count never continues; the bounds
    // only stop a corrupt stream from spinning
    for (int guard = 0; dgt == 15 && guard < 4; ++guard) {
        dgt = get4();
        nnib += dgt;
    }
    if (nnib > 64) nnib = 64;
    int32_t raw = 0;
    for (int k = 0; k < nnib; ++k) {
        dgt = get4();
        if (k < 8) raw |= dgt << (k * 4);
    }
    int32_t value = raw >> 1;
    if (raw & 1) value = -value - 1;
    else value += top;
    return value;
}

// Which streams a decode launch works on.  Stream s of the launch = (group s / per_group, member s % per_group):
//   symbols / indexes at group * group_stride + member * stream_stride + chunk_off,
//   word table slot (word_off / word_len / state index) = group * slot_group_stride + slot_base + member.
// The resumable single-stream layout is one group (per_group = n_streams, slot_base = 0); the multi-stream layout
// decodes, per coding step, `per_group` whole sub-streams of every image (fresh = 1: the state starts from the
// stream's first two words instead of state[]).
struct DecodeMap {
    int32_t per_group, slot_base, slot_group_stride, fresh;
    int64_t group_stride;
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rans_decode_kernel(const uint32_t *__restrict__ words, const int64_t *__restrict__ word_off,
                   const int64_t *__restrict__ word_len, int32_t n_streams,
                   rgbd_rans_dec_state *__restrict__ state, const uint8_t *__restrict__ idx,
                   int32_t *__restrict__ sym, int64_t stream_stride, int64_t chunk_off,
                   int32_t n_sym, rgbd_rans_tables t, DecodeMap map) {
    extern __shared__ __align__(16) uint16_t s_cdf[];
    __shared__ TableSmem meta;
    __shared__ int4 s_meta[kWarpsPerBlock][32];   // (base, length, offset, -) of the batch's tables
    for (int i = threadIdx.x; i < t.total; i += blockDim.x) s_cdf[i] = t.cdf[i];
    load_meta(t, meta);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int s = blockIdx.x * kWarpsPerBlock + wib;
    if (s >= n_streams) return;
    const int grp = s / map.per_group, mem = s - grp * map.per_group;
    const int slot = grp * map.slot_group_stride + map.slot_base + mem;
    const int64_t sym_off = (int64_t)grp * map.group_stride + (int64_t)mem * stream_stride + chunk_off;

    DecState d;
    const uint32_t *my_words = words + word_off[slot];
    if (map.fresh) {   // Rans64DecInit (rans64.h:107-115)
        d.x = (uint64_t)my_words[0] | ((uint64_t)my_words[1] << 32);
        d.pos = 2;
    } else {
        d.x = state[slot].x;
        d.pos = state[slot].pos;
    }
    WordFeed feed;
    feed.init(my_words, word_len[slot], d.pos, lane);
    d.wnext = feed.take(d.pos, lane);

    const uint8_t *my_idx = idx + sym_off;
    int32_t *my_sym = sym + sym_off;
    int4 *mrow = s_meta[wib];
    const uint32_t cdf_sa = (uint32_t)__cvta_generic_to_shared(s_cdf);
    const uint32_t lane2 = (uint32_t)lane * 2u;

    int32_t pf_idx = (lane < n_sym) ? (int32_t)my_idx[lane] : 0;  // prefetched one batch early
    for (int32_t lo_i = 0; lo_i < n_sym; lo_i += 32) {
        const int32_t i = lo_i + lane;
        const int32_t ti = pf_idx;
        if (i + 32 < n_sym) pf_idx = my_idx[i + 32];
        int4 mine = make_int4(0, 2, 0, 0);
        if (i < n_sym) mine = make_int4(meta.base[ti], meta.length[ti], meta.offset[ti], 0);
        mrow[lane] = mine;
        const uint32_t big = __ballot_sync(kFull, mine.y > 33);
        __syncwarp();
        int32_t my_out = 0;
        const int nvalid = (n_sym - lo_i) < 32 ? (n_sym - lo_i) : 32;
        int4 cur = mrow[0];
        if (big == 0) {
            // ---- fast path: every table of this batch fits one 32-lane window ----
            uint32_t c0r = lds_u16(cdf_sa + 2u * (uint32_t)cur.x + lane2);
            uint32_t c1r = lds_u16(cdf_sa + 2u * (uint32_t)cur.x + lane2 + 2u);
            for (int j = 0; j < nvalid; ++j) {
                const int4 nxt = mrow[(j + 1) & 31];
                const uint32_t n0 = lds_u16(cdf_sa + 2u * (uint32_t)nxt.x + lane2);
                const uint32_t n1 = lds_u16(cdf_sa + 2u * (uint32_t)nxt.x + lane2 + 2u);
                const int32_t L = cur.y;
                // cdf[L-1] == 65536 is stored as 0; lanes past the table never hit
                const uint32_t c0 = (lane >= L - 1) ? 65536u : c0r;
                const uint32_t c1 = (lane + 1 >= L - 1) ? 65536u : c1r;
                const uint32_t cf = (uint32_t)(d.x & 0xFFFFu);               // Rans64DecGet
                const bool hit = (c0 <= cf) && (cf < c1);
                // Rans64DecAdvance under this lane's hypothesis (rans64.h:126-142)
                const uint64_t xh = (uint64_t)(c1 - c0) * (d.x >> 16) + (uint64_t)(cf - c0);
                const uint32_t xlo = __reduce_or_sync(kFull, hit ? (uint32_t)xh : 0u);
                const uint32_t xhi = __reduce_or_sync(kFull, hit ? (uint32_t)(xh >> 32) : 0u);
                int32_t value = (int32_t)__reduce_or_sync(kFull, hit ? (uint32_t)lane : 0u);
                d.x = ((uint64_t)xhi << 32) | xlo;
                dec_renorm(d, feed, lane);
                if (value == L - 2) value = dec_escape(d, feed, lane, L - 2);
                if (lane == j) my_out = value + cur.z;
                cur = nxt;
                c0r = n0;
                c1r = n1;
            }
        } else {
            for (int j = 0; j < nvalid; ++j) {
                const int4 nxt = mrow[(j + 1) & 31];
                const int32_t b = cur.x, L = cur.y;
                const uint32_t cf = (uint32_t)(d.x & 0xFFFFu);
                // invariant cdf[lo] <= cf < cdf[hi]; cdf[L-1] == 65536 (stored as 0)
                int32_t lo = 0, hi = L - 1;
                while (hi - lo > 31) {
                    const int32_t stride = (hi - lo + 31) >> 5;
                    const int32_t p = lo + (lane + 1) * stride;
                    const bool above = (p >= hi) || ((uint32_t)s_cdf[b + p] > cf);
                    const uint32_t m = __ballot_sync(kFull, above);
                    const int f = __ffs(m) - 1;  // m != 0: lane 31 always reaches hi
                    const int32_t nhi = lo + (f + 1) * stride;
                    lo = lo + f * stride;
                    hi = nhi < hi ? nhi : hi;
                }
                const int32_t p = lo + lane;
                const uint32_t c0 = (p >= L - 1) ? 65536u : (uint32_t)s_cdf[b + p];
                const uint32_t c1 = (p + 1 >= L - 1) ? 65536u : (uint32_t)s_cdf[b + p + 1];
                const bool hit = (c0 <= cf) && (cf < c1);
                const uint64_t xh = (uint64_t)(c1 - c0) * (d.x >> 16) + (uint64_t)(cf - c0);
                const uint32_t xlo = __reduce_or_sync(kFull, hit ? (uint32_t)xh : 0u);
                const uint32_t xhi = __reduce_or_sync(kFull, hit ? (uint32_t)(xh >> 32) : 0u);
                int32_t value = (int32_t)__reduce_or_sync(kFull, hit ? (uint32_t)p : 0u);
                d.x = ((uint64_t)xhi << 32) | xlo;
                dec_renorm(d, feed, lane);
                if (value == L - 2) value = dec_escape(d, feed, lane, L - 2);
                if (lane == j) my_out = value + cur.z;
                cur = nxt;
            }
        }
        __syncwarp();
        if (i < n_sym) my_sym[i] = my_out;
    }
    if (lane == 0) {
        state[slot].x = d.x;
        state[slot].pos = d.pos;
    }
}

__global__ void rans_decode_init_kernel(const uint32_t *__restrict__ words,
                                        const int64_t *__restrict__ word_off, int32_t n_streams,
                                        rgbd_rans_dec_state *__restrict__ state) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const uint32_t *w = words + word_off[s];
    // Rans64DecInit (rans64.h:107-115)
    state[s].x = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
    state[s].pos = 2;
}

// Packs the tails of finished encoder streams into one buffer: dst = [n_total counts | words of stream 0 | stream 1 | ...].
// `dst` is normally pinned host memory (zero-copy stores): the compressed strings are on the host when the kernel ends.
__global__ void __launch_bounds__(256)
gather_streams_kernel(const uint32_t *__restrict__ a, int64_t a_cap, int32_t a_n, const uint32_t *__restrict__ b,
                      int64_t b_cap, int32_t b_n, const int32_t *__restrict__ nwords, uint32_t *__restrict__ dst,
                      int64_t dst_cap) {
    const int s = blockIdx.x, total = a_n + b_n;
    __shared__ int64_t s_off, s_all;
    if (threadIdx.x < 32) {
        int64_t before = 0, all = 0;
        for (int i = threadIdx.x; i < total; i += 32) {
            const int32_t n = nwords[i];
            const int64_t v = n > 0 ? n : 0;
            all += v;
            if (i < s) before += v;
        }
        for (int o = 16; o; o >>= 1) {
            before += __shfl_xor_sync(kFull, before, o);
            all += __shfl_xor_sync(kFull, all, o);
        }
        if (threadIdx.x == 0) {
            s_off = before;
            s_all = all;
        }
    }
    __syncthreads();
    const int32_t n = nwords[s];
    if (threadIdx.x == 0) dst[s] = (uint32_t)n;
    if (n <= 0 || s_all > dst_cap) return;   // overflowed stream, or the packed buffer is too small (the host falls back)
    const uint32_t *src = s < a_n ? a + (int64_t)s * a_cap + (a_cap - n) : b + (int64_t)(s - a_n) * b_cap + (b_cap - n);
    uint32_t *out = dst + total + s_off;
    for (int32_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = src[i];
}

size_t cdf_smem_bytes(const rgbd_rans_tables *t) {
    return (((size_t)t->total * 2 + 15) & ~(size_t)15) + 128;   // + slack for the 2-entry window prefetch
}

}  // namespace

#ifdef RGBD_TIMING_PROBES
// Timing experiments only (profiles/tools/skip_probe.py; NOT compiled into the shipped library — build with
// RGBD_BUILD_DEFINES=-DRGBD_TIMING_PROBES): RGBD_RANS_SKIP=1 drops the coder kernels, =2 replaces them by a one-warp
// kernel that only waits as long as the coder would (its latency without its shared memory / SM footprint).
__global__ void rans_sleep_kernel(long long ns) {
    const long long t0 = clock64();
    while (clock64() - t0 < ns * 19 / 10) __nanosleep(2000);
}
#endif

extern "C" int rgbd_rans_encode(const int32_t *sym, const uint8_t *idx, int64_t stream_stride,
                                int32_t n_sym, int32_t n_streams, const rgbd_rans_tables *t,
                                uint32_t *out, int64_t cap_words, int32_t *nwords, void *stream) {
    RGBD_CHECK_ARG(sym && idx && t && out && nwords, "null pointer");
    RGBD_CHECK_ARG(n_sym >= 0 && n_streams >= 0 && cap_words >= 2, "sizes");
    RGBD_CHECK_ARG(t->n_tables > 0 && t->n_tables <= kMaxTables, "n_tables must be in 1..256");
    RGBD_CHECK_ARG(t->enc_rec != nullptr, "tables have no encoder records (enc_rec)");
    if (n_streams == 0) return RGBD_OK;
#ifdef RGBD_TIMING_PROBES
    if (getenv("RGBD_RANS_SKIP")) {
        if (atoi(getenv("RGBD_RANS_SKIP")) == 2) rans_sleep_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long)n_sym * 74);
        return RGBD_OK;
    }
#endif
    const int grid = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
    rans_encode_kernel<<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        sym, idx, stream_stride, n_sym, n_streams, *t, out, cap_words, nwords);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_rans_decode_init(const uint32_t *words, const int64_t *word_off, int32_t n_streams,
                                     rgbd_rans_dec_state *state, void *stream) {
    RGBD_CHECK_ARG(words && word_off && state, "null pointer");
    if (n_streams <= 0) return RGBD_OK;
    rans_decode_init_kernel<<<(n_streams + 127) / 128, 128, 0, (cudaStream_t)stream>>>(words, word_off,
                                                                                       n_streams, state);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

static int launch_decode(const uint32_t *words, const int64_t *word_off, const int64_t *word_len, int32_t n_streams,
                         rgbd_rans_dec_state *state, const uint8_t *idx, int32_t *sym, int64_t stream_stride,
                         int64_t chunk_off, int32_t n_sym, const rgbd_rans_tables *t, const DecodeMap &map, void *stream) {
    const size_t smem = cdf_smem_bytes(t);
    if (smem > 200 * 1024) {
        rgbd_set_error("rans decode: CDF tables do not fit in shared memory");
        return RGBD_E_INVALID;
    }
#ifdef RGBD_TIMING_PROBES
    if (getenv("RGBD_RANS_SKIP")) {
        if (atoi(getenv("RGBD_RANS_SKIP")) == 2) rans_sleep_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((long long)n_sym * 150);
        return RGBD_OK;
    }
#endif
    static size_t configured[64] = {};   // per device: function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    {
        static std::mutex mu;     // several host threads may drive pipeline slots
        std::lock_guard<std::mutex> lock(mu);
        if (smem > configured[dev]) {
            cudaFuncSetAttribute(rans_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured[dev] = smem;
        }
    }
    const int grid = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
    rans_decode_kernel<<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        words, word_off, word_len, n_streams, state, idx, sym, stream_stride, chunk_off, n_sym, *t, map);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}

extern "C" int rgbd_rans_decode_chunk(const uint32_t *words, const int64_t *word_off,
                                      const int64_t *word_len, int32_t n_streams,
                                      rgbd_rans_dec_state *state, const uint8_t *idx, int32_t *sym,
                                      int64_t stream_stride, int64_t chunk_off, int32_t n_sym,
                                      const rgbd_rans_tables *t, void *stream) {
    RGBD_CHECK_ARG(words && word_off && word_len && state && idx && sym && t, "null pointer");
    RGBD_CHECK_ARG(n_sym >= 0 && n_streams >= 0, "sizes");
    RGBD_CHECK_ARG(t->n_tables > 0 && t->n_tables <= kMaxTables, "n_tables must be in 1..256");
    if (n_streams == 0 || n_sym == 0) return RGBD_OK;
    DecodeMap map;
    map.per_group = n_streams; map.slot_base = 0; map.slot_group_stride = 0; map.fresh = 0; map.group_stride = 0;
    return launch_decode(words, word_off, word_len, n_streams, state, idx, sym, stream_stride, chunk_off, n_sym, t, map, stream);
}

extern "C" int rgbd_rans_decode_streams(const uint32_t *words, const int64_t *word_off, const int64_t *word_len,
                                        int32_t n_groups, int32_t per_group, int32_t slot_base, int32_t slot_group_stride,
                                        rgbd_rans_dec_state *state, const uint8_t *idx, int32_t *sym, int64_t group_stride,
                                        int64_t stream_stride, int64_t chunk_off, int32_t n_sym, const rgbd_rans_tables *t,
                                        void *stream) {
    RGBD_CHECK_ARG(words && word_off && word_len && state && idx && sym && t, "null pointer");
    RGBD_CHECK_ARG(n_sym >= 0 && n_groups >= 0 && per_group >= 0 && slot_base >= 0 && slot_group_stride >= per_group, "sizes");
    RGBD_CHECK_ARG(t->n_tables > 0 && t->n_tables <= kMaxTables, "n_tables must be in 1..256");
    if (n_groups == 0 || per_group == 0 || n_sym == 0) return RGBD_OK;
    DecodeMap map;
    map.per_group = per_group; map.slot_base = slot_base; map.slot_group_stride = slot_group_stride; map.fresh = 1;
    map.group_stride = group_stride;
    return launch_decode(words, word_off, word_len, n_groups * per_group, state, idx, sym, stream_stride, chunk_off, n_sym, t, map,
                         stream);
}

extern "C" int rgbd_gather_streams(const uint32_t *out_a, int64_t cap_a, int32_t n_a, const uint32_t *out_b,
                                   int64_t cap_b, int32_t n_b, const int32_t *nwords, uint32_t *dst,
                                   int64_t dst_cap_words, void *stream) {
    RGBD_CHECK_ARG(nwords && dst && (n_a == 0 || out_a) && (n_b == 0 || out_b), "null pointer");
    RGBD_CHECK_ARG(n_a >= 0 && n_b >= 0 && cap_a >= 0 && cap_b >= 0 && dst_cap_words >= 0, "sizes");
    if (n_a + n_b == 0) return RGBD_OK;
    gather_streams_kernel<<<n_a + n_b, 256, 0, (cudaStream_t)stream>>>(out_a, cap_a, n_a, out_b, cap_b, n_b, nwords, dst,
                                                                      dst_cap_words);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
