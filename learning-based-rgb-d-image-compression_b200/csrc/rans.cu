// GPU rANS coder, byte-identical to compressai.ans
// (reference: CompressAI/compressai/cpp_exts/rans/rans_interface.cpp:99-351 and
//  CompressAI/third_party/ryg_rans/rans64.h:59-142).
//
// A rANS stream is one serial dependency chain, so the unit of parallelism is the stream:
// ONE WARP PER STREAM.  All 32 lanes carry the 64-bit coder state redundantly (uniform
// control flow, no divergence); the lanes differ in what they pre-compute for the chain:
//   encoder: lane j turns symbol (batch_end-1-j) into (start, range, exact reciprocal)
//            while the chain is still busy with the previous batch;
//   decoder: lanes pre-load the next 32 indexes / table descriptors / stream words and
//            probe 32 CDF entries at once (ballot search) for the symbol lookup.
// The compacted uint16 CDF tables (27 256 entries for the 64 Gaussian tables) live in
// shared memory for the whole kernel.
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 2;
constexpr uint64_t kRansL = 1ull << 31;
constexpr int kMaxTables = 256;

struct TableSmem {
    int32_t base[kMaxTables];
    int32_t length[kMaxTables];
    int32_t offset[kMaxTables];
};

__device__ __forceinline__ void load_tables(const rgbd_rans_tables &t, uint16_t *s_cdf, TableSmem &m) {
    for (int i = threadIdx.x; i < t.total; i += blockDim.x) s_cdf[i] = t.cdf[i];
    for (int i = threadIdx.x; i < t.n_tables; i += blockDim.x) {
        m.base[i] = t.base[i];
        m.length[i] = t.length[i];
        m.offset[i] = t.offset[i];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Encoder
// ---------------------------------------------------------------------------------------
struct EncSym {
    uint64_t rcp;      // exact reciprocal of range (Alverson), rans64.h:167-247
    uint32_t start;    // bias
    uint32_t range;
    uint32_t rcp_shift;
    uint32_t raw;      // escape payload (only meaningful when esc)
    uint32_t esc;      // 1 if the symbol sits in the escape bin
};

__device__ __forceinline__ void make_reciprocal(uint32_t freq, uint64_t &rcp, uint32_t &shift) {
    // freq >= 2 : rcp = ceil(2^(s+63) / freq), s = ceil(log2 freq); q = mulhi(x, rcp) >> (s-1)
    uint32_t s = 32 - __clz(freq - 1);  // ceil(log2(freq)) for freq >= 2
    uint64_t x1 = 1ull << (s + 31);
    uint64_t t1 = x1 / freq;
    uint64_t x0 = (uint64_t)(freq - 1) + ((x1 % freq) << 32);
    uint64_t t0 = x0 / freq;
    rcp = t0 + (t1 << 32);
    shift = s - 1;
}

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

// Per-symbol record staged in shared memory by the lane that prepared it; the serial chain reads it
// back as a broadcast (prefetched one symbol ahead, so the LDS latency is off the x chain).
struct __align__(16) EncRec {
    uint64_t rcp;     // exact reciprocal of range; ~0 for range == 1 (rans64.h:191-210)
    uint32_t bias;    // start (+ 65535 for range == 1)
    uint32_t packed;  // range | rcp_shift << 16
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rans_encode_kernel(const int32_t *__restrict__ sym, const uint8_t *__restrict__ idx,
                   int64_t stream_stride, int32_t n_sym, int32_t n_streams, rgbd_rans_tables t,
                   uint32_t *__restrict__ out, int64_t cap_words, int32_t *__restrict__ nwords) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ EncRec s_rec[kWarpsPerBlock][32];
    TableSmem &meta = *reinterpret_cast<TableSmem *>(smem_raw);
    uint16_t *s_cdf = reinterpret_cast<uint16_t *>(smem_raw + sizeof(TableSmem));
    load_tables(t, s_cdf, meta);

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int s = blockIdx.x * kWarpsPerBlock + wib;
    if (s >= n_streams) return;

    const int32_t *my_sym = sym + (int64_t)s * stream_stride;
    const uint8_t *my_idx = idx + (int64_t)s * stream_stride;
    uint32_t *out_base = out + (int64_t)s * cap_words;
    EncRec *rec = s_rec[wib];

    EncState st;
    st.x = kRansL;
    st.cur = cap_words;
    st.overflow = false;

    // software prefetch: the (sym, idx) pair of the NEXT batch is requested one batch early
    int32_t pf_sym = 0, pf_idx = 0;
    if (n_sym - 1 - lane >= 0) {
        pf_sym = my_sym[n_sym - 1 - lane];
        pf_idx = my_idx[n_sym - 1 - lane];
    }
    for (int32_t hi = n_sym; hi > 0; hi -= 32) {
        // lane j prepares symbol hi-1-j (the chain consumes lanes 0,1,2,... in that order)
        const int32_t i = hi - 1 - lane;
        const int32_t cur_sym = pf_sym, ti = pf_idx;
        if (i - 32 >= 0) {
            pf_sym = my_sym[i - 32];
            pf_idx = my_idx[i - 32];
        }
        EncSym e;
        e.rcp = ~0ull; e.start = 0; e.range = 1; e.rcp_shift = 0; e.raw = 0; e.esc = 0;
        if (i >= 0) {
            const int32_t top = meta.length[ti] - 2;
            int32_t v = cur_sym - meta.offset[ti];
            uint32_t raw = 0;
            if (v < 0) {
                raw = (uint32_t)(-2 * v - 1);
                v = top;
            } else if (v >= top) {
                raw = (uint32_t)(2 * (v - top));
                v = top;
            }
            const uint16_t c0 = s_cdf[meta.base[ti] + v];
            const uint16_t c1 = s_cdf[meta.base[ti] + v + 1];
            e.start = c0;
            e.range = (uint16_t)(c1 - c0);  // uint16 wrap == the reference's static_cast<uint16_t>
            e.raw = raw;
            e.esc = (v == top) ? 1u : 0u;
            if (e.range >= 2) make_reciprocal(e.range, e.rcp, e.rcp_shift);
        }
        const int nvalid = hi < 32 ? hi : 32;
        const uint32_t any_esc = __ballot_sync(0xffffffffu, e.esc != 0);
        if (any_esc == 0) {
            // ---- fast path: no escape symbol in this batch ----
            EncRec mine;
            mine.rcp = e.rcp;
            mine.bias = e.range >= 2 ? e.start : e.start + 65535u;   // range 1: q = x - 1 (rans64.h:191-210)
            mine.packed = e.range | (e.rcp_shift << 16);
            rec[lane] = mine;
            __syncwarp();
            EncRec cur = rec[0];
            for (int j = 0; j < nvalid; ++j) {
                const EncRec nxt = rec[(j + 1) & 31];
                const uint32_t range = cur.packed & 0xFFFFu;
                // Rans64EncPut (rans64.h:77-93): x_max = ((L >> 16) << 32) * range
                if (st.x >= ((uint64_t)range << 47)) enc_emit(st, out_base, lane);
                const uint64_t q = __umul64hi(st.x, cur.rcp) >> (cur.packed >> 16);   // == x / range (exact)
                st.x = st.x + cur.bias + q * (uint64_t)(65536u - range);
                cur = nxt;
            }
            __syncwarp();
            continue;
        }
        for (int j = 0; j < nvalid; ++j) {
            const uint32_t esc = __shfl_sync(0xffffffffu, e.esc, j);
            if (esc) {  // warp-uniform; rare
                const uint32_t raw = __shfl_sync(0xffffffffu, e.raw, j);
                int nnib = 0;
                while (nnib < 8 && (raw >> (nnib * 4)) != 0) ++nnib;
                for (int k = nnib - 1; k >= 0; --k) enc_put_bits4(st, (raw >> (k * 4)) & 15u, out_base, lane);
                // count field: forward order is [15]*(nnib/15) then nnib%15; emit reversed
                enc_put_bits4(st, (uint32_t)(nnib % 15), out_base, lane);
                for (int k = 0; k < nnib / 15; ++k) enc_put_bits4(st, 15u, out_base, lane);
            }
            const uint32_t range = __shfl_sync(0xffffffffu, e.range, j);
            const uint32_t start = __shfl_sync(0xffffffffu, e.start, j);
            const uint32_t rshift = __shfl_sync(0xffffffffu, e.rcp_shift, j);
            const uint64_t rcp = __shfl_sync(0xffffffffu, (unsigned long long)e.rcp, j);
            if (st.x >= ((uint64_t)range << 47)) enc_emit(st, out_base, lane);
            if (range >= 2) {
                const uint64_t q = __umul64hi(st.x, rcp) >> rshift;  // == x / range (exact)
                st.x = st.x + start + q * (uint64_t)(65536u - range);
            } else {
                // range 1: x/1 = x, x%1 = 0  (range 0 cannot occur: every bin has freq >= 1)
                st.x = (st.x << 16) + start;
            }
        }
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
    __device__ __forceinline__ uint32_t fetch(int64_t p, int lane) const {
        return p < len ? words[p] : 0u;
    }
    __device__ __forceinline__ void init(const uint32_t *w, int64_t l, int64_t pos, int lane) {
        words = w; len = l; base = pos;
        cur = fetch(base + lane, lane);
        nxt = fetch(base + 32 + lane, lane);
    }
    // word at absolute position p (p >= base, p < base + 64)
    __device__ __forceinline__ uint32_t take(int64_t p, int lane) {
        if (p - base >= 32) {
            cur = nxt;
            base += 32;
            nxt = fetch(base + 32 + lane, lane);
        }
        return __shfl_sync(0xffffffffu, cur, (int)(p - base));
    }
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rans_decode_kernel(const uint32_t *__restrict__ words, const int64_t *__restrict__ word_off,
                   const int64_t *__restrict__ word_len, int32_t n_streams,
                   rgbd_rans_dec_state *__restrict__ state, const uint8_t *__restrict__ idx,
                   int32_t *__restrict__ sym, int64_t stream_stride, int64_t chunk_off,
                   int32_t n_sym, rgbd_rans_tables t) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int4 s_meta[kWarpsPerBlock][32];   // (base, length, offset, -) of the batch's tables
    TableSmem &meta = *reinterpret_cast<TableSmem *>(smem_raw);
    uint16_t *s_cdf = reinterpret_cast<uint16_t *>(smem_raw + sizeof(TableSmem));
    load_tables(t, s_cdf, meta);

    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int s = blockIdx.x * kWarpsPerBlock + wib;
    if (s >= n_streams) return;

    uint64_t x = state[s].x;
    int64_t pos = state[s].pos;
    WordFeed feed;
    feed.init(words + word_off[s], word_len[s], pos, lane);
    uint32_t wnext = feed.take(pos, lane);   // the word a renormalisation would consume next

    const uint8_t *my_idx = idx + (int64_t)s * stream_stride + chunk_off;
    int32_t *my_sym = sym + (int64_t)s * stream_stride + chunk_off;
    int4 *mrow = s_meta[wib];

    int32_t pf_idx = (lane < n_sym) ? (int32_t)my_idx[lane] : 0;  // prefetched one batch early
    for (int32_t lo_i = 0; lo_i < n_sym; lo_i += 32) {
        const int32_t i = lo_i + lane;
        const int32_t ti = pf_idx;
        if (i + 32 < n_sym) pf_idx = my_idx[i + 32];
        int4 mine = make_int4(0, 2, 0, 0);
        if (i < n_sym) mine = make_int4(meta.base[ti], meta.length[ti], meta.offset[ti], 0);
        mrow[lane] = mine;
        __syncwarp();
        int32_t my_out = 0;
        const int nvalid = (n_sym - lo_i) < 32 ? (n_sym - lo_i) : 32;
        int4 cur = mrow[0];
        for (int j = 0; j < nvalid; ++j) {
            const int4 nxt = mrow[(j + 1) & 31];
            const int32_t b = cur.x, L = cur.y, off = cur.z;
            const uint32_t cf = (uint32_t)(x & 0xFFFFu);  // Rans64DecGet
            // invariant cdf[lo] <= cf < cdf[hi]; cdf[L-1] == 65536 (stored as 0)
            int32_t lo = 0, hi = L - 1;
            while (hi - lo > 31) {   // only tables with more than 32 entries
                const int32_t stride = (hi - lo + 31) >> 5;
                const int32_t p = lo + (lane + 1) * stride;
                const bool above = (p >= hi) || (s_cdf[b + p] > cf);
                const uint32_t m = __ballot_sync(0xffffffffu, above);
                const int f = __ffs(m) - 1;  // m != 0: lane 31 always reaches hi
                const int32_t nhi = lo + (f + 1) * stride;
                lo = lo + f * stride;
                hi = nhi < hi ? nhi : hi;
            }
            // every lane tests the hypothesis "the symbol is lo + lane" and pre-computes the state
            // that hypothesis leads to (Rans64DecAdvance, rans64.h:126-142); exactly one lane is right
            const int32_t p = lo + lane;
            const uint32_t c0 = (p >= L - 1) ? 65536u : (uint32_t)s_cdf[b + p];
            const uint32_t c1 = (p + 1 >= L - 1) ? 65536u : (uint32_t)s_cdf[b + p + 1];
            const bool hit = (c0 <= cf) && (cf < c1);
            const uint64_t xh = (uint64_t)(c1 - c0) * (x >> 16) + (uint64_t)(cf - c0);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            const int f = __ffs(m) - 1;
            x = __shfl_sync(0xffffffffu, (unsigned long long)xh, f);
            int32_t value = lo + f;
            if (x < kRansL) {
                x = (x << 32) | wnext;
                pos += 1;
                wnext = feed.take(pos, lane);
            }
            if (value == L - 2) {  // escape bin: bypass nibbles (rans_interface.cpp:323-344)
                auto get4 = [&]() -> int32_t {
                    const uint32_t v = (uint32_t)(x & 15u);
                    x >>= 4;
                    if (x < kRansL) {
                        x = (x << 32) | wnext;
                        pos += 1;
                        wnext = feed.take(pos, lane);
                    }
                    return (int32_t)v;
                };
                int32_t d = get4();
                int32_t nnib = d;
                // valid streams carry at most 8 payload nibbles, so the unary count never continues;
                // the bounds only stop a corrupt stream from spinning
                for (int guard = 0; d == 15 && guard < 4; ++guard) {
                    d = get4();
                    nnib += d;
                }
                if (nnib > 64) nnib = 64;
                int32_t raw = 0;
                for (int k = 0; k < nnib; ++k) {
                    d = get4();
                    if (k < 8) raw |= d << (k * 4);
                }
                value = raw >> 1;
                if (raw & 1) value = -value - 1;
                else value += L - 2;
            }
            if (lane == j) my_out = value + off;
            cur = nxt;
        }
        __syncwarp();
        if (i < n_sym) my_sym[i] = my_out;
    }
    if (lane == 0) {
        state[s].x = x;
        state[s].pos = pos;
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

size_t table_smem_bytes(const rgbd_rans_tables *t) {
    return sizeof(TableSmem) + (((size_t)t->total * 2 + 15) & ~(size_t)15) + 64;
}

}  // namespace

extern "C" int rgbd_rans_encode(const int32_t *sym, const uint8_t *idx, int64_t stream_stride,
                                int32_t n_sym, int32_t n_streams, const rgbd_rans_tables *t,
                                uint32_t *out, int64_t cap_words, int32_t *nwords, void *stream) {
    RGBD_CHECK_ARG(sym && idx && t && out && nwords, "null pointer");
    RGBD_CHECK_ARG(n_sym >= 0 && n_streams >= 0 && cap_words >= 2, "sizes");
    RGBD_CHECK_ARG(t->n_tables > 0 && t->n_tables <= kMaxTables, "n_tables must be in 1..256");
    const size_t smem = table_smem_bytes(t);
    RGBD_CHECK_ARG(smem <= 200 * 1024, "CDF tables do not fit in shared memory");
    if (n_streams == 0) return RGBD_OK;
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(rans_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const int grid = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
    rans_encode_kernel<<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
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

extern "C" int rgbd_rans_decode_chunk(const uint32_t *words, const int64_t *word_off,
                                      const int64_t *word_len, int32_t n_streams,
                                      rgbd_rans_dec_state *state, const uint8_t *idx, int32_t *sym,
                                      int64_t stream_stride, int64_t chunk_off, int32_t n_sym,
                                      const rgbd_rans_tables *t, void *stream) {
    RGBD_CHECK_ARG(words && word_off && word_len && state && idx && sym && t, "null pointer");
    RGBD_CHECK_ARG(n_sym >= 0 && n_streams >= 0, "sizes");
    RGBD_CHECK_ARG(t->n_tables > 0 && t->n_tables <= kMaxTables, "n_tables must be in 1..256");
    const size_t smem = table_smem_bytes(t);
    RGBD_CHECK_ARG(smem <= 200 * 1024, "CDF tables do not fit in shared memory");
    if (n_streams == 0 || n_sym == 0) return RGBD_OK;
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(rans_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const int grid = (n_streams + kWarpsPerBlock - 1) / kWarpsPerBlock;
    rans_decode_kernel<<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        words, word_off, word_len, n_streams, state, idx, sym, stream_stride, chunk_off, n_sym, *t);
    RGBD_LAUNCH_CHECK();
    return RGBD_OK;
}
