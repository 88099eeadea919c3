/*
 * ORACLE (test infrastructure, never shipped, never imported by the product path).
 *
 * Plain-C restatement of the reference's CPU entropy coder for the ELIC_united hot path:
 *
 *   - 64-bit-state rANS primitives          CompressAI/third_party/ryg_rans/rans64.h:59-142
 *   - symbol -> (start, range, bypass) list  CompressAI/compressai/cpp_exts/rans/rans_interface.cpp:99-165
 *   - reverse encode + 2-word flush          rans_interface.cpp:60-78, 167-192
 *   - resumable decode incl. bypass nibbles  rans_interface.cpp:80-96, 278-351
 *   - float pmf -> 16-bit quantised CDF      CompressAI/compressai/cpp_exts/ops/ops.cpp:24-81
 *
 * Parity pin: tests/test_oracle_rans.py checks this file byte-for-byte against the
 * reference's own compiled module (oracle/_ref/ans*.so, built by oracle/Makefile from the
 * sources under /root/reference) and against tests/golden/rans_kat.npz, which was produced
 * by that module (oracle/make_golden.py).
 *
 * The code is written from the specification in SURVEY.md Appendix B, with a flat
 * record array and explicit cursor arithmetic instead of the reference's std::vector
 * push/pop structure.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define PRECISION 16u
#define BYPASS_BITS 4u
#define BYPASS_MAX 15
#define RANS_LOW (1ull << 31)

typedef struct {
    uint16_t start;
    uint16_t range;
    uint8_t bypass;
} record_t;

/* number of records one symbol expands to (1 + escape payload) */
static int64_t expand_symbol(int32_t sym, int32_t idx, const int32_t *cdfs, int cdf_stride,
                             const int32_t *cdf_sizes, const int32_t *offsets, record_t *out)
{
    const int32_t *cdf = cdfs + (int64_t)idx * cdf_stride;
    const int32_t top = cdf_sizes[idx] - 2;
    int32_t v = sym - offsets[idx];
    uint32_t raw = 0;
    int64_t n = 0;
    if (v < 0) {
        raw = (uint32_t)(-2 * v - 1);
        v = top;
    } else if (v >= top) {
        raw = (uint32_t)(2 * (v - top));
        v = top;
    }
    if (out) {
        out[n].start = (uint16_t)cdf[v];
        out[n].range = (uint16_t)(cdf[v + 1] - cdf[v]);
        out[n].bypass = 0;
    }
    n++;
    if (v == top) {
        int32_t nib = 0;
        while ((raw >> (nib * BYPASS_BITS)) != 0)
            nib++;
        int32_t rem = nib;
        while (rem >= BYPASS_MAX) {
            if (out) {
                out[n].start = BYPASS_MAX;
                out[n].range = BYPASS_MAX + 1;
                out[n].bypass = 1;
            }
            n++;
            rem -= BYPASS_MAX;
        }
        if (out) {
            out[n].start = (uint16_t)rem;
            out[n].range = (uint16_t)(rem + 1);
            out[n].bypass = 1;
        }
        n++;
        for (int32_t j = 0; j < nib; ++j) {
            uint32_t d = (raw >> (j * BYPASS_BITS)) & BYPASS_MAX;
            if (out) {
                out[n].start = (uint16_t)d;
                out[n].range = (uint16_t)(d + 1);
                out[n].bypass = 1;
            }
            n++;
        }
    }
    return n;
}

/* Encode n symbols into `out` (capacity out_cap bytes). Returns the number of bytes, or
 * a negative value on error (-1 alloc, -2 capacity). The stream is the tail of the
 * word buffer exactly like the reference's flush(). */
int64_t rgbd_oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, int64_t n,
                                const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                                const int32_t *offsets, uint8_t *out, int64_t out_cap)
{
    int64_t nrec = 0;
    for (int64_t i = 0; i < n; ++i)
        nrec += expand_symbol(symbols[i], indexes[i], cdfs, cdf_stride, cdf_sizes, offsets, NULL);
    record_t *rec = (record_t *)malloc((size_t)(nrec > 0 ? nrec : 1) * sizeof(record_t));
    /* the reference's buffer holds one word per record; the final flush writes 2 more
     * words (below the buffer start if nothing was renormalised - UB there, fine here). */
    uint32_t *words = (uint32_t *)malloc((size_t)(nrec + 2) * sizeof(uint32_t));
    if (!rec || !words) {
        free(rec);
        free(words);
        return -1;
    }
    int64_t w = 0;
    for (int64_t i = 0; i < n; ++i)
        w += expand_symbol(symbols[i], indexes[i], cdfs, cdf_stride, cdf_sizes, offsets, rec + w);

    uint64_t x = RANS_LOW;
    int64_t cur = nrec + 2; /* write cursor, moves down */
    for (int64_t i = nrec - 1; i >= 0; --i) {
        if (!rec[i].bypass) {
            const uint64_t f = rec[i].range;
            const uint64_t lim = ((RANS_LOW >> PRECISION) << 32) * f;
            if (x >= lim) {
                words[--cur] = (uint32_t)x;
                x >>= 32;
            }
            x = ((x / f) << PRECISION) + (x % f) + rec[i].start;
        } else {
            const uint64_t f = 1u << (16 - BYPASS_BITS);
            const uint64_t lim = ((RANS_LOW >> 16) << 32) * f;
            if (x >= lim) {
                words[--cur] = (uint32_t)x;
                x >>= 32;
            }
            x = (x << BYPASS_BITS) | rec[i].start;
        }
    }
    cur -= 2;
    words[cur] = (uint32_t)x;
    words[cur + 1] = (uint32_t)(x >> 32);
    const int64_t nbytes = (nrec + 2 - cur) * 4;
    int64_t rv = nbytes;
    if (nbytes > out_cap)
        rv = -2;
    else
        memcpy(out, words + cur, (size_t)nbytes);
    free(rec);
    free(words);
    return rv;
}

/* resumable decoder state: x and the index of the next unread 32-bit word */
typedef struct {
    uint64_t x;
    int64_t pos;
} rgbd_oracle_dec_state;

void rgbd_oracle_rans_decode_init(rgbd_oracle_dec_state *st, const uint8_t *stream)
{
    uint32_t w0, w1;
    memcpy(&w0, stream, 4);
    memcpy(&w1, stream + 4, 4);
    st->x = (uint64_t)w0 | ((uint64_t)w1 << 32);
    st->pos = 2;
}

static inline uint32_t take_nibble(rgbd_oracle_dec_state *st, const uint8_t *stream)
{
    uint32_t v = (uint32_t)(st->x & BYPASS_MAX);
    st->x >>= BYPASS_BITS;
    if (st->x < RANS_LOW) {
        uint32_t w;
        memcpy(&w, stream + 4 * st->pos, 4);
        st->x = (st->x << 32) | w;
        st->pos++;
    }
    return v;
}

/* decode n symbols, continuing from *st (rans_interface.cpp:286-351) */
void rgbd_oracle_rans_decode_chunk(rgbd_oracle_dec_state *st, const uint8_t *stream,
                                   const int32_t *indexes, int64_t n, const int32_t *cdfs,
                                   int cdf_stride, const int32_t *cdf_sizes,
                                   const int32_t *offsets, int32_t *out)
{
    for (int64_t i = 0; i < n; ++i) {
        const int32_t idx = indexes[i];
        const int32_t *cdf = cdfs + (int64_t)idx * cdf_stride;
        const int32_t len = cdf_sizes[idx];
        const int32_t top = len - 2;
        const uint32_t cf = (uint32_t)(st->x & 0xFFFFu);
        int32_t s = 0;
        while (s < len && (uint32_t)cdf[s] <= cf)
            s++;
        s -= 1;
        const uint64_t start = (uint64_t)cdf[s], f = (uint64_t)(cdf[s + 1] - cdf[s]);
        st->x = f * (st->x >> PRECISION) + (st->x & 0xFFFFu) - start;
        if (st->x < RANS_LOW) {
            uint32_t w;
            memcpy(&w, stream + 4 * st->pos, 4);
            st->x = (st->x << 32) | w;
            st->pos++;
        }
        int32_t value = s;
        if (value == top) {
            int32_t d = (int32_t)take_nibble(st, stream);
            int32_t nib = d;
            while (d == BYPASS_MAX) {
                d = (int32_t)take_nibble(st, stream);
                nib += d;
            }
            int32_t raw = 0;
            for (int32_t j = 0; j < nib; ++j) {
                d = (int32_t)take_nibble(st, stream);
                raw |= d << (j * BYPASS_BITS);
            }
            value = raw >> 1;
            if (raw & 1)
                value = -value - 1;
            else
                value += top;
        }
        out[i] = value + offsets[idx];
    }
}

/* ops.cpp:24-81.  cdf has n+1 entries. Returns 0, or -1 if no bin can be stolen from. */
int rgbd_oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf)
{
    cdf[0] = 0;
    for (int i = 0; i < n; ++i)
        cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
    /* the reference accumulates into an `int` initial value */
    int total_i = 0;
    for (int i = 0; i <= n; ++i)
        total_i += (int)cdf[i];
    const uint32_t total = (uint32_t)total_i;
    for (int i = 0; i <= n; ++i)
        cdf[i] = (uint32_t)((((uint64_t)(1 << precision)) * cdf[i]) / total);
    for (int i = 1; i <= n; ++i)
        cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1])
            continue;
        uint32_t best = ~0u;
        int donor = -1;
        for (int j = 0; j < n; ++j) {
            uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < best) {
                best = f;
                donor = j;
            }
        }
        if (donor < 0)
            return -1;
        if (donor < i) {
            for (int j = donor + 1; j <= i; ++j)
                cdf[j]--;
        } else {
            for (int j = i + 1; j <= donor; ++j)
                cdf[j]++;
        }
    }
    return 0;
}
