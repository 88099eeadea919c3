// pmf -> 16-bit quantised CDF, bit-exact with compressai._CXX.pmf_to_quantized_cdf
// (reference CompressAI/compressai/cpp_exts/ops/ops.cpp:24-81).  Host code: it runs
// 64 + 2*192 times per update() and defines the tables both coder sides use.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../include/rgbd_b200.h"

extern "C" void rgbd_set_error(const char *fmt, ...);

extern "C" int rgbd_pmf_to_quantized_cdf(const float *pmf, int32_t n, int32_t precision, uint32_t *cdf) {
    if (!pmf || !cdf || n <= 0 || precision < 1 || precision > 16) {
        rgbd_set_error("rgbd_pmf_to_quantized_cdf: invalid argument");
        return RGBD_E_INVALID;
    }
    const uint32_t one = 1u << precision;
    std::vector<uint32_t> freq(n);
    // float product, round-half-away (std::round), then the reference's `int` accumulation
    int total_i = 0;
    for (int i = 0; i < n; ++i) {
        freq[i] = static_cast<uint32_t>(std::round(pmf[i] * static_cast<float>(one)));
        total_i += static_cast<int>(freq[i]);
    }
    const uint32_t total = static_cast<uint32_t>(total_i);
    if (total == 0) {
        rgbd_set_error("rgbd_pmf_to_quantized_cdf: pmf sums to zero");
        return RGBD_E_INVALID;
    }
    uint32_t run = 0;
    cdf[0] = 0;
    for (int i = 0; i < n; ++i) {
        run += static_cast<uint32_t>((static_cast<uint64_t>(one) * freq[i]) / total);
        cdf[i + 1] = run;
    }
    cdf[n] = one;
    // zero-width bins steal one count from the narrowest bin that can spare it
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        uint32_t narrowest = ~0u;
        int donor = -1;
        for (int j = 0; j < n; ++j) {
            const uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < narrowest) {
                narrowest = f;
                donor = j;
            }
        }
        if (donor < 0) {
            rgbd_set_error("rgbd_pmf_to_quantized_cdf: no bin to steal from");
            return RGBD_E_INVALID;
        }
        if (donor < i)
            for (int j = donor + 1; j <= i; ++j) cdf[j]--;
        else
            for (int j = i + 1; j <= donor; ++j) cdf[j]++;
    }
    return RGBD_OK;
}
