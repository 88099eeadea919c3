// Microtest: does a SWIZZLE_128B K-major UMMA descriptor accept a start address shifted by whole
// 128-byte rows (not 1024-aligned)?  mode 0: base_offset = 0; mode 1: base_offset = (addr >> 7) & 7.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int mode) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    if (mode == 1) d |= (uint64_t)((addr >> 7) & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr int kRows = 384, kN = 64;
__global__ void __launch_bounds__(128) test_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                   float *out, int r0, int mode) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t sa = base, sb = base + kRows * 128, bar = sb + kN * 128, bar2 = bar + 8, slot = bar + 16;
    uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(raw + (slot - smem_u32(raw)));
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, kRows * 128 + kN * 128);
        tma_load_2d(sa, &amap, bar, 0, 0);
        tma_load_2d(sa + 192 * 128, &amap, bar, 0, 192);
        tma_load_2d(sb, &bmap, bar, 0, 0);
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, make_desc(sa + r0 * 128 + k * 32, mode), make_desc(sb + k * 32, 0), idesc, k > 0);
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < kN; c += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[threadIdx.x * kN + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}
typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    std::vector<__nv_bfloat16> hA(kRows * 64), hB(kN * 64);
    std::vector<float> fA(kRows * 64), fB(kN * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = v; }
    __nv_bfloat16 *dA, *dB; float *dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * kN * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap am, bm; cuuint32_t ones[2] = {1, 1};
    { cuuint64_t dims[2] = {64, kRows}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, 192};
      CUresult r = enc(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, st, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc A %d\n", r); return 1; } }
    { cuuint64_t dims[2] = {64, kN}; cuuint64_t st[1] = {128}; cuuint32_t box[2] = {64, kN};
      CUresult r = enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, st, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); if (r) { printf("enc B %d\n", r); return 1; } }
    const size_t smem = kRows * 128 + kN * 128 + 1024 + 64;
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> hO(128 * kN);
    const int shifts[] = {0, 8, 1, 2, 3, 7, 9, 18, 33, 66, 100, 255};
    for (int mode = 0; mode < 2; ++mode)
        for (int r0 : shifts) {
            cudaMemset(dO, 0, 128 * kN * 4);
            test_kernel<<<1, 128, smem>>>(am, bm, dO, r0, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, r0, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0; int bad = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < kN; ++n) {
                float ref = 0; for (int k = 0; k < 64; ++k) ref += fA[(r0 + m) * 64 + k] * fB[n * 64 + k];
                double err = fabs(ref - hO[m * kN + n]); if (err > 1e-3) ++bad; if (err > maxerr) maxerr = err;
            }
            printf("mode %d (base_offset %s) row shift %3d: %s  (max err %.4f, %d bad)\n", mode, mode ? "=(addr>>7)&7" : "=0", r0, bad ? "MISMATCH" : "ok", maxerr, bad);
        }
    return 0;
}
