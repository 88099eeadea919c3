// Shared building blocks of the tcgen05 / TMA kernels (conv_halo.cu, conv_rb.cu): PTX wrappers, UMMA descriptors,
// bf16 pack / unpack helpers and the host-side tensor-map encoder.  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace {

constexpr int kBlockK = 64;                        // bf16 elements = 128 B = one swizzle row

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
#ifdef RGBD_MBAR_SOFT_TIMEOUT
        // debugging aid (RGBD_BUILD_DEFINES=-DRGBD_MBAR_SOFT_TIMEOUT): report the barrier and carry on, so that the kernel
        // ends and the message reaches the host (a trap discards the printf buffer)
        if (clock64() - t0 > 100000000LL) {
            if ((threadIdx.x & 31) == 0 || (threadIdx.x >> 5) < 2) {
                unsigned long long ns;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
                printf("T %llu rgbd tc: mbarrier timeout block %d warp %d bar# %u parity %u\n", ns, blockIdx.x, threadIdx.x >> 5, (bar & 1023u) >> 3, parity);
            }
            return;
        }
#else
        if (clock64() - t0 > 4000000000LL) {  // ~2 s
            printf("rgbd conv_halo: mbarrier timeout (block %d thread %d bar %u)\n", blockIdx.x, threadIdx.x, bar);
            __trap();
        }
#endif
    }
}
// wait that adds the stalled cycles to *acc when tracing
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, bool tr, long long &acc) {
    if (!tr) {
        mbar_wait(bar, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
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
// TMA store of a packed shared-memory box (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // writes are complete
// Programmatic dependent launch: a conv kernel is launched while its predecessor in the stream is still in its
// tail; everything up to griddep_wait() (TMEM allocation, barrier init, bias staging, descriptor fetch) overlaps
// that tail, and nothing that reads or writes activation memory happens before it.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
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
// 32 lanes x 16 consecutive fp32 columns; the caller waits (tmem_ld_wait) before reading v
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// the loaded registers are only defined after the wait: naming them as in/out operands keeps the compiler from
// scheduling any use of them above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t *r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// the MMAs of one weight stage into one accumulator: nk 16-channel steps (32 B = 2 descriptor units apart)
__device__ __forceinline__ void issue_mmas(uint32_t dcol, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t started, int nk) {
    if (nk == 4) {
        umma_bf16(dcol, adesc, bdesc, idesc, started);
        umma_bf16(dcol, adesc + 2, bdesc + 2, idesc, 1u);
        umma_bf16(dcol, adesc + 4, bdesc + 4, idesc, 1u);
        umma_bf16(dcol, adesc + 6, bdesc + 6, idesc, 1u);
    } else {
        for (int k = 0; k < nk; ++k)
            umma_bf16(dcol, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, started | (uint32_t)k);
    }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1),
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups), [46,48) version = 1,
//   [49,52) base offset = 0 (also for row-shifted starts: the swizzle uses absolute address bits),
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

// Branch-free activation: max(v, slope * v) with slope 1 (none), 0 (ReLU) or 0.01 (LeakyReLU).  A runtime
// switch on the activation kind compiles to three branches per element, which made the epilogue the
// bottleneck of every small-K layer.
__device__ __forceinline__ float act_slope(int act) {
    return act == RGBD_ACT_RELU ? 0.f : (act == RGBD_ACT_LEAKY ? 0.01f : 1.f);
}
__device__ __forceinline__ float act_fn(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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

__device__ __forceinline__ void unpack8(const uint4 &q, float *v) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162 *>(&w[i]);
        v[2 * i] = __low2float(h);
        v[2 * i + 1] = __high2float(h);
    }
}
__device__ __forceinline__ uint4 pack8(const float *v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t *>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// 32-byte global accesses (LDG.256 / STG.256).  In every epilogue a thread owns one pixel row, so the 32 accesses of a warp
// instruction land in 32 different 128-byte lines; the load / store unit spends one wavefront per line however many bytes
// it carries, so 32 bytes per lane halve the wavefronts of the 16-byte form.  The address must be 32-byte aligned.
struct alignas(32) U8 {
    uint32_t v[8];
};
__device__ __forceinline__ U8 ldg256(const void *p) {
    U8 r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
}
// 16 consecutive bf16 of one pixel row as two uint4: one 32-byte load when the address allows, else two 16-byte loads
__device__ __forceinline__ void ldg_2x128(const void *p, uint4 &a, uint4 &b) {
    if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
        const U8 w = ldg256(p);
        a = make_uint4(w.v[0], w.v[1], w.v[2], w.v[3]);
        b = make_uint4(w.v[4], w.v[5], w.v[6], w.v[7]);
    } else {
        a = reinterpret_cast<const uint4 *>(p)[0];
        b = reinterpret_cast<const uint4 *>(p)[1];
    }
}
__device__ __forceinline__ void stg256(void *p, const uint4 &a, const uint4 &b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
// 16 consecutive bf16 channels of one pixel -> fp32 (vector path when 16-byte aligned and complete)
__device__ __forceinline__ void load16_bf16(const __nv_bfloat16 *p, bool vec, int nvalid, float *v) {
    if (vec && nvalid >= 16) {
        uint4 a, b;
        ldg_2x128(p, a, b);       // one 32-byte load when the address allows
        unpack8(a, v);
        unpack8(b, v + 8);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = i < nvalid ? __bfloat162float(p[i]) : 0.f;
    }
}
template <typename T> __device__ __forceinline__ void store16(T *p, bool vec, int nvalid, const float *v);
template <> __device__ __forceinline__ void store16<__nv_bfloat16>(__nv_bfloat16 *p, bool vec, int nvalid, const float *v) {
    if (vec && nvalid >= 16) {
        if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {      // 16 channels = 32 bytes: one STG.256 when the row is 32-byte aligned
            stg256(p, pack8(v), pack8(v + 8));
        } else {
            reinterpret_cast<uint4 *>(p)[0] = pack8(v);
            reinterpret_cast<uint4 *>(p)[1] = pack8(v + 8);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nvalid) p[i] = __float2bfloat16_rn(v[i]);
    }
}
template <> __device__ __forceinline__ void store16<float>(float *p, bool vec, int nvalid, const float *v) {
    if (vec && nvalid >= 16) {
        if ((reinterpret_cast<uintptr_t>(p) & 31) == 0) {
            uint4 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                q[i] = make_uint4(__float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]), __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
            stg256(p, q[0], q[1]);
            stg256(p + 8, q[2], q[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4 *>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nvalid) p[i] = v[i];
    }
}

// position in a ring of n stages: stage index + phase parity, advanced without integer division
struct RingPos {
    int s;
    uint32_t ph;
    __device__ __forceinline__ void next(int n) {
        if (++s == n) {
            s = 0;
            ph ^= 1u;
        }
    }
};

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// one-time initialisation below is shared by host threads (the round-trip pipeline may drive each slot from its own thread)
std::mutex g_init_mutex;

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int encode_map(CUtensorMap *m, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides_bytes,
               const cuuint32_t *box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        rgbd_set_error("conv_tc: cuTensorMapEncodeTiled unavailable");
        return RGBD_E_CUDA;
    }
    cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), dims,
                     strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rgbd_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d) rank %d dims %llu %llu %llu box %u %u %u", (int)r,
                       rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                       box[0], box[1], box[2]);
        return RGBD_E_CUDA;
    }
    return RGBD_OK;
}

}  // namespace
