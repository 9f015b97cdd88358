// Device helpers shared by the specialised fused kernels (mbm_wta_fast.cu, mbm_wta_ws.cu).
#pragma once

#include "common.cuh"

namespace sd {
namespace mbm {

constexpr int BW = 64;            // tile width in pixels
constexpr int NCHUNK = 42;        // 16-byte chunks per cost-plane row: (BW + 20) cells x float2 / 16 B
constexpr int HALF = 21;          // even chunks [0,21), odd chunks [21,42)
constexpr int NSTRIP = 21;        // 4-column strips per cost-plane row
constexpr int LW = kBandLW;       // left band row pitch (floats): virtual columns c0-15 .. c0+80

// ---- TMA (bulk async copy) + mbarrier helpers ---------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ float2 lo2(const float4 &q) { return make_float2(q.x, q.y); }
__device__ __forceinline__ float2 hi2(const float4 &q) { return make_float2(q.z, q.w); }
// Shared-memory loads that the compiler keeps where they are written (volatile asm).
__device__ __forceinline__ float4 lds128(const float *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ float2 lds64(const float *p) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ float tap(float l, float r) { return __fsub_rn(255.0f, fabsf(__fsub_rn(l, r))); }

// Shared-memory position (in 16 B chunks) of logical chunk q within a cost-plane row.
__device__ __forceinline__ constexpr int chunk_pos(int q) { return (q >> 1) + (q & 1) * HALF; }


}  // namespace mbm
}  // namespace sd
