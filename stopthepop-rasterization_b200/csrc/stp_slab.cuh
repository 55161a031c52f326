// stp_slab.cuh -- per-tile Gaussian slabs: the render kernels' gathers, materialised once in tile order.
//
// The reference's depth-resorting render kernels (hierarchical_render.cuh:441-445,545-547,748-750,
// resorted_render.cuh:598-600) fetch, for every (consumer, instance) pair, the projected mean, the conic / opacity and
// the 48-byte inverse covariance of a Gaussian through its id: 72 bytes at three random addresses.  Here the epilogue of
// the per-tile depth sort (binning.cu) -- which touches every instance anyway -- writes ONE 64-byte record per instance
// in list order, so that the slab of a tile is one contiguous run of HBM that the render kernels pull into shared
// memory with 1-D bulk-async copies (cp.async.bulk = TMA, completion on an mbarrier) instead of per-entry gathers.
//
// Record (16 floats, four 16-byte chunks):
//   c0 = { mean.x, mean.y, conic.x, conic.y }        c1 = { conic.z, opacity, Gaussian id (bits), S00 }
//   c2 = { S01, S02, S11, S12 }                      c3 = { S22, u.x, u.y, u.z }       (S = Sigma^-1, u = Sigma^-1 (mu - o))
// Chunk c of instance j (j = absolute position in point_list) is stored at float4 index 4 j + (c ^ ((j >> 1) & 3)).  The
// XOR swizzle is baked into the global layout because a bulk copy cannot permute: with it, eight consecutive records
// read with 128-bit shared-memory loads (one quarter-warp) hit eight distinct 16-byte bank groups -- conflict-free for
// any alignment of the copied run -- and c0/c1 (everything the alpha test needs) as well as c2/c3 stay inside one
// 32-byte sector.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace stp {

constexpr int kSlabChunks = 4;               // float4 per record
constexpr int kSlabRecordBytes = 64;

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t slab_swizzle(uint32_t j) { return (j >> 1) & 3u; }

__device__ __forceinline__ void slab_store(float4* __restrict__ slab, uint32_t j, int id, float2 xy, float4 co, float4 ia,
                                           float4 ib, float4 ic) {
    float4* const r = slab + 4 * (size_t)j;
    const uint32_t s = slab_swizzle(j);
    r[0 ^ s] = make_float4(xy.x, xy.y, co.x, co.y);
    r[1 ^ s] = make_float4(co.z, co.w, __int_as_float(id), ia.x);
    r[2 ^ s] = make_float4(ia.y, ia.z, ib.x, ib.y);
    r[3 ^ s] = make_float4(ib.z, ic.x, ic.y, ic.z);
}

// alpha-test half of a record (c0, c1) / depth half (c2, c3; S00 travels in c1.w)
__device__ __forceinline__ void slab_load_head(const float4* rec, uint32_t j, float4& c0, float4& c1) {
    const uint32_t s = slab_swizzle(j);
    c0 = rec[0 ^ s];
    c1 = rec[1 ^ s];
}
__device__ __forceinline__ void slab_load_tail(const float4* rec, uint32_t j, float4& c2, float4& c3) {
    const uint32_t s = slab_swizzle(j);
    c2 = rec[2 ^ s];
    c3 = rec[3 ^ s];
}

// the same halves read from the GLOBAL slab through the read-only path (record j of the whole list)
__device__ __forceinline__ void slab_ldg_head(const float4* __restrict__ slab, uint32_t j, float4& c0, float4& c1) {
    const float4* const rec = slab + 4 * (size_t)j;
    const uint32_t s = slab_swizzle(j);
    c0 = __ldg(rec + (0 ^ s));
    c1 = __ldg(rec + (1 ^ s));
}
__device__ __forceinline__ void slab_ldg_inv(const float4* __restrict__ slab, uint32_t j, float* ic, float& ux, float& uy,
                                             float& uz) {
    const float4* const rec = slab + 4 * (size_t)j;
    const uint32_t s = slab_swizzle(j);
    ic[0] = __ldg(reinterpret_cast<const float*>(rec + (1 ^ s)) + 3);
    const float4 c2 = __ldg(rec + (2 ^ s)), c3 = __ldg(rec + (3 ^ s));
    ic[1] = c2.x; ic[2] = c2.y; ic[3] = c2.z; ic[4] = c2.w; ic[5] = c3.x;
    ux = c3.y; uy = c3.z; uz = c3.w;
}

// ---- bulk-async copy (TMA) + mbarrier, sm_90+ PTX ------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {  // make the initialised barriers visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {  // release.cta: everything this thread did before is visible to a waiter
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
#endif

}  // namespace stp
