// Shared device helpers of the tuber hot-path kernels: relaxed/streaming memory operations,
// presence-bit tests, row loads and the sort key of a row.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "compose.cuh"
#include "plan.h"

namespace tb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kInvalidDigit = 0xFFFFFFFFu;

// ---- small helpers ---------------------------------------------------------------------------

// Streaming (evict-first) 128-bit store: emitted buffers are written once and not re-read by
// this frame, so they should not displace the rows / keys that are.
__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) { __stcs(p, v); }

__device__ __forceinline__ bool mask_bit(const uint64_t* mask, uint32_t i) {
    return mask == nullptr || ((__ldg(mask + (i >> 6)) >> (i & 63u)) & 1ull);
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ Row load_row(const RowView& rv, uint32_t i) {
    Row r;
    r.a0 = rv.t[(size_t)i * rv.st];
    r.a1 = rv.a[(size_t)i * rv.st];
    r.b0 = rv.b[(size_t)i * rv.sb];
    r.b1 = rv.b[(size_t)i * rv.sb + 1];
    return r;
}

// Full 64-bit sort key of a flat row (chunks A0 A1 only when planar).
__device__ __forceinline__ uint64_t flat_key(const RowView& rv, uint32_t i, const float4 a0, const float4 a1) {
    const uint32_t tl = __float_as_uint(a1.w);
    float z;
    if (is_planar(a1)) {
        z = planar_z(a0, a1);
    } else {
        Row r;
        r.a0 = a0;
        r.a1 = a1;
        r.b0 = rv.b[(size_t)i * rv.sb];
        r.b1 = rv.b[(size_t)i * rv.sb + 1];
        z = local_z(r);
    }
    return sort_key(tl >> 16, tl & 0xFFFFu, z);
}

// World row written by the hierarchy path: {r0, r1, r2, {size.w, size.h, texlayer, 0}}.
struct WorldRow {
    float4 r0, r1, r2, s;
};

__device__ __forceinline__ uint64_t world_key(const float4 r2, const float4 s) {
    const uint32_t tl = __float_as_uint(s.z);
    return sort_key(tl >> 16, tl & 0xFFFFu, r2.w);
}

// ---- asynchronous copies into shared memory: mbarrier + TMA bulk (cp.async.bulk), cp.async ------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (1-D, no tensor map): size and both addresses multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

}  // namespace tb
