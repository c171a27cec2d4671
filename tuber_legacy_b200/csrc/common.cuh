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
constexpr uint32_t kFlagLocal = 1u << 30;   // decoupled look-back status words (tile_offsets == 0 only)
constexpr uint32_t kFlagIncl = 2u << 30;
constexpr uint32_t kCountMask = (1u << 30) - 1u;

// ---- small helpers ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Streaming (evict-first) 128-bit store: emitted buffers are written once and not re-read by
// this frame, so they should not displace the rows / keys that are.
__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) { return __ldg(p); }

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
    const size_t o = (size_t)i * rv.stride;
    r.a0 = rv.a[o];
    r.a1 = rv.a[o + 1];
    r.b0 = rv.b[o];
    r.b1 = rv.b[o + 1];
    return r;
}

// Full 64-bit sort key of a flat row (sector A only when planar).
__device__ __forceinline__ uint64_t flat_key(const RowView& rv, uint32_t i, const float4 a0, const float4 a1) {
    const uint32_t tl = __float_as_uint(a1.w);
    float z;
    if (is_planar(a1)) {
        z = planar_z(a0, a1);
    } else {
        Row r;
        r.a0 = a0;
        r.a1 = a1;
        const size_t o = (size_t)i * rv.stride;
        r.b0 = rv.b[o];
        r.b1 = rv.b[o + 1];
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

}  // namespace tb
