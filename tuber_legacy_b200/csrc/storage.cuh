// Upload / download conversion between the C ABI's host layouts and the device's two-sector rows,
// presence-bit maintenance, the standalone velocity system, and Ecs::query (presence AND -> packed
// ascending entity indices; crates/tuber-ecs/src/query.rs:94-129).
#pragma once
#include "common.cuh"

namespace tb {

// ---- upload conversion -----------------------------------------------------------------------

// tb_transform[first..first+n) (12 f32 each, transform.rs:5-11 order) -> row sectors.
__global__ void __launch_bounds__(kThreads) convert_transforms_kernel(const float* __restrict__ src, RowView rv,
                                                                       uint32_t first, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = src + (size_t)i * 12;
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = __ldg(t + k);
    const size_t e = first + i;
    // A0 = {t.x, t.y, t.z, c.z}
    rv.t[e * rv.st] = make_float4(v[0], v[1], v[2], v[8]);
    // A1 = {s.z, angle.x, angle.y, texlayer}: keep .w
    float* a1 = reinterpret_cast<float*>(rv.a + e * rv.st);
    a1[0] = v[11];
    a1[1] = v[3];
    a1[2] = v[4];
    // B0 = {sin(angle.z / 2), c.x, c.y, s.x}, B1 = {s.y, size.w, size.h, cos(angle.z / 2)}: keep B1 .y .z
    const float2 sc = yaw_sincos(v[5]);
    rv.b[e * rv.sb] = make_float4(sc.x, v[6], v[7], v[9]);
    float* b1 = reinterpret_cast<float*>(rv.b + e * rv.sb + 1);
    b1[0] = v[10];
    b1[3] = sc.y;
    rv.yaw[e] = v[5];
}

__global__ void __launch_bounds__(kThreads) convert_sprites_kernel(const uint32_t* __restrict__ src, RowView rv,
                                                                    uint32_t first, uint32_t n, uint32_t* range_err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(src) + i);
    if (s.z > 0xFFFFu || s.w > 0xFFFEu) atomicAdd(range_err, 1u);
    const size_t e = first + i;
    float* a1 = reinterpret_cast<float*>(rv.a + e * rv.st);
    a1[3] = __uint_as_float((s.z & 0xFFFFu) | (s.w << 16));
    float* b1 = reinterpret_cast<float*>(rv.b + e * rv.sb + 1);
    b1[1] = __uint_as_float(s.x);
    b1[2] = __uint_as_float(s.y);
}

// Copies every row from one layout to the other (split arrays <-> one 64-byte row per entity): one
// thread per 16-byte chunk.
__global__ void __launch_bounds__(kThreads) relayout_rows_kernel(RowView src, RowView dst, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t i = t >> 2;
    if (i >= n) return;
    const uint32_t ch = (uint32_t)t & 3u;
    const float4* from = ch == 0 ? src.t + i * src.st : ch == 1 ? src.a + i * src.st : src.b + i * src.sb + (ch & 1u);
    float4* to = ch == 0 ? dst.t + i * dst.st : ch == 1 ? dst.a + i * dst.st : dst.b + i * dst.sb + (ch & 1u);
    *to = *from;
}

// rows -> tb_transform (download for parity checks of the velocity system)
__global__ void __launch_bounds__(kThreads) unconvert_transforms_kernel(RowView rv, uint32_t first, uint32_t n,
                                                                         float* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Row r = load_row(rv, first + i);
    float* t = dst + (size_t)i * 12;
    t[0] = r.a0.x; t[1] = r.a0.y; t[2] = r.a0.z;
    t[3] = r.a1.y; t[4] = r.a1.z; t[5] = rv.yaw[first + i];
    t[6] = r.b0.y; t[7] = r.b0.z; t[8] = r.a0.w;
    t[9] = r.b0.w; t[10] = r.b1.x; t[11] = r.a1.x;
}

// Sets or clears presence bits [first, first+n) of one mask (whole words in the middle).
__global__ void __launch_bounds__(kThreads) mark_presence_kernel(uint64_t* mask, uint64_t first, uint64_t n, int set) {
    uint64_t w0 = first >> 6, w1 = (first + n - 1) >> 6;
    uint64_t w = w0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w > w1) return;
    uint64_t m = ~0ull;
    if (w == w0) m &= ~0ull << (first & 63);
    if (w == w1) {
        uint32_t last = (uint32_t)((first + n - 1) & 63);
        m &= (last == 63) ? ~0ull : ((1ull << (last + 1)) - 1ull);
    }
    if (set) atomicOr((unsigned long long*)&mask[w], m);
    else atomicAnd((unsigned long long*)&mask[w], ~m);
}

// Presence of the Parent component over [first, first+n): bit e == (parent[e] != TB_NO_PARENT).
__global__ void __launch_bounds__(kThreads) mark_parent_presence_kernel(uint64_t* mask, const uint32_t* __restrict__ parent,
                                                                         uint64_t first, uint64_t n) {
    const uint64_t w0 = first >> 6, w1 = (first + n - 1) >> 6;
    const uint64_t w = w0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (w > w1) return;
    uint64_t set = 0, range = 0;
    for (uint32_t b = 0; b < 64; ++b) {
        const uint64_t e = w * 64 + b;
        if (e < first || e >= first + n) continue;
        range |= 1ull << b;
        if (parent[e] != 0xFFFFFFFFu) set |= 1ull << b;
    }
    atomicAnd((unsigned long long*)&mask[w], ~range);
    atomicOr((unsigned long long*)&mask[w], set);
}

// ---- standalone velocity system (when no frame follows the tick) --------------------------------

__global__ void __launch_bounds__(kThreads) integrate_kernel(RowView rows, const uint64_t* mask_t, const uint64_t* mask_v,
                                                              const float* __restrict__ vel, float dt, uint32_t ticks,
                                                              uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!mask_bit(mask_t, i) || !mask_bit(mask_v, i)) return;
    const size_t o = (size_t)i * rows.st;
    float4 a0 = rows.t[o];
    const float* v = vel + (size_t)i * 3;
    integrate_ticks(a0, __ldg(v), __ldg(v + 1), __ldg(v + 2), dt, ticks);
    rows.t[o] = a0;
}

// ---- Ecs::query: presence AND -> packed ascending indices --------------------------------------

struct QueryParams {
    const uint64_t* masks[4];
    uint32_t nmasks;
    uint32_t nwords;
    uint32_t n;  // entity_count
};

__device__ __forceinline__ uint64_t query_word(const QueryParams& q, uint32_t w) {
    uint64_t m = ~0ull;
    for (uint32_t k = 0; k < q.nmasks; ++k) m &= q.masks[k] ? q.masks[k][w] : ~0ull;
    const uint32_t lo = w * 64u;
    if (lo + 64u > q.n) m &= (q.n > lo) ? ((q.n - lo == 64u) ? ~0ull : ((1ull << (q.n - lo)) - 1ull)) : 0ull;
    return m;
}

// popcount of 1024 mask words per block
__global__ void __launch_bounds__(kThreads) query_count_kernel(const QueryParams q, uint32_t* block_counts) {
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    uint32_t c = 0;
    for (uint32_t k = threadIdx.x; k < 1024; k += kThreads) {
        const uint32_t w = blockIdx.x * 1024 + k;
        if (w < q.nwords) c += __popcll(query_word(q, w));
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = s_c;
}

__global__ void __launch_bounds__(kThreads) query_write_kernel(const QueryParams q, const uint32_t* block_offsets,
                                                                uint32_t* out, uint32_t cap) {
    __shared__ uint32_t s_w[kWarps];
    __shared__ uint32_t s_run;
    if (threadIdx.x == 0) s_run = block_offsets[blockIdx.x];
    __syncthreads();
    for (uint32_t k0 = 0; k0 < 1024; k0 += kThreads) {
        const uint32_t w = blockIdx.x * 1024 + k0 + threadIdx.x;
        uint64_t m = w < q.nwords ? query_word(q, w) : 0ull;
        const uint32_t c = __popcll(m);
        uint32_t x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
        __syncthreads();
        uint32_t off = s_run + x - c;
        for (uint32_t ww = 0; ww < (threadIdx.x >> 5); ++ww) off += s_w[ww];
        while (m) {
            const uint32_t b = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (off < cap) out[off] = w * 64u + b;
            ++off;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int ww = 0; ww < kWarps; ++ww) t += s_w[ww];
            s_run += t;
        }
        __syncthreads();
    }
}

}  // namespace tb
