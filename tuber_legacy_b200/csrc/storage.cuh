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

// `first`, `skip_mask`: only entities [first, n) that are NOT in skip_mask (the scene-graph frames that run the velocity
// system of the renderable leaves inside the emit tick the remaining leaves here).
__global__ void __launch_bounds__(kThreads) integrate_kernel(RowView rows, const uint64_t* mask_t, const uint64_t* mask_v,
                                                              const float* __restrict__ vel, float dt, uint32_t ticks,
                                                              uint32_t n, uint32_t first = 0, const uint64_t* skip_mask = nullptr) {
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!mask_bit(mask_t, i) || !mask_bit(mask_v, i)) return;
    if (skip_mask != nullptr && mask_bit(skip_mask, i)) return;
    const size_t o = (size_t)i * rows.st;
    float4 a0 = rows.t[o];
    const float* v = vel + (size_t)i * 3;
    integrate_ticks(a0, __ldg(v), __ldg(v + 1), __ldg(v + 2), dt, ticks);
    rows.t[o] = a0;
}

// ---- [SPEC] the spin system: a second device system next to the velocity system --------------------
//
// angle.z += angular_velocity * dt for every entity with Transform and Spin (idiom of system.rs:132-137), multiply and
// add individually rounded. The row keeps sin / cos of the half yaw (compose.cuh), so the system re-evaluates them —
// with the host libm's sinf / cosf, bit for bit — for exactly the entities it turns.
// Access sets (tb_system_access): reads Spin, writes Transform.angle.z — disjoint from the velocity system's
// (reads Velocity, writes Transform.translation), so the two may run concurrently.
__global__ void __launch_bounds__(kThreads) spin_kernel(RowView rows, const uint64_t* mask_t, const uint64_t* mask_w,
                                                         const float* __restrict__ spin, float dt, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!mask_bit(mask_t, i) || !mask_bit(mask_w, i)) return;
    const float yaw = addr(rows.yaw[i], mulr(__ldg(spin + i), dt));
    rows.yaw[i] = yaw;
    const float2 sc = yaw_sincos(yaw);
    reinterpret_cast<float*>(rows.b + (size_t)i * rows.sb)[0] = sc.x;      // B0.x
    reinterpret_cast<float*>(rows.b + (size_t)i * rows.sb + 1)[3] = sc.y;  // B1.w
}

// Entities the velocity system moves but no frame draws (Transform and Velocity, no Sprite): the frames that run the
// velocity system inside the emit only visit drawn entities, so these are ticked by a side pass — when there are any.
__global__ void __launch_bounds__(kThreads) count_hidden_movers_kernel(const uint64_t* mask_t, const uint64_t* mask_v, const uint64_t* mask_s,
                                                                        uint32_t nwords, uint32_t n, unsigned long long* out) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (w < nwords) {
        uint64_t m = mask_t[w] & mask_v[w] & ~mask_s[w];
        const uint32_t lo = w * 64u;
        if (lo + 64u > n) m &= (n > lo) ? ((1ull << (n - lo)) - 1ull) : 0ull;
        c = __popcll(m);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// ---- Ecs::query: presence AND -> packed ascending indices --------------------------------------

struct QueryParams {
    const uint64_t* masks[8];
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

// ---- structural changes on the device (crates/tuber-ecs/src/ecs.rs:100-112) ----------------------------

struct AllMasks {
    uint64_t* m[8];
    uint32_t count;
};

// Ecs::delete_by_ids: every component of the listed entities is removed (presence bit cleared in every store); the
// slot's generation is bumped so that handles issued before the delete stop resolving. Ids outside [0, entity_count)
// are counted in *bad (the reference would panic on the out-of-range index).
__global__ void __launch_bounds__(kThreads) delete_ids_kernel(const uint32_t* __restrict__ ids, uint32_t n, uint32_t entity_count,
                                                               AllMasks masks, uint32_t* gen, uint32_t* bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = ids[i];
    if (e >= entity_count) {
        atomicAdd(bad, 1u);
        return;
    }
    const unsigned long long bit = 1ull << (e & 63u);
    for (uint32_t k = 0; k < masks.count; ++k)
        atomicAnd(reinterpret_cast<unsigned long long*>(masks.m[k] + (e >> 6)), ~bit);
    atomicAdd(gen + e, 1u);
}

// Ecs::delete_by_query: the same for every entity matching the presence AND of `q` (one thread per bitset word).
__global__ void __launch_bounds__(kThreads) delete_query_kernel(const QueryParams q, AllMasks masks, uint32_t* gen,
                                                                 unsigned long long* deleted) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (w < q.nwords) {
        uint64_t m = query_word(q, w);
        if (m) {
            for (uint32_t k = 0; k < masks.count; ++k) masks.m[k][w] &= ~m;
            cnt = __popcll(m);
            while (m) {
                const uint32_t b = __ffsll((long long)m) - 1;
                m &= m - 1;
                gen[w * 64u + b] += 1u;
            }
        }
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(deleted, (unsigned long long)cnt);
}

// Generation-checked handles resolved on the device: alive = the slot's generation is still the handle's.
__global__ void __launch_bounds__(kThreads) handles_alive_kernel(const uint32_t* __restrict__ ids, const uint32_t* __restrict__ gens,
                                                                  uint32_t n, uint32_t entity_count, const uint32_t* __restrict__ gen,
                                                                  uint8_t* alive) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = ids[i];
    alive[i] = (e < entity_count && gen[e] == gens[i]) ? 1 : 0;
}

}  // namespace tb
