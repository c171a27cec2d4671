// sm_100a kernels of the tuber hot path. HBM-bound integer/float streaming
// work: no tensor cores. What matters here is 128-bit coalesced access, one pass over each
// byte, grids sized in multiples of the SM count, and a single stable counting/radix
// placement fused with the emit so sorted payloads are written exactly once.
//
//   convert_*            AoS host layouts  -> two-sector device rows           (upload)
//   prepare_flat         velocity system + sort key + key masks + histograms   (frame pass A)
//   prepare_level        same for one hierarchy level, plus parent * local     (frame pass A, hierarchy)
//   scan_hist            digit histograms -> bucket bases
//   build_batches        (layer, texture) histogram -> batch table
//   radix_scatter        one stable LSD pass of (short key, entity) pairs, decoupled look-back
//   emit                 final stable pass fused with compose + instance/vertex emission
//   query_*              presence-mask AND -> packed ascending entity indices  (Ecs::query)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "compose.cuh"
#include "plan.h"

namespace tb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kInvalidDigit = 0xFFFFFFFFu;
constexpr uint32_t kFlagLocal = 1u << 30;
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
    const size_t o = (size_t)(first + i) * rv.stride;
    // A0 = {t.x, t.y, t.z, c.z}
    rv.a[o] = make_float4(v[0], v[1], v[2], v[8]);
    // A1 = {s.z, angle.x, angle.y, texlayer}: keep .w
    float* a1 = reinterpret_cast<float*>(rv.a + o + 1);
    a1[0] = v[11];
    a1[1] = v[3];
    a1[2] = v[4];
    // B0 = {angle.z, c.x, c.y, s.x}
    rv.b[o] = make_float4(v[5], v[6], v[7], v[9]);
    // B1 = {s.y, size.w, size.h, 0}: keep .y .z
    float* b1 = reinterpret_cast<float*>(rv.b + o + 1);
    b1[0] = v[10];
}

__global__ void __launch_bounds__(kThreads) convert_sprites_kernel(const uint32_t* __restrict__ src, RowView rv,
                                                                    uint32_t first, uint32_t n, uint32_t* range_err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(src) + i);
    if (s.z > 0xFFFFu || s.w > 0xFFFEu) atomicAdd(range_err, 1u);
    const size_t o = (size_t)(first + i) * rv.stride;
    float* a1 = reinterpret_cast<float*>(rv.a + o + 1);
    a1[3] = __uint_as_float((s.z & 0xFFFFu) | (s.w << 16));
    float* b1 = reinterpret_cast<float*>(rv.b + o + 1);
    b1[1] = __uint_as_float(s.x);
    b1[2] = __uint_as_float(s.y);
    b1[3] = 0.0f;
}

// rows -> tb_transform (download for parity checks of the velocity system)
__global__ void __launch_bounds__(kThreads) unconvert_transforms_kernel(RowView rv, uint32_t first, uint32_t n,
                                                                         float* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Row r = load_row(rv, first + i);
    float* t = dst + (size_t)i * 12;
    t[0] = r.a0.x; t[1] = r.a0.y; t[2] = r.a0.z;
    t[3] = r.a1.y; t[4] = r.a1.z; t[5] = r.b0.x;
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

// ---- frame pass A: velocity system + key + masks + histograms --------------------------------

struct PrepParams {
    uint32_t n;            // entity_count
    RowView rows;
    const uint64_t* mask_t;  // presence masks, nullptr == present on every entity
    const uint64_t* mask_s;
    const uint64_t* mask_v;
    const float* vel;      // tb_velocity AoS (3 f32), nullptr when no Velocity store
    float dt;
    uint32_t integrate;    // run the velocity system in this pass
    uint32_t have_plan;    // 0: only masks + count (plan discovery)
    uint32_t write_keys;   // short keys to keys_out (multi-pass plans)
    uint32_t lt_mode;      // 0: none (aliases digit pass 0), 1: shared-memory bins, 2: global bins
    void* keys_out;
    uint32_t* hist;        // [kMaxPasses][256]
    uint32_t* lt_hist;     // [1 << ltbits]
    unsigned long long* stats;  // [0] OR of keys, [1] OR of ~keys, [2] renderable count
    uint32_t* tile_counts;  // optional [1 << bits[0]][emit_tiles]: pass-0 digit counts per emit tile
    uint32_t emit_tiles;
    uint32_t emit_tile_shift;  // log2 of the emit tile: 9 (CTA tiles of 512) or 6 (warp tiles of 64)
    uint32_t* key_hist;     // optional [1 << nbits]: histogram over the whole short key (multi-GPU placement)
    FramePlan plan;
};

constexpr int kPrepItems = 4;
constexpr uint32_t kEmitTileConst = 512;  // == kEmitTile (two per prepare tile)

struct PrepShared {
    uint32_t hist[kMaxPasses][256];
    uint32_t tcnt[2][2][256];  // [iteration parity][half of the 1024-entity tile][pass-0 digit]
    uint32_t lt[2048];
    unsigned long long k_or, k_and;
    uint32_t count;
};

// Per-thread running key statistics, folded into the block's once at the end of the kernel.
struct PrepAcc {
    unsigned long long k_or = 0ull, k_and = ~0ull;
    uint32_t count = 0;
};

// Accounts one entity: key statistics in registers, digit histograms in shared memory, optional
// short key write. Must be called by all 32 lanes of a warp (it contains warp collectives).
//   small == true  (single pass, <= 32 buckets): one match.any per entity aggregates the warp's
//     histogram update, and the per-emit-tile counts go straight to global memory with one RED per
//     distinct digit per warp - no block barrier anywhere in the key pass;
//   small == false: plain shared atomics (many distinct digits, little contention), per-tile counts
//     through `tcnt` in shared memory when the caller wants them.
__device__ __forceinline__ void prep_account(const PrepParams& p, PrepShared& sh, PrepAcc& acc, bool renderable,
                                             uint64_t key, uint32_t i, bool small, uint32_t emit_tile, uint32_t* tcnt) {
    if (renderable) {
        acc.k_or |= key;
        acc.k_and &= key;
        acc.count += 1;
    }
    if (!p.have_plan) return;
    const uint64_t sk = pext_runs(key, p.plan.pext);
    if (small) {
        const uint32_t d = renderable ? (uint32_t)sk & ((1u << p.plan.bits[0]) - 1u) : kInvalidDigit;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (d != kInvalidDigit && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) {
            const uint32_t c = __popc(peers);
            atomicAdd(&sh.hist[0][d], c);
            if (p.tile_counts != nullptr) atomicAdd(&p.tile_counts[(size_t)d * p.emit_tiles + emit_tile], c);
            if (p.key_hist != nullptr) atomicAdd(&p.key_hist[d], c);  // single pass: the digit is the short key
            if (p.lt_mode == 1) atomicAdd(&sh.lt[d >> p.plan.zbits], c);
            else if (p.lt_mode == 2) atomicAdd(&p.lt_hist[d >> p.plan.zbits], c);
        }
        return;
    }
    if (!renderable) return;
    if (p.key_hist != nullptr) atomicAdd(&p.key_hist[(uint32_t)sk], 1u);
    for (uint32_t q = 0; q < p.plan.npasses; ++q) {
        const uint32_t d = (uint32_t)(sk >> p.plan.shift[q]) & ((1u << p.plan.bits[q]) - 1u);
        atomicAdd(&sh.hist[q][d], 1u);
        if (q == 0 && tcnt != nullptr) atomicAdd(&tcnt[d], 1u);
    }
    if (p.lt_mode == 1) atomicAdd(&sh.lt[(uint32_t)(sk >> p.plan.zbits)], 1u);
    else if (p.lt_mode == 2) atomicAdd(&p.lt_hist[(uint32_t)(sk >> p.plan.zbits)], 1u);
    if (p.write_keys) {
        if (p.plan.key64) reinterpret_cast<uint64_t*>(p.keys_out)[i] = sk;
        else reinterpret_cast<uint32_t*>(p.keys_out)[i] = (uint32_t)sk;
    }
}

// The single-pass few-bucket plan the `small` accounting path serves.
__device__ __forceinline__ bool prep_small(const PrepParams& p) {
    return p.have_plan && p.plan.npasses == 1 && p.plan.bits[0] <= 5 && !p.write_keys;
}

__device__ __forceinline__ void prep_init(const PrepParams& p, PrepShared& sh) {
    for (uint32_t k = threadIdx.x; k < kMaxPasses * 256; k += blockDim.x) (&sh.hist[0][0])[k] = 0;
    for (uint32_t k = threadIdx.x; k < 2048; k += blockDim.x) sh.lt[k] = 0;
    if (threadIdx.x == 0) {
        sh.k_or = 0ull;
        sh.k_and = ~0ull;
        sh.count = 0;
    }
    __syncthreads();
}

__device__ __forceinline__ void prep_flush(const PrepParams& p, PrepShared& sh, const PrepAcc& acc) {
    // fold the per-thread key statistics: warp reduce, then one shared atomic per warp
    {
        const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)acc.k_or);
        const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(acc.k_or >> 32));
        const uint32_t nlo = __reduce_and_sync(0xffffffffu, (uint32_t)acc.k_and);
        const uint32_t nhi = __reduce_and_sync(0xffffffffu, (uint32_t)(acc.k_and >> 32));
        const uint32_t cnt = __reduce_add_sync(0xffffffffu, acc.count);
        if ((threadIdx.x & 31) == 0 && cnt) {
            atomicOr(&sh.k_or, ((unsigned long long)hi << 32) | lo);
            atomicAnd(&sh.k_and, ((unsigned long long)nhi << 32) | nlo);
            atomicAdd(&sh.count, cnt);
        }
    }
    __syncthreads();
    if (p.have_plan) {
        for (uint32_t q = 0; q < p.plan.npasses; ++q) {
            const uint32_t R = 1u << p.plan.bits[q];
            for (uint32_t d = threadIdx.x; d < R; d += blockDim.x) {
                const uint32_t c = sh.hist[q][d];
                if (c) atomicAdd(&p.hist[q * 256 + d], c);
            }
        }
        if (p.lt_mode == 1) {
            const uint32_t B = 1u << p.plan.ltbits;
            for (uint32_t d = threadIdx.x; d < B; d += blockDim.x) {
                const uint32_t c = sh.lt[d];
                if (c) atomicAdd(&p.lt_hist[d], c);
            }
        }
    }
    if (threadIdx.x == 0 && sh.count) {
        atomicOr(&p.stats[0], sh.k_or);
        atomicOr(&p.stats[1], ~sh.k_and);  // OR of ~key: all-zero scratch is the neutral element
        atomicAdd(&p.stats[2], (unsigned long long)sh.count);
    }
}

// Flat scenes (no Parent anywhere): per entity, the velocity system
// (translation += velocity * dt, [SPEC]; idiom of system.rs:132-137) then the sort key of
// (&Transform, &Sprite) matches (query.rs:109-120 presence AND).
__global__ void __launch_bounds__(kThreads, 4) prepare_flat_kernel(const PrepParams p) {
    __shared__ PrepShared sh;
    prep_init(p, sh);
    const uint32_t tile_elems = kThreads * kPrepItems;
    const uint32_t ntiles = (p.n + tile_elems - 1) / tile_elems;
    const bool do_int = p.integrate && p.vel != nullptr;
    const bool small = prep_small(p);
    const bool tiles_out = !small && p.have_plan && p.tile_counts != nullptr;  // shared-memory per-tile counts
    const uint32_t R0 = 1u << p.plan.bits[0];
    PrepAcc acc;
    uint32_t par = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, par ^= 1u) {
        const uint32_t base = tile * tile_elems + threadIdx.x;
        if (tiles_out) {
            for (uint32_t d = threadIdx.x; d < 2 * 256; d += kThreads) (&sh.tcnt[par][0][0])[d] = 0;
        }
        float4 a0[kPrepItems], a1[kPrepItems];
        float vx[kPrepItems], vy[kPrepItems], vz[kPrepItems];
        bool has_t[kPrepItems], rend[kPrepItems], has_v[kPrepItems];
        // Every load of the tile is issued before anything is consumed: rows and velocities exist
        // (zero-filled) for every id below max_entities, so they do not wait for the presence bits.
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t i = base + k * kThreads;
            const bool inb = i < p.n;
            a0[k] = a1[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            vx[k] = vy[k] = vz[k] = 0.0f;
            has_t[k] = rend[k] = has_v[k] = false;
            if (inb) {
                const size_t o = (size_t)i * p.rows.stride;
                a0[k] = p.rows.a[o];
                a1[k] = p.rows.a[o + 1];
                if (do_int) {
                    const float* v = p.vel + (size_t)i * 3;
                    vx[k] = __ldg(v);
                    vy[k] = __ldg(v + 1);
                    vz[k] = __ldg(v + 2);
                    has_v[k] = mask_bit(p.mask_v, i);
                }
                has_t[k] = mask_bit(p.mask_t, i);
                rend[k] = mask_bit(p.mask_s, i);
            }
        }
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t i = base + k * kThreads;
            rend[k] = rend[k] && has_t[k];
            if (do_int && has_t[k] && has_v[k]) {
                a0[k].x = addr(a0[k].x, mulr(vx[k], p.dt));
                a0[k].y = addr(a0[k].y, mulr(vy[k], p.dt));
                a0[k].z = addr(a0[k].z, mulr(vz[k], p.dt));
                p.rows.a[(size_t)i * p.rows.stride] = a0[k];
            }
        }
        if (tiles_out) __syncthreads();
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t i = base + k * kThreads;
            uint64_t key = 0;
            if (rend[k]) key = flat_key(p.rows, i, a0[k], a1[k]);
            // this 1024-entity tile is two 512-entity emit tiles (shared-memory counts) or
            // sixteen 64-entity warp tiles (direct global counts of the small path)
            prep_account(p, sh, acc, rend[k], key, i, small, i >> p.emit_tile_shift, tiles_out ? sh.tcnt[par][k >> 1] : nullptr);
        }
        if (tiles_out) {
            __syncthreads();
            for (uint32_t d = threadIdx.x; d < 2 * R0; d += kThreads) {
                const uint32_t h = d / R0, dd = d - h * R0;
                const uint32_t et = tile * 2 + h;
                if (et < p.emit_tiles) p.tile_counts[(size_t)dd * p.emit_tiles + et] = sh.tcnt[par][h][dd];
            }
        }
    }
    prep_flush(p, sh, acc);
}

// ---- hierarchy: one level of world = world[parent] * local ------------------------------------

struct LevelParams {
    PrepParams prep;
    const uint32_t* parent;      // per entity, TB_NO_PARENT for roots (already sanitised: see levels)
    const uint32_t* list;        // entities of this level, nullptr when the level is an id range
    uint32_t first;              // first entity id (range mode) / first list slot
    uint32_t count;              // entities in this level
    uint32_t level;              // 0 == roots
    uint32_t global_first;       // id of this shard's entity 0 in the parents' index space
    WorldRow* world;             // per entity
};

// Level L of the scene graph. Siblings are usually contiguous, so lanes with the same parent
// elect one loader and share the parent's world row by warp shuffle.
__global__ void __launch_bounds__(kThreads) prepare_level_kernel(const LevelParams lp) {
    __shared__ PrepShared sh;
    const PrepParams& p = lp.prep;
    prep_init(p, sh);
    PrepAcc acc;
    const uint32_t nchunks = (lp.count + kThreads - 1) / kThreads;
    for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const uint32_t slot = chunk * kThreads + threadIdx.x;
        bool live = slot < lp.count;
        uint32_t e = 0;
        if (live) {
            e = lp.list ? __ldg(lp.list + lp.first + slot) : lp.first + slot;
            live = mask_bit(p.mask_t, e);  // level 0 also lists entities without a Transform
        }
        Row r;
        Affine w;
        uint32_t par = 0xFFFFFFFFu;
        if (live) {
            r = load_row(p.rows, e);
            if (p.integrate && p.vel != nullptr && mask_bit(p.mask_v, e)) {
                const float* v = p.vel + (size_t)e * 3;
                const float vx = __ldg(v), vy = __ldg(v + 1), vz = __ldg(v + 2);
                r.a0.x = addr(r.a0.x, mulr(vx, p.dt));
                r.a0.y = addr(r.a0.y, mulr(vy, p.dt));
                r.a0.z = addr(r.a0.z, mulr(vz, p.dt));
                p.rows.a[(size_t)e * p.rows.stride] = r.a0;
            }
            w = compose_local(r);
            if (lp.level > 0) par = __ldg(lp.parent + e) - lp.global_first;
        }
        if (lp.level > 0) {
            // warp-shuffle parent broadcast
            const uint32_t peers = __match_any_sync(0xffffffffu, par);
            const int leader = __ffs(peers) - 1;
            float4 p0 = make_float4(0, 0, 0, 0), p1 = p0, p2 = p0;
            if (live && (int)(threadIdx.x & 31) == leader) {
                const float4* pw = reinterpret_cast<const float4*>(lp.world + par);
                p0 = pw[0];
                p1 = pw[1];
                p2 = pw[2];
            }
            Affine P;
            P.r0.x = __shfl_sync(0xffffffffu, p0.x, leader); P.r0.y = __shfl_sync(0xffffffffu, p0.y, leader);
            P.r0.z = __shfl_sync(0xffffffffu, p0.z, leader); P.r0.w = __shfl_sync(0xffffffffu, p0.w, leader);
            P.r1.x = __shfl_sync(0xffffffffu, p1.x, leader); P.r1.y = __shfl_sync(0xffffffffu, p1.y, leader);
            P.r1.z = __shfl_sync(0xffffffffu, p1.z, leader); P.r1.w = __shfl_sync(0xffffffffu, p1.w, leader);
            P.r2.x = __shfl_sync(0xffffffffu, p2.x, leader); P.r2.y = __shfl_sync(0xffffffffu, p2.y, leader);
            P.r2.z = __shfl_sync(0xffffffffu, p2.z, leader); P.r2.w = __shfl_sync(0xffffffffu, p2.w, leader);
            if (live) w = mul_affine(P, w);
        }
        bool rend = false;
        uint64_t key = 0;
        if (live) {
            WorldRow wr;
            wr.r0 = w.r0;
            wr.r1 = w.r1;
            wr.r2 = w.r2;
            wr.s = make_float4(r.b1.y, r.b1.z, r.a1.w, 0.0f);
            float4* dst = reinterpret_cast<float4*>(lp.world + e);
            dst[0] = wr.r0;
            dst[1] = wr.r1;
            dst[2] = wr.r2;
            dst[3] = wr.s;
            rend = mask_bit(p.mask_s, e);
            if (rend) key = world_key(wr.r2, wr.s);
        }
        prep_account(p, sh, acc, rend, key, e, false, 0u, nullptr);
    }
    prep_flush(p, sh, acc);
}

// ---- histogram -> bucket bases; (layer, texture) histogram -> batch table -----------------------

// One block per radix pass: exclusive scan of its 256-bin digit histogram, in place.
__global__ void __launch_bounds__(256) scan_hist_kernel(uint32_t* hist) {
    __shared__ uint32_t wsum[8];
    uint32_t* h = hist + blockIdx.x * 256;
    const uint32_t v = h[threadIdx.x];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t add = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) add += wsum[w];
    h[threadIdx.x] = x + add - v;
}

// First output position of every (digit, tile) of one pass from its per-tile digit counts
// (reduce-then-scan radix placement: no inter-CTA waiting anywhere). counts and tile_off are
// digit-major [R][ntiles]; hist is the pass's raw 256-bin digit histogram. Two launches over a
// grid of (chunks of 1024 tiles, R): chunk sums, then every block adds the sums of the chunks
// before its own (a few hundred L2-resident words) and scans its chunk.
__global__ void __launch_bounds__(1024) tile_chunk_sums_kernel(const uint32_t* __restrict__ counts, uint32_t ntiles,
                                                                uint32_t nchunks, uint32_t* csum) {
    __shared__ uint32_t s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const uint32_t d = blockIdx.y, t = blockIdx.x * 1024u + threadIdx.x;
    uint32_t v = t < ntiles ? counts[(size_t)d * ntiles + t] : 0u;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s, v);
    __syncthreads();
    if (threadIdx.x == 0) csum[(size_t)d * nchunks + blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) tile_scan_kernel(const uint32_t* __restrict__ counts, uint32_t ntiles,
                                                          const uint32_t* __restrict__ hist, const uint32_t* __restrict__ csum,
                                                          uint32_t nchunks, uint32_t* tile_off) {
    __shared__ uint32_t s[1024];
    __shared__ uint32_t s_base;
    const uint32_t d = blockIdx.y, chunk = blockIdx.x;
    const uint32_t* row = counts + (size_t)d * ntiles;
    // base = all smaller digits (global) + this digit's chunks before this one
    uint32_t pre = 0;
    for (uint32_t k = threadIdx.x; k < d; k += 1024) pre += hist[k];
    for (uint32_t c = threadIdx.x; c < chunk; c += 1024) pre += csum[(size_t)d * nchunks + c];
    const uint32_t t = chunk * 1024u + threadIdx.x;
    const uint32_t v = t < ntiles ? row[t] : 0u;
    pre = __reduce_add_sync(0xffffffffu, pre);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && pre) atomicAdd(&s_base, pre);
    // inclusive scan of v over the block
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = s[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (threadIdx.x >= (uint32_t)o) w += y;
        }
        s[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t wbase = (threadIdx.x >> 5) ? s[(threadIdx.x >> 5) - 1] : 0u;
    if (t < ntiles) tile_off[(size_t)d * ntiles + t] = s_base + wbase + x - v;
}

// Per-tile digit counts of one pass's input (positions in the pass's input order).
// counts is digit-major [R][ntiles]. One block per tile.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) tile_hist_kernel(const KeyT* __restrict__ keys, uint32_t n,
                                                              const unsigned long long* count_ptr,
                                                              const uint64_t* mask_t, const uint64_t* mask_s, uint32_t first,
                                                              uint32_t tile_elems, uint32_t shift, uint32_t bits,
                                                              uint32_t ntiles, uint32_t* counts) {
    __shared__ uint32_t h[256];
    const uint32_t R = 1u << bits, mask = R - 1u;
    const uint32_t n_eff = count_ptr ? min(n, (uint32_t)*count_ptr) : n;
    for (uint32_t d = threadIdx.x; d < R; d += kThreads) h[d] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * tile_elems;
    for (uint32_t k = threadIdx.x; k < tile_elems; k += kThreads) {
        const uint32_t pos = base + k;
        bool valid = pos < n_eff;
        if (valid && first) valid = mask_bit(mask_t, pos) && mask_bit(mask_s, pos);
        if (valid) atomicAdd(&h[(uint32_t)(keys[pos] >> shift) & mask], 1u);
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < R; d += kThreads) counts[(size_t)d * ntiles + blockIdx.x] = h[d];
}

struct BatchOut {
    uint32_t layer, texture, first_instance, count;
};

// Every non-empty (layer, texture) bin is one batch: first = exclusive prefix, count = bin.
// `bins` has 1 << ltbits entries (when it aliases digit pass 0 it must be read BEFORE
// scan_hist_kernel rewrites it: the host orders the launches). Single block.
__global__ void __launch_bounds__(1024) build_batches_kernel(const uint32_t* __restrict__ bins, uint32_t nbins,
                                                              FramePlan plan, BatchOut* out, uint32_t max_batches,
                                                              uint32_t* num_batches) {
    __shared__ uint32_t s_cnt[1024], s_sum[1024];
    const uint32_t per = (nbins + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per;
    uint32_t nz = 0, sum = 0;
    for (uint32_t b = b0; b < b0 + per && b < nbins; ++b) {
        const uint32_t c = bins[b];
        nz += c != 0;
        sum += c;
    }
    s_cnt[threadIdx.x] = nz;
    s_sum[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        uint32_t c = 0, s = 0;
        if (threadIdx.x >= o) {
            c = s_cnt[threadIdx.x - o];
            s = s_sum[threadIdx.x - o];
        }
        __syncthreads();
        s_cnt[threadIdx.x] += c;
        s_sum[threadIdx.x] += s;
        __syncthreads();
    }
    uint32_t slot = s_cnt[threadIdx.x] - nz;
    uint32_t first = s_sum[threadIdx.x] - sum;
    for (uint32_t b = b0; b < b0 + per && b < nbins; ++b) {
        const uint32_t c = bins[b];
        if (c) {
            if (slot < max_batches) {
                const uint64_t key = plan.constant | pdep_runs((uint64_t)b << plan.zbits, plan.pext);
                BatchOut o;
                o.layer = (uint32_t)(key >> 48);
                o.texture = (uint32_t)(key >> 32) & 0xFFFFu;
                o.first_instance = first;
                o.count = c;
                out[slot] = o;
            }
            ++slot;
            first += c;
        }
    }
    if (threadIdx.x == 1023) *num_batches = s_cnt[1023];
}

// ---- stable tile ranking with decoupled look-back (shared by scatter and emit) ------------------

struct RankShared {
    uint32_t wcnt[kWarps * 256];  // [warp][R] per-warp digit counters, then per-warp exclusive prefixes
    uint32_t dstart[256];         // first tile slot of each digit
    uint32_t dbase[256];          // destination of digit d's slot s is dbase[d] + s
    uint32_t wsum[kWarps];
    uint32_t tile;
    uint32_t tile_count;
};

struct PassParams {
    uint32_t n;              // positions to visit
    const unsigned long long* count_ptr;  // when non-null only positions < *count_ptr hold elements
                                          // (passes after the first: the renderable count lives on the device)
    uint32_t shift, bits;    // this pass's digit of the short key
    const uint32_t* gbase;   // exclusive scan of this pass's digit histogram
    uint32_t* status;        // [tiles][1 << bits] look-back words, zeroed per frame
    uint32_t* tile_counter;  // zeroed per frame
    const uint32_t* tile_off;  // optional [1 << bits][ntiles]: first output position of each (digit, tile),
                               // precomputed from per-tile counts; replaces the look-back
    uint32_t ntiles;           // tiles of this pass (row length of tile_off)
};

// Zeroes the per-warp digit counters of the next rank_tile call; a __syncthreads() must
// separate it from that call.
__device__ __forceinline__ void rank_reset(RankShared& rs, uint32_t bits) {
    const uint32_t R = 1u << bits;
    for (uint32_t k = threadIdx.x; k < kWarps * R; k += kThreads) rs.wcnt[k] = 0;
}

// Ranks ITEMS digits per thread (warp-striped order: item k of lane l in warp w is element
// w*ITEMS*32 + k*32 + l of the tile) into stable tile slots and leaves dbase[] ready: from the
// precomputed tile offsets when the pass has them, else by publishing the tile's digit counts
// and looking back over the preceding tiles (decoupled look-back). Invalid items (kInvalidDigit)
// get no slot. Needs rank_reset() + __syncthreads() before, and a __syncthreads() before the
// next rank_reset().
template <int ITEMS>
__device__ __forceinline__ void rank_tile(RankShared& rs, const PassParams& pp, const uint32_t tile,
                                          const uint32_t (&digit)[ITEMS], uint32_t (&slot)[ITEMS]) {
    const uint32_t R = 1u << pp.bits;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* wc = rs.wcnt + warp * R;
    const uint32_t lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = digit[k];
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)lane == leader && d != kInvalidDigit) {
            base = wc[d];
            wc[d] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        slot[k] = base + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, tile total
    uint32_t total = 0;
    if (threadIdx.x < R) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = rs.wcnt[w * R + threadIdx.x];
            rs.wcnt[w * R + threadIdx.x] = total;
            total += c;
        }
    }
    const bool lookback = pp.tile_off == nullptr;
    // publish the local count as early as possible
    uint32_t* st = pp.status + (size_t)tile * R + threadIdx.x;
    if (lookback && threadIdx.x < R) st_relaxed_u32(st, (tile == 0 ? kFlagIncl : kFlagLocal) | total);
    // block exclusive scan of the totals -> dstart (all digits live in warp 0 when R <= 32)
    uint32_t x = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    uint32_t add = 0;
    if (R > 32) {
        if (lane == 31) rs.wsum[warp] = x;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            if ((uint32_t)w < warp) add += rs.wsum[w];
        if (threadIdx.x == kThreads - 1) rs.tile_count = x + add;
    } else if (threadIdx.x == 31) {
        rs.tile_count = x;
    }
    const uint32_t dstart = x + add - total;
    if (threadIdx.x < R) {
        rs.dstart[threadIdx.x] = dstart;
        if (lookback) {
            uint32_t excl = 0;
            if (tile > 0) {
                for (int t = (int)tile - 1; t >= 0; --t) {
                    uint32_t v;
                    do {
                        v = ld_relaxed_u32(pp.status + (size_t)t * R + threadIdx.x);
                    } while ((v >> 30) == 0u);
                    excl += v & kCountMask;
                    if ((v >> 30) == 2u) break;
                }
                st_relaxed_u32(st, kFlagIncl | (excl + total));
            }
            rs.dbase[threadIdx.x] = pp.gbase[threadIdx.x] + excl - dstart;
        } else {
            rs.dbase[threadIdx.x] = __ldg(pp.tile_off + (size_t)threadIdx.x * pp.ntiles + tile) - dstart;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = digit[k];
        if (d != kInvalidDigit) slot[k] += rs.dstart[d] + rs.wcnt[warp * R + d];
    }
}

__device__ __forceinline__ void claim_tile(RankShared& rs, const PassParams& pp) {
    if (threadIdx.x == 0) rs.tile = atomicAdd(pp.tile_counter, 1u);
    rank_reset(rs, pp.bits);
    __syncthreads();
}

// ---- one LSD scatter pass of (short key, entity) pairs ------------------------------------------

template <typename KeyT>
struct ScatterParams {
    PassParams pass;
    const KeyT* keys_in;
    const uint32_t* idx_in;  // nullptr on the first pass: the entity is the position
    const uint64_t* mask_t;  // first pass only: which positions are renderable
    const uint64_t* mask_s;
    KeyT* keys_out;
    uint32_t* idx_out;
};

constexpr int kScatterItems = 12;

template <typename KeyT>
__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(const ScatterParams<KeyT> sp) {
    constexpr int ITEMS = kScatterItems;
    constexpr uint32_t TILE = kThreads * ITEMS;
    __shared__ RankShared rs;
    __shared__ KeyT s_key[TILE];
    __shared__ uint32_t s_idx[TILE];
    claim_tile(rs, sp.pass);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile_base = rs.tile * TILE;
    const uint32_t mask = (1u << sp.pass.bits) - 1u;
    KeyT key[ITEMS];
    uint32_t idx[ITEMS], digit[ITEMS], slot[ITEMS];
    const uint32_t n_eff = sp.pass.count_ptr ? min(sp.pass.n, (uint32_t)*sp.pass.count_ptr) : sp.pass.n;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t pos = tile_base + warp * (ITEMS * 32) + k * 32 + lane;
        bool valid = pos < n_eff;
        if (valid && sp.idx_in == nullptr) valid = mask_bit(sp.mask_t, pos) && mask_bit(sp.mask_s, pos);
        key[k] = 0;
        idx[k] = pos;
        if (valid) {
            key[k] = sp.keys_in[pos];
            if (sp.idx_in != nullptr) idx[k] = sp.idx_in[pos];
        }
        digit[k] = valid ? ((uint32_t)(key[k] >> sp.pass.shift) & mask) : kInvalidDigit;
    }
    rank_tile<ITEMS>(rs, sp.pass, rs.tile, digit, slot);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (digit[k] != kInvalidDigit) {
            s_key[slot[k]] = key[k];
            s_idx[slot[k]] = idx[k];
        }
    }
    __syncthreads();
    const uint32_t cnt = rs.tile_count;
    for (uint32_t s = threadIdx.x; s < cnt; s += kThreads) {
        const KeyT kk = s_key[s];
        const uint32_t d = (uint32_t)(kk >> sp.pass.shift) & mask;
        const uint32_t dst = rs.dbase[d] + s;
        sp.keys_out[dst] = kk;
        sp.idx_out[dst] = s_idx[s];
    }
}

// ---- final pass fused with compose + emit -------------------------------------------------------

struct InstanceOut {
    float4 c0, c1, c2, c3, s;  // column-major model (matrix.rs:219-251) + {size.w, size.h, texture, layer}
};
static_assert(sizeof(InstanceOut) == 80, "instance is 80 bytes");

template <typename KeyT>
struct EmitParams {
    PassParams pass;
    const KeyT* keys_in;     // nullptr when `first` and keys are recomputed from the rows
    const uint32_t* idx_in;  // nullptr: the entity is the position (single-pass plans)
    const uint64_t* mask_t;
    const uint64_t* mask_s;
    RowView rows;
    const WorldRow* world;   // non-null: hierarchy path, world rows instead of composing
    FramePlan plan;
    float4* instances;       // 5 float4 per position
    float4* vertices;        // 4 float4 per position
    uint32_t* order;         // entity per position
    uint32_t* texlayer;      // optional per-position texture|layer<<16 (generic batch path)
    uint32_t global_first;   // added to the entity ids written to `order`
    uint32_t staged;         // stage outputs through shared memory for full-line stores
    const uint32_t* remap;   // optional [1 << nbits]: added to the shard-local position of every element
                             // with that short key -> position in the global (all ranks) draw order
    // single-kernel frames (emit_small_kernel<KeyT, true>): the velocity system and the key
    // statistics run inside the emit, placement comes from the cached per-tile offsets
    const float* vel;
    const uint64_t* mask_v;
    float dt;
    uint32_t integrate;
    unsigned long long* stats;  // [0] OR of keys, [1] OR of ~keys, [2] renderable count
};

constexpr int kEmitItems = 2;
constexpr uint32_t kEmitTile = kThreads * kEmitItems;  // 512 entities per tile
constexpr int kStageF4 = 5;  // per-lane staging slot: 80 B (one instance, or 64 B of vertices)

// Dynamic shared memory of emit_kernel: ~108 KB, two persistent CTAs per SM.
struct EmitShared {
    float4 rows[2][kEmitTile * 4];      // 2 x 32 KB: this tile's rows and the prefetched next tile's,
                                        // 16-byte chunks XOR-swizzled (row_slot) against bank conflicts
    unsigned long long keys[2][kEmitTile];  // short keys of in-order passes that read them (hierarchy)
    unsigned long long masks[2][2][8];  // presence words (Transform, Sprite) of the tile
    RankShared rs;
    uint32_t s_dst[kEmitTile];          // output position per sorted tile slot
    uint32_t s_ent[kEmitTile];          // entity id per sorted tile slot
    uint16_t s_li[kEmitTile];           // tile-local item index per sorted tile slot
    float4 stage[kWarps][32 * kStageF4];  // per-warp staging for full-line stores
    unsigned long long mbar[2];         // completion barriers of the TMA bulk copies, one per buffer
    uint32_t s_tile[3];                 // ring of claimed tile ids
};

// 16-byte chunk c (0..3) of tile-local row li lives at rows[row_slot(li, c)]: the 128-byte
// swizzle (chunk index XOR line index), so both consecutive rows and the scattered rows of one
// sort bucket spread over all eight 16-byte bank groups.
__device__ __forceinline__ uint32_t row_slot(uint32_t li, uint32_t c) { return ((li << 2) | c) ^ ((li >> 1) & 7u); }

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

__device__ __forceinline__ void make_outputs(const Affine& w, float sw, float shh, uint32_t tl, InstanceOut& inst,
                                             float4 (&vtx)[4]) {
    inst.c0 = make_float4(w.r0.x, w.r1.x, w.r2.x, 0.0f);
    inst.c1 = make_float4(w.r0.y, w.r1.y, w.r2.y, 0.0f);
    inst.c2 = make_float4(w.r0.z, w.r1.z, w.r2.z, 0.0f);
    inst.c3 = make_float4(w.r0.w, w.r1.w, w.r2.w, 1.0f);
    inst.s = make_float4(sw, shh, __uint_as_float(tl & 0xFFFFu), __uint_as_float(tl >> 16));
    // corners (0,0) (w,0) (w,h) (0,h), transform_vec3 (matrix.rs:160-171)
    const float cx[4] = {0.0f, sw, sw, 0.0f};
    const float cy[4] = {0.0f, 0.0f, shh, shh};
    const uint32_t uv[4] = {0x00000000u, 0x00000001u, 0x00010001u, 0x00010000u};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        vtx[k] = make_float4(corner_coord(w.r0.x, w.r0.y, w.r0.w, cx[k], cy[k]),
                             corner_coord(w.r1.x, w.r1.y, w.r1.w, cx[k], cy[k]),
                             corner_coord(w.r2.x, w.r2.y, w.r2.w, cx[k], cy[k]), __uint_as_float(uv[k]));
    }
}

// Writes one warp's 32 results (lane i holds the instance and vertices of output position dst)
// as whole 128-byte lines: staged in the warp's shared buffer, then streamed out 16 bytes per lane.
// Instances use a 20-word lane stride (conflict-free 128-bit stores; the read-out is linear);
// vertices a 16-word stride with the chunk index XORed by bits of the lane (conflict-free both ways).
__device__ __forceinline__ void staged_store(float4* stage, uint32_t lane, bool live, uint32_t nlive, uint32_t dst,
                                             const InstanceOut& inst, const float4 (&vtx)[4], float4* instances,
                                             float4* vertices) {
    float4* my = stage + lane * kStageF4;
    if (live) {
        my[0] = inst.c0; my[1] = inst.c1; my[2] = inst.c2; my[3] = inst.c3; my[4] = inst.s;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const uint32_t q = j * 32 + lane;  // 16-byte chunk of the warp's 2560-byte instance stream
        const uint32_t item = q / 5, part = q - item * 5;
        const uint32_t d = __shfl_sync(0xffffffffu, dst, item & 31);
        if (item < nlive) st_cs_f4(instances + (size_t)d * 5 + part, stage[q]);
    }
    __syncwarp();
    const uint32_t sw = (lane >> 1) & 3u;
    if (live) {
        float4* mv = stage + lane * 4;
        mv[0 ^ sw] = vtx[0]; mv[1 ^ sw] = vtx[1]; mv[2 ^ sw] = vtx[2]; mv[3 ^ sw] = vtx[3];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t q = j * 32 + lane;
        const uint32_t item = q >> 2, part = q & 3u;
        const uint32_t d = __shfl_sync(0xffffffffu, dst, item & 31);
        if (item < nlive) st_cs_f4(vertices + (size_t)d * 4 + part, stage[item * 4 + (part ^ ((item >> 1) & 3u))]);
    }
    __syncwarp();
}

// Final radix pass fused with compose + emit. Persistent CTAs (two per SM) claim 512-position
// tiles in order from an atomic counter:
//   1. the tile's rows come into shared memory asynchronously (cp.async, 16-byte chunks into a
//      bank-swizzled layout), so no register holds a row while the tile is ranked: when the pass
//      reads positions in entity order (single-pass plans) the NEXT tile's rows, presence words
//      and keys are prefetched while this tile is processed; otherwise every row is gathered
//      while the tile is ranked;
//   2. stable rank of the tile's digits -> final output position per item, from precomputed
//      per-tile offsets (reduce-then-scan) or by decoupled look-back;
//   3. warps walk the tile in sorted order (consecutive slots of a digit are consecutive output
//      positions), compose from shared memory, stage 32 results and write whole 128-byte lines.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads, 2) emit_kernel(const EmitParams<KeyT> ep) {
    constexpr int ITEMS = kEmitItems;
    constexpr uint32_t TILE = kEmitTile;
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    EmitShared& sm = *reinterpret_cast<EmitShared*>(emit_smem_raw);
    RankShared& rs = sm.rs;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool in_order = ep.idx_in == nullptr;  // positions are entity ids: contiguous rows
    const bool from_world = ep.world != nullptr;
    const bool keys_given = ep.keys_in != nullptr;
    const uint32_t n_eff = ep.pass.count_ptr ? min(ep.pass.n, (uint32_t)*ep.pass.count_ptr) : ep.pass.n;
    const uint32_t ntiles = (ep.pass.n + TILE - 1) / TILE;
    const uint32_t mask = (1u << ep.pass.bits) - 1u;

    // Shared-memory row layouts. In-order passes land rows with TMA bulk copies, i.e. linearly
    // (sector A of item li at rows[li*sstride], sector B at rows[boff + li*sstride]); gathered rows
    // are placed chunk by chunk with cp.async into the bank-swizzled layout (row_slot).
    const bool split = in_order && !from_world && ep.rows.stride == 2;
    const uint32_t sstride = split ? 2u : 4u;
    const uint32_t boff = split ? 2u * TILE : 2u;
    auto chunk = [&](uint32_t li, uint32_t c) -> uint32_t {
        if (!in_order) return row_slot(li, c);
        return (c < 2 ? 0u : boff - 2u) + li * sstride + c;
    };

    // in-order passes: TMA bulk copies of one tile's rows / presence words / keys into buffer `buf`
    // (issued by one thread, completion counted in bytes on the buffer's mbarrier)
    auto fetch_tile = [&](uint32_t tile, uint32_t buf) {
        const uint32_t base = tile * TILE;
        const uint32_t cnt = min(TILE, n_eff - base);
        const uint32_t mwords = ((cnt + 127u) >> 7) * 2u;  // presence words, padded to 16 bytes
        const uint32_t kbytes = keys_given ? ((cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u) : 0u;
        const uint32_t mbytes = ((ep.mask_t ? 1u : 0u) + (ep.mask_s ? 1u : 0u)) * mwords * 8u;
        float4* dst = sm.rows[buf];
        mbar_expect_tx(&sm.mbar[buf], cnt * 64u + mbytes + kbytes);
        if (from_world) {
            bulk_g2s(dst, ep.world + base, cnt * 64u, &sm.mbar[buf]);
        } else if (split) {
            bulk_g2s(dst, ep.rows.a + (size_t)base * 2, cnt * 32u, &sm.mbar[buf]);
            bulk_g2s(dst + 2 * TILE, ep.rows.b + (size_t)base * 2, cnt * 32u, &sm.mbar[buf]);
        } else {
            bulk_g2s(dst, ep.rows.a + (size_t)base * 4, cnt * 64u, &sm.mbar[buf]);
        }
        if (ep.mask_t) bulk_g2s(sm.masks[buf][0], ep.mask_t + (base >> 6), mwords * 8u, &sm.mbar[buf]);
        if (ep.mask_s) bulk_g2s(sm.masks[buf][1], ep.mask_s + (base >> 6), mwords * 8u, &sm.mbar[buf]);
        if (keys_given) bulk_g2s(sm.keys[buf], ep.keys_in + base, kbytes, &sm.mbar[buf]);
    };

    if (threadIdx.x == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        const uint32_t t0 = atomicAdd(ep.pass.tile_counter, 1u);
        sm.s_tile[0] = t0;
        sm.s_tile[1] = atomicAdd(ep.pass.tile_counter, 1u);
        if (in_order && t0 < ntiles) fetch_tile(t0, 0);
    }
    rank_reset(rs, ep.pass.bits);
    __syncthreads();
    uint32_t cur = 0, ring = 0, parity0 = 0, parity1 = 0;
    for (;;) {
        const uint32_t tile = sm.s_tile[ring];
        if (tile >= ntiles) break;
        const uint32_t next = sm.s_tile[ring == 2 ? 0 : ring + 1];
        if (threadIdx.x == 0) {
            sm.s_tile[ring == 0 ? 2 : ring - 1] = atomicAdd(ep.pass.tile_counter, 1u);
            if (in_order && next < ntiles) fetch_tile(next, cur ^ 1u);  // prefetch; buffer freed by the last barrier
        }
        const float4* rows = sm.rows[cur];
        const uint32_t tile_base = tile * TILE;
        uint32_t ent[ITEMS], digit[ITEMS], slot[ITEMS], shift_g[ITEMS];
        bool valid[ITEMS];
        if (in_order) {
            mbar_wait(&sm.mbar[cur], cur ? parity1 : parity0);
            if (cur) parity1 ^= 1u; else parity0 ^= 1u;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t li = warp * (ITEMS * 32) + k * 32 + lane;
                const uint32_t pos = tile_base + li;
                ent[k] = pos;
                valid[k] = pos < n_eff;
                if (valid[k]) {
                    const bool t_ok = ep.mask_t == nullptr || ((sm.masks[cur][0][li >> 6] >> (li & 63u)) & 1ull);
                    const bool s_ok = ep.mask_s == nullptr || ((sm.masks[cur][1][li >> 6] >> (li & 63u)) & 1ull);
                    valid[k] = t_ok && s_ok;
                }
                uint64_t sk = 0;
                if (valid[k]) {
                    if (keys_given) {
                        sk = (uint64_t) reinterpret_cast<const KeyT*>(sm.keys[cur])[li];
                    } else if (from_world) {
                        sk = pext_runs(world_key(rows[chunk(li, 2)], rows[chunk(li, 3)]), ep.plan.pext);
                    } else {
                        const float4 a0 = rows[chunk(li, 0)], a1 = rows[chunk(li, 1)];
                        float z;
                        if (is_planar(a1)) {
                            z = planar_z(a0, a1);
                        } else {
                            Row r;
                            r.a0 = a0;
                            r.a1 = a1;
                            r.b0 = rows[chunk(li, 2)];
                            r.b1 = rows[chunk(li, 3)];
                            z = local_z(r);
                        }
                        const uint32_t tl = __float_as_uint(a1.w);
                        sk = pext_runs(sort_key(tl >> 16, tl & 0xFFFFu, z), ep.plan.pext);
                    }
                }
                digit[k] = valid[k] ? ((uint32_t)(sk >> ep.pass.shift) & mask) : kInvalidDigit;
                shift_g[k] = (ep.remap != nullptr && valid[k]) ? __ldg(ep.remap + (uint32_t)sk) : 0u;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t li = warp * (ITEMS * 32) + k * 32 + lane;
                const uint32_t pos = tile_base + li;
                valid[k] = pos < n_eff;
                ent[k] = pos;
                digit[k] = kInvalidDigit;
                shift_g[k] = 0u;
                if (valid[k]) {
                    const KeyT key = ep.keys_in[pos];
                    ent[k] = ep.idx_in[pos];
                    digit[k] = (uint32_t)(key >> ep.pass.shift) & mask;
                    if (ep.remap != nullptr) shift_g[k] = __ldg(ep.remap + (uint32_t)key);
                    // gather this entity's row into the item's shared-memory slot
                    float4* d = sm.rows[cur];
                    if (from_world) {
                        const float4* src = reinterpret_cast<const float4*>(ep.world + ent[k]);
                        cp_async16(d + row_slot(li, 0), src);
                        cp_async16(d + row_slot(li, 1), src + 1);
                        cp_async16(d + row_slot(li, 2), src + 2);
                        cp_async16(d + row_slot(li, 3), src + 3);
                    } else {
                        const size_t o = (size_t)ent[k] * ep.rows.stride;
                        cp_async16(d + row_slot(li, 0), ep.rows.a + o);
                        cp_async16(d + row_slot(li, 1), ep.rows.a + o + 1);
                        cp_async16(d + row_slot(li, 2), ep.rows.b + o);
                        cp_async16(d + row_slot(li, 3), ep.rows.b + o + 1);
                    }
                }
            }
        }
        rank_tile<ITEMS>(rs, ep.pass, tile, digit, slot);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (digit[k] != kInvalidDigit) {
                sm.s_li[slot[k]] = (uint16_t)(warp * (ITEMS * 32) + k * 32 + lane);
                sm.s_ent[slot[k]] = ent[k];
                sm.s_dst[slot[k]] = rs.dbase[digit[k]] + slot[k] + shift_g[k];
            }
        }
        if (!in_order) cp_async_wait_all();
        __syncthreads();
        const uint32_t cnt = rs.tile_count;
        float4* stage = sm.stage[warp];
        for (uint32_t s0 = warp * 32; s0 < cnt; s0 += kThreads) {
            const uint32_t s = s0 + lane;
            const bool live = s < cnt;
            uint32_t dst = 0;
            InstanceOut inst;
            float4 vtx[4];
            if (live) {
                const uint32_t li = sm.s_li[s];
                const uint32_t e = sm.s_ent[s];
                dst = sm.s_dst[s];
                Affine w;
                float sw, shh;
                uint32_t tl;
                if (from_world) {
                    w.r0 = rows[chunk(li, 0)];
                    w.r1 = rows[chunk(li, 1)];
                    w.r2 = rows[chunk(li, 2)];
                    const float4 sv = rows[chunk(li, 3)];
                    sw = sv.x;
                    shh = sv.y;
                    tl = __float_as_uint(sv.z);
                } else {
                    Row r;
                    r.a0 = rows[chunk(li, 0)];
                    r.a1 = rows[chunk(li, 1)];
                    r.b0 = rows[chunk(li, 2)];
                    r.b1 = rows[chunk(li, 3)];
                    w = compose_local(r);
                    sw = r.b1.y;
                    shh = r.b1.z;
                    tl = __float_as_uint(r.a1.w);
                }
                make_outputs(w, sw, shh, tl, inst, vtx);
                ep.order[dst] = e + ep.global_first;
                if (ep.texlayer != nullptr) ep.texlayer[dst] = tl;
            }
            if (!ep.staged) {
                if (live) {
                    float4* ip = ep.instances + (size_t)dst * 5;
                    st_cs_f4(ip, inst.c0);
                    st_cs_f4(ip + 1, inst.c1);
                    st_cs_f4(ip + 2, inst.c2);
                    st_cs_f4(ip + 3, inst.c3);
                    st_cs_f4(ip + 4, inst.s);
                    float4* vp = ep.vertices + (size_t)dst * 4;
                    st_cs_f4(vp, vtx[0]);
                    st_cs_f4(vp + 1, vtx[1]);
                    st_cs_f4(vp + 2, vtx[2]);
                    st_cs_f4(vp + 3, vtx[3]);
                }
            } else {
                staged_store(stage, lane, live, min(32u, cnt - s0), dst, inst, vtx, ep.instances, ep.vertices);
            }
        }
        rank_reset(rs, ep.pass.bits);
        __syncthreads();  // releases rows[cur], the rank scratch and the tile ring slot for reuse
        cur ^= 1u;
        ring = ring == 2 ? 0 : ring + 1;
    }
}

// ---- few-bucket frames: every warp is its own pipeline -----------------------------------------------
//
// Single-pass plans with at most 8 buckets (e.g. a particle system with a handful of textures)
// need no block at all: a warp takes 64 consecutive entities, ranks them with a few ballots in
// registers, and — because pass A counted the buckets of every 64-entity tile and tile_scan turned
// the counts into offsets — knows every output position without talking to any other warp. There is
// not a single block barrier in this kernel; warps drift apart and hide each other's latencies.
constexpr uint32_t kWarpTile = 64;          // entities per warp step (2 per lane)
constexpr uint32_t kSmallMaxBits = 3;       // up to 8 buckets: runs of >= 8 instances per warp tile

struct WarpShared {
    float4 rows[2][kWarpTile * 4];          // 2 x 4 KB: this tile's rows and the prefetched next tile's
    unsigned long long keys[2][kWarpTile];  // short keys when the pass reads them (hierarchy)
    float4 vel[2][kWarpTile * 3 / 4];       // velocities of the tile (single-kernel frames)
    float4 stage[32 * kStageF4];            // 2.5 KB staging
    uint32_t s_dst[kWarpTile];
    uint8_t s_li[kWarpTile];
    unsigned long long mbar[2];
};
struct SmallShared {
    WarpShared w[kWarps];
};

// FUSED: the whole frame in this one kernel. When the bucket of an entity depends only on its
// Sprite (the plan has no varying z bit), per-tile bucket offsets stay valid from frame to frame
// until a component is uploaded, so the steady-state frame needs no separate key pass: this kernel
// also runs the velocity system (writing sector A back), and accumulates the key statistics with
// which the host validates, every frame, that the cached plan and offsets still hold.
template <typename KeyT, bool FUSED>
__global__ void __launch_bounds__(kThreads, 2) emit_small_kernel(const EmitParams<KeyT> ep) {
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpShared& sm = reinterpret_cast<SmallShared*>(emit_smem_raw)->w[warp];
    const bool do_int = FUSED && ep.integrate && ep.vel != nullptr;
    PrepAcc acc;
    const bool from_world = ep.world != nullptr;
    const bool keys_given = ep.keys_in != nullptr;
    const uint32_t n = ep.pass.n;
    const uint32_t ntiles = ep.pass.ntiles;  // == ceil(n / 64)
    const uint32_t R = 1u << ep.pass.bits;
    const bool split = !from_world && ep.rows.stride == 2;
    const uint32_t sstride = split ? 2u : 4u;
    const uint32_t boff = split ? 2u * kWarpTile : 2u;
    const uint32_t lt = lanemask_lt();
    const uint32_t nwarps = gridDim.x * kWarps;
    uint32_t tile = blockIdx.x * kWarps + warp;

    auto fetch = [&](uint32_t t, uint32_t buf) {  // lane 0: TMA bulk copies of tile t's rows (and keys)
        const uint32_t base = t * kWarpTile;
        const uint32_t cnt = min(kWarpTile, n - base);
        const uint32_t kbytes = keys_given ? ((cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u) : 0u;
        const uint32_t vbytes = do_int ? ((cnt * 12u + 15u) & ~15u) : 0u;
        mbar_expect_tx(&sm.mbar[buf], cnt * 64u + kbytes + vbytes);
        if (from_world) {
            bulk_g2s(sm.rows[buf], ep.world + base, cnt * 64u, &sm.mbar[buf]);
        } else if (split) {
            bulk_g2s(sm.rows[buf], ep.rows.a + (size_t)base * 2, cnt * 32u, &sm.mbar[buf]);
            bulk_g2s(sm.rows[buf] + 2 * kWarpTile, ep.rows.b + (size_t)base * 2, cnt * 32u, &sm.mbar[buf]);
        } else {
            bulk_g2s(sm.rows[buf], ep.rows.a + (size_t)base * 4, cnt * 64u, &sm.mbar[buf]);
        }
        if (keys_given) bulk_g2s(sm.keys[buf], ep.keys_in + base, kbytes, &sm.mbar[buf]);
        if (do_int) bulk_g2s(sm.vel[buf], ep.vel + (size_t)base * 3, vbytes, &sm.mbar[buf]);
    };
    // Presence words of a tile, prefetched one tile ahead WITHOUT being consumed: lane 0 holds the
    // Transform word, lane 1 the Sprite word, lane 2 the Velocity word; they are combined (by shuffle)
    // only when the tile is processed, so the loads never stall the warp.
    auto load_raw = [&](uint32_t t) -> unsigned long long {
        const uint64_t* src = lane == 0 ? ep.mask_t : lane == 1 ? ep.mask_s : (lane == 2 && do_int) ? ep.mask_v : nullptr;
        return src != nullptr ? __ldg(src + t) : ~0ull;
    };

    if (lane == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        if (tile < ntiles) fetch(tile, 0);
    }
    __syncwarp();
    unsigned long long raw = tile < ntiles ? load_raw(tile) : 0ull, raw_next = 0ull;
    uint32_t off = (tile < ntiles && lane < R) ? __ldg(ep.pass.tile_off + (size_t)lane * ntiles + tile) : 0u;
    uint32_t cur = 0, par0 = 0, par1 = 0;
    for (; tile < ntiles; tile += nwarps) {
        // prefetch the next tile of this warp: rows by TMA, presence word and bucket offsets in registers
        const uint32_t next = tile + nwarps;
        uint32_t off_next = 0u;
        if (next < ntiles) {
            if (lane == 0) fetch(next, cur ^ 1u);
            raw_next = load_raw(next);
            if (lane < R) off_next = __ldg(ep.pass.tile_off + (size_t)lane * ntiles + next);
        }
        float4* rows = sm.rows[cur];
        const uint32_t base = tile * kWarpTile;
        mbar_wait(&sm.mbar[cur], cur ? par1 : par0);
        if (cur) par1 ^= 1u; else par0 ^= 1u;
        // (Transform and Sprite) = renderable; (Transform and Velocity) = moved by the tick
        const unsigned long long tail = (n - base < kWarpTile) ? ((1ull << (n - base)) - 1ull) : ~0ull;
        const unsigned long long mt = __shfl_sync(0xffffffffu, raw, 0);
        const unsigned long long pres = mt & __shfl_sync(0xffffffffu, raw, 1) & tail;
        const unsigned long long moving = do_int ? (mt & __shfl_sync(0xffffffffu, raw, 2) & tail) : 0ull;
        if (do_int) {
            // the velocity system on this tile: translation += velocity * dt, in shared memory (the
            // compose below reads it) and written back to sector A
            const float* vel = reinterpret_cast<const float*>(sm.vel[cur]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t li = k * 32 + lane;
                if ((moving >> li) & 1ull) {
                    float4 a0 = rows[li * sstride];
                    a0.x = addr(a0.x, mulr(vel[li * 3], ep.dt));
                    a0.y = addr(a0.y, mulr(vel[li * 3 + 1], ep.dt));
                    a0.z = addr(a0.z, mulr(vel[li * 3 + 2], ep.dt));
                    rows[li * sstride] = a0;
                    ep.rows.a[(size_t)(base + li) * ep.rows.stride] = a0;
                }
            }
        }
        // bucket of this lane's two entities (li = lane and lane + 32)
        uint32_t digit[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t li = k * 32 + lane;
            digit[k] = kInvalidDigit;
            if ((pres >> li) & 1ull) {
                uint64_t sk;
                if (keys_given) {
                    sk = (uint64_t) reinterpret_cast<const KeyT*>(sm.keys[cur])[li];
                } else if (from_world) {
                    sk = pext_runs(world_key(rows[li * 4 + 2], rows[li * 4 + 3]), ep.plan.pext);
                } else {
                    const float4 a0 = rows[li * sstride], a1 = rows[li * sstride + 1];
                    float z;
                    if (is_planar(a1)) {
                        z = planar_z(a0, a1);
                    } else {
                        Row r;
                        r.a0 = a0;
                        r.a1 = a1;
                        r.b0 = rows[boff + li * sstride];
                        r.b1 = rows[boff + li * sstride + 1];
                        z = local_z(r);
                    }
                    const uint32_t tl = __float_as_uint(a1.w);
                    const uint64_t key = sort_key(tl >> 16, tl & 0xFFFFu, z);
                    if (FUSED) {
                        acc.k_or |= key;
                        acc.k_and &= key;
                        acc.count += 1;
                    }
                    sk = pext_runs(key, ep.plan.pext);
                }
                digit[k] = (uint32_t)sk & (R - 1u);
            }
        }
        // stable rank inside the warp tile, in registers: per bucket two ballots
        uint32_t slot[2] = {0, 0}, dpos[2] = {0, 0};
        uint32_t run = 0;
        for (uint32_t d = 0; d < R; ++d) {
            const uint32_t b0 = __ballot_sync(0xffffffffu, digit[0] == d);
            const uint32_t b1 = __ballot_sync(0xffffffffu, digit[1] == d);
            const uint32_t o = __shfl_sync(0xffffffffu, off, d);
            if (digit[0] == d) {
                const uint32_t r = __popc(b0 & lt);
                slot[0] = run + r;
                dpos[0] = o + r;
            }
            if (digit[1] == d) {
                const uint32_t r = __popc(b0) + __popc(b1 & lt);
                slot[1] = run + r;
                dpos[1] = o + r;
            }
            run += __popc(b0) + __popc(b1);
        }
        const uint32_t cnt = run;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (digit[k] != kInvalidDigit) {
                sm.s_li[slot[k]] = (uint8_t)(k * 32 + lane);
                sm.s_dst[slot[k]] = dpos[k] + (ep.remap != nullptr ? __ldg(ep.remap + digit[k]) : 0u);
            }
        }
        __syncwarp();
        for (uint32_t s0 = 0; s0 < cnt; s0 += 32) {
            const uint32_t s = s0 + lane;
            const bool live = s < cnt;
            uint32_t dst = 0;
            InstanceOut inst;
            float4 vtx[4];
            if (live) {
                const uint32_t li = sm.s_li[s];
                dst = sm.s_dst[s];
                Affine w;
                float sw, shh;
                uint32_t tl;
                if (from_world) {
                    w.r0 = rows[li * 4];
                    w.r1 = rows[li * 4 + 1];
                    w.r2 = rows[li * 4 + 2];
                    const float4 sv = rows[li * 4 + 3];
                    sw = sv.x;
                    shh = sv.y;
                    tl = __float_as_uint(sv.z);
                } else {
                    Row r;
                    r.a0 = rows[li * sstride];
                    r.a1 = rows[li * sstride + 1];
                    r.b0 = rows[boff + li * sstride];
                    r.b1 = rows[boff + li * sstride + 1];
                    w = compose_local(r);
                    sw = r.b1.y;
                    shh = r.b1.z;
                    tl = __float_as_uint(r.a1.w);
                }
                make_outputs(w, sw, shh, tl, inst, vtx);
                ep.order[dst] = base + li + ep.global_first;
                if (ep.texlayer != nullptr) ep.texlayer[dst] = tl;
            }
            if (!ep.staged) {
                if (live) {
                    float4* ip = ep.instances + (size_t)dst * 5;
                    st_cs_f4(ip, inst.c0);
                    st_cs_f4(ip + 1, inst.c1);
                    st_cs_f4(ip + 2, inst.c2);
                    st_cs_f4(ip + 3, inst.c3);
                    st_cs_f4(ip + 4, inst.s);
                    float4* vp = ep.vertices + (size_t)dst * 4;
                    st_cs_f4(vp, vtx[0]);
                    st_cs_f4(vp + 1, vtx[1]);
                    st_cs_f4(vp + 2, vtx[2]);
                    st_cs_f4(vp + 3, vtx[3]);
                }
            } else {
                staged_store(sm.stage, lane, live, min(32u, cnt - s0), dst, inst, vtx, ep.instances, ep.vertices);
            }
        }
        __syncwarp();  // every lane is done with rows[cur], s_li, s_dst before they are reused
        raw = raw_next;
        off = off_next;
        cur ^= 1u;
    }
    if (FUSED) {
        // key statistics of everything this warp emitted: the host validates the cached plan with them
        const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)acc.k_or);
        const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(acc.k_or >> 32));
        const uint32_t nlo = __reduce_or_sync(0xffffffffu, ~(uint32_t)acc.k_and);
        const uint32_t nhi = __reduce_or_sync(0xffffffffu, ~(uint32_t)(acc.k_and >> 32));
        const uint32_t c = __reduce_add_sync(0xffffffffu, acc.count);
        if (lane == 0 && c) {
            atomicOr(&ep.stats[0], ((unsigned long long)hi << 32) | lo);
            atomicOr(&ep.stats[1], ((unsigned long long)nhi << 32) | nlo);
            atomicAdd(&ep.stats[2], (unsigned long long)c);
        }
    }
}

// ---- generic batch boundaries (more than 2^16 distinct (layer, texture) bins) -------------------

// Counts run heads of texlayer[0..n) per block of 1024 positions.
__global__ void __launch_bounds__(kThreads) batch_heads_count_kernel(const uint32_t* __restrict__ tl, uint32_t n,
                                                                      uint32_t* block_counts) {
    __shared__ uint32_t s_c;
    if (threadIdx.x == 0) s_c = 0;
    __syncthreads();
    uint32_t c = 0;
    const uint32_t base = blockIdx.x * 1024;
    for (uint32_t k = threadIdx.x; k < 1024; k += kThreads) {
        const uint32_t i = base + k;
        if (i < n) c += (i == 0) || (tl[i] != tl[i - 1]);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
    __syncthreads();
    if (threadIdx.x == 0) block_counts[blockIdx.x] = s_c;
}

// Exclusive scan of up to 2^20 block counts by one block (in place), total to *total.
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t* counts, uint32_t nblocks, uint32_t* total) {
    __shared__ uint32_t s[1024];
    const uint32_t per = (nblocks + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per;
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b0 + per && b < nblocks; ++b) sum += counts[b];
    s[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        uint32_t v = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s[threadIdx.x] - sum;
    for (uint32_t b = b0; b < b0 + per && b < nblocks; ++b) {
        const uint32_t c = counts[b];
        counts[b] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *total = s[1023];
}

// Writes run heads in order; one block per 1024 positions, ordered inside the block by a
// ballot scan. The count of each batch is filled by batch_counts_kernel.
__global__ void __launch_bounds__(kThreads) batch_heads_write_kernel(const uint32_t* __restrict__ tl, uint32_t n,
                                                                      const uint32_t* __restrict__ block_offsets,
                                                                      BatchOut* out, uint32_t max_batches) {
    __shared__ uint32_t s_w[kWarps];
    __shared__ uint32_t s_run;
    if (threadIdx.x == 0) s_run = block_offsets[blockIdx.x];
    __syncthreads();
    const uint32_t base = blockIdx.x * 1024;
    for (uint32_t k0 = 0; k0 < 1024; k0 += kThreads) {
        const uint32_t i = base + k0 + threadIdx.x;
        const bool head = i < n && ((i == 0) || (tl[i] != tl[i - 1]));
        const uint32_t b = __ballot_sync(0xffffffffu, head);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = __popc(b);
        __syncthreads();
        uint32_t off = s_run;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) off += s_w[w];
        if (head) {
            const uint32_t slot = off + __popc(b & lanemask_lt());
            if (slot < max_batches) {
                BatchOut o;
                o.layer = tl[i] >> 16;
                o.texture = tl[i] & 0xFFFFu;
                o.first_instance = i;
                o.count = 0;
                out[slot] = o;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kWarps; ++w) t += s_w[w];
            s_run += t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) batch_counts_kernel(BatchOut* out, const uint32_t* num_batches,
                                                                 uint32_t max_batches, uint32_t n) {
    const uint32_t nb = min(*num_batches, max_batches);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const uint32_t next = (i + 1 < *num_batches && i + 1 < max_batches) ? out[i + 1].first_instance : n;
    out[i].count = next - out[i].first_instance;
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

// ---- hierarchy planning (on upload of parents, not per frame) -----------------------------------

// hop[e] = sanitised parent (TB_NO_PARENT when the entity or its parent has no Transform, or
// the parent is outside this shard); dist[e] = 1 for a child, 0 for a root.
__global__ void __launch_bounds__(kThreads) level_init_kernel(const uint32_t* __restrict__ parent,
                                                               const uint64_t* mask_t, const uint64_t* mask_p,
                                                               uint32_t n, uint32_t global_first, uint32_t* hop,
                                                               uint32_t* dist, uint32_t* sane_parent) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint32_t p = 0xFFFFFFFFu;
    if (parent != nullptr && mask_bit(mask_t, e) && mask_bit(mask_p, e)) {
        const uint32_t g = parent[e];
        if (g != 0xFFFFFFFFu && g >= global_first && g - global_first < n && mask_bit(mask_t, g - global_first))
            p = g - global_first;
    }
    hop[e] = p;
    dist[e] = p != 0xFFFFFFFFu;
    sane_parent[e] = p == 0xFFFFFFFFu ? 0xFFFFFFFFu : p + global_first;
}

// One pointer-jumping round (Wyllie): dist += dist[hop]; hop = hop[hop]. `pending` counts
// entities that have not reached a root yet.
__global__ void __launch_bounds__(kThreads) level_jump_kernel(const uint32_t* __restrict__ hop_in,
                                                               const uint32_t* __restrict__ dist_in, uint32_t n,
                                                               uint32_t* hop_out, uint32_t* dist_out, uint32_t* pending) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    uint32_t h = hop_in[e], d = dist_in[e];
    if (h != 0xFFFFFFFFu) {
        d += dist_in[h];
        h = hop_in[h];
        if (h != 0xFFFFFFFFu) atomicAdd(pending, 1u);
    }
    hop_out[e] = h;
    dist_out[e] = d;
}

// max level, and whether levels are non-decreasing with the entity id (then every level is
// an id range and no list is needed). stats[0] = max level, stats[1] = number of descents.
__global__ void __launch_bounds__(kThreads) level_stats_kernel(const uint32_t* __restrict__ level, const uint64_t* mask_t,
                                                                uint32_t n, uint32_t* stats) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t l = level[e];
    atomicMax(&stats[0], l);
    if (e > 0 && level[e - 1] > l) atomicAdd(&stats[1], 1u);
    (void)mask_t;
}

__global__ void __launch_bounds__(kThreads) level_hist_kernel(const uint32_t* __restrict__ level, const uint64_t* mask_t,
                                                               uint32_t n, uint32_t* hist, uint32_t nlevels) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (!mask_bit(mask_t, e)) return;
    const uint32_t l = level[e];
    if (l < nlevels) atomicAdd(&hist[l], 1u);
}

// world rows -> row-major 4x4 (identity where the entity has no Transform)
__global__ void __launch_bounds__(kThreads) world_to_mat4_kernel(const WorldRow* __restrict__ world, const uint64_t* mask_t,
                                                                  uint32_t first, uint32_t n, float* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 r0 = make_float4(1, 0, 0, 0), r1 = make_float4(0, 1, 0, 0), r2 = make_float4(0, 0, 1, 0);
    if (mask_bit(mask_t, first + i)) {
        const float4* w = reinterpret_cast<const float4*>(world + first + i);
        r0 = w[0];
        r1 = w[1];
        r2 = w[2];
    }
    float4* o = reinterpret_cast<float4*>(dst + (size_t)i * 16);
    o[0] = r0;
    o[1] = r1;
    o[2] = r2;
    o[3] = make_float4(0, 0, 0, 1);
}

// Transform -> local world row for flat scenes (tb_download_world_matrices / tb_as_matrix4)
__global__ void __launch_bounds__(kThreads) compose_world_kernel(RowView rows, const uint64_t* mask_t, uint32_t n,
                                                                  WorldRow* world) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (!mask_bit(mask_t, e)) return;
    const Row r = load_row(rows, e);
    const Affine w = compose_local(r);
    float4* dst = reinterpret_cast<float4*>(world + e);
    dst[0] = w.r0;
    dst[1] = w.r1;
    dst[2] = w.r2;
    dst[3] = make_float4(r.b1.y, r.b1.z, r.a1.w, 0.0f);
}

// ---- standalone velocity system (when no frame follows the tick) --------------------------------

__global__ void __launch_bounds__(kThreads) integrate_kernel(RowView rows, const uint64_t* mask_t, const uint64_t* mask_v,
                                                              const float* __restrict__ vel, float dt, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!mask_bit(mask_t, i) || !mask_bit(mask_v, i)) return;
    const size_t o = (size_t)i * rows.stride;
    float4 a0 = rows.a[o];
    const float* v = vel + (size_t)i * 3;
    a0.x = addr(a0.x, mulr(__ldg(v), dt));
    a0.y = addr(a0.y, mulr(__ldg(v + 1), dt));
    a0.z = addr(a0.z, mulr(__ldg(v + 2), dt));
    rows.a[o] = a0;
}

// ---- generic digit histograms of a u32 key array (hierarchy level lists) -------------------------

struct DigitSplit {
    uint32_t npasses;
    uint32_t shift[kMaxPasses];
    uint32_t bits[kMaxPasses];
};

__global__ void __launch_bounds__(kThreads) key_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                             const DigitSplit ds, uint32_t* hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = keys[i];
    for (uint32_t q = 0; q < ds.npasses; ++q)
        atomicAdd(&hist[q * 256 + ((k >> ds.shift[q]) & ((1u << ds.bits[q]) - 1u))], 1u);
}

// ---- multi-GPU placement --------------------------------------------------------------

// hists: [nranks][nbins] short-key histograms of every shard. Per block of 1024 bins: the sum over
// all ranks and this rank's own sum.
__global__ void __launch_bounds__(kThreads) place_sums_kernel(const uint32_t* __restrict__ hists, uint32_t nranks,
                                                               uint32_t rank, uint32_t nbins, uint32_t* bsum_total,
                                                               uint32_t* bsum_local) {
    __shared__ uint32_t s_t, s_l;
    if (threadIdx.x == 0) s_t = s_l = 0;
    __syncthreads();
    uint32_t t = 0, l = 0;
    for (uint32_t k = blockIdx.x * 1024 + threadIdx.x; k < min(nbins, (blockIdx.x + 1) * 1024); k += kThreads) {
        for (uint32_t r = 0; r < nranks; ++r) t += hists[(size_t)r * nbins + k];
        l += hists[(size_t)rank * nbins + k];
    }
    t = __reduce_add_sync(0xffffffffu, t);
    l = __reduce_add_sync(0xffffffffu, l);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_t, t);
        atomicAdd(&s_l, l);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bsum_total[blockIdx.x] = s_t;
        bsum_local[blockIdx.x] = s_l;
    }
}

// remap[k] = (elements of all ranks with a smaller key) + (elements of lower ranks with key k)
//            - (elements of this rank with a smaller key):
// added to an element's shard-local sorted position it gives its position in the global order
// (id ranges ascend with the rank, so ties between ranks resolve like ties between entity ids).
// Also accumulates the global (layer, texture) bins for the batch table.
__global__ void __launch_bounds__(kThreads) place_write_kernel(const uint32_t* __restrict__ hists, uint32_t nranks,
                                                                uint32_t rank, uint32_t nbins,
                                                                const uint32_t* __restrict__ boff_total,
                                                                const uint32_t* __restrict__ boff_local, uint32_t zbits,
                                                                uint32_t* remap, uint32_t* lt_global) {
    __shared__ uint32_t s_wt[kWarps], s_wl[kWarps];
    const uint32_t k0 = blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t t[4], l[4], lower[4];
    uint32_t st = 0, sl = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t k = k0 + j;
        t[j] = l[j] = lower[j] = 0;
        if (k < nbins) {
            for (uint32_t r = 0; r < nranks; ++r) {
                const uint32_t h = hists[(size_t)r * nbins + k];
                t[j] += h;
                if (r < rank) lower[j] += h;
                if (r == rank) l[j] = h;
            }
            if (t[j]) atomicAdd(&lt_global[k >> zbits], t[j]);
        }
        st += t[j];
        sl += l[j];
    }
    uint32_t xt = st, xl = sl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t yt = __shfl_up_sync(0xffffffffu, xt, o), yl = __shfl_up_sync(0xffffffffu, xl, o);
        if ((threadIdx.x & 31) >= (uint32_t)o) {
            xt += yt;
            xl += yl;
        }
    }
    if ((threadIdx.x & 31) == 31) {
        s_wt[threadIdx.x >> 5] = xt;
        s_wl[threadIdx.x >> 5] = xl;
    }
    __syncthreads();
    uint32_t bt = boff_total[blockIdx.x] + xt - st, bl = boff_local[blockIdx.x] + xl - sl;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) {
        bt += s_wt[w];
        bl += s_wl[w];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t k = k0 + j;
        if (k < nbins) remap[k] = bt + lower[j] - bl;
        bt += t[j];
        bl += l[j];
    }
}

}  // namespace tb
