// Frame pass A: the velocity system, the sort key of every renderable, the key statistics the
// frame plan is built from / validated with, digit histograms and per-tile bucket counts.
// prepare_flat for scenes without Parent links, prepare_level (one launch per hierarchy level,
// world = world[parent] * local) otherwise; plus the level planning that runs when parents change.
#pragma once
#include "common.cuh"

namespace tb {

// ---- frame pass A: velocity system + key + masks + histograms --------------------------------

struct PrepParams {
    uint32_t n;            // entity_count
    RowView rows;
    const uint64_t* mask_t;  // presence masks, nullptr == present on every entity
    const uint64_t* mask_s;
    const uint64_t* mask_v;
    const float* vel;      // tb_velocity AoS (3 f32), nullptr when no Velocity store
    float dt;
    uint32_t integrate;    // engine ticks of `dt` the velocity system runs in this pass (0: none)
    uint32_t have_plan;    // 0: only masks + count (plan discovery)
    uint32_t write_keys;   // short keys to keys_out (multi-pass plans)
    uint32_t check_keys;   // keys_out still holds the previous frame's keys: compare before overwriting
                           // (stats[5] != 0 afterwards <=> some key changed <=> the cached sort is stale)
    uint32_t lt_mode;      // 0: none (aliases digit pass 0), 1: shared-memory bins, 2: global bins
    void* keys_out;
    uint32_t* hist;        // [kMaxPasses][256]
    uint32_t* lt_hist;     // [1 << ltbits]
    unsigned long long* stats;  // [0] OR of keys, [1] OR of ~keys, [2] renderable count, [4] z-dictionary misses ([3] holds the batch count)
    uint32_t* tile_counts;  // optional [1 << bits[0]][emit_tiles]: pass-0 digit counts per emit tile
    uint32_t emit_tiles;
    uint32_t emit_tile_shift;  // log2 of pass 0's tile: 6 (warp tiles of 64), 9 (emit tiles of 512), 11 (scatter tiles of 2048)
    uint32_t* key_hist;     // optional [1 << nbits]: histogram over the whole short key (multi-GPU placement)
    uint32_t* chunk_hist;   // optional [nchunks][1 << nbits]: the same per chunk of 1 << chunk_shift entity ids (chunked.cuh)
    uint32_t chunk_shift;
    unsigned long long* zset;  // plan discovery: open-addressing set of the distinct orderable z values
    uint32_t* zset_overflow;   // ... set when the scene has more distinct depths than the set is meant for
    FramePlan plan;
};

constexpr int kPrepItems = 4;

struct PrepShared {
    uint32_t hist[kMaxPasses][256];
    uint32_t tcnt[2][2][256];  // [iteration parity][half of the 1024-entity tile][pass-0 digit]
    uint32_t lt[2048];
    uint32_t zdict[kMaxZDict];  // the plan's z dictionary
    unsigned long long mwords[2][3][16];  // [iteration parity][Transform, Sprite, Velocity][word]: presence words of a 1024-entity tile
    unsigned long long k_or, k_and;
    uint32_t count;
};

// Per-thread running key statistics, folded into the block's once at the end of the kernel.
struct PrepAcc {
    unsigned long long k_or = 0ull, k_and = ~0ull;
    uint32_t count = 0;
    bool miss = false;  // a depth that the plan's z dictionary does not hold
    bool changed = false;  // an entity whose short key differs from the previous frame's
};

// Plan discovery: remember this depth (warp-aggregated insert into a small open-addressing set).
__device__ __forceinline__ void zset_insert(unsigned long long* zset, uint32_t* overflow, bool renderable, uint32_t z) {
    const uint32_t peers = __match_any_sync(0xffffffffu, renderable ? z : 0xFFFFFFFFu);
    if (!renderable || (threadIdx.x & 31) != (uint32_t)(__ffs(peers) - 1)) return;
    if (*reinterpret_cast<volatile uint32_t*>(overflow)) return;
    const unsigned long long want = (1ull << 32) | z;
    uint32_t h = (z * 2654435761u) >> 23;
    for (uint32_t probe = 0; probe < 4u * kMaxZDict; ++probe, ++h) {
        unsigned long long* slot = zset + (h & (kZSetSlots - 1));
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(slot);
        if (cur == want) return;
        if (cur == 0ull) {
            cur = atomicCAS(slot, 0ull, want);
            if (cur == 0ull || cur == want) return;
        }
    }
    atomicExch(overflow, 1u);  // far more distinct depths than a dictionary can hold
}

// Accounts one entity: key statistics in registers, digit histograms in shared memory, optional
// short key write. Must be called by all 32 lanes of a warp (it contains warp collectives).
//   small == true  (single pass, <= 32 buckets): one match.any per entity aggregates the warp's
//     histogram update, and the per-emit-tile counts go straight to global memory with one RED per
//     distinct digit per warp - no block barrier anywhere in the key pass;
//   small == false: plain shared atomics (many distinct digits, little contention), per-tile counts
//     through `tcnt` in shared memory when the caller wants them.
// Every entity gets a key slot in entity order (all ones: not renderable), so that the next frame's pass can
// tell whether anything moved in the draw order by comparing slot by slot before it overwrites them.
// `old32`: the slot's previous 32-bit value when the caller fetched it ahead of time (32-bit keys, key check on).
__device__ __forceinline__ void store_key(const PrepParams& p, PrepAcc& acc, uint32_t i, uint64_t v, bool have_old, uint32_t old32) {
    if (p.plan.key64) {
        uint64_t* slot = reinterpret_cast<uint64_t*>(p.keys_out) + i;
        if (p.check_keys && *slot != v) acc.changed = true;
        *slot = v;
    } else {
        uint32_t* slot = reinterpret_cast<uint32_t*>(p.keys_out) + i;
        if (p.check_keys && (have_old ? old32 : *slot) != (uint32_t)v) acc.changed = true;
        *slot = (uint32_t)v;
    }
}
// The previous frame's 32-bit key of entity i, read through the non-coherent path: only this thread ever touches
// the slot in this kernel and it reads before it writes, so the load may be issued long before the compare.
__device__ __forceinline__ bool fetch_old_key(const PrepParams& p, bool inb, uint32_t i, uint32_t* old32) {
    if (!(p.check_keys && p.write_keys && !p.plan.key64 && inb)) return false;
    *old32 = __ldg(reinterpret_cast<const uint32_t*>(p.keys_out) + i);
    return true;
}

// `inb`: i is a real entity of this shard (out-of-range lanes take part in the warp collectives only).
__device__ __forceinline__ void prep_account(const PrepParams& p, PrepShared& sh, PrepAcc& acc, bool inb, bool renderable,
                                             uint64_t key, uint32_t i, bool small, uint32_t emit_tile, uint32_t* tcnt,
                                             bool have_old = false, uint32_t old32 = 0u) {
    if (renderable) {
        acc.k_or |= key;
        acc.k_and &= key;
        acc.count += 1;
    }
    if (!p.have_plan) {
        if (p.zset != nullptr) zset_insert(p.zset, p.zset_overflow, renderable, (uint32_t)key);
        return;
    }
    bool miss = false;
    const uint64_t sk = plan_short_key(p.plan, sh.zdict, key, &miss);
    if (renderable && miss) acc.miss = true;
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
    if (!renderable) {
        if (p.write_keys && inb) store_key(p, acc, i, ~0ull, have_old, old32);  // all ones: not renderable
        return;
    }
    if (p.key_hist != nullptr) atomicAdd(&p.key_hist[(uint32_t)sk], 1u);
    if (p.chunk_hist != nullptr) atomicAdd(&p.chunk_hist[((size_t)(i >> p.chunk_shift) << p.plan.pext.nbits) + (uint32_t)sk], 1u);
    for (uint32_t q = 0; q < p.plan.npasses; ++q) {
        const uint32_t d = (uint32_t)(sk >> p.plan.shift[q]) & ((1u << p.plan.bits[q]) - 1u);
        atomicAdd(&sh.hist[q][d], 1u);
        if (q == 0 && tcnt != nullptr) atomicAdd(&tcnt[d], 1u);
    }
    if (p.lt_mode == 1) atomicAdd(&sh.lt[(uint32_t)(sk >> p.plan.zbits)], 1u);
    else if (p.lt_mode == 2) atomicAdd(&p.lt_hist[(uint32_t)(sk >> p.plan.zbits)], 1u);
    if (p.write_keys) store_key(p, acc, i, sk, have_old, old32);
}

// The single-pass few-bucket plan the `small` accounting path serves.
__device__ __forceinline__ bool prep_small(const PrepParams& p) {
    return p.have_plan && p.plan.npasses == 1 && p.plan.bits[0] <= 5 && !p.write_keys;
}

__device__ __forceinline__ void prep_init(const PrepParams& p, PrepShared& sh) {
    for (uint32_t k = threadIdx.x; k < kMaxPasses * 256; k += blockDim.x) (&sh.hist[0][0])[k] = 0;
    for (uint32_t k = threadIdx.x; k < 2048; k += blockDim.x) sh.lt[k] = 0;
    for (uint32_t k = threadIdx.x; k < (uint32_t)kMaxZDict; k += blockDim.x) sh.zdict[k] = p.have_plan ? p.plan.zdict[k] : 0u;
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
        if (__any_sync(0xffffffffu, acc.miss) && (threadIdx.x & 31) == 0) atomicOr(&p.stats[4], 1ull);
        if (__any_sync(0xffffffffu, acc.changed) && (threadIdx.x & 31) == 0) atomicOr(&p.stats[5], 1ull);
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
__global__ void __launch_bounds__(kThreads, 3) prepare_flat_kernel(const PrepParams p) {
    __shared__ PrepShared sh;
    prep_init(p, sh);
    const uint32_t tile_elems = kThreads * kPrepItems;
    const uint32_t ntiles = (p.n + tile_elems - 1) / tile_elems;
    const bool do_int = p.integrate && p.vel != nullptr;
    const bool small = prep_small(p);
    const bool tiles_out = !small && p.have_plan && p.tile_counts != nullptr;  // shared-memory per-tile counts
    const uint32_t R0 = 1u << p.plan.bits[0];
    const bool wide_tiles = p.emit_tile_shift == 11;  // a 2048-pair scatter tile is two of this kernel's tiles
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
        // Every load of the tile is issued before anything is consumed: rows and velocities exist (zero-filled) for
        // every id below max_entities, so they do not wait for the presence bits. The tile's presence words (16 per
        // mask) are fetched once by 48 threads and shared: a presence test consumed inside the load loop would stall
        // on its own load before the next item's loads are even issued (ncu: the top stall of the round-1 kernel).
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t i = base + k * kThreads;
            a0[k] = a1[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            vx[k] = vy[k] = vz[k] = 0.0f;
            if (i < p.n) {
                const size_t o = (size_t)i * p.rows.st;
                a0[k] = p.rows.t[o];
                a1[k] = p.rows.a[o];
                if (do_int) {
                    const float* v = p.vel + (size_t)i * 3;
                    vx[k] = __ldg(v);
                    vy[k] = __ldg(v + 1);
                    vz[k] = __ldg(v + 2);
                }
            }
        }
        uint32_t old32[kPrepItems];
        bool have_old[kPrepItems];
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {  // the previous frame's keys come with the tile's other loads
            old32[k] = 0u;
            have_old[k] = fetch_old_key(p, base + k * kThreads < p.n, base + k * kThreads, &old32[k]);
        }
        if (threadIdx.x < 48) {
            const uint32_t m = threadIdx.x >> 4, w = threadIdx.x & 15u;
            const uint64_t* mask = m == 0 ? p.mask_t : m == 1 ? p.mask_s : (do_int ? p.mask_v : nullptr);
            const uint32_t gw = tile * 16u + w;
            unsigned long long word = ~0ull;  // no mask: present on every entity
            if (m == 2 && !do_int) word = 0ull;
            else if (mask != nullptr) word = gw < ((p.n + 63u) >> 6) ? __ldg(mask + gw) : 0ull;
            sh.mwords[par][m][w] = word;
        }
        __syncthreads();  // also orders the zeroing of tcnt[par] above before this tile's counting
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t li = k * kThreads + threadIdx.x;
            const bool inb = base + k * kThreads < p.n;
            has_t[k] = inb && ((sh.mwords[par][0][li >> 6] >> (li & 63u)) & 1ull);
            rend[k] = inb && ((sh.mwords[par][1][li >> 6] >> (li & 63u)) & 1ull);
            has_v[k] = inb && ((sh.mwords[par][2][li >> 6] >> (li & 63u)) & 1ull);
        }
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t i = base + k * kThreads;
            rend[k] = rend[k] && has_t[k];
            if (do_int && has_t[k] && has_v[k]) {
                integrate_ticks(a0[k], vx[k], vy[k], vz[k], p.dt, p.integrate);
                p.rows.t[(size_t)i * p.rows.st] = a0[k];
            }
        }
#pragma unroll
        for (int k = 0; k < kPrepItems; ++k) {
            const uint32_t i = base + k * kThreads;
            uint64_t key = 0;
            if (rend[k]) key = flat_key(p.rows, i, a0[k], a1[k]);
            // this 1024-entity tile is two 512-entity emit tiles (shared-memory counts) or
            // sixteen 64-entity warp tiles (direct global counts of the small path)
            prep_account(p, sh, acc, i < p.n, rend[k], key, i, small, i >> p.emit_tile_shift,
                         tiles_out ? sh.tcnt[par][wide_tiles ? 0 : (k >> 1)] : nullptr, have_old[k], old32[k]);
        }
        if (tiles_out) {
            __syncthreads();
            if (wide_tiles) {
                // half of a scatter tile: add to its (zeroed) counts
                for (uint32_t d = threadIdx.x; d < R0; d += kThreads) {
                    const uint32_t cnt = sh.tcnt[par][0][d];
                    if (cnt) atomicAdd(&p.tile_counts[(size_t)d * p.emit_tiles + (tile >> 1)], cnt);
                }
            } else {
                for (uint32_t d = threadIdx.x; d < 2 * R0; d += kThreads) {
                    const uint32_t h = d / R0, dd = d - h * R0;
                    const uint32_t et = tile * 2 + h;
                    if (et < p.emit_tiles) p.tile_counts[(size_t)dd * p.emit_tiles + et] = sh.tcnt[par][h][dd];
                }
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
// elect one loader and share the parent's world row by warp shuffle. One CTA's share of the level: chunks
// first_chunk, first_chunk + chunk_stride, ...
__device__ __forceinline__ void level_pass(const LevelParams& lp, PrepShared& sh, PrepAcc& acc, uint32_t first_chunk, uint32_t chunk_stride) {
    const PrepParams& p = lp.prep;
    const uint32_t nchunks = (lp.count + kThreads - 1) / kThreads;
    for (uint32_t chunk = first_chunk; chunk < nchunks; chunk += chunk_stride) {
        const uint32_t slot = chunk * kThreads + threadIdx.x;
        const bool inb = slot < lp.count;
        uint32_t e = 0;
        if (inb) e = lp.list ? __ldg(lp.list + lp.first + slot) : lp.first + slot;
        // Every load of the entity is issued before any is consumed: rows, velocities and parent
        // links exist (zero / none) for every id, so they do not wait for the presence bits.
        Row r;
        r.a0 = r.a1 = r.b0 = r.b1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float vx = 0.0f, vy = 0.0f, vz = 0.0f;
        uint32_t par = 0xFFFFFFFFu;
        bool live = false, moves = false;
        const bool do_int = p.integrate && p.vel != nullptr;
        if (inb) {
            r = load_row(p.rows, e);
            if (do_int) {
                const float* v = p.vel + (size_t)e * 3;
                vx = __ldg(v);
                vy = __ldg(v + 1);
                vz = __ldg(v + 2);
                moves = mask_bit(p.mask_v, e);
            }
            if (lp.level > 0) par = __ldg(lp.parent + e) - lp.global_first;
            live = mask_bit(p.mask_t, e);  // level 0 also lists entities without a Transform
        }
        Affine w;
        if (live) {
            if (do_int && moves) {
                integrate_ticks(r.a0, vx, vy, vz, p.dt, p.integrate);
                p.rows.t[(size_t)e * p.rows.st] = r.a0;
            }
            w = compose_local(r);
        } else {
            par = 0xFFFFFFFFu;
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
        prep_account(p, sh, acc, inb, rend, key, e, false, 0u, nullptr);
    }
}

__global__ void __launch_bounds__(kThreads) prepare_level_kernel(const LevelParams lp) {
    __shared__ PrepShared sh;
    prep_init(lp.prep, sh);
    PrepAcc acc;
    level_pass(lp, sh, acc, blockIdx.x, gridDim.x);
    prep_flush(lp.prep, sh, acc);
}

// The top of a scene graph is a handful of tiny levels (1, 8, 74, ... entities), each waiting for the one above: as
// separate launches they cost a launch plus a chain of dependent loads apiece. ONE CTA walks a span of consecutive small
// levels instead; a block barrier (the world rows a level wrote are read by the next through plain loads) stands where
// the kernel boundary stood.
constexpr int kMaxSpanLevels = 8;
constexpr uint32_t kSpanLevelEntities = 1024;  // levels up to this size join a span
struct LevelSpan {
    uint32_t n;
    uint32_t first[kMaxSpanLevels], count[kMaxSpanLevels], level[kMaxSpanLevels];
};
__global__ void __launch_bounds__(kThreads) prepare_level_span_kernel(LevelParams lp, const LevelSpan span) {
    __shared__ PrepShared sh;
    prep_init(lp.prep, sh);
    PrepAcc acc;
    for (uint32_t i = 0; i < span.n; ++i) {
        lp.first = span.first[i];
        lp.count = span.count[i];
        lp.level = span.level[i];
        level_pass(lp, sh, acc, 0u, 1u);
        __threadfence();
        __syncthreads();
    }
    prep_flush(lp.prep, sh, acc);
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

}  // namespace tb
