// Sorted frames whose row reads stay inside the L2: the chunked sort.
//
// A frame whose draw order is not the storage order has to read every 64-byte row once, and a row read at random
// costs a whole 128-byte DRAM fetch (profiles/README.md: 0.73 ms per 2^24 rows, whatever the kernel does). But the
// order is (key, entity id), so inside one key the entities are ascending ids: cut the id range into chunks of
// C = 2^17..2^20 entities (8-64 MB of rows, a fraction of the 126 MB L2) and the global order is, per key, chunk 0's
// entities, then chunk 1's, ... Each (chunk, key) pair is one contiguous RUN of the output. So the frame
//   1. counts H[chunk][key] in the key pass                                  (nchunks x 2^bits counters, L2-resident)
//   2. places: run start G(key, chunk) = exclusive scan of H in (key, chunk) order, remap = G - (chunk-local start)
//   3. sorts every chunk on its own (stable LSD passes over (key, id) pairs, all chunks in one launch; the last pass
//      writes the list of (entity, destination) in chunk-major order)
//   4. emits chunk by chunk: gathers rows inside the chunk's window (both halves of every 128-byte fetch get used
//      while the line is still in L2: measured 0.23-0.29 ms per 2^24 rows instead of 0.67, profiles/r2_window_gather.jsonl)
//      and writes every run (2^log2(C) / 2^bits entities on average) as whole lines.
// The same placement arithmetic as the multi-GPU frame (shard.cuh), with chunks in the role of ranks.
// The list survives the frame: a frame whose keys did not change is step 4 alone (emit_list_kernel<VALIDATE> re-checks
// every key while it emits, and runs the velocity system on the rows it gathers).
#pragma once
#include "emit.cuh"

namespace tb {

struct ChunkGeom {
    uint32_t shift;    // log2 of the chunk size in entities
    uint32_t nchunks;
    uint32_t nbits;    // short-key bits: 1 << nbits counters per chunk
};

// Per key: total over the chunks (-> keytot) and the (layer, texture) bins of the batch table.
__global__ void __launch_bounds__(kThreads) chunk_key_totals_kernel(const uint32_t* __restrict__ H, ChunkGeom g, uint32_t zbits,
                                                                     uint32_t* __restrict__ keytot, uint32_t* lt_bins) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t K = 1u << g.nbits;
    if (k >= K) return;
    uint32_t t = 0;
    for (uint32_t c = 0; c < g.nchunks; ++c) t += H[(size_t)c * K + k];
    keytot[k] = t;
    if (t) atomicAdd(&lt_bins[k >> zbits], t);
}

// remap[c][k] = G(k, c): first position of run (chunk c, key k) in the global draw order. keystart = exclusive scan
// of keytot.
__global__ void __launch_bounds__(kThreads) chunk_place_kernel(const uint32_t* __restrict__ H, const uint32_t* __restrict__ keystart,
                                                                ChunkGeom g, uint32_t* __restrict__ remap,
                                                                const unsigned long long* skip) {
    if (sort_is_cached(skip)) return;
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t K = 1u << g.nbits;
    if (k >= K) return;
    uint32_t run = keystart[k];
    for (uint32_t c = 0; c < g.nchunks; ++c) {
        remap[(size_t)c * K + k] = run;
        run += H[(size_t)c * K + k];
    }
}

// One block per chunk: remap[c][k] -= (first position of key k in chunk c's own sorted list), chunk_cnt[c] = renderables
// of the chunk. Afterwards destination = (position in the chunk's sorted list) + remap[c][key] (mod 2^32).
__global__ void __launch_bounds__(1024) chunk_local_kernel(const uint32_t* __restrict__ H, ChunkGeom g, uint32_t* __restrict__ remap,
                                                            uint32_t* __restrict__ chunk_cnt, const unsigned long long* skip) {
    if (sort_is_cached(skip)) return;
    __shared__ uint32_t s[1024];
    const uint32_t K = 1u << g.nbits, c = blockIdx.x;
    const uint32_t per = (K + 1023u) / 1024u;
    const uint32_t k0 = threadIdx.x * per;
    const uint32_t* h = H + (size_t)c * K;
    uint32_t sum = 0;
    for (uint32_t k = k0; k < k0 + per && k < K; ++k) sum += h[k];
    s[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = s[threadIdx.x] - sum;
    for (uint32_t k = k0; k < k0 + per && k < K; ++k) {
        remap[(size_t)c * K + k] -= run;
        run += h[k];
    }
    if (threadIdx.x == 1023) chunk_cnt[c] = s[1023];
}

// Tile offsets of one pass of the per-chunk sorts: counts / tile_off are digit-major [R][ntiles] like every pass;
// the chunk's tiles are columns [c * tpc, (c + 1) * tpc). Output order inside a chunk: digit, then tile, then rank;
// chunk c's list starts at c << shift. One block per chunk, one warp per digit row.
__global__ void __launch_bounds__(1024) chunk_tile_scan_kernel(const uint32_t* __restrict__ counts, uint32_t ntiles, uint32_t tpc,
                                                                uint32_t bits, uint32_t chunk_shift, uint32_t* __restrict__ tile_off,
                                                                const unsigned long long* skip) {
    if (sort_is_cached(skip)) return;
    __shared__ uint32_t rowsum[256];
    const uint32_t R = 1u << bits, c = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t t0 = c * tpc;
    const uint32_t nt = min(tpc, ntiles > t0 ? ntiles - t0 : 0u);
    for (uint32_t d = warp; d < R; d += 32) {
        uint32_t v = 0;
        for (uint32_t t = lane; t < nt; t += 32) v += counts[(size_t)d * ntiles + t0 + t];
        v = __reduce_add_sync(0xffffffffu, v);
        if (lane == 0) rowsum[d] = v;
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of up to 256 row sums
        uint32_t carry = 0;
        for (uint32_t d0 = 0; d0 < R; d0 += 32) {
            const uint32_t v = d0 + lane < R ? rowsum[d0 + lane] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= (uint32_t)o) x += y;
            }
            if (d0 + lane < R) rowsum[d0 + lane] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
    __syncthreads();
    for (uint32_t d = warp; d < R; d += 32) {
        uint32_t carry = (c << chunk_shift) + rowsum[d];
        for (uint32_t tb0 = 0; tb0 < nt; tb0 += 32) {
            const uint32_t t = tb0 + lane;
            const uint32_t v = t < nt ? counts[(size_t)d * ntiles + t0 + t] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= (uint32_t)o) x += y;
            }
            if (t < nt) tile_off[(size_t)d * ntiles + t0 + t] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
}

// ---- step 4: emit over the chunk-major (entity, destination) list ---------------------------------------------------------
//
// Every warp is its own pipeline (as in emit_ordered_kernel): 64 consecutive list positions per step; the ids,
// destinations (and kept keys) of tile k+2 arrive by TMA bulk copy, the rows (and velocities) of tile k+1 are
// gathered by id with cp.async into the 128-byte-swizzled buffer, tile k is composed and written to its
// destinations as whole 16-byte chunks (consecutive list positions of one run have consecutive destinations).
//   VALIDATE   flat frames without a key pass: every row's key is recomputed and compared with the kept one;
//              key statistics, dictionary misses and the "changed" flag go to `stats` for the host to confirm
//   integrate  the velocity system runs on the gathered rows (written back to the row storage) before the compose
struct ListParams {
    const uint32_t* ids;        // [npad] global entity id per list position (chunk-padded)
    const uint32_t* dst;        // [npad] position in the global draw order
    const uint32_t* keys;       // [npad] kept short keys (VALIDATE)
    const uint32_t* chunk_cnt;  // [nchunks]
    uint32_t chunk_shift;
    uint32_t npad;              // nchunks << chunk_shift
    RowView rows;
    const WorldRow* world;      // non-null: hierarchy frames emit from the world rows
    const float* vel;
    const uint64_t* mask_v;
    float dt;
    uint32_t integrate;         // ticks
    float4* instances;
    float4* vertices;
    uint32_t* order;
    uint32_t global_first;
    unsigned long long* stats;
    float4 vp[4];
    uint32_t has_vp;
    uint32_t l2_hint;           // 1: row gathers carry the L2 evict_last policy
    FramePlan plan;
};

struct ListWarpShared {
    float4 rows[2][kWarpTile * 4];
    float4 stage[32 * kStageF4];
    float vel[2][kWarpTile * 3];
    uint32_t ids[3][kWarpTile];
    uint32_t dst[3][kWarpTile];
    unsigned long long kbar[3];
};
struct ListShared {
    ListWarpShared w[kWarps];
};

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// L2 cache policies: the gathered rows of a chunk are touched again within microseconds (the other entities of the
// same 128-byte line), while 148 bytes per entity of output stream through the same L2 — without a hint that stream
// evicts the rows before their neighbours are read (ncu: 147-228 B of DRAM reads per entity for 72 B needed).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_hint(void* dst, const void* src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(policy) : "memory");
}

template <bool VALIDATE>
__global__ void __launch_bounds__(kThreads, 2) emit_list_kernel(const ListParams ep) {
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    __shared__ uint32_t s_zdict[VALIDATE ? kMaxZDict : 1];
    if (VALIDATE) {
        for (uint32_t k = threadIdx.x; k < (uint32_t)kMaxZDict; k += kThreads) s_zdict[k] = ep.plan.zdict[k];
        __syncthreads();  // the only block barrier, before the warps go their own ways
    }
    PrepAcc acc;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ListWarpShared& sm = reinterpret_cast<ListShared*>(emit_smem_raw)->w[warp];
    const bool from_world = ep.world != nullptr;
    const bool do_int = ep.integrate != 0 && ep.vel != nullptr && !from_world;
    const uint32_t ntiles = ep.npad / kWarpTile;
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t first = blockIdx.x * kWarps + warp;
    const uint32_t cmask = (1u << ep.chunk_shift) - 1u;
    const uint64_t keep = l2_policy_evict_last();
    const bool hint = ep.l2_hint != 0;

    auto tile_cnt = [&](uint32_t t) -> uint32_t {
        if (t >= ntiles) return 0u;
        const uint32_t base = t * kWarpTile, local = base & cmask;
        const uint32_t cnt = __ldg(ep.chunk_cnt + (base >> ep.chunk_shift));
        return cnt > local ? min(kWarpTile, cnt - local) : 0u;
    };
    auto fetch_list = [&](uint32_t t, uint32_t ring) {  // lane 0
        const uint32_t cnt = tile_cnt(t);
        if (cnt == 0) return;
        const uint32_t bytes = (cnt * 4u + 15u) & ~15u;
        mbar_expect_tx(&sm.kbar[ring], bytes * 2u);
        bulk_g2s(sm.ids[ring], ep.ids + (size_t)t * kWarpTile, bytes, &sm.kbar[ring]);
        bulk_g2s(sm.dst[ring], ep.dst + (size_t)t * kWarpTile, bytes, &sm.kbar[ring]);
    };
    auto gather = [&](uint32_t t, uint32_t buf, uint32_t ring) {  // all lanes: two rows each
        const uint32_t cnt = tile_cnt(t);
        float4* d = sm.rows[buf];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t li = k * 32 + lane;
            if (li < cnt) {
                const uint32_t e = sm.ids[ring][li] - ep.global_first;
                const float4 *s0, *s1, *s2, *s3;
                if (from_world) {
                    s0 = reinterpret_cast<const float4*>(ep.world + e);
                    s1 = s0 + 1;
                    s2 = s0 + 2;
                    s3 = s0 + 3;
                } else {
                    s0 = ep.rows.t + (size_t)e * ep.rows.st;
                    s1 = ep.rows.a + (size_t)e * ep.rows.st;
                    s2 = ep.rows.b + (size_t)e * ep.rows.sb;
                    s3 = s2 + 1;
                }
                if (hint) {
                    cp_async16_hint(d + row_slot(li, 0), s0, keep);
                    cp_async16_hint(d + row_slot(li, 1), s1, keep);
                    cp_async16_hint(d + row_slot(li, 2), s2, keep);
                    cp_async16_hint(d + row_slot(li, 3), s3, keep);
                } else {
                    cp_async16(d + row_slot(li, 0), s0);
                    cp_async16(d + row_slot(li, 1), s1);
                    cp_async16(d + row_slot(li, 2), s2);
                    cp_async16(d + row_slot(li, 3), s3);
                }
                if (!from_world) {
                    if (do_int) {
                        const float* v = ep.vel + (size_t)e * 3;
                        float* sv = sm.vel[buf] + li * 3;
                        cp_async4(sv, v);
                        cp_async4(sv + 1, v + 1);
                        cp_async4(sv + 2, v + 2);
                    }
                }
            }
        }
    };

    if (lane == 0) {
        mbar_init(&sm.kbar[0], 1);
        mbar_init(&sm.kbar[1], 1);
        mbar_init(&sm.kbar[2], 1);
    }
    __syncwarp();
    uint32_t kpar = 0;  // phase parity of each ring slot (bit r)
    if (lane == 0) {
        fetch_list(first, 0);
        fetch_list(first + nwarps, 1);
    }
    if (tile_cnt(first)) {
        mbar_wait(&sm.kbar[0], 0);
        kpar ^= 1u;
    }
    gather(first, 0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    uint32_t ring = 0, cur = 0;
    for (uint32_t tile = first; tile < ntiles; tile += nwarps) {
        const uint32_t ring1 = ring == 2 ? 0u : ring + 1u, ring2 = ring1 == 2 ? 0u : ring1 + 1u;
        const uint32_t next = tile + nwarps;
        if (lane == 0) fetch_list(next + nwarps, ring2);  // slot ring2 was consumed two iterations ago
        if (tile_cnt(next)) {
            mbar_wait(&sm.kbar[ring1], (kpar >> ring1) & 1u);
            kpar ^= 1u << ring1;
        }
        gather(next, cur ^ 1u, ring1);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
        __syncwarp();  // every lane's gathers of this tile have landed
        float4* rows = sm.rows[cur];
        const uint32_t cnt = tile_cnt(tile);
        if (cnt == 0) {
            cur ^= 1u;
            ring = ring1;
            continue;
        }
        uint32_t kept[2] = {0u, 0u};
        if (VALIDATE) {  // issued here, consumed after the compose
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if ((uint32_t)k * 32u + lane < cnt) kept[k] = __ldg(ep.keys + (size_t)tile * kWarpTile + k * 32 + lane);
        }
        if (do_int) {
            // the velocity system on the gathered rows: shared memory (the compose reads it) and the row storage
            const float* vel = sm.vel[cur];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t li = k * 32 + lane;
                if (li < cnt) {
                    const uint32_t e = sm.ids[ring][li] - ep.global_first;
                    if (mask_bit(ep.mask_v, e)) {
                        float4 a0 = rows[row_slot(li, 0)];
                        integrate_ticks(a0, vel[li * 3], vel[li * 3 + 1], vel[li * 3 + 2], ep.dt, ep.integrate);
                        rows[row_slot(li, 0)] = a0;
                        ep.rows.t[(size_t)e * ep.rows.st] = a0;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t li = k * 32 + lane;
            const bool live = li < cnt;
            uint32_t dst = 0;
            InstanceOut inst;
            float4 vtx[4];
            if (live) {
                dst = sm.dst[ring][li];
                Affine w;
                float sw, shh;
                uint32_t tl;
                if (from_world) {
                    w.r0 = rows[row_slot(li, 0)];
                    w.r1 = rows[row_slot(li, 1)];
                    w.r2 = rows[row_slot(li, 2)];
                    const float4 sv = rows[row_slot(li, 3)];
                    sw = sv.x;
                    shh = sv.y;
                    tl = __float_as_uint(sv.z);
                } else {
                    Row r;
                    r.a0 = rows[row_slot(li, 0)];
                    r.a1 = rows[row_slot(li, 1)];
                    r.b0 = rows[row_slot(li, 2)];
                    r.b1 = rows[row_slot(li, 3)];
                    w = compose_local(r);
                    sw = r.b1.y;
                    shh = r.b1.z;
                    tl = __float_as_uint(r.a1.w);
                    if (VALIDATE) {
                        const float z = is_planar(r.a1) ? planar_z(r.a0, r.a1) : local_z(r);
                        const uint64_t key = sort_key(tl >> 16, tl & 0xFFFFu, z);
                        acc.k_or |= key;
                        acc.k_and &= key;
                        acc.count += 1;
                        bool miss = false;
                        if ((uint32_t)plan_short_key(ep.plan, s_zdict, key, &miss) != kept[k]) acc.changed = true;
                        if (miss) acc.miss = true;
                    }
                }
                make_outputs(ep, w, sw, shh, tl, inst, vtx);
                ep.order[dst] = sm.ids[ring][li];
            }
            const uint32_t nlive = cnt > (uint32_t)k * 32u ? min(32u, cnt - k * 32u) : 0u;
            staged_store(sm.stage, lane, live, nlive, dst, inst, vtx, ep.instances, ep.vertices);
        }
        __syncwarp();  // the row buffer and the list slot are free for the next iterations
        cur ^= 1u;
        ring = ring1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (VALIDATE) {
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
        if (__any_sync(0xffffffffu, acc.miss) && lane == 0) atomicOr(&ep.stats[4], 1ull);
        if (__any_sync(0xffffffffu, acc.changed) && lane == 0) atomicOr(&ep.stats[5], 1ull);
    }
}

}  // namespace tb
