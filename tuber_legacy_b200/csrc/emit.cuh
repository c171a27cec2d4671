// The last radix pass fused with compose + instance / quad-vertex emission:
//   emit_kernel        persistent CTAs over 512-position tiles, any plan (multi-pass frames gather rows)
//   emit_small_kernel  single-pass plans with <= 8 buckets: every warp is its own TMA-fed pipeline,
//                      no block barrier; the FUSED variant is the whole steady-state frame in one kernel
#pragma once
#include "placement.cuh"

namespace tb {

// ---- final pass fused with compose + emit -------------------------------------------------------

struct InstanceOut {
    float4 c0, c1, c2, c3, s;  // column-major model (matrix.rs:219-251) + {size.w, size.h, texture, layer}
};
static_assert(sizeof(InstanceOut) == 80, "instance is 80 bytes");

template <typename KeyT>
struct EmitParams {
    PassParams pass;
    const KeyT* keys_in;     // short keys in input order; nullptr: recomputed from the rows (single-pass flat frames)
    const uint32_t* idx_in;  // entity per input position; nullptr: the entity is the position (single-pass plans)
    const uint64_t* mask_t;
    const uint64_t* mask_s;
    RowView rows;
    const WorldRow* world;   // non-null: hierarchy path, world rows instead of composing
    float4* instances;       // 5 float4 per position
    float4* vertices;        // 4 float4 per position; nullptr: do not emit vertices (a presenter that
                             // receives instances from other GPUs regenerates them: vertices_from_instances)
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
    uint32_t integrate;      // engine ticks of `dt` to run (0: none)
    unsigned long long* stats;  // [0] OR of keys, [1] OR of ~keys, [2] renderable count
    // camera ([SPEC]): rows of a row-major view-projection matrix folded into the
    // emitted geometry (instance model = vp * world, corners transformed by it); has_vp == 0: none
    float4 vp[4];
    uint32_t has_vp;
    // sorted frames that keep the previous frame's sort: `changed` points at the key pass's
    // "some key changed" flag. run_when 1: this kernel runs only if it is set (the ranking emit of a re-sorted
    // frame), 2: only if it is clear (emit_ordered_kernel over the cached draw order), 0: always.
    const unsigned long long* changed;
    uint32_t run_when;
    // the ranking emit of a (re)sorted frame also leaves the draw order and its short keys for the frames
    // that will keep them: final_idx[pos] = entity id (global), final_keys[pos] = short key
    uint32_t* final_idx;
    KeyT* final_keys;
    // scene graphs on the draw-ordered working storage (emit_sequential_kernel<KeyT, true> with world != nullptr): draw
    // positions whose entity has a local id >= leaf_first belong to the LAST level (leaves: nobody's parent). Their LOCAL
    // rows, velocities and parent ids live in draw order (sequential streams); the kernel ticks them, composes
    // world = world[parent] * local from the parent's world row (the parents are a small, L2-resident set, written by
    // the level pass of the inner levels just before) and re-checks their keys. The other positions (inner levels, a
    // small minority) read the world row the level pass wrote for them.
    const uint32_t* seq_par;  // parent (global id) per draw position
    uint32_t leaf_first;
    // emit_sequential_kernel: the draw-ordered copy. World rows (static scene graphs): 64 bytes per position. Local rows
    // (flat scenes, scene-graph leaves; seq_split != 0): a dense block of translation chunks A0 (16 B per position: what
    // the velocity system writes back, as a dense stream) followed, from float4 index seq_split, by A1 B0 B1 (48 B per
    // position: a 3-float4 stride is bank-conflict-free in shared memory)
    const float4* seq_rows;
    float4* seq_rows_rw;     // ... the same storage, writable
    uint32_t seq_split;
    const float* seq_vel;    // ... velocities in draw order (zero where the entity has none)
    // ... one launch may serve a range of the copy's 64-position tiles only (list frames: one launch per chunk of the
    // shard's draw order, each followed by a completion signal to the presenting rank); tile_end == 0: all tiles
    uint32_t tile_first, tile_end;
    // ... and may write, instead of instance + quad vertices, the WORLD ROW of every position (3x4 affine, then size /
    // texture|layer: a WorldRow, 64 bytes) into `vertices` (`instances` == nullptr): what a shard ships to the presenting
    // rank in steady list frames — 68 instead of 84 bytes per element cross the link, and the presenter expands the rows
    // (expand_world_rows_kernel: the same make_outputs, bit-identical)
    uint32_t world_rows_out;
    FramePlan plan;             // last: its z dictionary is the cold tail of the parameter block
};

constexpr int kEmitItems = 2;
constexpr uint32_t kEmitTile = kThreads * kEmitItems;  // 512 entities per tile
constexpr int kStageF4 = 5;  // per-lane staging slot: 80 B (one instance, or 64 B of vertices)

// Dynamic shared memory of emit_kernel: ~110 KB with 32-bit short keys, two persistent CTAs per SM.
template <typename KeyT>
struct EmitShared {
    float4 rows[2][kEmitTile * 4];      // 2 x 32 KB: this tile's rows and the prefetched next tile's
                                        // (gathered rows: 16-byte chunks XOR-swizzled, see row_slot)
    KeyT keys[3][kEmitTile];            // ring of short keys: tiles k, k+1, k+2 of a sorted stream
    uint32_t idx[3][kEmitTile];         // ring of entity ids of the same tiles
    unsigned long long masks[2][2][8];  // presence words (Transform, Sprite) of an in-order tile
    RankShared rs;
    uint32_t s_dst[kEmitTile];          // output position per sorted tile slot
    uint16_t s_li[kEmitTile];           // tile-local item index per sorted tile slot
    float4 stage[kWarps][32 * kStageF4];  // per-warp staging for full-line stores
    unsigned long long mbar[2];         // completion barriers of the row copies (in-order passes)
    unsigned long long kbar[3];         // completion barriers of the (key, entity) ring
    uint32_t zdict[kMaxZDict];          // the plan's z dictionary (passes that recompute keys from rows)
};

// 16-byte chunk c (0..3) of tile-local row li lives at rows[row_slot(li, c)]: the 128-byte
// swizzle (chunk index XOR line index), so both consecutive rows and the scattered rows of one
// sort bucket spread over all eight 16-byte bank groups.
__device__ __forceinline__ uint32_t row_slot(uint32_t li, uint32_t c) { return ((li << 2) | c) ^ ((li >> 1) & 7u); }

// address of 16-byte chunk c (0..3 = A0 A1 B0 B1) of entity e's row, in the row storage or the world-row array
template <typename KeyT>
__device__ __forceinline__ const float4* row_chunk(const EmitParams<KeyT>& ep, bool from_world, uint32_t e, uint32_t c) {
    if (from_world) return reinterpret_cast<const float4*>(ep.world + e) + c;
    return c == 0   ? ep.rows.t + (size_t)e * ep.rows.st
           : c == 1 ? ep.rows.a + (size_t)e * ep.rows.st
                    : ep.rows.b + (size_t)e * ep.rows.sb + (c - 2u);
}

template <typename Params>
__device__ __forceinline__ void make_outputs(const Params& ep, const Affine& world, float sw, float shh, uint32_t tl,
                                             InstanceOut& inst, float4 (&vtx)[4]) {
    Affine w = world;
    float4 r3 = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    if (ep.has_vp) {  // Matrix4::mul(view_proj, world), matrix.rs:174-194
        w.r0 = mul_row(ep.vp[0], world);
        w.r1 = mul_row(ep.vp[1], world);
        w.r2 = mul_row(ep.vp[2], world);
        r3 = mul_row(ep.vp[3], world);
    }
    inst.c0 = make_float4(w.r0.x, w.r1.x, w.r2.x, r3.x);
    inst.c1 = make_float4(w.r0.y, w.r1.y, w.r2.y, r3.y);
    inst.c2 = make_float4(w.r0.z, w.r1.z, w.r2.z, r3.z);
    inst.c3 = make_float4(w.r0.w, w.r1.w, w.r2.w, r3.w);
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

// make_outputs, or (ROWS: the list-frame instantiations, EmitParams::world_rows_out) the world row itself in its place. A
// compile-time choice: as a run-time branch it cost the single-GPU headline kernel 2 % (0.656 -> 0.669 ms at 2^24 entities).
template <bool ROWS, typename Params>
__device__ __forceinline__ void make_outputs_or_row(const Params& ep, const Affine& world, float sw, float shh, uint32_t tl,
                                                    InstanceOut& inst, float4 (&vtx)[4]) {
    if (ROWS) {
        vtx[0] = world.r0;
        vtx[1] = world.r1;
        vtx[2] = world.r2;
        vtx[3] = make_float4(sw, shh, __uint_as_float(tl), 0.0f);
    } else {
        make_outputs(ep, world, sw, shh, tl, inst, vtx);
    }
}

// Writes one warp's 32 results (lane i holds the instance and vertices of output position dst)
// as whole 128-byte lines: staged in the warp's shared buffer, then streamed out 16 bytes per lane.
// Instances use a 20-word lane stride (conflict-free 128-bit stores; the read-out is linear);
// vertices a 16-word stride with the chunk index XORed by bits of the lane (conflict-free both ways).
template <bool INST = true>
__device__ __forceinline__ void staged_store(float4* stage, uint32_t lane, bool live, uint32_t nlive, uint32_t dst,
                                             const InstanceOut& inst, const float4 (&vtx)[4], float4* instances,
                                             float4* vertices) {
    if (INST) {  // false: world rows only (through the 64-byte path below), see EmitParams::world_rows_out
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
    }
    if (vertices == nullptr) return;  // warp-uniform
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

// Final radix pass fused with compose + emit, any plan. Persistent CTAs (two per SM) walk
// 512-position tiles (tile = blockIdx.x + k * gridDim.x); every tile's output positions come from
// the precomputed per-tile digit offsets, so no CTA ever waits for another.
//   1. The tile's rows reach shared memory asynchronously, one tile ahead of the compute:
//      - passes that read positions in entity order (single-pass plans): TMA bulk copies of the
//        next tile's sector blocks, presence words and keys;
//      - passes over a sorted (key, entity) stream: the (key, entity) pairs of tile k+2 arrive by TMA
//        while the rows of tile k+1 are gathered by entity id with cp.async (into a 128-byte-swizzled
//        layout) and tile k is ranked and emitted — a three-deep software pipeline, all in shared memory.
//   2. Stable rank of the tile's digits (match.any per warp, per-warp counters, one block scan).
//   3. Warps walk the tile in sorted order (consecutive slots of a digit are consecutive output
//      positions), compose from shared memory, stage 32 results and write whole 128-byte lines.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads, 2) emit_kernel(const EmitParams<KeyT> ep) {
    constexpr int ITEMS = kEmitItems;
    constexpr uint32_t TILE = kEmitTile;
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    EmitShared<KeyT>& sm = *reinterpret_cast<EmitShared<KeyT>*>(emit_smem_raw);
    RankShared& rs = sm.rs;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ep.run_when == 1u && *reinterpret_cast<const volatile unsigned long long*>(ep.changed) == 0ull) return;
    const bool in_order = ep.idx_in == nullptr;  // positions are entity ids: contiguous rows
    const bool from_world = ep.world != nullptr;
    const bool keys_given = ep.keys_in != nullptr;
    const uint32_t n_eff = ep.pass.count_ptr ? min(ep.pass.n, (uint32_t)*ep.pass.count_ptr) : ep.pass.n;
    const uint32_t ntiles = ep.pass.ntiles;
    const uint32_t mask = (1u << ep.pass.bits) - 1u;
    const uint32_t step = gridDim.x;

    // Shared-memory row layouts. In-order passes land rows with TMA bulk copies, i.e. as they lie in
    // global memory (split arrays: T block, A block, B block of the tile; else 64-byte rows); gathered
    // rows are placed chunk by chunk with cp.async into the bank-swizzled layout (row_slot).
    const bool split = in_order && !from_world && ep.rows.st == 1;
    auto chunk = [&](uint32_t li, uint32_t c) -> uint32_t {
        if (!in_order) return row_slot(li, c);
        if (!split) return li * 4u + c;
        return c == 0 ? li : c == 1 ? TILE + li : 2u * TILE + li * 2u + (c - 2u);
    };
    auto tile_count = [&](uint32_t tile) -> uint32_t {
        const uint32_t base = tile * TILE;
        return (tile < ntiles && base < n_eff) ? min(TILE, n_eff - base) : 0u;
    };
    // in-order passes (thread 0): TMA bulk copies of one tile's rows / presence words / keys
    auto fetch_rows = [&](uint32_t tile, uint32_t buf, uint32_t ring) {
        const uint32_t base = tile * TILE;
        const uint32_t cnt = tile_count(tile);
        if (cnt == 0) return;
        const uint32_t mwords = ((cnt + 127u) >> 7) * 2u;  // presence words, padded to 16 bytes
        const uint32_t kbytes = keys_given ? ((cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u) : 0u;
        const uint32_t mbytes = ((ep.mask_t ? 1u : 0u) + (ep.mask_s ? 1u : 0u)) * mwords * 8u;
        float4* dst = sm.rows[buf];
        mbar_expect_tx(&sm.mbar[buf], cnt * 64u + mbytes + kbytes);
        if (from_world) {
            bulk_g2s(dst, ep.world + base, cnt * 64u, &sm.mbar[buf]);
        } else if (split) {
            bulk_g2s(dst, ep.rows.t + base, cnt * 16u, &sm.mbar[buf]);
            bulk_g2s(dst + TILE, ep.rows.a + base, cnt * 16u, &sm.mbar[buf]);
            bulk_g2s(dst + 2 * TILE, ep.rows.b + (size_t)base * 2, cnt * 32u, &sm.mbar[buf]);
        } else {
            bulk_g2s(dst, ep.rows.t + (size_t)base * 4, cnt * 64u, &sm.mbar[buf]);
        }
        if (ep.mask_t) bulk_g2s(sm.masks[buf][0], ep.mask_t + (base >> 6), mwords * 8u, &sm.mbar[buf]);
        if (ep.mask_s) bulk_g2s(sm.masks[buf][1], ep.mask_s + (base >> 6), mwords * 8u, &sm.mbar[buf]);
        if (keys_given) bulk_g2s(sm.keys[ring], ep.keys_in + base, kbytes, &sm.mbar[buf]);
    };
    // sorted-stream passes (thread 0): TMA bulk copies of one tile's (key, entity) pairs
    auto fetch_pairs = [&](uint32_t tile, uint32_t ring) {
        const uint32_t base = tile * TILE;
        const uint32_t cnt = tile_count(tile);
        if (cnt == 0) return;
        const uint32_t kbytes = (cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u;
        const uint32_t ibytes = (cnt * 4u + 15u) & ~15u;
        mbar_expect_tx(&sm.kbar[ring], kbytes + ibytes);
        bulk_g2s(sm.keys[ring], ep.keys_in + base, kbytes, &sm.kbar[ring]);
        bulk_g2s(sm.idx[ring], ep.idx_in + base, ibytes, &sm.kbar[ring]);
    };
    // sorted-stream passes (all threads): gather the rows of one tile whose pairs are in ring slot `ring`
    auto gather_rows = [&](uint32_t tile, uint32_t buf, uint32_t ring) {
        const uint32_t cnt = tile_count(tile);
        float4* d = sm.rows[buf];
        // the four lanes of a quad fetch the four 16-byte chunks of one row: an interleaved (or world) row is ONE
        // 64-byte request, not four 16-byte ones (random rows over 1 GiB: 39 vs 23 G rows/s, tools/microbench_pages.cu)
        const uint32_t c = lane & 3u;
#pragma unroll
        for (int k = 0; k < ITEMS * 4; ++k) {
            const uint32_t li = warp * (ITEMS * 32) + k * 8 + (lane >> 2);
            if (li < cnt) cp_async16(d + row_slot(li, c), row_chunk(ep, from_world, sm.idx[ring][li], c));
        }
    };

    uint32_t tile = blockIdx.x;
    uint32_t kpar = 0;  // phase parities of the three kbar slots (bit r)
    for (uint32_t k = threadIdx.x; k < (uint32_t)kMaxZDict; k += kThreads) sm.zdict[k] = ep.plan.zdict[k];
    if (threadIdx.x == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        mbar_init(&sm.kbar[0], 1);
        mbar_init(&sm.kbar[1], 1);
        mbar_init(&sm.kbar[2], 1);
        if (in_order) {
            fetch_rows(tile, 0, 0);
        } else {
            fetch_pairs(tile, 0);
            fetch_pairs(tile + step, 1);
        }
    }
    rank_reset(rs, ep.pass.bits);
    __syncthreads();
    if (!in_order) {
        if (tile_count(tile)) {
            mbar_wait(&sm.kbar[0], 0);
            kpar ^= 1u;
        }
        gather_rows(tile, 0, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t cur = 0, ring = 0, parity0 = 0, parity1 = 0;
    for (; tile < ntiles; tile += step) {
        const uint32_t next = tile + step;
        const uint32_t ring1 = ring == 2 ? 0 : ring + 1, ring2 = ring == 0 ? 2 : ring - 1;
        const float4* rows = sm.rows[cur];
        const uint32_t tile_base = tile * TILE;
        const uint32_t tcnt = tile_count(tile);
        uint32_t ent[ITEMS], digit[ITEMS], slot[ITEMS], shift_g[ITEMS];
        bool valid[ITEMS];
        if (in_order) {
            if (threadIdx.x == 0) fetch_rows(next, cur ^ 1u, ring1);  // prefetch; buffer freed by the last barrier
            if (tcnt) {
                mbar_wait(&sm.mbar[cur], cur ? parity1 : parity0);
                if (cur) parity1 ^= 1u; else parity0 ^= 1u;
            }
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t li = warp * (ITEMS * 32) + k * 32 + lane;
                const uint32_t pos = tile_base + li;
                ent[k] = pos;
                valid[k] = li < tcnt;
                if (valid[k]) {
                    const bool t_ok = ep.mask_t == nullptr || ((sm.masks[cur][0][li >> 6] >> (li & 63u)) & 1ull);
                    const bool s_ok = ep.mask_s == nullptr || ((sm.masks[cur][1][li >> 6] >> (li & 63u)) & 1ull);
                    valid[k] = t_ok && s_ok;
                }
                uint64_t sk = 0;
                if (valid[k]) {
                    if (keys_given) {  // hierarchy frames always carry the keys pass A wrote
                        sk = (uint64_t)sm.keys[ring][li];
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
                        bool miss = false;  // pass A of the same frame reports misses
                        sk = plan_short_key(ep.plan, sm.zdict, sort_key(tl >> 16, tl & 0xFFFFu, z), &miss);
                    }
                }
                digit[k] = valid[k] ? ((uint32_t)(sk >> ep.pass.shift) & mask) : kInvalidDigit;
                shift_g[k] = (ep.remap != nullptr && valid[k]) ? __ldg(ep.remap + (uint32_t)sk) : 0u;
            }
        } else {
            // pairs of tile k+2 by TMA, rows of tile k+1 by cp.async, then wait for the rows of tile k
            if (threadIdx.x == 0) fetch_pairs(next + step, ring2);
            if (tile_count(next)) {
                mbar_wait(&sm.kbar[ring1], (kpar >> ring1) & 1u);
                kpar ^= 1u << ring1;
            }
            gather_rows(next, cur ^ 1u, ring1);
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                const uint32_t li = warp * (ITEMS * 32) + k * 32 + lane;
                valid[k] = li < tcnt;
                ent[k] = 0;
                digit[k] = kInvalidDigit;
                shift_g[k] = 0u;
                if (valid[k]) {
                    const KeyT key = sm.keys[ring][li];
                    ent[k] = sm.idx[ring][li];
                    digit[k] = (uint32_t)(key >> ep.pass.shift) & mask;
                    if (ep.remap != nullptr) shift_g[k] = __ldg(ep.remap + (uint32_t)key);
                }
            }
        }
        rank_tile<ITEMS>(rs, ep.pass, tile, digit, slot);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (digit[k] != kInvalidDigit) {
                sm.s_li[slot[k]] = (uint16_t)(warp * (ITEMS * 32) + k * 32 + lane);
                sm.s_dst[slot[k]] = rs.dbase[digit[k]] + slot[k] + shift_g[k];
            }
        }
        __syncthreads();  // also publishes every thread's finished cp.async gathers of this tile
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
                const uint32_t e = in_order ? tile_base + li : sm.idx[ring][li];
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
                make_outputs(ep, w, sw, shh, tl, inst, vtx);
                ep.order[dst] = e + ep.global_first;
                if (ep.final_idx != nullptr) {
                    // the kept order is the SHARD's: local position = global position - the bucket's shift
                    const uint32_t lpos = dst - (ep.remap != nullptr ? __ldg(ep.remap + (uint32_t)sm.keys[ring][li]) : 0u);
                    ep.final_idx[lpos] = e + ep.global_first;
                    ep.final_keys[lpos] = sm.keys[ring][li];
                }
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
                    if (ep.vertices != nullptr) {
                        float4* vp = ep.vertices + (size_t)dst * 4;
                        st_cs_f4(vp, vtx[0]);
                        st_cs_f4(vp + 1, vtx[1]);
                        st_cs_f4(vp + 2, vtx[2]);
                        st_cs_f4(vp + 3, vtx[3]);
                    }
                }
            } else {
                staged_store(stage, lane, live, min(32u, cnt - s0), dst, inst, vtx, ep.instances, ep.vertices);
            }
        }
        rank_reset(rs, ep.pass.bits);
        __syncthreads();  // releases rows[cur], the rank scratch and the pair ring slot for reuse
        cur ^= 1u;
        ring = ring1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
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
// also runs the velocity system (writing the translation chunk back), and accumulates the key statistics with
// which the host validates, every frame, that the cached plan and offsets still hold.
template <typename KeyT, bool FUSED, bool ROWS = false>
__global__ void __launch_bounds__(kThreads, 2) emit_small_kernel(const EmitParams<KeyT> ep) {
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpShared& sm = reinterpret_cast<SmallShared*>(emit_smem_raw)->w[warp];
    const bool do_int = FUSED && ep.integrate && ep.vel != nullptr;
    PrepAcc acc;
    // z dictionary of the plan. The single-kernel frame (FUSED) only runs under plans without z bits.
    __shared__ uint32_t s_zdict[FUSED ? 1 : kMaxZDict];
    if (!FUSED) {
        for (uint32_t k = threadIdx.x; k < (uint32_t)kMaxZDict; k += kThreads) s_zdict[k] = ep.plan.zdict[k];
        __syncthreads();  // the only block barrier of the kernel, before the warps go their own ways
    }
    const bool from_world = ep.world != nullptr;
    const bool keys_given = ep.keys_in != nullptr;
    const uint32_t n = ep.pass.n;
    const uint32_t ntiles = ep.pass.ntiles;  // == ceil(n / 64)
    const uint32_t R = 1u << ep.pass.bits;
    // shared-memory position of chunk c of tile item li: as the TMA bulk copies land it (split
    // arrays: T block, A block, B block of the tile; else 64-byte rows)
    const bool split = !from_world && ep.rows.st == 1;
    const uint32_t s01 = split ? 1u : 4u, sb = split ? 2u : 4u;
    const uint32_t off1 = split ? kWarpTile : 1u, offb = split ? 2u * kWarpTile : 2u;
    const uint32_t lt = lanemask_lt();
    const uint32_t nwarps = gridDim.x * kWarps;
    // a launch may serve a range of the tiles only (list frames in chunks, see EmitParams::tile_first)
    const uint32_t tend = ep.tile_end != 0 ? min(ep.tile_end, ntiles) : ntiles;
    uint32_t tile = ep.tile_first + blockIdx.x * kWarps + warp;

    auto fetch = [&](uint32_t t, uint32_t buf) {  // lane 0: TMA bulk copies of tile t's rows (and keys)
        const uint32_t base = t * kWarpTile;
        const uint32_t cnt = min(kWarpTile, n - base);
        const uint32_t kbytes = keys_given ? ((cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u) : 0u;
        const uint32_t vbytes = do_int ? ((cnt * 12u + 15u) & ~15u) : 0u;
        mbar_expect_tx(&sm.mbar[buf], cnt * 64u + kbytes + vbytes);
        if (from_world) {
            bulk_g2s(sm.rows[buf], ep.world + base, cnt * 64u, &sm.mbar[buf]);
        } else if (split) {
            bulk_g2s(sm.rows[buf], ep.rows.t + base, cnt * 16u, &sm.mbar[buf]);
            bulk_g2s(sm.rows[buf] + kWarpTile, ep.rows.a + base, cnt * 16u, &sm.mbar[buf]);
            bulk_g2s(sm.rows[buf] + 2 * kWarpTile, ep.rows.b + (size_t)base * 2, cnt * 32u, &sm.mbar[buf]);
        } else {
            bulk_g2s(sm.rows[buf], ep.rows.t + (size_t)base * 4, cnt * 64u, &sm.mbar[buf]);
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
        if (tile < tend) fetch(tile, 0);
    }
    __syncwarp();
    unsigned long long raw = tile < tend ? load_raw(tile) : 0ull, raw_next = 0ull;
    uint32_t off = (tile < tend && lane < R) ? __ldg(ep.pass.tile_off + (size_t)lane * ntiles + tile) : 0u;
    uint32_t cur = 0, par0 = 0, par1 = 0;
    for (; tile < tend; tile += nwarps) {
        // prefetch the next tile of this warp: rows by TMA, presence word and bucket offsets in registers
        const uint32_t next = tile + nwarps;
        uint32_t off_next = 0u;
        if (next < tend) {
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
            // compose below reads it) and written back to the translation array
            const float* vel = reinterpret_cast<const float*>(sm.vel[cur]);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t li = k * 32 + lane;
                if ((moving >> li) & 1ull) {
                    float4 a0 = rows[li * s01];
                    integrate_ticks(a0, vel[li * 3], vel[li * 3 + 1], vel[li * 3 + 2], ep.dt, ep.integrate);
                    rows[li * s01] = a0;
                    ep.rows.t[(size_t)(base + li) * ep.rows.st] = a0;
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
                if (keys_given) {  // hierarchy frames always carry the keys pass A wrote
                    sk = (uint64_t) reinterpret_cast<const KeyT*>(sm.keys[cur])[li];
                } else {
                    const float4 a0 = rows[li * s01], a1 = rows[off1 + li * s01];
                    float z;
                    if (is_planar(a1)) {
                        z = planar_z(a0, a1);
                    } else {
                        Row r;
                        r.a0 = a0;
                        r.a1 = a1;
                        r.b0 = rows[offb + li * sb];
                        r.b1 = rows[offb + li * sb + 1];
                        z = local_z(r);
                    }
                    const uint32_t tl = __float_as_uint(a1.w);
                    const uint64_t key = sort_key(tl >> 16, tl & 0xFFFFu, z);
                    if (FUSED) {
                        acc.k_or |= key;
                        acc.k_and &= key;
                        acc.count += 1;
                    }
                    if (FUSED) {
                        sk = pext_runs(key, ep.plan.pext);
                    } else {
                        bool miss = false;  // pass A of the same frame reports misses
                        sk = plan_short_key(ep.plan, s_zdict, key, &miss);
                    }
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
                    r.a0 = rows[li * s01];
                    r.a1 = rows[off1 + li * s01];
                    r.b0 = rows[offb + li * sb];
                    r.b1 = rows[offb + li * sb + 1];
                    w = compose_local(r);
                    sw = r.b1.y;
                    shh = r.b1.z;
                    tl = __float_as_uint(r.a1.w);
                }
                make_outputs_or_row<ROWS>(ep, w, sw, shh, tl, inst, vtx);
                ep.order[dst] = base + li + ep.global_first;
                if (ep.texlayer != nullptr) ep.texlayer[dst] = tl;
            }
            if (!ROWS && !ep.staged) {
                if (live) {
                    float4* ip = ep.instances + (size_t)dst * 5;
                    st_cs_f4(ip, inst.c0);
                    st_cs_f4(ip + 1, inst.c1);
                    st_cs_f4(ip + 2, inst.c2);
                    st_cs_f4(ip + 3, inst.c3);
                    st_cs_f4(ip + 4, inst.s);
                    if (ep.vertices != nullptr) {
                        float4* vp = ep.vertices + (size_t)dst * 4;
                        st_cs_f4(vp, vtx[0]);
                        st_cs_f4(vp + 1, vtx[1]);
                        st_cs_f4(vp + 2, vtx[2]);
                        st_cs_f4(vp + 3, vtx[3]);
                    }
                }
            } else {
                staged_store<!ROWS>(sm.stage, lane, live, min(32u, cnt - s0), dst, inst, vtx, ep.instances, ep.vertices);
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

// ---- sorted frames whose draw order is cached: gather, compose, write ------------------------------------
//
// When the key pass finds every entity's key unchanged, the draw order of the previous frame (its `order`
// output, kept in an internal buffer) is this frame's draw order. The emit then needs no digit, no rank and no
// tile offsets: a warp takes 64 consecutive OUTPUT positions, the entity ids of tile k+2 arrive by TMA bulk
// copy, the rows of tile k+1 are gathered by id with cp.async (128-byte-swizzled), tile k is composed and
// streamed out as whole lines. No block barrier; bounded by the random 64-byte row reads alone.
struct OrderedWarpShared {
    float4 rows[2][kWarpTile * 4];
    float4 stage[32 * kStageF4];
    uint32_t idx[3][kWarpTile];
    unsigned long long kbar[3];
};
struct OrderedShared {
    OrderedWarpShared w[kWarps];
};

//
// VALIDATE: the whole steady-state frame of a flat sorted scene in this one kernel (no key pass at all). Every
// gathered row's key is recomputed and compared with the kept short key of its position (keys_in); the key
// statistics, dictionary misses and a "changed" flag go to `stats`, with which the host confirms after the
// frame that the kept plan, sort and draw order still hold — or redoes the frame the long way.
template <typename KeyT, bool VALIDATE>
__global__ void __launch_bounds__(kThreads, 2) emit_ordered_kernel(const EmitParams<KeyT> ep) {
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    if (ep.run_when == 2u && *reinterpret_cast<const volatile unsigned long long*>(ep.changed) != 0ull) return;
    __shared__ uint32_t s_zdict[VALIDATE ? kMaxZDict : 1];
    if (VALIDATE) {
        for (uint32_t k = threadIdx.x; k < (uint32_t)kMaxZDict; k += kThreads) s_zdict[k] = ep.plan.zdict[k];
        __syncthreads();  // the only block barrier, before the warps go their own ways
    }
    PrepAcc acc;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    OrderedWarpShared& sm = reinterpret_cast<OrderedShared*>(emit_smem_raw)->w[warp];
    const bool from_world = ep.world != nullptr;
    const uint32_t count = ep.pass.count_ptr ? min(ep.pass.n, (uint32_t)*ep.pass.count_ptr) : ep.pass.n;
    const uint32_t ntiles = (count + kWarpTile - 1) / kWarpTile;
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t first = blockIdx.x * kWarps + warp;

    auto tile_cnt = [&](uint32_t t) -> uint32_t { return t < ntiles ? min(kWarpTile, count - t * kWarpTile) : 0u; };
    auto fetch_ids = [&](uint32_t t, uint32_t ring) {  // lane 0: the cached draw order of tile t
        const uint32_t cnt = tile_cnt(t);
        if (cnt == 0) return;
        const uint32_t bytes = (cnt * 4u + 15u) & ~15u;
        mbar_expect_tx(&sm.kbar[ring], bytes);
        bulk_g2s(sm.idx[ring], ep.idx_in + (size_t)t * kWarpTile, bytes, &sm.kbar[ring]);
    };
    auto gather = [&](uint32_t t, uint32_t buf, uint32_t ring) {  // all lanes: two rows each
        const uint32_t cnt = tile_cnt(t);
        float4* d = sm.rows[buf];
        const uint32_t c = lane & 3u;  // one 64-byte request per row: a quad of lanes takes its four chunks
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t li = k * 8 + (lane >> 2);
            if (li < cnt) cp_async16(d + row_slot(li, c), row_chunk(ep, from_world, sm.idx[ring][li] - ep.global_first, c));
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
        fetch_ids(first, 0);
        fetch_ids(first + nwarps, 1);
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
        if (lane == 0) fetch_ids(next + nwarps, ring2);  // slot ring2 was consumed two iterations ago
        if (tile_cnt(next)) {
            mbar_wait(&sm.kbar[ring1], (kpar >> ring1) & 1u);
            kpar ^= 1u << ring1;
        }
        gather(next, cur ^ 1u, ring1);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
        __syncwarp();  // every lane's gathers of this tile have landed
        const float4* rows = sm.rows[cur];
        const uint32_t base = tile * kWarpTile, cnt = tile_cnt(tile);
        KeyT kept[2] = {(KeyT)0, (KeyT)0};
        if (VALIDATE) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if ((uint32_t)k * 32u + lane < cnt) kept[k] = __ldg(ep.keys_in + base + k * 32 + lane);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t li = k * 32 + lane;
            const bool live = li < cnt;
            const uint32_t dst = base + li;
            InstanceOut inst;
            float4 vtx[4];
            if (live) {
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
                        if ((KeyT)plan_short_key(ep.plan, s_zdict, key, &miss) != kept[k]) acc.changed = true;
                        if (miss) acc.miss = true;
                    }
                }
                make_outputs(ep, w, sw, shh, tl, inst, vtx);
                ep.order[dst] = sm.idx[ring][li];
            }
            const uint32_t nlive = cnt > (uint32_t)k * 32u ? min(32u, cnt - k * 32u) : 0u;
            staged_store(sm.stage, lane, live, nlive, dst, inst, vtx, ep.instances, ep.vertices);
        }
        __syncwarp();  // the row buffer and the id slot are free for the next iterations
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

// ---- static scenes: rows kept in draw order ---------------------------------------------------------------
//
// A scene nothing has touched for two frames (no upload, no tick) gets a copy of its rows in draw order
// (gather_rows_kernel, once). From then on a frame — e.g. under a moving camera — streams: a warp takes 64
// consecutive output positions, TMA bulk copies bring the next tile's 4 KB of rows and its entity ids, the
// current tile is composed and written out in whole lines. Sequential reads instead of one random 64-byte row
// per entity (which costs about 128 bytes of DRAM traffic each).
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(RowView rows, const WorldRow* world, const uint32_t* __restrict__ idx,
                                                                uint32_t global_first, uint32_t count, float4* __restrict__ out,
                                                                const float* __restrict__ vel, const uint64_t* mask_v,
                                                                float* __restrict__ vel_out, const uint32_t* __restrict__ parent = nullptr,
                                                                uint32_t* __restrict__ par_out = nullptr, uint32_t split = 0u) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 16-byte chunk: a row's four are adjacent lanes
    const uint32_t pos = t >> 2, ch = t & 3u;
    if (pos >= count) return;
    const uint32_t e = __ldg(idx + pos) - global_first;
    if (vel_out != nullptr && ch < 3u)  // velocities in draw order; zero where the entity has no Velocity
        vel_out[(size_t)pos * 3 + ch] = (vel != nullptr && mask_bit(mask_v, e)) ? __ldg(vel + (size_t)e * 3 + ch) : 0.0f;
    if (par_out != nullptr && ch == 3u) par_out[pos] = __ldg(parent + e);
    if (split != 0u) {  // local rows: dense A0 block, then A1 B0 B1 per position
        const float4 v = ch == 0 ? rows.t[(size_t)e * rows.st] : ch == 1 ? rows.a[(size_t)e * rows.st] : rows.b[(size_t)e * rows.sb + (ch & 1u)];
        out[ch == 0 ? (size_t)pos : (size_t)split + (size_t)pos * 3 + (ch - 1u)] = v;
        return;
    }
    float4 v;
    if (world != nullptr) v = reinterpret_cast<const float4*>(world + e)[ch];
    else v = ch == 0 ? rows.t[(size_t)e * rows.st] : ch == 1 ? rows.a[(size_t)e * rows.st] : rows.b[(size_t)e * rows.sb + (ch & 1u)];
    out[(size_t)pos * 4 + ch] = v;
}

// Rows of tile-local item li of a draw-ordered tile: plain 64-byte rows as the TMA bulk copy lands them.
struct SeqWarpShared {
    float4 rows[2][kWarpTile * 4];
    float4 stage[32 * kStageF4];
    float vel[2][kWarpTile * 3];
    unsigned long long keys[2][kWarpTile];  // kept short keys (KeyT)
    uint32_t idx[2][kWarpTile];
    uint32_t par[2][kWarpTile];             // parent ids (scene graphs)
    unsigned long long kbar[2];
};
struct SeqShared {
    SeqWarpShared w[kWarps];
};

// VALIDATE: the steady-state frame of a flat sorted scene whose rows LIVE in draw order (the copy is the working
// storage): the velocity system runs on the streamed rows (ep.integrate ticks; translation chunk written back in
// place, sequentially), every key is recomputed and compared with the kept one, statistics go to `stats` for the host
// to confirm — all of it on sequential reads. Entity-order rows are brought up to date lazily (sync_rows_kernel).
template <typename KeyT, bool VALIDATE, bool GRAPH = false, bool ROWS = false>
__global__ void __launch_bounds__(kThreads, 2) emit_sequential_kernel(const EmitParams<KeyT> ep) {
    extern __shared__ __align__(128) unsigned char emit_smem_raw[];
    __shared__ uint32_t s_zdict[VALIDATE ? kMaxZDict : 1];
    if (VALIDATE) {
        for (uint32_t k = threadIdx.x; k < (uint32_t)kMaxZDict; k += kThreads) s_zdict[k] = ep.plan.zdict[k];
        __syncthreads();  // the only block barrier, before the warps go their own ways
    }
    PrepAcc acc;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SeqWarpShared& sm = reinterpret_cast<SeqShared*>(emit_smem_raw)->w[warp];
    constexpr bool graph = VALIDATE && GRAPH;  // scene graph on the working storage (see EmitParams::seq_par)
    const bool from_world = ep.world != nullptr && !graph;
    const bool do_int = VALIDATE && ep.integrate != 0 && ep.seq_vel != nullptr && !from_world;
    const bool split = ep.seq_split != 0;
    const uint32_t count = ep.pass.n;
    const uint32_t all_tiles = (count + kWarpTile - 1) / kWarpTile;
    const uint32_t ntiles = ep.tile_end != 0 ? min(ep.tile_end, all_tiles) : all_tiles;
    const uint32_t nwarps = gridDim.x * kWarps;
    // shared-memory slot of chunk c of tile item li
    auto slot = [&](uint32_t li, uint32_t c) -> uint32_t { return split ? (c == 0 ? li : kWarpTile + li * 3u + (c - 1u)) : li * 4u + c; };
    auto fetch = [&](uint32_t t, uint32_t buf) {  // lane 0
        const uint32_t cnt = min(kWarpTile, count - t * kWarpTile);
        const uint32_t ibytes = (cnt * 4u + 15u) & ~15u;
        const uint32_t kbytes = VALIDATE ? ((cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u) : 0u;
        const uint32_t vbytes = do_int ? ((cnt * 12u + 15u) & ~15u) : 0u;
        mbar_expect_tx(&sm.kbar[buf], cnt * 64u + ibytes + kbytes + vbytes + (graph ? ibytes : 0u));
        if (split) {
            bulk_g2s(sm.rows[buf], ep.seq_rows + (size_t)t * kWarpTile, cnt * 16u, &sm.kbar[buf]);
            bulk_g2s(sm.rows[buf] + kWarpTile, ep.seq_rows + ep.seq_split + (size_t)t * kWarpTile * 3, cnt * 48u, &sm.kbar[buf]);
        } else {
            bulk_g2s(sm.rows[buf], ep.seq_rows + (size_t)t * kWarpTile * 4, cnt * 64u, &sm.kbar[buf]);
        }
        bulk_g2s(sm.idx[buf], ep.idx_in + (size_t)t * kWarpTile, ibytes, &sm.kbar[buf]);
        if (graph) bulk_g2s(sm.par[buf], ep.seq_par + (size_t)t * kWarpTile, ibytes, &sm.kbar[buf]);
        if (VALIDATE) bulk_g2s(sm.keys[buf], ep.keys_in + (size_t)t * kWarpTile, kbytes, &sm.kbar[buf]);
        if (do_int) bulk_g2s(sm.vel[buf], ep.seq_vel + (size_t)t * kWarpTile * 3, vbytes, &sm.kbar[buf]);
    };
    uint32_t tile = ep.tile_first + blockIdx.x * kWarps + warp;
    if (lane == 0) {
        mbar_init(&sm.kbar[0], 1);
        mbar_init(&sm.kbar[1], 1);
        if (tile < ntiles) fetch(tile, 0);
    }
    __syncwarp();
    uint32_t cur = 0, par0 = 0, par1 = 0;
    for (; tile < ntiles; tile += nwarps) {
        const uint32_t next = tile + nwarps;
        if (lane == 0 && next < ntiles) fetch(next, cur ^ 1u);
        mbar_wait(&sm.kbar[cur], cur ? par1 : par0);
        if (cur) par1 ^= 1u; else par0 ^= 1u;
        float4* rows = sm.rows[cur];
        const uint32_t base = tile * kWarpTile, cnt = min(kWarpTile, count - base);
        // scene graphs: the world rows this lane's two positions need (a leaf's parent, or an inner node's own) are
        // requested before anything else of the tile is consumed
        float4 wr[2][4];
        bool leaf[2] = {false, false};
        if (graph) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t li = k * 32 + lane;
                if (li < cnt) {
                    const uint32_t e = sm.idx[cur][li] - ep.global_first;
                    leaf[k] = e >= ep.leaf_first;
                    const float4* p = reinterpret_cast<const float4*>(ep.world + (leaf[k] ? sm.par[cur][li] - ep.global_first : e));
                    wr[k][0] = __ldg(p);
                    wr[k][1] = __ldg(p + 1);
                    wr[k][2] = __ldg(p + 2);
                    if (!leaf[k]) wr[k][3] = __ldg(p + 3);
                }
            }
        }
        if (do_int) {
            const float* vel = sm.vel[cur];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint32_t li = k * 32 + lane;
                if (li < cnt && (!graph || leaf[k])) {
                    const float vx = vel[li * 3], vy = vel[li * 3 + 1], vz = vel[li * 3 + 2];
                    if (vx != 0.0f || vy != 0.0f || vz != 0.0f) {  // no Velocity (or a zero one): the value stays
                        float4 a0 = rows[slot(li, 0)];
                        integrate_ticks(a0, vx, vy, vz, ep.dt, ep.integrate);
                        rows[slot(li, 0)] = a0;
                        ep.seq_rows_rw[split ? (size_t)(base + li) : (size_t)(base + li) * 4] = a0;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t li = k * 32 + lane;
            const bool live = li < cnt;
            uint32_t dst = base + li;
            // list frames: the shard-local position becomes the position in the global draw order (the kept short key
            // names the bucket; should the key have changed, the whole frame is redone anyway)
            if (VALIDATE && ep.remap != nullptr && live) dst += __ldg(ep.remap + (uint32_t) reinterpret_cast<const KeyT*>(sm.keys[cur])[li]);
            InstanceOut inst;
            float4 vtx[4];
            if (live) {
                Affine w;
                float sw, shh;
                uint32_t tl;
                if (graph && !leaf[k]) {  // an inner node: the level pass wrote (and key-checked) its world row
                    w.r0 = wr[k][0];
                    w.r1 = wr[k][1];
                    w.r2 = wr[k][2];
                    sw = wr[k][3].x;
                    shh = wr[k][3].y;
                    tl = __float_as_uint(wr[k][3].z);
                } else if (from_world) {
                    w.r0 = rows[li * 4];
                    w.r1 = rows[li * 4 + 1];
                    w.r2 = rows[li * 4 + 2];
                    const float4 sv = rows[li * 4 + 3];
                    sw = sv.x;
                    shh = sv.y;
                    tl = __float_as_uint(sv.z);
                } else {
                    Row r;
                    r.a0 = rows[slot(li, 0)];
                    r.a1 = rows[slot(li, 1)];
                    r.b0 = rows[slot(li, 2)];
                    r.b1 = rows[slot(li, 3)];
                    w = compose_local(r);
                    sw = r.b1.y;
                    shh = r.b1.z;
                    tl = __float_as_uint(r.a1.w);
                    if (graph) {  // a leaf: world = world[parent] * local
                        Affine P;
                        P.r0 = wr[k][0];
                        P.r1 = wr[k][1];
                        P.r2 = wr[k][2];
                        w = mul_affine(P, w);
                    }
                    if (VALIDATE) {
                        const float z = graph ? w.r2.w : (is_planar(r.a1) ? planar_z(r.a0, r.a1) : local_z(r));
                        const uint64_t key = sort_key(tl >> 16, tl & 0xFFFFu, z);
                        acc.k_or |= key;
                        acc.k_and &= key;
                        acc.count += 1;
                        bool miss = false;
                        if ((KeyT)plan_short_key(ep.plan, s_zdict, key, &miss) != reinterpret_cast<const KeyT*>(sm.keys[cur])[li]) acc.changed = true;
                        if (miss) acc.miss = true;
                    }
                }
                make_outputs_or_row<ROWS>(ep, w, sw, shh, tl, inst, vtx);
                ep.order[dst] = sm.idx[cur][li];
            }
            const uint32_t nlive = cnt > (uint32_t)k * 32u ? min(32u, cnt - k * 32u) : 0u;
            staged_store<!ROWS>(sm.stage, lane, live, nlive, dst, inst, vtx, ep.instances, ep.vertices);
        }
        __syncwarp();  // the buffers are free for the fetch of the tile after next
        cur ^= 1u;
    }
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

// The translation chunks of the draw-ordered working storage back into the entity-order rows (lazily: before anything
// reads or replaces those rows).
// `first`: only entities with a local id >= first own their copy (scene graphs: the leaves).
__global__ void __launch_bounds__(kThreads) sync_rows_kernel(const float4* __restrict__ seq_rows, const uint32_t* __restrict__ idx,
                                                              uint32_t global_first, uint32_t count, RowView rows, uint32_t first,
                                                              uint32_t stride) {
    const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= count) return;
    const uint32_t e = __ldg(idx + pos) - global_first;
    if (e < first) return;
    rows.t[(size_t)e * rows.st] = __ldcs(seq_rows + (size_t)pos * stride);  // stride 1: the copy's dense translation block
}

// ---- list frames: completion per chunk of the draw order ------------------------------
//
// A shard's steady-state list frame is emitted as several launches over consecutive ranges of ITS draw order; local
// draw order is monotone in the global one, so after launch k everything this rank owns below the global position of its
// next element is in place. One thread says so to the presenting rank: flag = frame sequence number << 32 | that position.
// The kernel boundary orders the flag after the (peer) stores of the launches before it.
template <typename KeyT>
__global__ void list_signal_kernel(unsigned long long* flag, uint32_t seq, const KeyT* __restrict__ keys,
                                   const uint32_t* __restrict__ remap, uint32_t next_pos, uint32_t count, uint32_t total) {
    const uint32_t g = next_pos >= count ? total : next_pos + remap[(uint32_t)keys[next_pos]];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(flag) = ((unsigned long long)seq << 32) | g;
}
// Plans of one pass emit in ENTITY order (emit_small_kernel): after the tiles before `next_tile`, this rank's part of
// bucket b is complete up to the position where that tile's bucket-b elements start — the cached per-tile offsets say
// where. One flag per bucket (thread b).
__global__ void list_signal_buckets_kernel(unsigned long long* flags, uint32_t seq, const uint32_t* __restrict__ tile_off, uint32_t ntiles,
                                           uint32_t next_tile, const uint32_t* __restrict__ remap, const uint32_t* __restrict__ hist) {
    const uint32_t b = threadIdx.x;
    uint32_t g = 0;
    if (ntiles != 0) g = remap[b] + (next_tile < ntiles ? tile_off[(size_t)b * ntiles + next_tile] : tile_off[(size_t)b * ntiles] + hist[b]);
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(flags + b) = ((unsigned long long)seq << 32) | g;
}

// State of the presenting rank's expansion (device memory, kListStateWords words): [0] a rank stayed silent,
// [1 ...] the range table vertices_from_instances_kernel reads {n, first_0, end_0, ...}, [kListDone ...] where each
// (rank, bucket) stream had come to before this chunk.
constexpr uint32_t kListSlots = 8;      // flags per (rank, chunk): buckets of a one-pass plan (sorted plans use slot 0)
constexpr uint32_t kListMaxStreams = 64;  // ranks x buckets the range table holds
constexpr uint32_t kListTable = 1, kListDone = 2 + 2 * kListMaxStreams + 1, kListStateWords = kListDone + kListMaxStreams + 1;

// The presenting rank, before regenerating the vertices of chunk k: wait until every rank has signalled chunk k of frame
// `seq` (flags[(r * nchunks + k) * kListSlots + b], b < nslots); the table then lists, per (rank, slot) stream, what has
// arrived since the previous chunk. Sorted plans (nslots == 1): every rank's stream is the whole list in order, so the
// single range is [previous end, the lowest position some rank has not passed yet). Plans of one pass: stream (r, b) is
// rank r's segment of bucket b, which starts at starts[r * nslots + b]. state[0] is set (and the table left empty) if a
// rank stays silent for `patience` clock cycles, so a failed peer cannot hang the device.
__global__ void list_wait_kernel(const unsigned long long* flags, uint32_t nranks, uint32_t nchunks, uint32_t k, uint32_t seq,
                                 uint32_t nslots, const uint32_t* __restrict__ starts, uint32_t* state, long long patience) {
    uint32_t* table = state + kListTable;
    uint32_t* done = state + kListDone;
    const long long t0 = clock64();
    table[0] = 0;
    uint32_t lowest = 0xFFFFFFFFu;
    for (uint32_t r = 0; r < nranks; ++r) {
        for (uint32_t b = 0; b < nslots; ++b) {
            const volatile unsigned long long* f = flags + ((size_t)r * nchunks + k) * kListSlots + b;
            unsigned long long v = *f;
            while ((uint32_t)(v >> 32) != seq) {
                if (clock64() - t0 > patience) {
                    table[0] = 0;
                    state[0] = 1u;
                    return;
                }
                __nanosleep(200);
                v = *f;
            }
            if (starts == nullptr) {
                lowest = min(lowest, (uint32_t)v);
            } else {
                const uint32_t j = r * nslots + b;
                const uint32_t lo = k == 0 ? starts[j] : done[j], hi = max((uint32_t)v, lo);
                table[1 + 2 * j] = lo;
                table[2 + 2 * j] = hi;
                done[j] = hi;
            }
        }
    }
    __threadfence_system();
    if (starts == nullptr) {
        const uint32_t lo = k == 0 ? 0u : done[0], hi = max(lowest, lo);
        table[1] = lo;
        table[2] = hi;
        done[0] = hi;
        table[0] = 1;
    } else {
        table[0] = nranks * nslots;
    }
}

// Where every (rank, bucket) stream of a one-pass plan starts in the global list, from the ranks' histograms
// (hists[r * nbins + b], the all-gather of the placed frame): buckets in key order, ranks in order inside a bucket.
__global__ void list_starts_kernel(const uint32_t* __restrict__ hists, uint32_t nranks, uint32_t nbins, uint32_t* starts) {
    uint32_t pos = 0;
    for (uint32_t b = 0; b < nbins; ++b)
        for (uint32_t r = 0; r < nranks; ++r) {
            starts[r * nbins + b] = pos;
            pos += hists[(size_t)r * nbins + b];
        }
}

// ---- quad vertices from instances ------------------------------------------------------------------
//
// The four quad corners are a pure function of an instance (column-major model + size), computed
// with the same individually rounded operations as in the emit kernels, so regenerating them on the
// presenting GPU gives bit-identical vertices while only instances cross NVLink. One warp per 32
// instances: coalesced 16-byte loads of the 2560-byte instance block into shared memory, compute,
// XOR-swizzled staging, coalesced 16-byte stores.
// `ranges` (optional, device memory): {n, first_0, end_0, first_1, end_1, ...} read at run time instead of the two arguments
// — a list frame's presenter learns on the device how far every rank's stores have come (list_wait_kernel): one range of
// the list for sorted plans, one per (rank, bucket) for plans of one pass.
__global__ void __launch_bounds__(kThreads) vertices_from_instances_kernel(const float4* __restrict__ instances,
                                                                            uint32_t first, uint32_t count,
                                                                            float4* __restrict__ vertices,
                                                                            const uint32_t* __restrict__ ranges) {
    __shared__ float4 s_in[kWarps][32 * 5];
    __shared__ float4 s_out[kWarps][32 * 4];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t nranges = ranges != nullptr ? ranges[0] : 1u;
    for (uint32_t blk = blockIdx.x * kWarps + warp;; blk += nwarps) {
        // 32-instance block number blk of the ranges laid end to end
        uint32_t base = 0, m = 0;
        if (ranges == nullptr) {
            if (blk * 32u < count) {
                base = first + blk * 32u;
                m = min(32u, count - blk * 32u);
            }
        } else {
            uint32_t skipped = 0;
            for (uint32_t j = 0; j < nranges; ++j) {
                const uint32_t lo = ranges[1 + 2 * j], hi = ranges[2 + 2 * j];
                const uint32_t nb = hi > lo ? (hi - lo + 31u) / 32u : 0u;
                if (blk - skipped < nb) {
                    base = lo + (blk - skipped) * 32u;
                    m = min(32u, hi - base);
                    break;
                }
                skipped += nb;
            }
        }
        if (m == 0) break;
        const float4* src = instances + (size_t)base * 5;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const uint32_t q = j * 32 + lane;
            if (q < m * 5) s_in[warp][q] = __ldcs(src + q);
        }
        __syncwarp();
        const uint32_t sw = (lane >> 1) & 3u;
        if (lane < m) {
            const float4 c0 = s_in[warp][lane * 5], c1 = s_in[warp][lane * 5 + 1], c3 = s_in[warp][lane * 5 + 3];
            const float4 sz = s_in[warp][lane * 5 + 4];
            const float cx[4] = {0.0f, sz.x, sz.x, 0.0f};
            const float cy[4] = {0.0f, 0.0f, sz.y, sz.y};
            const uint32_t uv[4] = {0x00000000u, 0x00000001u, 0x00010001u, 0x00010000u};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                s_out[warp][lane * 4 + (k ^ sw)] =
                    make_float4(corner_coord(c0.x, c1.x, c3.x, cx[k], cy[k]), corner_coord(c0.y, c1.y, c3.y, cx[k], cy[k]),
                                corner_coord(c0.z, c1.z, c3.z, cx[k], cy[k]), __uint_as_float(uv[k]));
        }
        __syncwarp();
        float4* dst = vertices + (size_t)base * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = j * 32 + lane;
            const uint32_t item = q >> 2, part = q & 3u;
            if (item < m) st_cs_f4(dst + q, s_out[warp][item * 4 + (part ^ ((item >> 1) & 3u))]);
        }
        __syncwarp();
    }
}

// ---- instances and quad vertices from world rows -------------------------------------------------------
//
// The presenting rank of a steady list frame receives WORLD ROWS (EmitParams::world_rows_out) and turns them into
// instances and quad vertices with the same make_outputs the emit kernels use (camera included), so the list is
// bit-identical to one emitted directly. Ranges as in vertices_from_instances_kernel.
struct CameraParams {
    float4 vp[4];
    uint32_t has_vp;
};
__global__ void __launch_bounds__(kThreads) expand_world_rows_kernel(const float4* __restrict__ rows, uint32_t first, uint32_t count,
                                                                      float4* __restrict__ instances, float4* __restrict__ vertices,
                                                                      const uint32_t* __restrict__ ranges, const CameraParams cam) {
    __shared__ float4 s_stage[kWarps][32 * kStageF4];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * kWarps;
    const uint32_t nranges = ranges != nullptr ? ranges[0] : 1u;
    for (uint32_t blk = blockIdx.x * kWarps + warp;; blk += nwarps) {
        uint32_t base = 0, m = 0;
        if (ranges == nullptr) {
            if (blk * 32u < count) {
                base = first + blk * 32u;
                m = min(32u, count - blk * 32u);
            }
        } else {
            uint32_t skipped = 0;
            for (uint32_t j = 0; j < nranges; ++j) {
                const uint32_t lo = ranges[1 + 2 * j], hi = ranges[2 + 2 * j];
                const uint32_t nb = hi > lo ? (hi - lo + 31u) / 32u : 0u;
                if (blk - skipped < nb) {
                    base = lo + (blk - skipped) * 32u;
                    m = min(32u, hi - base);
                    break;
                }
                skipped += nb;
            }
        }
        if (m == 0) break;
        // the block's 2 KB of rows through shared memory: coalesced 16-byte loads, then each lane reads its own row
        float4* st = s_stage[warp];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t q = j * 32 + lane;
            if (q < m * 4) st[q ^ ((q >> 3) & 3u)] = __ldcs(rows + (size_t)base * 4 + q);
        }
        __syncwarp();
        const bool live = lane < m;
        InstanceOut inst;
        float4 vtx[4];
        if (live) {
            const uint32_t q0 = lane * 4, sw4 = (q0 >> 3) & 3u;  // the four chunks of a row share one swizzle term
            Affine w;
            w.r0 = st[(q0 + 0) ^ sw4];
            w.r1 = st[(q0 + 1) ^ sw4];
            w.r2 = st[(q0 + 2) ^ sw4];
            const float4 sv = st[(q0 + 3) ^ sw4];
            make_outputs(cam, w, sv.x, sv.y, __float_as_uint(sv.z), inst, vtx);
        }
        __syncwarp();
        staged_store(st, lane, live, m, base + lane, inst, vtx, instances, vertices);
    }
}

}  // namespace tb
