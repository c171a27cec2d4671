// Stable radix placement without inter-CTA waiting: digit histograms -> bucket bases, per-tile
// digit counts -> per-tile offsets (reduce-then-scan), tile ranking, one LSD scatter pass, and the
// batch table from the (layer, texture) histogram (or, for very wide scenes, run-head detection).
#pragma once
#include "common.cuh"

namespace tb {

// ---- histogram -> bucket bases; (layer, texture) histogram -> batch table -----------------------

// First output position of every (digit, tile) of one pass from its per-tile digit counts
// (reduce-then-scan radix placement: no inter-CTA waiting anywhere). counts and tile_off are
// digit-major [R][ntiles]; hist is the pass's raw 256-bin digit histogram. Two launches over a
// grid of (chunks of 1024 tiles, R): chunk sums, then every block adds the sums of the chunks
// before its own (a few hundred L2-resident words) and scans its chunk.
// `skip` (all sort-pass kernels): non-null when the sorted arrays and tile offsets of the previous frame
// are still in place; *skip == 0 then means the key pass found every entity's short key unchanged, so
// this frame's sort would reproduce them bit for bit and the kernel returns at once.
// One thread per block reads the flag (every thread of a 30 000-block grid hammering one L2 line costs more than
// the work being skipped); must be called by all threads of the block.
__device__ __forceinline__ bool sort_is_cached(const unsigned long long* skip) {
    if (skip == nullptr) return false;
    __shared__ uint32_t s_cached;
    if (threadIdx.x == 0) s_cached = *reinterpret_cast<const volatile unsigned long long*>(skip) == 0ull ? 1u : 0u;
    __syncthreads();
    return s_cached != 0u;
}

__global__ void __launch_bounds__(1024) tile_chunk_sums_kernel(const uint32_t* __restrict__ counts, uint32_t ntiles,
                                                                uint32_t nchunks, uint32_t* csum,
                                                                const unsigned long long* skip) {
    if (sort_is_cached(skip)) return;
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
                                                          uint32_t nchunks, uint32_t* tile_off, const unsigned long long* skip) {
    if (sort_is_cached(skip)) return;
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
// counts is digit-major [R][ntiles]. Blocks walk the tiles (t = blockIdx.x + k * gridDim.x); four independent loads in flight per thread.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads) tile_hist_kernel(const KeyT* __restrict__ keys, uint32_t n,
                                                              const unsigned long long* count_ptr,
                                                              const uint64_t* mask_t, const uint64_t* mask_s, uint32_t first,
                                                              uint32_t tile_elems, uint32_t shift, uint32_t bits,
                                                              uint32_t ntiles, uint32_t* counts, const unsigned long long* skip,
                                                              const uint32_t* chunk_cnt = nullptr, uint32_t chunk_shift = 0) {
    if (sort_is_cached(skip)) return;
    __shared__ uint32_t h[256];
    const uint32_t R = 1u << bits, mask = R - 1u;
    uint32_t n_eff = count_ptr ? min(n, (uint32_t)*count_ptr) : n;
    const uint32_t n_all = n_eff;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {  // persistent blocks: a skipped launch costs ~1000 exits
        for (uint32_t d = threadIdx.x; d < R; d += kThreads) h[d] = 0;
        __syncthreads();
        const uint32_t base = t * tile_elems;
        if (chunk_cnt != nullptr) {  // chunk-padded list: positions of this tile below the chunk's count hold data
            const uint32_t c0 = (base >> chunk_shift) << chunk_shift;
            n_eff = min(n_all, c0 + __ldg(chunk_cnt + (base >> chunk_shift)));
        }
        for (uint32_t k0 = threadIdx.x; k0 < tile_elems; k0 += 4 * kThreads) {
            KeyT key[4];
            bool valid[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t k = k0 + j * kThreads, pos = base + k;
                valid[j] = k < tile_elems && pos < n_eff;
                key[j] = valid[j] ? keys[pos] : (KeyT)0;
                if (valid[j] && first) valid[j] = mask_bit(mask_t, pos) && mask_bit(mask_s, pos);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (valid[j]) atomicAdd(&h[(uint32_t)(key[j] >> shift) & mask], 1u);
        }
        __syncthreads();
        for (uint32_t d = threadIdx.x; d < R; d += kThreads) counts[(size_t)d * ntiles + t] = h[d];
        __syncthreads();  // h is reused by the next tile
    }
}

struct BatchOut {
    uint32_t layer, texture, first_instance, count;
};

// Every non-empty (layer, texture) bin is one batch: first = exclusive prefix, count = bin.
// `bins` has 1 << ltbits entries (it may alias the raw digit histogram of pass 0). Single block.
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
                // with a z dictionary the pext runs cover the (layer, texture) half only, from bit 0
                const uint64_t key = plan.constant | pdep_runs(plan.zdict_n ? (uint64_t)b : ((uint64_t)b << plan.zbits), plan.pext);
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

// ---- stable tile ranking (shared by scatter and emit) ------------------------------------------------

struct RankShared {
    uint32_t wcnt[kWarps * 256];  // [warp][R] per-warp digit counters, then per-warp exclusive prefixes
    uint32_t dstart[256];         // first tile slot of each digit
    uint32_t dbase[256];          // destination of digit d's slot s is dbase[d] + s
    uint32_t wsum[kWarps];
    uint32_t tile_count;
};

struct PassParams {
    uint32_t n;              // positions to visit
    const unsigned long long* count_ptr;  // when non-null only positions < *count_ptr hold elements
                                          // (passes after the first: the renderable count lives on the device)
    uint32_t shift, bits;    // this pass's digit of the short key
    const uint32_t* tile_off;  // [1 << bits][ntiles]: first output position of each (digit, tile), from the
                               // per-tile digit counts (tile_chunk_sums + tile_scan)
    uint32_t ntiles;           // tiles of this pass (row length of tile_off)
    // chunked sort (chunked.cuh): the pass's input is a list padded per chunk of 1 << chunk_shift entities — chunk c
    // occupies positions [c << chunk_shift, (c << chunk_shift) + chunk_cnt[c]); nullptr: one dense list
    const uint32_t* chunk_cnt;
    uint32_t chunk_shift;
};

// Elements of tile `tile` (tile_elems positions each) that hold data.
__device__ __forceinline__ uint32_t pass_tile_count(const PassParams& pp, uint32_t tile, uint32_t tile_elems, uint32_t n_eff) {
    if (tile >= pp.ntiles) return 0u;
    const uint32_t base = tile * tile_elems;
    if (pp.chunk_cnt != nullptr) {
        const uint32_t local = base & ((1u << pp.chunk_shift) - 1u);
        const uint32_t cnt = __ldg(pp.chunk_cnt + (base >> pp.chunk_shift));
        return cnt > local ? min(tile_elems, cnt - local) : 0u;
    }
    return base < n_eff ? min(tile_elems, n_eff - base) : 0u;
}

// Zeroes the per-warp digit counters of the next rank_tile call; a __syncthreads() must
// separate it from that call.
__device__ __forceinline__ void rank_reset(RankShared& rs, uint32_t bits) {
    const uint32_t R = 1u << bits;
    for (uint32_t k = threadIdx.x; k < kWarps * R; k += kThreads) rs.wcnt[k] = 0;
}

// Ranks ITEMS digits per thread (warp-striped order: item k of lane l in warp w is element
// w*ITEMS*32 + k*32 + l of the tile) into stable tile slots, and leaves dbase[] ready from the
// pass's precomputed tile offsets (no CTA ever waits for another). Invalid items (kInvalidDigit)
// get no slot. Needs rank_reset() + __syncthreads() before, and a __syncthreads() before the next
// rank_reset().
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
    // per digit: exclusive prefix over warps, tile total; the tile offset load overlaps the scan
    uint32_t total = 0, toff = 0;
    if (threadIdx.x < R) {
        toff = __ldg(pp.tile_off + (size_t)threadIdx.x * pp.ntiles + tile);
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = rs.wcnt[w * R + threadIdx.x];
            rs.wcnt[w * R + threadIdx.x] = total;
            total += c;
        }
    }
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
        rs.dbase[threadIdx.x] = toff - dstart;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t d = digit[k];
        if (d != kInvalidDigit) slot[k] += rs.dstart[d] + rs.wcnt[warp * R + d];
    }
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
    const unsigned long long* skip;  // see sort_is_cached
    // last pass of a chunked sort: idx_out receives GLOBAL entity ids (+ id_offset) and dst_out the position of every
    // element in the global draw order: (position inside its chunk's sorted list) + remap[chunk][short key]
    uint32_t* dst_out;
    const uint32_t* remap;
    uint32_t key_bits;
    uint32_t id_offset;
};

constexpr int kScatterItems = 8;
constexpr uint32_t kScatterTile = kThreads * kScatterItems;  // 2048 pairs per tile

// Dynamic shared memory of radix_scatter_kernel: 58 KB with 32-bit keys, three CTAs per SM.
template <typename KeyT>
struct ScatterShared {
    KeyT kin[2][kScatterTile];          // this tile's keys and the prefetched next tile's (TMA bulk copies)
    uint32_t iin[2][kScatterTile];      // ... entity ids
    unsigned long long masks[2][2][kScatterTile / 64 + 2];  // ... presence words (first pass)
    KeyT s_key[kScatterTile];           // the tile in sorted order
    uint32_t s_idx[kScatterTile];
    RankShared rs;
    unsigned long long mbar[2];
};

// One stable LSD pass. Persistent CTAs walk 2048-pair tiles (tile = blockIdx.x + k * gridDim.x); the
// next tile's keys / ids / presence words are prefetched by TMA bulk copies while the current tile is
// ranked, ordered in shared memory and written out in runs of consecutive destinations.
template <typename KeyT>
__global__ void __launch_bounds__(kThreads, 3) radix_scatter_kernel(const ScatterParams<KeyT> sp) {
    constexpr int ITEMS = kScatterItems;
    constexpr uint32_t TILE = kScatterTile;
    extern __shared__ __align__(128) unsigned char scatter_smem_raw[];
    ScatterShared<KeyT>& sm = *reinterpret_cast<ScatterShared<KeyT>*>(scatter_smem_raw);
    RankShared& rs = sm.rs;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (sort_is_cached(sp.skip)) return;
    const bool first = sp.idx_in == nullptr;
    const uint32_t n_eff = sp.pass.count_ptr ? min(sp.pass.n, (uint32_t)*sp.pass.count_ptr) : sp.pass.n;
    const uint32_t ntiles = sp.pass.ntiles;
    const uint32_t mask = (1u << sp.pass.bits) - 1u;
    const uint32_t step = gridDim.x;
    auto tile_count = [&](uint32_t tile) -> uint32_t {
        // the first pass of a chunked sort reads keys in entity order (dense); later passes the chunk-padded list
        if (first) {
            const uint32_t base = tile * TILE;
            return (tile < ntiles && base < n_eff) ? min(TILE, n_eff - base) : 0u;
        }
        return pass_tile_count(sp.pass, tile, TILE, n_eff);
    };
    auto fetch = [&](uint32_t tile, uint32_t buf) {  // thread 0
        const uint32_t base = tile * TILE;
        const uint32_t cnt = tile_count(tile);
        if (cnt == 0) return;
        const uint32_t kbytes = (cnt * (uint32_t)sizeof(KeyT) + 15u) & ~15u;
        const uint32_t ibytes = first ? 0u : ((cnt * 4u + 15u) & ~15u);
        const uint32_t mwords = ((cnt + 127u) >> 7) * 2u;
        const uint32_t mbytes = first ? ((sp.mask_t ? 1u : 0u) + (sp.mask_s ? 1u : 0u)) * mwords * 8u : 0u;
        mbar_expect_tx(&sm.mbar[buf], kbytes + ibytes + mbytes);
        bulk_g2s(sm.kin[buf], sp.keys_in + base, kbytes, &sm.mbar[buf]);
        if (!first) bulk_g2s(sm.iin[buf], sp.idx_in + base, ibytes, &sm.mbar[buf]);
        if (first && sp.mask_t) bulk_g2s(sm.masks[buf][0], sp.mask_t + (base >> 6), mwords * 8u, &sm.mbar[buf]);
        if (first && sp.mask_s) bulk_g2s(sm.masks[buf][1], sp.mask_s + (base >> 6), mwords * 8u, &sm.mbar[buf]);
    };
    uint32_t tile = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        fetch(tile, 0);
    }
    rank_reset(rs, sp.pass.bits);
    __syncthreads();
    uint32_t cur = 0, parity0 = 0, parity1 = 0;
    for (; tile < ntiles; tile += step) {
        if (threadIdx.x == 0) fetch(tile + step, cur ^ 1u);  // buffer freed by the last barrier
        const uint32_t tile_base = tile * TILE;
        const uint32_t tcnt = tile_count(tile);
        if (tcnt) {
            mbar_wait(&sm.mbar[cur], cur ? parity1 : parity0);
            if (cur) parity1 ^= 1u; else parity0 ^= 1u;
        }
        KeyT key[ITEMS];
        uint32_t idx[ITEMS], digit[ITEMS], slot[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t li = warp * (ITEMS * 32) + k * 32 + lane;
            bool valid = li < tcnt;
            if (valid && first) {
                const bool t_ok = sp.mask_t == nullptr || ((sm.masks[cur][0][li >> 6] >> (li & 63u)) & 1ull);
                const bool s_ok = sp.mask_s == nullptr || ((sm.masks[cur][1][li >> 6] >> (li & 63u)) & 1ull);
                valid = t_ok && s_ok;
            }
            key[k] = valid ? sm.kin[cur][li] : (KeyT)0;
            idx[k] = first ? tile_base + li : (valid ? sm.iin[cur][li] : 0u);
            digit[k] = valid ? ((uint32_t)(key[k] >> sp.pass.shift) & mask) : kInvalidDigit;
        }
        rank_tile<ITEMS>(rs, sp.pass, tile, digit, slot);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            if (digit[k] != kInvalidDigit) {
                sm.s_key[slot[k]] = key[k];
                sm.s_idx[slot[k]] = idx[k];
            }
        }
        __syncthreads();
        const uint32_t cnt = rs.tile_count;
        for (uint32_t s = threadIdx.x; s < cnt; s += kThreads) {
            const KeyT kk = sm.s_key[s];
            const uint32_t d = (uint32_t)(kk >> sp.pass.shift) & mask;
            const uint32_t dst = rs.dbase[d] + s;
            sp.keys_out[dst] = kk;
            if (sp.dst_out != nullptr) {
                const uint32_t chunk = dst >> sp.pass.chunk_shift;
                sp.idx_out[dst] = sm.s_idx[s] + sp.id_offset;
                sp.dst_out[dst] = (dst - (chunk << sp.pass.chunk_shift)) + __ldg(sp.remap + ((size_t)chunk << sp.key_bits) + (uint32_t)kk);
            } else {
                sp.idx_out[dst] = sm.s_idx[s];
            }
        }
        rank_reset(rs, sp.pass.bits);
        __syncthreads();  // releases the input buffer, the sorted tile and the rank scratch
        cur ^= 1u;
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

}  // namespace tb
