// Multi-GPU placement: from every rank's short-key histogram to the shift that turns
// a shard-local sorted position into the position in the global draw order.
#pragma once
#include "common.cuh"

namespace tb {

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

// ---- wide keys: exact global order by co-ranking sorted key arrays ------------------------------------
//
// Short keys wider than TB_SHARD_MAX_KEY_BITS (continuous z) cannot be placed from a histogram. Every rank then sorts
// its shard as usual, the ranks all-gather their sorted FULL 64-bit keys (8 B per element, to every rank), and an
// element's global position is its local position plus, for every other rank, the number of that rank's elements that
// come before it: keys <= k on lower ranks (their ids are smaller: the entity-id tie-break), keys < k on higher ranks.
namespace tb {

// Full sort key of every local draw position: from the world row (scene graphs) or the row.
__global__ void __launch_bounds__(kThreads) draw_keys_kernel(const uint32_t* __restrict__ order, uint32_t global_first, uint32_t count,
                                                              RowView rows, const WorldRow* world, unsigned long long* keys, uint32_t padded) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= padded) return;
    if (p >= count) {
        keys[p] = ~0ull;  // padding up to the longest rank's count
        return;
    }
    const uint32_t e = __ldg(order + p) - global_first;
    if (world != nullptr) {
        const float4* w = reinterpret_cast<const float4*>(world + e);
        keys[p] = world_key(w[2], w[3]);
    } else {
        const float4 a0 = rows.t[(size_t)e * rows.st], a1 = rows.a[(size_t)e * rows.st];
        keys[p] = flat_key(rows, e, a0, a1);
    }
}

__global__ void __launch_bounds__(kThreads) corank_kernel(const unsigned long long* __restrict__ all_keys, const unsigned long long* __restrict__ counts,
                                                           uint32_t stride, uint32_t nranks, uint32_t rank, uint32_t count, uint32_t* dst) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    const unsigned long long k = all_keys[(size_t)rank * stride + p];
    uint32_t pos = p;
    for (uint32_t r = 0; r < nranks; ++r) {
        if (r == rank) continue;
        const unsigned long long* a = all_keys + (size_t)r * stride;
        uint32_t lo = 0, hi = (uint32_t)counts[r];
        if (r < rank) {  // upper bound: elements with key <= k come first
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(a + mid) <= k) lo = mid + 1;
                else hi = mid;
            }
        } else {         // lower bound: only elements with key < k come first
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(a + mid) < k) lo = mid + 1;
                else hi = mid;
            }
        }
        pos += lo;
    }
    dst[p] = pos;
}

}  // namespace tb
