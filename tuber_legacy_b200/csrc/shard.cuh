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
