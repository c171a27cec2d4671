// What one B200 can push into a peer's memory over NVLink: per-lane 16-byte stores (a warp writes 512 contiguous
// bytes per instruction, tiles of 5 KB) against TMA bulk stores (cp.async.bulk shared -> global) of the same tiles,
// and the peer copy engine. Needs 2 GPUs. Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_peer_store tools/microbench_peer_store.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

constexpr int kTile = 5120;  // bytes per warp tile: 64 instances of 80 bytes
constexpr int kWarps = 8;

__global__ void __launch_bounds__(256) store_lanes(float4* __restrict__ dst, uint32_t ntiles) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)lane);
    for (uint32_t t = blockIdx.x * kWarps + warp; t < ntiles; t += gridDim.x * kWarps) {
        float4* d = dst + (size_t)t * (kTile / 16);
#pragma unroll
        for (int k = 0; k < kTile / 512; ++k) d[k * 32 + lane] = v;
    }
}

__global__ void __launch_bounds__(256) store_bulk(float4* __restrict__ dst, uint32_t ntiles, int depth) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4* mine = reinterpret_cast<float4*>(smem) + warp * (kTile / 16);
    for (int k = lane; k < kTile / 16; k += 32) mine[k] = make_float4(1.f, 2.f, 3.f, (float)lane);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    for (uint32_t t = blockIdx.x * kWarps + warp; t < ntiles; t += gridDim.x * kWarps) {
        if (lane == 0) {
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(mine);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)t * (kTile / 16)), "r"(s), "r"(kTile)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        __syncwarp();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 2) {
        printf("{\"error\": \"needs 2 GPUs\"}\n");
        return 0;
    }
    const size_t bytes = (size_t)2 << 30;
    const uint32_t ntiles = (uint32_t)(bytes / kTile);
    float4 *local, *remote, *src;
    cudaSetDevice(0);
    cudaMalloc(&remote, bytes);
    cudaSetDevice(1);
    cudaMalloc(&local, bytes);
    cudaMalloc(&src, bytes);
    int can = 0;
    cudaDeviceCanAccessPeer(&can, 1, 0);
    cudaDeviceEnablePeerAccess(0, 0);
    cudaFuncSetAttribute(store_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, kWarps * kTile);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto timed = [&](const char* what, const char* where, auto fn) {
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            fn();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = std::min(best, ms);
        }
        printf("{\"how\": \"%s\", \"to\": \"%s\", \"ms\": %.3f, \"GBps\": %.1f}\n", what, where, best, bytes / (best * 1e-3) / 1e9);
        fflush(stdout);
    };
    printf("{\"peer_access\": %d}\n", can);
    for (int where = 0; where < 2; ++where) {
        float4* d = where ? remote : local;
        const char* w = where ? "peer" : "local";
        for (int ctas = 1; ctas <= 4; ctas *= 2) {
            char name[64];
            snprintf(name, sizeof name, "lane stores, %d CTAs/SM", ctas);
            timed(name, w, [&] { store_lanes<<<148 * ctas, 256>>>(d, ntiles); });
        }
        timed("bulk stores, 2 CTAs/SM, 1 in flight per warp", w, [&] { store_bulk<<<148 * 2, 256, kWarps * kTile>>>(d, ntiles, 1); });
        timed("bulk stores, 4 CTAs/SM, 1 in flight per warp", w, [&] { store_bulk<<<148 * 4, 256, kWarps * kTile>>>(d, ntiles, 1); });
        timed("cudaMemcpyAsync", w, [&] { cudaMemcpyAsync(d, src, bytes, cudaMemcpyDeviceToDevice, 0); });
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
