// What bounds the sorted (gathering) emit: the rate at which a B200 delivers 64-byte rows read in random
// order from an array far larger than L2, next to the same rows read in order. Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_gather tools/microbench_gather.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

template <int ROWS>
__global__ void gather_kernel(const float4* __restrict__ rows, const uint32_t* __restrict__ idx, uint32_t n, float4* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = t; i + (ROWS - 1) * stride < n; i += ROWS * stride) {
        float4 v[ROWS][4];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float4* p = rows + (size_t)idx[i + r * stride] * 4;
#pragma unroll
            for (int c = 0; c < 4; ++c) v[r][c] = p[c];
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float4 s = v[r][0];
            s.x += v[r][1].x + v[r][2].y + v[r][3].z;
            out[i + r * stride] = s;
        }
    }
}

int main(int argc, char** argv) {
    if (argc > 1) {  // optional: L2 fetch granularity in bytes (32, 64, 128)
        size_t g = (size_t)atoi(argv[1]);
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"l2_fetch_granularity_requested\": %zu, \"status\": \"%s\", \"now\": %zu}\n", g, cudaGetErrorString(e), got);
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("{\"l2_fetch_granularity_default\": %zu}\n", got);
    }
    const uint32_t n = 1u << 24;  // 1 GiB of 64-byte rows
    float4 *rows, *out;
    uint32_t* idx;
    cudaMalloc(&rows, (size_t)n * 64);
    cudaMalloc(&out, (size_t)n * 16);
    cudaMalloc(&idx, (size_t)n * 4);
    cudaMemset(rows, 0, (size_t)n * 64);
    std::vector<uint32_t> h(n);
    for (uint32_t i = 0; i < n; ++i) h[i] = i;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        if (mode == 1) {
            std::mt19937 rng(1);
            std::shuffle(h.begin(), h.end(), rng);
        }
        cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
        for (int rows_per_thread = 1; rows_per_thread <= 4; rows_per_thread *= 2) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (rows_per_thread == 1) gather_kernel<1><<<148 * 16, 256>>>(rows, idx, n, out);
                if (rows_per_thread == 2) gather_kernel<2><<<148 * 16, 256>>>(rows, idx, n, out);
                if (rows_per_thread == 4) gather_kernel<4><<<148 * 8, 256>>>(rows, idx, n, out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = std::min(best, ms);
            }
            // bytes: 64 gathered + 4 index + 16 written per row
            printf("{\"order\": \"%s\", \"rows_in_flight_per_thread\": %d, \"ms\": %.4f, \"rows_per_s\": %.3e, \"gathered_GBps\": %.1f, \"total_GBps\": %.1f}\n",
                   mode ? "random" : "sequential", rows_per_thread, best, n / (best * 1e-3), n * 64.0 / (best * 1e-3) / 1e9,
                   n * 84.0 / (best * 1e-3) / 1e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
