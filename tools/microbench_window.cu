// Does L2 turn a window-local random gather of 64-byte rows into streaming DRAM traffic? Rows are read through an
// index array that is a random permutation INSIDE consecutive windows of W rows (W * 64 B against the 126 MB L2),
// while 144 bytes per row are streamed out with evict-first stores (what the emit writes). Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_window tools/microbench_window.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

__global__ void gather_emit_kernel(const float4* __restrict__ rows, const uint32_t* __restrict__ idx, uint32_t n, float4* out, int write) {
    // warp per 32 consecutive output positions; blocks walk positions in order so that concurrently running
    // blocks work inside the same window
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w * 32 < n; w += warps) {
        const uint32_t i = w * 32 + lane;
        const float4* p = rows + (size_t)idx[i] * 4;
        float4 a = p[0], b = p[1], c = p[2], d = p[3];
        float4 s = make_float4(a.x + b.y, c.z + d.w, a.y, b.x);
        if (write) {
            float4* o = out + (size_t)i * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) __stcs(o + k, s);
        } else if (s.x == 12345.f) {
            out[i] = s;
        }
    }
}

int main() {
    const uint32_t n = 1u << 24;
    float4 *rows, *out;
    uint32_t* idx;
    cudaMalloc(&rows, (size_t)n * 64);
    cudaMalloc(&out, (size_t)n * 144);
    cudaMalloc(&idx, (size_t)n * 4);
    cudaMemset(rows, 0, (size_t)n * 64);
    std::vector<uint32_t> h(n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    std::mt19937 rng(1);
    const uint32_t windows[] = {1u, 1u << 14, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 22, 1u << 24};
    for (uint32_t W : windows) {
        for (uint32_t i = 0; i < n; ++i) h[i] = i;
        if (W > 1)
            for (uint32_t b = 0; b < n; b += W) std::shuffle(h.begin() + b, h.begin() + b + W, rng);
        cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
        for (int write = 0; write < 2; ++write) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                gather_emit_kernel<<<148 * 8, 256>>>(rows, idx, n, out, write);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = std::min(best, ms);
            }
            printf("{\"window_rows\": %u, \"window_MB\": %.1f, \"writes_144B\": %d, \"ms\": %.4f, \"alg_GBps\": %.0f}\n", W, W * 64.0 / 1e6, write,
                   best, n * (68.0 + (write ? 144.0 : 0.0)) / (best * 1e-3) / 1e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
