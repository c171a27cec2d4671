// Throughput of the bit-exact libm sinf/cosf restatement (double precision) against CUDA's sincosf and against
// raw DFMA / FFMA issue rates. Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
#include "../tuber_legacy_b200/csrc/libm_sincosf.h"

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a = seed + threadIdx.x * 1e-3f + blockIdx.x * 1e-5f, acc = 0.f;
    double d = a, e = 1.000001;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { float s, c; tbm::sincosf_ref<true>(a, &s, &c); acc += s + c; a += 0.37f; if (a > 6.f) a -= 6.f; }
        if (MODE == 1) { float s, c; sincosf(a, &s, &c); acc += s + c; a += 0.37f; if (a > 6.f) a -= 6.f; }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) d = __fma_rn(d, e, 0.5);
        }
        if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a = __fmaf_rn(a, 1.000001f, 0.5f);
        }
        if (MODE == 4) {  // 4 independent DFMA chains
            double d1 = d + 1, d2 = d + 2, d3 = d + 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) { d = __fma_rn(d, e, 0.5); d1 = __fma_rn(d1, e, 0.5); d2 = __fma_rn(d2, e, 0.5); d3 = __fma_rn(d3, e, 0.5); }
            d += d1 + d2 + d3;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)d + a;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 16 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096;
    const char* names[] = {"libm_sincosf_ref (FP64)", "cuda sincosf", "DFMA x16 dependent", "FFMA x16 dependent", "DFMA 4 chains x4"};
    for (int mode = 0; mode < 5; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.1f);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters, 0.1f);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters, 0.1f);
            if (mode == 3) k<3><<<148 * 8, 256>>>(out, iters, 0.1f);
            if (mode == 4) k<4><<<148 * 8, 256>>>(out, iters, 0.1f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double calls = 148.0 * 8 * 256 * iters * ((mode >= 2) ? 16 : 1);
        printf("{\"what\": \"%s\", \"ms\": %.3f, \"per_second\": %.3e, \"per_clk_per_SM\": %.2f}\n", names[mode], best, calls / (best * 1e-3),
               calls / (best * 1e-3) / 148 / 1.965e9);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
