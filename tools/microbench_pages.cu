// Is the random 64-byte row gather bound by address translation? Random rows over footprints from 256 MB to 4 GiB,
// from a cudaMalloc allocation and from a VMM allocation (cuMemCreate + a 512 MB-aligned reservation) that the
// driver is free to back with larger pages. Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_pages tools/microbench_pages.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void gather_kernel(const float4* __restrict__ rows, const uint32_t* __restrict__ idx, uint32_t n, float4* out) {
    // four lanes per row, 16 bytes each (one coalesced 64-byte request per row)
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t c = t & 3;
    float4 acc = make_float4(0, 0, 0, 0);
    for (uint32_t i = t >> 2; i < n; i += stride >> 2) {
        const float4 v = rows[(size_t)idx[i] * 4 + c];
        acc.x += v.x;
        acc.y += v.y;
    }
    if (acc.x == 123.f) out[t] = acc;
}

static uint32_t lcg(uint32_t& s) { return s = s * 1664525u + 1013904223u; }

static void run(const char* what, const float4* rows, size_t footprint, uint32_t* idx, float4* out) {
    const uint32_t nrows = (uint32_t)(footprint / 64);
    const uint32_t n = 1u << 24;  // gathers per launch
    std::vector<uint32_t> h(n);
    uint32_t s = 12345;
    for (uint32_t i = 0; i < n; ++i) h[i] = (uint32_t)(((uint64_t)(lcg(s) >> 4) * nrows) >> 28);
    cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        gather_kernel<<<148 * 8, 256>>>(rows, idx, n, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    printf("{\"alloc\": \"%s\", \"footprint_MB\": %zu, \"ms\": %.4f, \"rows_per_s\": %.3e, \"gathered_GBps\": %.1f}\n", what, footprint >> 20, best,
           n / (best * 1e-3), n * 64.0 / (best * 1e-3) / 1e9);
    fflush(stdout);
}

int main() {
    cudaFree(0);
    uint32_t* idx;
    float4* out;
    cudaMalloc(&idx, (size_t)(1u << 24) * 4);
    cudaMalloc(&out, (size_t)148 * 8 * 256 * 16);
    const size_t maxb = (size_t)4 << 30;
    {
        float4* rows;
        cudaMalloc(&rows, maxb);
        cudaMemset(rows, 0, maxb);
        printf("{\"cudaMalloc_ptr_low_bits\": \"%llx\"}\n", (unsigned long long)((uintptr_t)rows & ((1ull << 29) - 1)));
        for (size_t f = (size_t)64 << 20; f <= maxb; f <<= 1) run("cudaMalloc", rows, f, idx, out);
        cudaFree(rows);
    }
    {
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = 0;
        size_t gmin = 0, grec = 0;
        cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
        cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
        printf("{\"vmm_granularity_min\": %zu, \"recommended\": %zu}\n", gmin, grec);
        CUdeviceptr va = 0;
        CUresult r = cuMemAddressReserve(&va, maxb, (size_t)512 << 20, 0, 0);
        CUmemGenericAllocationHandle hnd;
        CUresult r2 = cuMemCreate(&hnd, maxb, &prop, 0);
        CUresult r3 = cuMemMap(va, maxb, 0, hnd, 0);
        CUmemAccessDesc acc = {};
        acc.location = prop.location;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CUresult r4 = cuMemSetAccess(va, maxb, &acc, 1);
        printf("{\"vmm\": [%d, %d, %d, %d], \"va_low_bits\": \"%llx\"}\n", (int)r, (int)r2, (int)r3, (int)r4, (unsigned long long)(va & ((1ull << 29) - 1)));
        if (r == 0 && r2 == 0 && r3 == 0 && r4 == 0) {
            cudaMemset((void*)va, 0, maxb);
            for (size_t f = (size_t)64 << 20; f <= maxb; f <<= 1) run("vmm_512MB_aligned", (const float4*)va, f, idx, out);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
