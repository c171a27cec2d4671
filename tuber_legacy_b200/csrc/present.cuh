// Headless presenter: what sits behind GraphicsAPI::render_scene once the frame has produced its
// draw list (crates/tuber-graphics/src/lib.rs:134-156 submits an empty encoder and presents; a real renderer draws the
// batches in order). This one consumes the frame's OWN outputs — the clip-space quad vertices in draw order — and
// rasterises them with the painter's rule (a later draw position covers an earlier one) into an id buffer, so that
// the draw list can be verified pixel by pixel against a CPU rasterisation of the oracle's list, and resolves the id
// buffer to RGBA through the instances' (texture, layer). Quads entirely outside the clip square are culled here, per
// quad, on the presenter side: the frame itself never culls, so its cached placement survives camera moves.
//
// Coverage is decided by four edge functions evaluated at the pixel centre with individually rounded f32 operations
// (no FMA), so a numpy float32 restatement makes the same decision for every pixel.
#pragma once
#include "common.cuh"

namespace tb {

struct RasterParams {
    const float4* vertices;  // 4 per instance: {x, y, z, uv}, clip space
    uint32_t count;          // instances, in draw order
    uint32_t width, height;
    uint32_t* ids;           // [height][width]: 1 + draw position of the topmost quad, 0 = background
    unsigned long long* counters;  // [0] quads with at least one candidate pixel, [1] quads culled (outside the clip square)
};

__device__ __forceinline__ float pixel_center(uint32_t i, float step) { return subr(mulr(addr((float)i, 0.5f), step), 1.0f); }

// edge function of (a -> b) at p
__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float px, float py) {
    return subr(mulr(subr(bx, ax), subr(py, ay)), mulr(subr(by, ay), subr(px, ax)));
}

// One warp per quad: the lanes share the quad's bounding box of pixels.
__global__ void __launch_bounds__(kThreads) raster_quads_kernel(const RasterParams rp) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * kWarps;
    const float sx = 2.0f / (float)rp.width, sy = 2.0f / (float)rp.height;
    unsigned long long drawn = 0, culled = 0;
    for (uint32_t q = blockIdx.x * kWarps + (threadIdx.x >> 5); q < rp.count; q += nwarps) {
        const float4 v0 = __ldg(rp.vertices + (size_t)q * 4), v1 = __ldg(rp.vertices + (size_t)q * 4 + 1);
        const float4 v2 = __ldg(rp.vertices + (size_t)q * 4 + 2), v3 = __ldg(rp.vertices + (size_t)q * 4 + 3);
        const float xmin = fminf(fminf(v0.x, v1.x), fminf(v2.x, v3.x)), xmax = fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x));
        const float ymin = fminf(fminf(v0.y, v1.y), fminf(v2.y, v3.y)), ymax = fmaxf(fmaxf(v0.y, v1.y), fmaxf(v2.y, v3.y));
        if (!(xmax >= -1.0f && xmin <= 1.0f && ymax >= -1.0f && ymin <= 1.0f)) {  // also rejects NaN
            culled += lane == 0;
            continue;
        }
        // conservative pixel box (one pixel of slack on every side; the edge functions decide)
        const int i0 = max(0, (int)floorf((xmin + 1.0f) * 0.5f * (float)rp.width) - 1);
        const int i1 = min((int)rp.width - 1, (int)floorf((xmax + 1.0f) * 0.5f * (float)rp.width) + 1);
        const int j0 = max(0, (int)floorf((ymin + 1.0f) * 0.5f * (float)rp.height) - 1);
        const int j1 = min((int)rp.height - 1, (int)floorf((ymax + 1.0f) * 0.5f * (float)rp.height) + 1);
        if (i1 < i0 || j1 < j0) {
            culled += lane == 0;
            continue;
        }
        drawn += lane == 0;
        const uint32_t w = (uint32_t)(i1 - i0 + 1), npix = w * (uint32_t)(j1 - j0 + 1);
        for (uint32_t p = lane; p < npix; p += 32) {
            const uint32_t i = (uint32_t)i0 + p % w, j = (uint32_t)j0 + p / w;
            const float px = pixel_center(i, sx), py = pixel_center(j, sy);
            const float e0 = edge_fn(v0.x, v0.y, v1.x, v1.y, px, py), e1 = edge_fn(v1.x, v1.y, v2.x, v2.y, px, py);
            const float e2 = edge_fn(v2.x, v2.y, v3.x, v3.y, px, py), e3 = edge_fn(v3.x, v3.y, v0.x, v0.y, px, py);
            const bool in = (e0 >= 0.0f && e1 >= 0.0f && e2 >= 0.0f && e3 >= 0.0f) || (e0 <= 0.0f && e1 <= 0.0f && e2 <= 0.0f && e3 <= 0.0f);
            if (in) atomicMax(rp.ids + (size_t)j * rp.width + i, q + 1u);  // painter's rule: the later draw wins
        }
    }
    drawn = __reduce_add_sync(0xffffffffu, (uint32_t)drawn);
    culled = __reduce_add_sync(0xffffffffu, (uint32_t)culled);
    if (lane == 0) {
        if (drawn) atomicAdd(&rp.counters[0], drawn);
        if (culled) atomicAdd(&rp.counters[1], culled);
    }
}

// id buffer -> RGBA8: a fixed colour per (texture, layer) of the instance on top, opaque; background 0.
__global__ void __launch_bounds__(kThreads) raster_resolve_kernel(const uint32_t* __restrict__ ids, const float4* __restrict__ instances,
                                                                   uint32_t npix, uint32_t* __restrict__ rgba) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const uint32_t id = ids[p];
    uint32_t c = 0;
    if (id) {
        const float4 s = __ldg(instances + (size_t)(id - 1) * 5 + 4);  // {size.w, size.h, texture, layer}
        uint32_t h = __float_as_uint(s.z) * 0x9E3779B1u ^ (__float_as_uint(s.w) + 0x7F4A7C15u) * 0x85EBCA6Bu;
        h ^= h >> 15;
        c = 0xFF000000u | (h & 0x00FFFFFFu);
    }
    rgba[p] = c;
}

}  // namespace tb
