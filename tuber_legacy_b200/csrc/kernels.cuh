// sm_100a kernels of the tuber hot path. HBM-bound integer / float streaming work:
// no tensor cores. What matters here is 128-bit coalesced access, one pass over each byte, TMA /
// shared-memory staging, grids sized in multiples of the SM count, and one stable radix placement
// fused with the emit so sorted payloads are written exactly once.
//
//   storage.cuh    convert_*, mark_*presence, integrate, query_*        uploads, Ecs::query
//   prepare.cuh    prepare_flat, prepare_level, level_*                 frame pass A, hierarchy planning
//   placement.cuh  scan_hist, tile_hist, tile_chunk_sums, tile_scan,    reduce-then-scan radix placement,
//                  rank_tile, radix_scatter, build_batches, batch_*     batch table
//   emit.cuh       emit_kernel, emit_small_kernel                       last pass fused with compose + emit
//   chunked.cuh    chunk_place, chunk_tile_scan, emit_list                 sorted frames whose row reads stay in L2
//   shard.cuh      place_sums, place_write                              multi-GPU global draw order
//   present.cuh    raster_quads, raster_resolve                         headless presenter of the draw list
//   mathops.cuh    quat_*, mat4_*                                       the rest of tuber-math, batched
#pragma once
#include "common.cuh"
#include "storage.cuh"
#include "prepare.cuh"
#include "placement.cuh"
#include "emit.cuh"
#include "chunked.cuh"
#include "shard.cuh"
#include "mathops.cuh"
#include "present.cuh"
