// libtuber_b200.so — host side of the C ABI in include/tuber_b200.h.
//
// One tb_ctx per GPU owns the device-resident ECS columns (two-sector rows, presence bitsets,
// velocity and parent columns), the frame scratch (key statistics, histograms, tile offset
// tables, sort ping-pong buffers) and the emitted buffers. A frame is a short fixed sequence
// of kernel launches on the context's stream; the only host round trip is the
// read-back of the counts and key statistics that tb_frame_transform_batch returns.
//
// There is no CPU fallback anywhere in this file: without an sm_100 device tb_ctx_create
// fails with TB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/tuber_b200.h"
#include "comm.h"
#include "kernels.cuh"

using namespace tb;

namespace {

struct ProfEntry {
    const char* name;
    cudaEvent_t e0, e1;
    uint64_t bytes;
};

// per-kernel-name totals since tb_profile_enable
struct ProfTotal {
    const char* name;
    double ms;
    uint64_t bytes;
    uint32_t launches;
};

}  // namespace

// Everything the launches of a full frame read on the host side. Two frames with equal signatures
// issue identical launches, so the second can replay the first one's graph.
struct FrameSig {
    uint64_t n = 0, max_batches = 0, global_first = 0, level_epoch = 0;
    const void* ptr[26] = {};
    FramePlan plan{};
    uint32_t flags[18] = {};
    uint32_t vp_bits[17] = {};  // has_vp, then the 16 matrix entries
    bool operator==(const FrameSig& o) const {
        return memcmp(vp_bits, o.vp_bits, sizeof(vp_bits)) == 0 && n == o.n && max_batches == o.max_batches && global_first == o.global_first && level_epoch == o.level_epoch &&
               memcmp(ptr, o.ptr, sizeof(ptr)) == 0 && memcmp(flags, o.flags, sizeof(flags)) == 0 && plan_equal(plan, o.plan);
    }
};

// View of `cap` rows stored at `mem` (cap x 64 bytes) in either layout (compose.cuh).
RowView make_row_view(float4* mem, uint64_t cap, bool interleaved, float* yaw) {
    RowView v;
    v.yaw = yaw;
    v.t = mem;
    v.a = interleaved ? mem + 1 : mem + cap;
    v.b = interleaved ? mem + 2 : mem + 2 * cap;
    v.st = interleaved ? 4 : 1;
    v.sb = interleaved ? 4 : 2;
    return v;
}

struct tb_ctx {
    int device = 0;
    int sm_count = 148;
    int libm_fma = 1;  // which build of sinf / cosf the host libm uses (mirrored in the device's c_libm_fma)
    cudaStream_t own_stream = nullptr, stream = nullptr;
    uint64_t cap = 0, n = 0;
    std::string err;

    // storage
    int interleaved = 0;
    bool layout_auto = true;  // frames that gather rows in sorted order switch to 64-byte rows on their own
    bool rows_touched = false;
    float4* rows_mem = nullptr;
    float* yaw_mem = nullptr;  // angle.z per entity (the rows keep sin / cos of its half)
    RowView rows{};
    uint64_t* mask[TB_COMPONENT_COUNT] = {};
    bool store[TB_COMPONENT_COUNT] = {};
    float* vel = nullptr;
    float* spin = nullptr;       // [cap] angular velocity about z
    uint32_t* gen = nullptr;     // [cap] generation of every entity slot (bumped by deletes)
    cudaStream_t side_stream[TB_SYSTEM_COUNT] = {};  // stages of tb_bundle_step: one stream per concurrent system
    cudaEvent_t fork_event = nullptr, join_event[TB_SYSTEM_COUNT] = {};
    // a bundle stepped again with the same systems, dt and storage replays its captured graph (fork / join included)
    cudaGraphExec_t bundle_graph = nullptr;
    std::vector<int> bundle_sig_systems;
    uint64_t bundle_sig[8] = {};
    bool bundle_graph_valid = false;
    uint32_t* parent = nullptr;
    void* stage = nullptr;
    size_t stage_bytes = 0;

    // frame scratch
    uint32_t* scratch = nullptr;
    size_t scratch_words = 0;
    void* keys[2] = {nullptr, nullptr};
    void* ekeys = nullptr;  // short keys in entity order (what the key pass writes; kept from frame to frame)
    uint32_t* idx[2] = {nullptr, nullptr};
    size_t sort_cap = 0, sort_keybytes = 0;
    uint32_t* texlayer = nullptr;
    size_t texlayer_cap = 0;
    WorldRow* world = nullptr;
    size_t world_cap = 0;
    bool world_valid = false;

    // outputs
    float4* inst = nullptr;
    float4* vtx = nullptr;
    uint32_t* order = nullptr;
    size_t out_cap = 0;
    BatchOut* batches = nullptr;
    uint64_t max_batches = 65536, batches_cap = 0;
    void* ext_inst = nullptr;
    void* ext_vtx = nullptr;
    void* ext_order = nullptr;
    uint64_t ext_cap = 0;
    uint64_t last_instances = 0, last_batches = 0;
    bool frame_valid = false;

    // camera folded into the emitted geometry (tb_frame_set_view_projection)
    float vp[16] = {};
    bool has_vp = false;

    // plan
    FramePlan plan{};
    bool plan_valid = false;
    // single-kernel steady-state frames: per-tile bucket offsets cached from the last full frame
    bool static_valid = false;
    // sorted frames: the previous frame's sorted (key, id) arrays and tile offsets are still in place; a frame
    // whose key pass finds every entity's key unchanged skips its sort passes (decided on the device)
    bool sorted_valid = false;
    bool fast_clean = false;  // no component was uploaded since the last frame that ran a key pass
    int sorted_cache = 1;
    // ... and the draw order itself (a copy of the last re-sorted frame's `order` output): frames that keep
    // the sort emit straight from it (emit_ordered_kernel)
    uint32_t* final_idx = nullptr;   // [n] entity id (global) per draw position
    void* final_keys = nullptr;      // [n] short key per draw position
    bool final_valid = false;
    // chunked sort (chunked.cuh): final_idx / final_keys are then the chunk-major list (padded per chunk) and final_dst
    // the destination of every list position in the global draw order
    uint32_t* final_dst = nullptr;
    int chunked_sort = 0;            // option; off by default: measured slower than the kept-order paths on B200 (profiles/README.md)
    int chunk_shift_opt = 0;         // option: log2 of the chunk size (0: chosen from the key width)
    int chunk_l2_hint = 1;           // option: row gathers of the list emit carry the L2 evict_last policy
    int hier_fused = 1;              // option: kept-order scene-graph frames compose the last level inside the emit
    bool chunked = false;            // the current plan runs as a chunked sort
    ChunkGeom cg{};
    size_t chunk_cnt_off = 0, chunk_remap_off = 0;  // where the kept chunk tables live in the scratch buffer
    uint64_t sorted_count = 0, sorted_batches = 0;
    // static scenes (no upload, no tick for a few frames): a copy of the rows (hierarchy: world rows) in draw order,
    // from which frames stream instead of gathering
    float4* seq_rows = nullptr;
    size_t seq_cap = 0;
    bool seq_valid = false, seq_world = false;
    // ... which is the WORKING storage of moving flat scenes: velocities are kept in draw order too and the velocity
    // system runs on the copy (emit_sequential_kernel<VALIDATE>); seq_live: the copy holds newer translations than the
    // entity-order rows, which are brought up to date lazily (ensure_rows_current)
    float* seq_vel = nullptr;
    size_t seq_vel_cap = 0;
    bool seq_has_vel = false, seq_live = false;
    uint64_t hidden_movers = 0;  // entities with Transform and Velocity but no Sprite, counted when the copy is made
    // scene graphs: the copy holds the LOCAL rows (+ velocities + parent ids) of the drawn entities; the last level's
    // (the leaves') are the working storage, the inner levels stay in the entity-order rows (hierarchy_seq_frame)
    uint32_t* seq_par = nullptr;
    size_t seq_par_cap = 0;
    bool seq_graph = false;
    uint32_t seq_leaf_first = 0;
    uint32_t seq_split = 0;  // float4 index where the A1 B0 B1 block of a local-row copy starts (0: 64-byte world rows)
    uint32_t hier_streak = 0;  // consecutive scene-graph frames that kept their sort
    // A single-kernel frame whose validation fails costs its own time plus the full frame's. After a failure the next
    // frames go straight to the full path (whose key pass skips the sort on the device when nothing changed) and the
    // single-kernel attempts resume once this many full frames have passed: keys that change every frame cost one
    // full frame per frame, not a failed attempt on top.
    uint32_t fast_backoff = 0;
    uint32_t clean_streak = 0;  // consecutive frames served by a kept draw order with nothing touched
    int static_buckets = 1;
    uint64_t static_count = 0, static_batches = 0;
    size_t static_offs_off = 0;
    uint32_t static_tiles = 0;
    uint32_t force_path = TB_PATH_NONE;
    int digit_bits = kMaxDigitBits;
    int staged_emit = 1;
    int small_emit = 1;
    int lt_limit = (int)kMaxLtBits;
    int z_dictionary = 1;

    // plan discovery: the distinct depths of the scene (device set + pinned host copy)
    unsigned long long* zset = nullptr;  // [kZSetSlots] entries + 1 overflow word
    unsigned long long* h_zset = nullptr;
    bool zset_active = false;

    // steady-state full frames replay a captured CUDA graph (memset, every kernel, header copy)
    int use_graphs = 1;
    cudaGraphExec_t frame_graph = nullptr;
    FrameSig frame_sig{};
    bool frame_graph_valid = false;
    bool frame_graph_seen = false;  // frame_sig holds the previous steady frame's signature
    uint32_t frame_graph_launches = 0;
    uint64_t level_epoch = 0;  // bumped whenever the level structure is rebuilt

    // hierarchy
    bool levels_dirty = true;
    uint32_t nlevels = 1;
    bool level_ranges = true;
    std::vector<uint32_t> level_first, level_count;
    uint32_t* level_list = nullptr;
    uint32_t* sane_parent = nullptr;
    size_t level_cap = 0;

    // velocity system
    bool pending = false;
    float pending_dt = 0.0f;
    uint32_t pending_ticks = 0;  // consecutive ticks of the same dt no frame has consumed yet: run as one pass

    // shard
    int rank = 0, nranks = 1;
    uint64_t global_first = 0;
    int shard_state = 0;  // 0 idle, 1 stats taken, 2 global plan set, 3 pass A done (histogram out), 4 placed
    uint64_t shard_count = 0, shard_global_count = 0;
    uint32_t* remap = nullptr;      // [1 << nbits]
    uint32_t* lt_global = nullptr;  // [1 << ltbits]
    uint32_t* place_tmp = nullptr;  // block sums of the placement scans
    size_t remap_cap = 0, lt_global_cap = 0, place_tmp_cap = 0;

    // library-owned multi-GPU frame (tb_shard_init / tb_frame_sharded): the exchange transport, the presenting
    // rank's draw-list buffers (owned by the root, peer-mapped by everyone else) and small exchange buffers
    std::unique_ptr<Comm> comm;
    int list_root = -1;
    void* list_inst = nullptr;   // [list_cap] tb_instance
    void* list_vtx = nullptr;    // [list_cap] 4 x tb_quad_vertex (root only)
    void* list_order = nullptr;  // [list_cap] u32
    uint64_t list_cap = 0;
    bool list_ipc = false;       // list_inst / list_order were opened with cudaIpcOpenMemHandle
    uint64_t* d_xchg = nullptr;  // device: [4] send + [4 * nranks] receive words
    uint64_t* h_xchg = nullptr;  // pinned mirror
    // list frames that keep the previous frame's order (tb_frame_sharded)
    bool list_kept = false;    // final_idx / final_keys / remap still describe this shard's part of the global list
    int list_backoff = 0;      // full frames to run before the kept order is tried again (the same on every rank)
    uint32_t list_seq = 0;     // sequence number carried by the completion flags (the same on every rank)
    uint64_t list_batches = 0;
    unsigned long long* list_flags = nullptr;  // in the presenting rank's memory, mapped by all: [nranks][kListChunks]
    void* list_rows = nullptr;                 // in the presenting rank's memory, mapped by all: a WorldRow per list position
    uint32_t* list_state = nullptr;            // presenting rank: expansion state (kListStateWords) + stream starts (kListMaxStreams)
    unsigned long long* merge_keys = nullptr;  // wide keys: this rank's sorted full keys, padded to the longest rank
    unsigned long long* merge_all = nullptr;   // ... and every rank's
    size_t merge_keys_cap = 0, merge_all_cap = 0;
    BatchOut* merge_batches = nullptr;         // every rank's local batch table
    size_t merge_batches_cap = 0;
    uint32_t* shard_hist = nullptr;      // [nbins] this rank's short-key histogram
    uint32_t* shard_all_hists = nullptr; // [nranks * nbins]
    size_t shard_hist_cap = 0, shard_all_cap = 0;

    // pinned host mirror of the scratch header
    uint64_t* h_hdr = nullptr;
    void* h_peek = nullptr;  // pinned, kPeekBytes: small device -> host reads land here (peek_async)
    cudaEvent_t copied = nullptr;  // marks the last copy of a bulk upload (LinkTurn)

    // instrumentation
    bool prof = false;
    std::vector<ProfEntry> prof_entries;   // launched, not yet harvested
    std::vector<ProfTotal> prof_totals;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pool;
    uint64_t launches = 0;
    uint32_t frame_launches = 0;
};

// Everything a later frame would reuse without a key pass of its own (per-tile offsets of few-bucket scenes, the
// kept sort of sorted scenes) is dropped; the next frame computes all of it again.
inline void drop_kept_frames(tb_ctx* c) { c->static_valid = c->sorted_valid = c->list_kept = false; }


namespace {

// scratch header layout (32-bit words). Words [0, kHdrFrame) outlive a frame; everything from
// kHdrFrame up to ScratchLayout::total_words is zeroed by ONE memset at the start of every frame.
constexpr size_t kHdrRangeErr = 0;     // u32: sprite range violations since the last check
constexpr size_t kHdrPending = 1;      // u32: scratch of the level planner
constexpr size_t kHdrLevelStats = 2;   // 2 x u32
constexpr size_t kHdrFrame = 4;
constexpr size_t kHdrStats = 4;        // 3 x u64: OR(key), OR(~key), renderable count
constexpr size_t kHdrNumBatches = 10;  // u32
constexpr size_t kHdrZMiss = 12;       // u64 (stats[4]): a depth outside the plan's z dictionary was seen
constexpr size_t kHdrKeysChanged = 14;  // u64 (stats[5]): an entity's short key differs from the previous frame's
constexpr size_t kHdrReadWords = 12;   // what a frame copies back: stats, batch count, dictionary misses, key changes
constexpr size_t kHdrHist = 16;        // kMaxPasses * 256 u32: raw digit histograms
constexpr size_t kHdrWords = kHdrHist + kMaxPasses * 256;  // then (layer,texture) bins, tile tables

tb_status fail(tb_ctx* c, tb_status s, const std::string& msg) {
    if (c) c->err = msg;
    return s;
}

tb_status cuda_fail(tb_ctx* c, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    return fail(c, e == cudaErrorMemoryAllocation ? TB_ERR_OOM : TB_ERR_CUDA, m);
}

#define TB_CUDA(c, call)                                   \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail((c), e__, #call); \
    } while (0)

void prof_begin(tb_ctx* c, const char* name, uint64_t bytes) {
    if (!c->prof) return;
    ProfEntry pe{name, nullptr, nullptr, bytes};
    if (!c->prof_pool.empty()) {
        pe.e0 = c->prof_pool.back().first;
        pe.e1 = c->prof_pool.back().second;
        c->prof_pool.pop_back();
    } else {
        cudaEventCreate(&pe.e0);
        cudaEventCreate(&pe.e1);
    }
    cudaEventRecord(pe.e0, c->stream);
    c->prof_entries.push_back(pe);
}
void prof_end(tb_ctx* c) {
    if (!c->prof) return;
    cudaEventRecord(c->prof_entries.back().e1, c->stream);
}
// Folds finished launches into the per-name totals. Call only when the stream is idle.
void prof_harvest(tb_ctx* c) {
    for (auto& pe : c->prof_entries) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, pe.e0, pe.e1) != cudaSuccess) {
            cudaGetLastError();
            ms = 0.0f;
        }
        ProfTotal* t = nullptr;
        for (auto& pt : c->prof_totals)
            if (pt.name == pe.name || strcmp(pt.name, pe.name) == 0) t = &pt;
        if (!t) {
            c->prof_totals.push_back(ProfTotal{pe.name, 0.0, 0, 0});
            t = &c->prof_totals.back();
        }
        t->ms += ms;
        t->bytes += pe.bytes;
        t->launches += 1;
        c->prof_pool.emplace_back(pe.e0, pe.e1);
    }
    c->prof_entries.clear();
}
void prof_clear(tb_ctx* c) {
    prof_harvest(c);
    for (auto& ev : c->prof_pool) {
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    c->prof_pool.clear();
    c->prof_totals.clear();
}

// Makes the context's device current for the duration of one C-ABI call and restores the caller's afterwards: a
// process may hold one context per GPU, or call from a thread whose current device is another one.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (device < 0) return;
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
        if (prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
    ~DeviceGuard() {
        if (switched && prev >= 0) cudaSetDevice(prev);
    }
};
#define TB_ENTER(c) DeviceGuard device_guard__((c) ? (c)->device : -1)

// Device scratch that lives for one call.
struct DeviceTemp {
    void* p = nullptr;
    DeviceTemp() = default;
    DeviceTemp(const DeviceTemp&) = delete;
    DeviceTemp& operator=(const DeviceTemp&) = delete;
    ~DeviceTemp() { cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
};

// NVTX range (header-only NVTX3: a call through an empty function table unless a tool such as nsys / ncu is attached).
// Every frame entry point and every kernel launch carries one, named like the profiler's rows (tb_profile_*); the
// reference's only tracing is `trace!("Starting scene render")` / `"Render finished"` (tuber-graphics/src/lib.rs:136,152).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
    ~NvtxRange() { nvtxRangePop(); }
};

#define TB_LAUNCH(c, name, bytes, kern, grid, block, smem, ...)        \
    do {                                                               \
        NvtxRange nvtx_range__(name);                                  \
        prof_begin((c), (name), (bytes));                              \
        kern<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__);   \
        prof_end((c));                                                 \
        (c)->launches++;                                               \
        (c)->frame_launches++;                                         \
    } while (0)

inline uint32_t div_up(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// Small device -> host reads (the frame's result header, counters, a batch table) do not go through a copy engine:
// a cudaMemcpyAsync queues on the device's D2H engine behind whatever bulk download ANOTHER context has in flight
// (measured: a 64-byte read waiting 25 ms behind a 1.3 GB copy, tools/profile_e2e.py). One warp stores the words
// into pinned host memory instead (cudaMallocHost memory is device-addressable under unified addressing).
__global__ void peek_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t nwords) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
constexpr size_t kPeekBytes = 64u << 10;
// pinned_dst: pinned host memory; bytes: a multiple of 4. Counted as a launch of the context, not as a kernel of the frame
// (it stands where a memcpy node stood).
inline void peek_async(tb_ctx* c, void* pinned_dst, const void* d_src, size_t bytes) {
    const uint32_t nwords = (uint32_t)(bytes / 4);
    peek_kernel<<<nwords > 1024 ? 8 : 1, nwords > 32 ? 256 : 32, 0, c->stream>>>((uint32_t*)pinned_dst, (const uint32_t*)d_src, nwords);
    c->launches++;
}
// One BULK host transfer per direction per device at a time, process-wide. Several contexts (a streaming user keeps
// two or three in flight so that one step's download overlaps the next step's upload) otherwise put their copies on the
// link together, and same-direction copies sharing PCIe measured worse than taking turns: 59.8 ms per 3.76 GB step
// against 50.8 ms (tools/profile_e2e.py; the link's duplex bound is 49 ms). A turn lasts until the copy has finished.
constexpr size_t kBulkTransferBytes = 8u << 20;
inline std::mutex& link_mutex(int device, int direction) {
    static std::mutex m[64][2];
    return m[(unsigned)device & 63u][direction];
}
struct LinkTurn {
    enum { kUp = 0, kDown = 1 };
    std::unique_lock<std::mutex> lk;
    LinkTurn(const tb_ctx* c, int direction, size_t bytes) {
        if (bytes >= kBulkTransferBytes) lk = std::unique_lock<std::mutex>(link_mutex(c->device, direction));
    }
    bool held() const { return lk.owns_lock(); }
};
// end of an upload's turn: wait for the copies enqueued so far (not for the kernels that follow them)
inline cudaError_t finish_upload_turn(tb_ctx* c, const LinkTurn& turn) {
    if (!turn.held()) return cudaSuccess;
    cudaError_t e = cudaEventRecord(c->copied, c->stream);
    return e == cudaSuccess ? cudaEventSynchronize(c->copied) : e;
}

// ... and the blocking form for a few words into any host memory
inline cudaError_t peek_sync(tb_ctx* c, void* dst, const void* d_src, size_t bytes) {
    peek_async(c, c->h_peek, d_src, bytes);
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) memcpy(dst, c->h_peek, bytes);
    return e;
}

template <typename T>
tb_status grow(tb_ctx* c, T** p, size_t* cap, size_t need, size_t elem) {
    if (*cap >= need && *p) return TB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, need * elem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc");
    *p = (T*)q;
    *cap = need;
    return TB_OK;
}

tb_status ensure_stage(tb_ctx* c, size_t bytes) {
    size_t cap = c->stage_bytes;
    char* p = (char*)c->stage;
    tb_status s = grow(c, &p, &cap, bytes, 1);
    c->stage = p;
    c->stage_bytes = cap;
    return s;
}

const uint64_t* mask_or_null(tb_ctx* c, int comp) { return c->mask[comp]; }

tb_status check_range(tb_ctx* c, uint64_t first, uint64_t n, const void* p) {
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    if (!p) return fail(c, TB_ERR_INVALID, "null pointer");
    if (first + n > c->n) return fail(c, TB_ERR_INVALID, "range outside entity_count");
    return TB_OK;
}

constexpr size_t kStageChunk = 64u << 20;

tb_status mark_presence(tb_ctx* c, int comp, uint64_t first, uint64_t n, int set) {
    if (n == 0) return TB_OK;
    uint64_t nw = ((first + n - 1) >> 6) - (first >> 6) + 1;
    TB_LAUNCH(c, "mark_presence", 0, mark_presence_kernel, div_up(nw, kThreads), kThreads, 0, c->mask[comp], first, n, set);
    return TB_OK;
}

// The entity-order rows hold what every tick so far has produced: when the draw-ordered copy is the working storage
// (seq_live), its translation chunks are written back first.
tb_status ensure_rows_current(tb_ctx* c) {
    if (!c->seq_live) return TB_OK;
    c->seq_live = false;
    TB_LAUNCH(c, "sync_rows", c->sorted_count * 36ull, sync_rows_kernel, div_up(c->sorted_count, kThreads), kThreads, 0,
              (const float4*)c->seq_rows, (const uint32_t*)c->final_idx, (uint32_t)c->global_first, (uint32_t)c->sorted_count, c->rows,
              c->seq_graph ? c->seq_leaf_first : 0u, c->seq_split ? 1u : 4u);
    return TB_OK;
}

tb_status flush_integrate(tb_ctx* c) {
    tb_status sync = ensure_rows_current(c);
    if (sync) return sync;
    if (!c->pending) return TB_OK;
    c->pending = false;
    if (c->n == 0 || !c->store[TB_TRANSFORM] || !c->store[TB_VELOCITY]) return TB_OK;
    TB_LAUNCH(c, "integrate", c->n * 36ull, integrate_kernel, div_up(c->n, kThreads), kThreads, 0, c->rows,
              mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_VELOCITY), c->vel, c->pending_dt, c->pending_ticks, (uint32_t)c->n);
    c->world_valid = false;
    c->frame_valid = false;
    // the rows moved outside a frame: every frame kind that runs WITHOUT a key pass of its own (hierarchy replay,
    // draw-ordered copy, single-kernel frames placed by cached tile offsets) needs one clean key pass first
    c->fast_clean = c->seq_valid = false;
    c->static_valid = false;
    return TB_OK;
}

// The velocity system for the entities no frame draws (Transform and Velocity, no Sprite), on the entity-order rows: the
// frames that tick inside their emit only visit drawn entities.
void tick_hidden_movers(tb_ctx* c) {
    TB_LAUNCH(c, "integrate", c->n, integrate_kernel, div_up(c->n, kThreads), kThreads, 0, c->rows, mask_or_null(c, TB_TRANSFORM),
              mask_or_null(c, TB_VELOCITY), c->vel, c->pending_dt, c->pending_ticks, (uint32_t)c->n, 0u, mask_or_null(c, TB_SPRITE));
}

// Scratch layout for one frame under a plan.
struct ScratchLayout {
    size_t lt_off = 0, lt_words = 0;
    uint32_t tiles[kMaxPasses];
    uint32_t tile_elems[kMaxPasses];
    size_t counts_off[kMaxPasses];   // per-tile digit counts  [R][tiles]
    size_t offs_off[kMaxPasses];     // per-tile digit offsets [R][tiles]
    size_t csum_off[kMaxPasses];     // per-chunk (1024 tiles) sums [R][chunks]
    bool small = false;              // last pass runs the warp-autonomous few-bucket emit
    size_t total_words = 0, alloc_words = 0;
    // chunked sort: H[nchunks][K] (zeroed every frame), keytot[K], remap[nchunks][K], chunk_cnt[nchunks]
    bool chunked = false;
    size_t H_off = 0, keytot_off = 0, remap_off = 0, cnt_off = 0;
};

bool use_small_emit(const tb_ctx* c, const FramePlan& plan) {
    return c->small_emit != 0 && plan.npasses == 1 && plan.bits[0] <= kSmallMaxBits;
}

// Scratch of one frame under a plan: [header | (layer,texture) bins | pass-0 tile counts when the key
// pass accumulates them] is zeroed every frame; the tile count / offset tables behind it are fully
// overwritten by tile_hist / tile_scan.
ScratchLayout layout_scratch(const FramePlan& plan, uint64_t n, bool use_lt, bool zero_counts0, bool small = false,
                             const ChunkGeom* cg = nullptr) {
    ScratchLayout L;
    L.small = small;
    L.chunked = cg != nullptr;
    size_t off = kHdrWords;
    L.lt_off = off;
    L.lt_words = use_lt ? ((size_t)1 << plan.ltbits) : 0;
    off += L.lt_words;
    for (uint32_t p = 0; p < plan.npasses; ++p) {
        const bool last = p + 1 == plan.npasses;
        L.tile_elems[p] = cg ? kScatterTile : last ? (small ? kWarpTile : kEmitTile) : kScatterTile;
        L.tiles[p] = cg ? (uint32_t)(((uint64_t)cg->nchunks << cg->shift) / kScatterTile) : div_up(n, L.tile_elems[p]);
    }
    if (cg) {
        L.H_off = off;
        off += (size_t)cg->nchunks << cg->nbits;
    }
    if (zero_counts0) {
        // the key pass accumulates pass-0 per-tile counts with REDs: they start from zero
        L.counts_off[0] = off;
        off += (size_t)L.tiles[0] << plan.bits[0];
    }
    L.total_words = off;  // everything up to here is zeroed every frame
    for (uint32_t p = 0; p < plan.npasses; ++p) {
        if (!(p == 0 && zero_counts0)) {
            L.counts_off[p] = off;
            off += (size_t)L.tiles[p] << plan.bits[p];
        }
        L.offs_off[p] = off;
        off += (size_t)L.tiles[p] << plan.bits[p];
        L.csum_off[p] = off;
        off += (size_t)div_up(L.tiles[p], 1024) << plan.bits[p];
    }
    if (cg) {
        L.keytot_off = off;
        off += (size_t)1 << cg->nbits;
        L.remap_off = off;
        off += (size_t)cg->nchunks << cg->nbits;
        L.cnt_off = off;
        off += cg->nchunks;
    }
    L.alloc_words = off;
    return L;
}

tb_status ensure_scratch(tb_ctx* c, size_t words) {
    if (c->scratch_words >= words) return TB_OK;
    // keep the header values that outlive a frame (range error counter)
    uint32_t* old = c->scratch;
    size_t old_words = c->scratch_words;
    void* q = nullptr;
    size_t want = std::max(words, kHdrWords);
    cudaError_t e = cudaMalloc(&q, want * 4);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(scratch)");
    cudaMemsetAsync(q, 0, want * 4, c->stream);
    if (old) {
        cudaMemcpyAsync(q, old, std::min(old_words, kHdrWords) * 4, cudaMemcpyDeviceToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaFree(old);
    }
    c->scratch = (uint32_t*)q;
    c->scratch_words = want;
    drop_kept_frames(c);  // cached tile offsets lived in the old buffer
    return TB_OK;
}

tb_status realloc_dev(tb_ctx* c, void** p, size_t bytes) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc");
    return TB_OK;
}

tb_status ensure_outputs(tb_ctx* c, uint64_t n) {
    tb_status s;
    if (c->ext_inst) {
        if (c->ext_cap < n) return fail(c, TB_ERR_CAPACITY, "external output buffers too small");
    } else if (c->out_cap < n || !c->inst) {
        c->out_cap = 0;
        if ((s = realloc_dev(c, (void**)&c->inst, (size_t)n * 80))) return s;
        if ((s = realloc_dev(c, (void**)&c->vtx, (size_t)n * 64))) return s;
        if ((s = realloc_dev(c, (void**)&c->order, (size_t)n * 4))) return s;
        c->out_cap = n;
    }
    if (!c->batches || c->batches_cap < c->max_batches) {
        c->batches_cap = 0;
        drop_kept_frames(c);  // the single-kernel frames reuse the table that lived in the old buffer
        if ((s = realloc_dev(c, (void**)&c->batches, (size_t)c->max_batches * sizeof(BatchOut)))) return s;
        c->batches_cap = c->max_batches;
    }
    return TB_OK;
}

tb_status ensure_sort_buffers(tb_ctx* c, uint64_t n, size_t keybytes) {
    if (c->sort_cap >= n && c->sort_keybytes >= keybytes && c->keys[0]) return TB_OK;
    for (int k = 0; k < 2; ++k) {
        if (c->keys[k]) cudaFree(c->keys[k]);
        if (c->idx[k]) cudaFree(c->idx[k]);
        c->keys[k] = nullptr;
        c->idx[k] = nullptr;
    }
    if (c->ekeys) cudaFree(c->ekeys);
    if (c->final_idx) cudaFree(c->final_idx);
    if (c->final_keys) cudaFree(c->final_keys);
    if (c->final_dst) cudaFree(c->final_dst);
    c->ekeys = c->final_keys = nullptr;
    c->final_idx = c->final_dst = nullptr;
    c->sort_cap = 0;
    c->sorted_valid = c->list_kept = c->final_valid = false;
    {
        cudaError_t e = cudaMalloc(&c->ekeys, (size_t)n * keybytes + 16);
        if (e == cudaSuccess) e = cudaMalloc((void**)&c->final_idx, (size_t)n * 4 + 16);
        if (e == cudaSuccess) e = cudaMalloc(&c->final_keys, (size_t)n * keybytes + 16);
        if (e == cudaSuccess) e = cudaMalloc((void**)&c->final_dst, (size_t)n * 4 + 16);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(entity keys / draw order)");
    }
    for (int k = 0; k < 2; ++k) {
        // +16 bytes: the emit kernel prefetches keys in 16-byte chunks
        cudaError_t e = cudaMalloc(&c->keys[k], (size_t)n * keybytes + 16);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(sort keys)");
        e = cudaMalloc((void**)&c->idx[k], (size_t)n * 4);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaMalloc(sort idx)");
    }
    c->sort_cap = n;
    c->sort_keybytes = keybytes;
    return TB_OK;
}

tb_status ensure_world(tb_ctx* c, uint64_t n) {
    size_t cap = c->world ? c->world_cap : 0;
    tb_status s = grow(c, &c->world, &cap, (size_t)n, sizeof(WorldRow));
    c->world_cap = cap;
    return s;
}

// ---- hierarchy planning ---------------------------------------------------------------------------

void split_digits(uint32_t nbits, int digit_bits, DigitSplit* ds) {
    uint32_t np = (nbits + digit_bits - 1) / digit_bits;
    if (np == 0) np = 1;
    ds->npasses = np;
    uint32_t base = nbits / np, rem = nbits % np, sh = 0;
    for (uint32_t p = 0; p < (uint32_t)kMaxPasses; ++p) ds->shift[p] = ds->bits[p] = 0;
    for (uint32_t p = 0; p < np; ++p) {
        ds->shift[p] = sh;
        ds->bits[p] = base + (p < rem ? 1u : 0u);
        sh += ds->bits[p];
    }
}

// Stable sort of entity ids by a u32 key (hierarchy level), using the frame's own radix
// machinery. keys are consumed; the sorted ids land in *out (one of c->idx[]).
tb_status sort_ids_by_key(tb_ctx* c, uint32_t* d_keys, uint32_t n, uint32_t nbits, uint32_t** out) {
    DigitSplit ds;
    split_digits(nbits, kMaxDigitBits, &ds);
    tb_status s = ensure_sort_buffers(c, n, 4);
    if (s) return s;
    c->list_kept = c->sorted_valid = false;  // this sort reuses the frame's ping-pong buffers and scratch
    const uint32_t tile = kScatterTile;
    const uint32_t tiles = div_up(n, tile);
    size_t words = kHdrWords;
    size_t counts_off[kMaxPasses], offs_off[kMaxPasses], csum_off[kMaxPasses];
    for (uint32_t p = 0; p < ds.npasses; ++p) {
        counts_off[p] = words;
        words += (size_t)tiles << ds.bits[p];
        offs_off[p] = words;
        words += (size_t)tiles << ds.bits[p];
        csum_off[p] = words;
        words += (size_t)div_up(tiles, 1024) << ds.bits[p];
    }
    s = ensure_scratch(c, words);
    if (s) return s;
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrHist, 0, (kHdrWords - kHdrHist) * 4, c->stream));
    uint32_t* hist = c->scratch + kHdrHist;
    TB_LAUNCH(c, "level_key_hist", 0, key_hist_kernel, div_up(n, kThreads), kThreads, 0, d_keys, n, ds, hist);
    const uint32_t* kin = d_keys;
    const uint32_t* iin = nullptr;
    int cur = 0;
    for (uint32_t p = 0; p < ds.npasses; ++p) {
        uint32_t* counts = c->scratch + counts_off[p];
        uint32_t* offs = c->scratch + offs_off[p];
        TB_LAUNCH(c, "tile_hist", 0, tile_hist_kernel<uint32_t>, std::min<uint32_t>(tiles, 16u * (uint32_t)c->sm_count), kThreads, 0, kin, n,
                  (const unsigned long long*)nullptr,
                  (const uint64_t*)nullptr, (const uint64_t*)nullptr, 0u, tile, ds.shift[p], ds.bits[p], tiles, counts,
                  (const unsigned long long*)nullptr);
        {
            const uint32_t nchunks = div_up(tiles, 1024);
            uint32_t* csum = c->scratch + csum_off[p];
            if (nchunks > 1)
                TB_LAUNCH(c, "tile_chunk_sums", 0, tile_chunk_sums_kernel, dim3(nchunks, 1u << ds.bits[p]), 1024, 0, counts, tiles,
                          nchunks, csum, (const unsigned long long*)nullptr);
            TB_LAUNCH(c, "tile_scan", 0, tile_scan_kernel, dim3(nchunks, 1u << ds.bits[p]), 1024, 0, counts, tiles, hist + p * 256,
                      (const uint32_t*)csum, nchunks, offs, (const unsigned long long*)nullptr);
        }
        ScatterParams<uint32_t> sp{};
        sp.pass.n = n;
        sp.pass.count_ptr = nullptr;
        sp.pass.shift = ds.shift[p];
        sp.pass.bits = ds.bits[p];
        sp.pass.tile_off = offs;
        sp.pass.ntiles = tiles;
        sp.keys_in = kin;
        sp.idx_in = iin;
        sp.mask_t = nullptr;
        sp.mask_s = nullptr;
        sp.keys_out = (uint32_t*)c->keys[cur];
        sp.idx_out = c->idx[cur];
        TB_LAUNCH(c, "level_sort", 0, radix_scatter_kernel<uint32_t>, std::min<uint32_t>(tiles, 3u * (uint32_t)c->sm_count), kThreads,
                  sizeof(ScatterShared<uint32_t>), sp);
        kin = (const uint32_t*)c->keys[cur];
        iin = c->idx[cur];
        cur ^= 1;
    }
    *out = (uint32_t*)iin;
    return TB_OK;
}

tb_status plan_levels(tb_ctx* c) {
    c->levels_dirty = true;  // until this function succeeds
    c->level_epoch++;
    c->nlevels = 1;
    c->level_ranges = true;
    c->level_first.clear();
    c->level_count.clear();
    const uint32_t n = (uint32_t)c->n;
    if (!c->store[TB_PARENT] || n == 0) {
        c->levels_dirty = false;
        return TB_OK;
    }
    // hop / dist ping-pong + sane_parent + list
    if (c->level_cap < n) {
        if (c->level_list) cudaFree(c->level_list);
        if (c->sane_parent) cudaFree(c->sane_parent);
        c->level_list = c->sane_parent = nullptr;
        TB_CUDA(c, cudaMalloc((void**)&c->level_list, (size_t)n * 4));
        TB_CUDA(c, cudaMalloc((void**)&c->sane_parent, (size_t)n * 4));
        c->level_cap = n;
    }
    const size_t stride = ((size_t)n + 3) & ~(size_t)3;  // 16-byte aligned columns: the level sort reads them by TMA
    DeviceTemp tmp_owner;  // freed on every return path
    TB_CUDA(c, tmp_owner.alloc(stride * 4 * 4 + 16));
    uint32_t* tmp = static_cast<uint32_t*>(tmp_owner.p);
    uint32_t* hop[2] = {tmp, tmp + stride};
    uint32_t* dist[2] = {tmp + 2 * stride, tmp + 3 * stride};
    tb_status s = ensure_scratch(c, kHdrWords);
    if (s) {
        return s;
    }
    const uint32_t grid = div_up(n, kThreads);
    TB_LAUNCH(c, "level_init", 0, level_init_kernel, grid, kThreads, 0, c->parent, mask_or_null(c, TB_TRANSFORM),
              mask_or_null(c, TB_PARENT), n, (uint32_t)c->global_first, hop[0], dist[0], c->sane_parent);
    int cur = 0;
    bool done = false;
    for (int round = 0; round < 33 && !done; ++round) {
        cudaMemsetAsync(c->scratch + kHdrPending, 0, 4, c->stream);
        TB_LAUNCH(c, "level_jump", 0, level_jump_kernel, grid, kThreads, 0, hop[cur], dist[cur], n, hop[cur ^ 1],
                  dist[cur ^ 1], c->scratch + kHdrPending);
        cur ^= 1;
        uint32_t pending = 0;
        cudaError_t e = peek_sync(c, &pending, c->scratch + kHdrPending, 4);
        if (e != cudaSuccess) {
            return cuda_fail(c, e, "plan_levels");
        }
        done = pending == 0;
    }
    if (!done) {
        c->levels_dirty = true;
        return fail(c, TB_ERR_CYCLE, "the Parent links contain a cycle");
    }
    uint32_t* level = dist[cur];
    cudaMemsetAsync(c->scratch + kHdrLevelStats, 0, 8, c->stream);
    TB_LAUNCH(c, "level_stats", 0, level_stats_kernel, grid, kThreads, 0, level, mask_or_null(c, TB_TRANSFORM), n,
              c->scratch + kHdrLevelStats);
    uint32_t stats[2] = {0, 0};
    cudaError_t e = peek_sync(c, stats, c->scratch + kHdrLevelStats, 8);
    if (e != cudaSuccess) {
        return cuda_fail(c, e, "plan_levels");
    }
    c->nlevels = stats[0] + 1;
    c->level_ranges = stats[1] == 0;
    if (c->nlevels > 1) {
        if (c->nlevels > 65536) {
            return fail(c, TB_ERR_UNSUPPORTED, "hierarchy deeper than 65536 levels");
        }
        DeviceTemp hist_owner;
        TB_CUDA(c, hist_owner.alloc((size_t)c->nlevels * 4));
        uint32_t* d_hist = static_cast<uint32_t*>(hist_owner.p);
        cudaMemsetAsync(d_hist, 0, (size_t)c->nlevels * 4, c->stream);
        TB_LAUNCH(c, "level_hist", 0, level_hist_kernel, grid, kThreads, 0, level, (const uint64_t*)nullptr, n, d_hist,
                  c->nlevels);
        c->level_count.resize(c->nlevels);
        cudaMemcpyAsync(c->level_count.data(), d_hist, (size_t)c->nlevels * 4, cudaMemcpyDeviceToHost, c->stream);
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return cuda_fail(c, e, "plan_levels");
        c->level_first.resize(c->nlevels);
        uint32_t run = 0;
        for (uint32_t l = 0; l < c->nlevels; ++l) {
            c->level_first[l] = run;
            run += c->level_count[l];
        }
        if (!c->level_ranges) {
            uint32_t nbits = 0;
            while ((1u << nbits) < c->nlevels) ++nbits;
            uint32_t* sorted = nullptr;
            if ((s = sort_ids_by_key(c, level, n, nbits, &sorted))) return s;
            TB_CUDA(c, cudaMemcpyAsync(c->level_list, sorted, (size_t)n * 4, cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    TB_CUDA(c, cudaStreamSynchronize(c->stream));  // the temporaries die with this scope
    c->levels_dirty = false;
    return TB_OK;
}

// Switches the row layout, converting whatever the rows hold (one pass over the storage, out of place).
tb_status relayout_rows(tb_ctx* c, bool interleaved) {
    if ((c->interleaved != 0) == interleaved) return TB_OK;
    if (tb_status sy = ensure_rows_current(c)) return sy;
    RowView dst = c->rows;
    float4* fresh = nullptr;
    if (c->rows_touched) {
        cudaError_t e = cudaMalloc((void**)&fresh, c->cap * 64);
        if (e != cudaSuccess) return cuda_fail(c, e, "relayout_rows");
    } else {
        fresh = c->rows_mem;  // nothing uploaded yet: the rows are all zero in either layout
    }
    dst = make_row_view(fresh, c->cap, interleaved, c->yaw_mem);
    if (c->rows_touched) {
        TB_LAUNCH(c, "relayout_rows", c->cap * 128, relayout_rows_kernel, (uint32_t)div_up(c->cap * 4, (uint64_t)kThreads), kThreads, 0,
                  c->rows, dst, (uint64_t)c->cap);
        const cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            cudaFree(fresh);
            return cuda_fail(c, e, "relayout_rows");
        }
        cudaFree(c->rows_mem);
        c->rows_mem = fresh;
    }
    c->rows = dst;
    c->interleaved = interleaved ? 1 : 0;
    drop_kept_frames(c);
    return TB_OK;
}

// ---- the frame --------------------------------------------------------------------------------

void make_plan(tb_ctx* c, uint64_t key_or, uint64_t key_and, const uint32_t* zvals = nullptr, uint32_t nz = 0) {
    if (nz) build_frame_plan_dict(key_or, key_and, zvals, nz, &c->plan);
    else build_frame_plan(key_or, key_and, &c->plan);
    // optional narrower digits (tests, tb_frame_force_path(TB_PATH_SORT))
    int db = c->digit_bits;
    if (c->force_path == TB_PATH_SORT && c->plan.pext.nbits >= 2) db = std::min<int>(db, (c->plan.pext.nbits + 1) / 2);
    if (db < kMaxDigitBits && c->plan.pext.nbits > 0) {
        DigitSplit ds;
        uint32_t np = (c->plan.pext.nbits + db - 1) / db;
        if (np > (uint32_t)kMaxPasses) db = (c->plan.pext.nbits + kMaxPasses - 1) / kMaxPasses;
        split_digits(c->plan.pext.nbits, db, &ds);
        c->plan.npasses = ds.npasses;
        for (int p = 0; p < kMaxPasses; ++p) {
            c->plan.shift[p] = ds.shift[p];
            c->plan.bits[p] = ds.bits[p];
        }
    }
    c->plan_valid = true;
    drop_kept_frames(c);
}

struct FrameMode {
    bool hier;
    bool lt_alias;    // (layer, texture) bins are digit pass 0 itself
    bool lt_hist;     // separate (layer, texture) histogram
    bool lt_generic;  // run-head detection over an emitted texlayer array
};

tb_status run_prepare(tb_ctx* c, bool have_plan, bool integrate, const FrameMode& fm, const ScratchLayout* L,
                      uint32_t* key_hist = nullptr, uint32_t max_levels = 0xFFFFFFFFu) {
    const uint32_t n = (uint32_t)c->n;
    PrepParams p{};
    p.n = n;
    p.rows = c->rows;
    p.mask_t = mask_or_null(c, TB_TRANSFORM);
    p.mask_s = mask_or_null(c, TB_SPRITE);
    p.mask_v = mask_or_null(c, TB_VELOCITY);
    p.vel = c->store[TB_VELOCITY] ? c->vel : nullptr;
    p.dt = c->pending_dt;
    p.integrate = (integrate && p.vel != nullptr) ? c->pending_ticks : 0u;
    p.have_plan = have_plan;
    p.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    p.hist = c->scratch + kHdrHist;
    p.key_hist = key_hist;
    if (!have_plan && c->zset_active) {
        p.zset = c->zset;
        p.zset_overflow = reinterpret_cast<uint32_t*>(c->zset + kZSetSlots);
    }
    if (have_plan) {
        p.plan = c->plan;
        p.write_keys = c->plan.npasses > 1 || fm.hier || L->chunked;
        p.keys_out = c->ekeys;
        p.check_keys = (c->sorted_valid && p.write_keys) ? 1u : 0u;
        p.lt_mode = (fm.lt_hist && !L->chunked) ? (c->plan.ltbits <= 11 ? 1u : 2u) : 0u;  // chunked: bins come from H
        p.lt_hist = c->scratch + L->lt_off;
        if (L->chunked) {
            p.write_keys = 1u;
            p.chunk_hist = c->scratch + L->H_off;
            p.chunk_shift = c->cg.shift;
        }
        if (!fm.hier) {
            // pass 0 reads the keys in entity order, which is how this pass produces them: it counts
            // pass 0's per-tile digits on the way (no tile_hist launch for that pass)
            static_assert(kWarpTile == 64 && kEmitTile == 512 && kScatterTile == 2 * kThreads * kPrepItems, "tile shifts");
            p.tile_counts = c->scratch + L->counts_off[0];
            p.emit_tiles = L->tiles[0];
            p.emit_tile_shift = L->small ? 6u : ((c->plan.npasses == 1 && !L->chunked) ? 9u : 11u);
        }
    }
    const uint64_t bytes_in = (uint64_t)n * (32 + (p.integrate ? 12 + 16 : 0) + (p.write_keys ? (c->plan.key64 ? 8 : 4) : 0));
    if (!fm.hier) {
        const uint32_t tiles = div_up(n, kThreads * kPrepItems);
        const uint32_t grid = std::min<uint32_t>(tiles, (uint32_t)c->sm_count * 8u);
        TB_LAUNCH(c, have_plan ? "prepare_flat" : "discover_flat", bytes_in, prepare_flat_kernel, grid, kThreads, 0, p);
    } else {
        const uint32_t nl = std::min(c->nlevels, max_levels);
        for (uint32_t l = 0; l < nl; ++l) {
            LevelParams lp{};
            lp.prep = p;
            lp.parent = c->sane_parent;
            lp.list = c->level_ranges ? nullptr : c->level_list;
            lp.first = c->level_first[l];
            lp.count = c->level_count[l];
            lp.level = l;
            lp.global_first = (uint32_t)c->global_first;
            lp.world = c->world;
            if (lp.count == 0) continue;
            // consecutive small levels (the top of the tree): one launch of one CTA walks them in order
            LevelSpan span{};
            uint32_t l2 = l;
            while (l2 < nl && span.n < (uint32_t)kMaxSpanLevels && c->level_count[l2] <= kSpanLevelEntities) {
                if (c->level_count[l2] > 0) {
                    span.first[span.n] = c->level_first[l2];
                    span.count[span.n] = c->level_count[l2];
                    span.level[span.n] = l2;
                    span.n += 1;
                }
                l2 += 1;
            }
            if (span.n >= 2) {
                uint64_t ents = 0;
                for (uint32_t i = 0; i < span.n; ++i) ents += span.count[i];
                TB_LAUNCH(c, have_plan ? "prepare_level" : "discover_level", ents * (64 + 4 + 64 + (p.integrate ? 44 : 0)), prepare_level_span_kernel, 1,
                          kThreads, 0, lp, span);
                l = l2 - 1;
                continue;
            }
            const uint32_t chunks = div_up(lp.count, kThreads);
            const uint32_t grid = std::min<uint32_t>(chunks, (uint32_t)c->sm_count * 8u);
            TB_LAUNCH(c, have_plan ? "prepare_level" : "discover_level", (uint64_t)lp.count * (64 + 4 + 64 + (p.integrate ? 44 : 0)),
                      prepare_level_kernel, grid, kThreads, 0, lp);
        }
        c->world_valid = max_levels >= c->nlevels;
    }
    return TB_OK;
}

tb_status read_header(tb_ctx* c) {
    peek_async(c, c->h_hdr, c->scratch + kHdrFrame, kHdrReadWords * sizeof(uint32_t));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

// The context's camera as kernel parameters.
template <typename Params>
void set_camera(const tb_ctx* c, Params* ep) {
    ep->has_vp = c->has_vp ? 1u : 0u;
    for (int j = 0; j < 4; ++j)
        ep->vp[j] = c->has_vp ? make_float4(c->vp[4 * j], c->vp[4 * j + 1], c->vp[4 * j + 2], c->vp[4 * j + 3])
                              : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

template <typename KeyT>
tb_status run_sort_and_emit(tb_ctx* c, const FrameMode& fm, const ScratchLayout& L, bool prepare_made_counts,
                            float4* d_inst, float4* d_vtx, uint32_t* d_order) {
    const uint32_t n = (uint32_t)c->n;
    const FramePlan& plan = c->plan;
    uint32_t* hist = c->scratch + kHdrHist;
    const unsigned long long* d_count = reinterpret_cast<const unsigned long long*>(c->scratch + kHdrStats) + 2;
    const bool have_keys = plan.npasses > 1 || fm.hier;  // pass A wrote short keys (entity order)
    const KeyT* kin = have_keys ? (const KeyT*)c->ekeys : nullptr;
    // the previous frame's sort is still in place: the sort-pass kernels return at once when the key pass found
    // no changed key (stats[5] == 0)
    const unsigned long long* skip =
        c->sorted_valid ? reinterpret_cast<const unsigned long long*>(c->scratch + kHdrKeysChanged) : nullptr;
    const uint32_t* iin = nullptr;
    const uint64_t kb = sizeof(KeyT);
    int cur = 1;
    for (uint32_t p = 0; p < plan.npasses; ++p) {
        const bool last = p + 1 == plan.npasses;
        const uint32_t tile = L.tile_elems[p];
        PassParams pp{};
        pp.n = n;  // passes after the first only hold the renderables: bounded by the device-side count
        pp.count_ptr = iin ? d_count : nullptr;
        pp.shift = plan.shift[p];
        pp.bits = plan.bits[p];
        pp.ntiles = L.tiles[p];
        {
            uint32_t* counts = c->scratch + L.counts_off[p];
            uint32_t* offs = c->scratch + L.offs_off[p];
            if (!(p == 0 && prepare_made_counts))
                TB_LAUNCH(c, "tile_hist", (uint64_t)n * kb, tile_hist_kernel<KeyT>,
                          std::min<uint32_t>(L.tiles[p], 16u * (uint32_t)c->sm_count), kThreads, 0, kin, n, pp.count_ptr,
                          mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_SPRITE), iin ? 0u : 1u, tile, pp.shift, pp.bits,
                          L.tiles[p], counts, skip);
            const uint32_t nchunks = div_up(L.tiles[p], 1024);
            uint32_t* csum = c->scratch + L.csum_off[p];
            if (nchunks > 1)
                TB_LAUNCH(c, "tile_chunk_sums", 0, tile_chunk_sums_kernel, dim3(nchunks, 1u << pp.bits), 1024, 0, counts, L.tiles[p],
                          nchunks, csum, skip);
            TB_LAUNCH(c, "tile_scan", 0, tile_scan_kernel, dim3(nchunks, 1u << pp.bits), 1024, 0, counts, L.tiles[p], hist + p * 256,
                      (const uint32_t*)csum, nchunks, offs, skip);
            pp.tile_off = offs;
        }
        if (!last) {
            ScatterParams<KeyT> sp{};
            sp.pass = pp;
            sp.keys_in = kin;
            sp.idx_in = iin;
            sp.mask_t = mask_or_null(c, TB_TRANSFORM);
            sp.mask_s = mask_or_null(c, TB_SPRITE);
            sp.keys_out = (KeyT*)c->keys[cur];
            sp.idx_out = c->idx[cur];
            sp.skip = skip;
            TB_LAUNCH(c, "radix_scatter", (uint64_t)n * (kb + (iin ? 4 : 0) + kb + 4), radix_scatter_kernel<KeyT>,
                      std::min<uint32_t>(L.tiles[p], 3u * (uint32_t)c->sm_count), kThreads, sizeof(ScatterShared<KeyT>), sp);
            kin = (const KeyT*)c->keys[cur];
            iin = c->idx[cur];
            cur ^= 1;
        } else {
            EmitParams<KeyT> ep{};
            ep.pass = pp;
            ep.keys_in = kin;
            ep.idx_in = iin;
            ep.mask_t = mask_or_null(c, TB_TRANSFORM);
            ep.mask_s = mask_or_null(c, TB_SPRITE);
            ep.rows = c->rows;
            ep.world = fm.hier ? c->world : nullptr;
            ep.plan = plan;
            ep.instances = d_inst;
            ep.vertices = d_vtx;
            ep.order = d_order;
            ep.texlayer = fm.lt_generic ? c->texlayer : nullptr;
            ep.global_first = (uint32_t)c->global_first;
            ep.staged = c->staged_emit;
            ep.remap = c->shard_state == 4 ? c->remap : nullptr;
            set_camera(c, &ep);
            if (L.small && iin == nullptr) {
                TB_LAUNCH(c, "emit", (uint64_t)n * (64 + 80 + 64 + 4 + (kin ? kb : 0)), (emit_small_kernel<KeyT, false>),
                          std::min<uint32_t>(div_up(L.tiles[p], kWarps), 2u * (uint32_t)c->sm_count), kThreads, sizeof(SmallShared), ep);
            } else {
                // a frame that keeps the previous frame's sort also keeps its draw order: exactly one of the two
                // emit kernels does the work, chosen on the device by the key pass's "changed" flag
                const bool ordered = skip != nullptr && c->final_valid && !fm.lt_generic;
                if (ordered) {
                    ep.changed = skip;
                    ep.run_when = 1u;
                }
                if (c->sorted_cache && iin != nullptr && !fm.lt_generic) {
                    ep.final_idx = c->final_idx;  // whenever this kernel ranks, it leaves the draw order behind
                    ep.final_keys = (KeyT*)c->final_keys;
                }
                TB_LAUNCH(c, "emit", (uint64_t)n * (64 + 80 + 64 + 4 + (kin ? kb : 0) + (iin ? 4 : 0)), emit_kernel<KeyT>,
                          std::min<uint32_t>(L.tiles[p], 2u * (uint32_t)c->sm_count), kThreads, sizeof(EmitShared<KeyT>), ep);
                if (ordered) {
                    EmitParams<KeyT> eo = ep;
                    eo.run_when = 2u;
                    eo.keys_in = nullptr;
                    eo.idx_in = c->final_idx;
                    eo.final_idx = nullptr;
                    eo.final_keys = nullptr;
                    eo.pass.n = n;
                    eo.pass.count_ptr = d_count;
                    TB_LAUNCH(c, "emit_ordered", (uint64_t)n * (64 + 80 + 64 + 4 + 4), (emit_ordered_kernel<KeyT, false>),
                              2u * (uint32_t)c->sm_count, kThreads, sizeof(OrderedShared), eo);
                }
            }
        }
    }
    return TB_OK;
}

// The emit over the kept chunk-major list (chunked.cuh). VALIDATE: flat frames without a key pass.
void fill_list_params(tb_ctx* c, const FrameMode& fm, bool integrate, float4* d_inst, float4* d_vtx, uint32_t* d_order, ListParams* out) {
    ListParams& lp = *out;
    lp = ListParams{};
    lp.ids = c->final_idx;
    lp.dst = c->final_dst;
    lp.keys = (const uint32_t*)c->final_keys;
    lp.chunk_cnt = c->scratch + c->chunk_cnt_off;
    lp.chunk_shift = c->cg.shift;
    lp.npad = c->cg.nchunks << c->cg.shift;
    lp.rows = c->rows;
    lp.world = fm.hier ? c->world : nullptr;
    lp.vel = c->store[TB_VELOCITY] ? c->vel : nullptr;
    lp.mask_v = mask_or_null(c, TB_VELOCITY);
    lp.dt = c->pending_dt;
    lp.integrate = (integrate && lp.vel != nullptr && c->store[TB_TRANSFORM]) ? c->pending_ticks : 0u;
    lp.instances = d_inst;
    lp.vertices = d_vtx;
    lp.order = d_order;
    lp.global_first = (uint32_t)c->global_first;
    lp.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    lp.plan = c->plan;
    lp.l2_hint = (uint32_t)c->chunk_l2_hint;
    set_camera(c, &lp);
}

// Steps 2-4 of a chunked sorted frame (chunked.cuh) after the key pass: batch bins and run starts from H, the
// per-chunk LSD passes (all chunks in one launch; every sort kernel returns at once when the key pass found no changed
// key), then the emit over the chunk-major list.
tb_status run_chunked_sort_and_emit(tb_ctx* c, const FrameMode& fm, const ScratchLayout& L, bool prepare_made_counts,
                                    float4* d_inst, float4* d_vtx, uint32_t* d_order) {
    const uint32_t n = (uint32_t)c->n;
    const FramePlan& plan = c->plan;
    const ChunkGeom g = c->cg;
    const uint32_t K = 1u << g.nbits;
    const uint32_t npad = g.nchunks << g.shift;
    const uint32_t tpc = (1u << g.shift) / kScatterTile;
    const unsigned long long* skip =
        c->sorted_valid ? reinterpret_cast<const unsigned long long*>(c->scratch + kHdrKeysChanged) : nullptr;
    uint32_t* H = c->scratch + L.H_off;
    uint32_t* keytot = c->scratch + L.keytot_off;
    uint32_t* remap = c->scratch + L.remap_off;
    uint32_t* chunk_cnt = c->scratch + L.cnt_off;
    // batch table bins + key starts: recomputed every frame (H is), cheap
    TB_LAUNCH(c, "chunk_key_totals", (uint64_t)g.nchunks * K * 4, chunk_key_totals_kernel, div_up(K, kThreads), kThreads, 0, (const uint32_t*)H, g,
              plan.zbits, keytot, c->scratch + L.lt_off);
    TB_LAUNCH(c, "build_batches", 0, build_batches_kernel, 1, 1024, 0, (const uint32_t*)(c->scratch + L.lt_off), 1u << plan.ltbits, plan,
              c->batches, (uint32_t)c->max_batches, c->scratch + kHdrNumBatches);
    TB_LAUNCH(c, "chunk_key_scan", 0, scan_blocks_kernel, 1, 1024, 0, keytot, K, c->scratch + kHdrPending);
    TB_LAUNCH(c, "chunk_place", (uint64_t)g.nchunks * K * 8, chunk_place_kernel, div_up(K, kThreads), kThreads, 0, (const uint32_t*)H,
              (const uint32_t*)keytot, g, remap, skip);
    TB_LAUNCH(c, "chunk_local", (uint64_t)g.nchunks * K * 12, chunk_local_kernel, g.nchunks, 1024, 0, (const uint32_t*)H, g, remap, chunk_cnt, skip);
    const uint32_t* kin = (const uint32_t*)c->ekeys;
    const uint32_t* iin = nullptr;
    int cur = 1;
    for (uint32_t p = 0; p < plan.npasses; ++p) {
        const bool last = p + 1 == plan.npasses;
        const uint32_t ntiles = L.tiles[p];
        uint32_t* counts = c->scratch + L.counts_off[p];
        uint32_t* offs = c->scratch + L.offs_off[p];
        if (!(p == 0 && prepare_made_counts))
            TB_LAUNCH(c, "tile_hist", (uint64_t)n * 4, tile_hist_kernel<uint32_t>, std::min<uint32_t>(ntiles, 16u * (uint32_t)c->sm_count), kThreads, 0,
                      kin, iin ? npad : n, (const unsigned long long*)nullptr, mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_SPRITE),
                      iin ? 0u : 1u, kScatterTile, plan.shift[p], plan.bits[p], ntiles, counts, skip,
                      iin ? (const uint32_t*)chunk_cnt : (const uint32_t*)nullptr, g.shift);
        TB_LAUNCH(c, "chunk_tile_scan", 0, chunk_tile_scan_kernel, g.nchunks, 1024, 0, (const uint32_t*)counts, ntiles, tpc, plan.bits[p], g.shift, offs,
                  skip);
        ScatterParams<uint32_t> sp{};
        sp.pass.n = iin ? npad : n;
        sp.pass.count_ptr = nullptr;
        sp.pass.shift = plan.shift[p];
        sp.pass.bits = plan.bits[p];
        sp.pass.tile_off = offs;
        sp.pass.ntiles = ntiles;
        sp.pass.chunk_cnt = chunk_cnt;
        sp.pass.chunk_shift = g.shift;
        sp.keys_in = kin;
        sp.idx_in = iin;
        sp.mask_t = mask_or_null(c, TB_TRANSFORM);
        sp.mask_s = mask_or_null(c, TB_SPRITE);
        sp.skip = skip;
        if (last) {
            sp.keys_out = (uint32_t*)c->final_keys;
            sp.idx_out = c->final_idx;
            sp.dst_out = c->final_dst;
            sp.remap = remap;
            sp.key_bits = g.nbits;
            sp.id_offset = (uint32_t)c->global_first;
        } else {
            sp.keys_out = (uint32_t*)c->keys[cur];
            sp.idx_out = c->idx[cur];
        }
        TB_LAUNCH(c, "radix_scatter", (uint64_t)n * (last ? 20 : 16), radix_scatter_kernel<uint32_t>,
                  std::min<uint32_t>(ntiles, 3u * (uint32_t)c->sm_count), kThreads, sizeof(ScatterShared<uint32_t>), sp);
        kin = (const uint32_t*)c->keys[cur];
        iin = c->idx[cur];
        cur ^= 1;
    }
    ListParams lp;
    fill_list_params(c, fm, false, d_inst, d_vtx, d_order, &lp);
    TB_LAUNCH(c, "emit_list", (uint64_t)n * (64 + 80 + 64 + 4 + 8), emit_list_kernel<false>, 2u * (uint32_t)c->sm_count, kThreads, sizeof(ListShared), lp);
    return TB_OK;
}

// The steady-state frame of a flat scene under a chunked sort whose list is kept: one kernel — the velocity system on
// the gathered rows, every key re-checked against the kept one, compose, emit. *ok as in sorted_fast_frame.
tb_status chunk_fast_frame(tb_ctx* c, bool integrate, float4* d_inst, float4* d_vtx, uint32_t* d_order, bool* ok) {
    *ok = false;
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
    FrameMode fm{};
    ListParams lp;
    fill_list_params(c, fm, integrate, d_inst, d_vtx, d_order, &lp);
    if (lp.integrate) tick_hidden_movers(c);  // the list only holds drawn entities
    TB_LAUNCH(c, "emit_list", c->sorted_count * (64 + 80 + 64 + 4 + 12 + (lp.integrate ? 28 : 0)), emit_list_kernel<true>,
              2u * (uint32_t)c->sm_count, kThreads, sizeof(ListShared), lp);
    tb_status s = read_header(c);
    if (s) return s;
    *ok = c->h_hdr[(kHdrKeysChanged - kHdrFrame) / 2] == 0 && c->h_hdr[(kHdrZMiss - kHdrFrame) / 2] == 0 &&
          c->h_hdr[2] == c->sorted_count && plan_covers(c->plan, c->h_hdr[0], ~c->h_hdr[1]);
    return TB_OK;
}

// The steady-state frame of a scene graph whose draw order is kept and whose entities move: the level pass runs for
// every level but the last (velocity system, world rows, key comparison: sequential in entity order), then ONE streaming
// kernel serves the frame from the draw-ordered working storage — the leaves (typically ~90 % of the entities) are ticked
// there, composed as world[parent] * local from the parents' L2-resident world rows, and key-checked; their world rows
// are never written and never re-read. *ok as in sorted_fast_frame.
template <typename KeyT>
tb_status hierarchy_seq_frame(tb_ctx* c, bool integrate, const ScratchLayout& L, float4* d_inst, float4* d_vtx, uint32_t* d_order, bool* ok) {
    *ok = false;
    tb_status s;
    FrameMode fm{};
    fm.hier = true;
    fm.lt_hist = true;
    const uint32_t last = c->nlevels - 1;
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, (L.total_words - kHdrFrame) * 4, c->stream));
    if ((s = run_prepare(c, true, integrate, fm, &L, nullptr, last))) return s;
    const bool tick = integrate && c->store[TB_VELOCITY] && c->store[TB_TRANSFORM];
    if (tick && c->level_count[last] > 0 && c->hidden_movers > 0) {
        // leaves that are not drawn (no Sprite) are not in the working storage: their tick runs on the entity-order rows
        const uint32_t first = c->level_first[last], end = first + c->level_count[last];
        TB_LAUNCH(c, "integrate", (uint64_t)c->level_count[last], integrate_kernel, div_up(end - first, kThreads), kThreads, 0, c->rows,
                  mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_VELOCITY), c->vel, c->pending_dt, c->pending_ticks, end, first,
                  mask_or_null(c, TB_SPRITE));
    }
    EmitParams<KeyT> ep{};
    ep.pass.n = (uint32_t)c->sorted_count;
    ep.idx_in = c->final_idx;
    ep.keys_in = (const KeyT*)c->final_keys;
    ep.seq_rows = c->seq_rows;
    ep.seq_rows_rw = c->seq_rows;
    ep.seq_split = c->seq_split;
    ep.seq_vel = c->seq_has_vel ? c->seq_vel : nullptr;
    ep.seq_par = c->seq_par;
    ep.leaf_first = c->seq_leaf_first;
    ep.world = c->world;
    ep.dt = c->pending_dt;
    ep.integrate = (tick && c->seq_has_vel) ? c->pending_ticks : 0u;
    ep.plan = c->plan;
    ep.instances = d_inst;
    ep.vertices = d_vtx;
    ep.order = d_order;
    ep.global_first = (uint32_t)c->global_first;
    ep.staged = c->staged_emit;
    ep.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    set_camera(c, &ep);
    if (ep.integrate) c->seq_live = true;
    TB_LAUNCH(c, "emit_sequential", c->sorted_count * (64 + 80 + 64 + 4 + 4 + 4 + sizeof(KeyT) + (tick ? 28 : 0)),
              (emit_sequential_kernel<KeyT, true, true>), 2u * (uint32_t)c->sm_count, kThreads, sizeof(SeqShared), ep);
    if ((s = read_header(c))) return s;
    *ok = c->h_hdr[(kHdrKeysChanged - kHdrFrame) / 2] == 0 && c->h_hdr[(kHdrZMiss - kHdrFrame) / 2] == 0 &&
          c->h_hdr[2] == c->sorted_count && plan_covers(c->plan, c->h_hdr[0], ~c->h_hdr[1]);
    c->world_valid = false;  // the leaves' world rows were not written
    return TB_OK;
}

// A scene graph nothing has touched since the frame that left its world rows and list: the emit over those.
tb_status chunk_replay_hierarchy(tb_ctx* c, float4* d_inst, float4* d_vtx, uint32_t* d_order) {
    FrameMode fm{};
    fm.hier = true;
    ListParams lp;
    fill_list_params(c, fm, false, d_inst, d_vtx, d_order, &lp);
    TB_LAUNCH(c, "emit_list", c->sorted_count * (64 + 80 + 64 + 4 + 8), emit_list_kernel<false>, 2u * (uint32_t)c->sm_count, kThreads,
              sizeof(ListShared), lp);
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

// Every launch of a full frame under the cached plan, in stream order, ending with the header copy.
// Launch-only (no allocation, no synchronisation), so it can be captured into a graph.
tb_status record_full_frame(tb_ctx* c, const FrameMode& fm, const ScratchLayout& L, bool integrate, bool prepare_made_counts,
                            float4* d_inst, float4* d_vtx, uint32_t* d_order) {
    const FramePlan& plan = c->plan;
    tb_status s;
    // zero: stats, batch count, histograms, (layer,texture) bins, pass-0 tile counts
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, (L.total_words - kHdrFrame) * 4, c->stream));
    if ((s = run_prepare(c, true, integrate, fm, &L))) return s;
    uint32_t* hist = c->scratch + kHdrHist;
    if (!fm.lt_generic && !L.chunked) {
        const uint32_t* bins = fm.lt_alias ? hist : c->scratch + L.lt_off;
        TB_LAUNCH(c, "build_batches", 0, build_batches_kernel, 1, 1024, 0, bins, 1u << plan.ltbits, plan, c->batches,
                  (uint32_t)c->max_batches, c->scratch + kHdrNumBatches);
    }
    if (L.chunked) s = run_chunked_sort_and_emit(c, fm, L, prepare_made_counts, d_inst, d_vtx, d_order);
    else if (plan.key64) s = run_sort_and_emit<uint64_t>(c, fm, L, prepare_made_counts, d_inst, d_vtx, d_order);
    else s = run_sort_and_emit<uint32_t>(c, fm, L, prepare_made_counts, d_inst, d_vtx, d_order);
    if (s) return s;
    peek_async(c, c->h_hdr, c->scratch + kHdrFrame, kHdrReadWords * sizeof(uint32_t));
    return TB_OK;
}

void frame_signature(tb_ctx* c, const FrameMode& fm, bool integrate, const void* d_inst, const void* d_vtx, const void* d_order,
                     FrameSig* out) {
    FrameSig& g = *out;
    g.n = c->n;
    g.max_batches = c->max_batches;
    g.global_first = c->global_first;
    g.level_epoch = c->level_epoch;
    const void* ptrs[] = {c->rows.t, c->rows.b, c->scratch, c->keys[0], c->keys[1], c->idx[0], c->idx[1], c->world, c->vel,
                          c->mask[0], c->mask[1], c->mask[2], c->mask[3], c->batches, c->texlayer, c->level_list,
                          c->sane_parent, c->remap, d_inst, d_vtx, d_order, c->h_hdr, c->ekeys, c->final_idx, c->final_dst,
                          c->final_keys};
    static_assert(sizeof(ptrs) <= sizeof(g.ptr), "FrameSig::ptr too small");
    memcpy(g.ptr, ptrs, sizeof(ptrs));
    g.plan = c->plan;
    float dt = c->pending_dt;
    uint32_t dt_bits;
    memcpy(&dt_bits, &dt, 4);
    const uint32_t flags[] = {integrate ? c->pending_ticks : 0u, integrate ? dt_bits : 0u, fm.hier ? 1u : 0u, fm.lt_alias ? 1u : 0u,
                              fm.lt_hist ? 1u : 0u, (uint32_t)c->staged_emit, (uint32_t)c->small_emit, c->nlevels,
                              c->level_ranges ? 1u : 0u, (uint32_t)c->shard_state, (uint32_t)c->rows.st,
                              (uint32_t)(c->store[TB_VELOCITY] ? 1 : 0), c->sorted_valid ? 1u : 0u,
                              (c->sorted_valid && c->final_valid) ? 1u : 0u, (uint32_t)c->sorted_cache,
                              (uint32_t)c->static_buckets, c->chunked ? 1u + c->cg.shift : 0u, c->cg.nchunks};
    static_assert(sizeof(flags) == sizeof(g.flags), "FrameSig::flags size");
    memcpy(g.flags, flags, sizeof(flags));
    g.vp_bits[0] = c->has_vp ? 1u : 0u;
    if (c->has_vp) memcpy(g.vp_bits + 1, c->vp, sizeof(c->vp));
}


// Frames served by the draw-ordered copy of the rows (emit_sequential_kernel). validate == false: nothing has touched
// the scene since the frame that validated it (camera moves). validate == true: the copy is the working storage of a
// flat sorted scene — a pending tick runs on it, every key is re-checked; *ok as in sorted_fast_frame.
template <typename KeyT>
tb_status sequential_frame(tb_ctx* c, bool validate, bool integrate, float4* d_inst, float4* d_vtx, uint32_t* d_order, bool* ok) {
    *ok = false;
    EmitParams<KeyT> ep{};
    ep.pass.n = (uint32_t)c->sorted_count;
    ep.idx_in = c->final_idx;
    ep.keys_in = (const KeyT*)c->final_keys;
    ep.seq_rows = c->seq_rows;
    ep.seq_rows_rw = c->seq_rows;
    ep.seq_split = c->seq_split;
    ep.seq_vel = c->seq_has_vel ? c->seq_vel : nullptr;
    ep.world = c->seq_world ? c->world : nullptr;  // only tells the kernel which kind of row it reads
    ep.plan = c->plan;
    ep.instances = d_inst;
    ep.vertices = d_vtx;
    ep.order = d_order;
    ep.global_first = (uint32_t)c->global_first;
    ep.staged = c->staged_emit;
    ep.dt = c->pending_dt;
    ep.integrate = (validate && integrate && c->seq_has_vel) ? c->pending_ticks : 0u;
    ep.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    set_camera(c, &ep);
    if (!validate) {
        TB_LAUNCH(c, "emit_sequential", c->sorted_count * (64 + 80 + 64 + 4 + 4), (emit_sequential_kernel<KeyT, false>),
                  2u * (uint32_t)c->sm_count, kThreads, sizeof(SeqShared), ep);
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        *ok = true;
        return TB_OK;
    }
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
    if (ep.integrate && c->hidden_movers > 0) tick_hidden_movers(c);
    if (ep.integrate) c->seq_live = true;  // the copy now holds newer translations than the entity-order rows
    TB_LAUNCH(c, "emit_sequential", c->sorted_count * (64 + 80 + 64 + 4 + 4 + sizeof(KeyT) + (ep.integrate ? 28 : 0)),
              (emit_sequential_kernel<KeyT, true>), 2u * (uint32_t)c->sm_count, kThreads, sizeof(SeqShared), ep);
    tb_status s = read_header(c);
    if (s) return s;
    *ok = c->h_hdr[(kHdrKeysChanged - kHdrFrame) / 2] == 0 && c->h_hdr[(kHdrZMiss - kHdrFrame) / 2] == 0 &&
          c->h_hdr[2] == c->sorted_count && plan_covers(c->plan, c->h_hdr[0], ~c->h_hdr[1]);
    return TB_OK;
}

// After two validated frames in the kept draw order: copy the rows (world rows for a scene graph) — and, for flat
// scenes, the velocities — into draw order, once.
tb_status build_sequential_rows(tb_ctx* c, bool world, bool graph = false) {
    c->seq_valid = false;
    c->seq_graph = false;
    size_t cap = c->seq_rows ? c->seq_cap : 0;
    if (grow(c, &c->seq_rows, &cap, (size_t)c->sorted_count * 4 + 80, sizeof(float4)) != TB_OK) {
        c->seq_cap = 0;
        cudaGetLastError();
        return TB_OK;  // no memory for the copy: keep gathering
    }
    c->seq_cap = cap;
    const bool with_vel = !world && c->store[TB_VELOCITY] && c->vel != nullptr;
    // local rows are kept as a dense translation block + the rest (see EmitParams::seq_split); 16-byte TMA alignment of
    // every tile of both blocks holds for any float4 index
    // ... for scenes that do not fit the L2 (the dense write-back saves DRAM traffic); smaller ones keep 64-byte rows, which
    // measured faster while everything is L2-resident (1 M sprites: 0.060 vs 0.068 ms)
    static const int split_env = getenv("TB_SEQ_SPLIT") ? atoi(getenv("TB_SEQ_SPLIT")) : -1;
    const bool want_split = split_env >= 0 ? split_env != 0 : c->sorted_count * 64ull > (96ull << 20);
    const uint32_t split = (world || !want_split) ? 0u : (uint32_t)((c->sorted_count + 63) & ~(uint64_t)63);
    if (graph) {
        size_t pcap = c->seq_par ? c->seq_par_cap : 0;
        if (grow(c, &c->seq_par, &pcap, (size_t)c->sorted_count + 4, sizeof(uint32_t)) != TB_OK) {
            c->seq_par_cap = 0;
            cudaGetLastError();
            return TB_OK;
        }
        c->seq_par_cap = pcap;
    }
    if (with_vel) {
        size_t vcap = c->seq_vel ? c->seq_vel_cap : 0;
        if (grow(c, &c->seq_vel, &vcap, (size_t)c->sorted_count * 3 + 4, sizeof(float)) != TB_OK) {
            c->seq_vel_cap = 0;
            cudaGetLastError();
            return TB_OK;
        }
        c->seq_vel_cap = vcap;
    }
    TB_LAUNCH(c, "gather_rows", c->sorted_count * (64 + 64 + 4 + (with_vel ? 24 : 0)), gather_rows_kernel, div_up(c->sorted_count * 4, kThreads),
              kThreads, 0, c->rows, world ? c->world : (const WorldRow*)nullptr, (const uint32_t*)c->final_idx, (uint32_t)c->global_first,
              (uint32_t)c->sorted_count, c->seq_rows, with_vel ? (const float*)c->vel : (const float*)nullptr, mask_or_null(c, TB_VELOCITY),
              with_vel ? c->seq_vel : (float*)nullptr, graph ? (const uint32_t*)c->sane_parent : (const uint32_t*)nullptr,
              graph ? c->seq_par : (uint32_t*)nullptr, split);
    c->seq_split = split;
    c->hidden_movers = 0;
    if (with_vel) {  // scene graphs: counted over the whole scene (a superset of the undrawn leaves)
        unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(c->scratch + kHdrLevelStats);
        TB_CUDA(c, cudaMemsetAsync(d_cnt, 0, 8, c->stream));
        const uint32_t nwords = div_up(c->n, 64);
        TB_LAUNCH(c, "count_hidden_movers", nwords * 24ull, count_hidden_movers_kernel, div_up(nwords, kThreads), kThreads, 0,
                  mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_VELOCITY), mask_or_null(c, TB_SPRITE), nwords, (uint32_t)c->n, d_cnt);
        unsigned long long h = 0;
        TB_CUDA(c, peek_sync(c, &h, d_cnt, 8));
        c->hidden_movers = h;
    }
    c->seq_valid = true;
    c->seq_graph = graph;
    if (graph) c->seq_leaf_first = c->level_first[c->nlevels - 1];
    c->seq_world = world;
    c->seq_has_vel = with_vel;
    c->seq_live = false;
    return TB_OK;
}

// The steady-state frame of a flat SORTED scene whose previous sort and draw order are still in place: the
// velocity system as its own streaming pass (when a tick is pending), then one kernel that gathers, composes
// and writes in the kept draw order while recomputing every key and comparing it with the kept one. *ok tells
// whether the frame's own statistics confirm the kept plan, keys and order; if not the caller redoes the
// frame the long way (the velocity system has run exactly once either way).
template <typename KeyT>
tb_status sorted_fast_frame(tb_ctx* c, bool integrate, float4* d_inst, float4* d_vtx, uint32_t* d_order, bool* ok) {
    *ok = false;
    const FramePlan& plan = c->plan;
    if (integrate && c->store[TB_VELOCITY] && c->store[TB_TRANSFORM])
        TB_LAUNCH(c, "integrate", c->n * 44ull, integrate_kernel, div_up(c->n, kThreads), kThreads, 0, c->rows,
                  mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_VELOCITY), c->vel, c->pending_dt, c->pending_ticks, (uint32_t)c->n);
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
    EmitParams<KeyT> ep{};
    ep.pass.n = (uint32_t)c->sorted_count;
    ep.idx_in = c->final_idx;
    ep.keys_in = (const KeyT*)c->final_keys;
    ep.rows = c->rows;
    ep.plan = plan;
    ep.instances = d_inst;
    ep.vertices = d_vtx;
    ep.order = d_order;
    ep.global_first = (uint32_t)c->global_first;
    ep.staged = c->staged_emit;
    ep.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    set_camera(c, &ep);
    TB_LAUNCH(c, "emit_ordered", c->sorted_count * (64 + 80 + 64 + 4 + 4 + sizeof(KeyT)), (emit_ordered_kernel<KeyT, true>),
              2u * (uint32_t)c->sm_count, kThreads, sizeof(OrderedShared), ep);
    tb_status s = read_header(c);
    if (s) return s;
    *ok = c->h_hdr[(kHdrKeysChanged - kHdrFrame) / 2] == 0 && c->h_hdr[(kHdrZMiss - kHdrFrame) / 2] == 0 &&
          c->h_hdr[2] == c->sorted_count && plan_covers(plan, c->h_hdr[0], ~c->h_hdr[1]);
    return TB_OK;
}

// A hierarchy frame whose inputs are those of the previous frame: world rows, sort and draw order are all
// in place, so the frame is emit_ordered_kernel over the world rows.
template <typename KeyT>
tb_status replay_hierarchy_frame(tb_ctx* c, float4* d_inst, float4* d_vtx, uint32_t* d_order) {
    EmitParams<KeyT> ep{};
    ep.pass.n = (uint32_t)c->sorted_count;
    ep.idx_in = c->final_idx;
    ep.rows = c->rows;
    ep.world = c->world;
    ep.plan = c->plan;
    ep.instances = d_inst;
    ep.vertices = d_vtx;
    ep.order = d_order;
    ep.global_first = (uint32_t)c->global_first;
    ep.staged = c->staged_emit;
    set_camera(c, &ep);
    TB_LAUNCH(c, "emit_ordered", c->sorted_count * (64 + 80 + 64 + 4 + 4), (emit_ordered_kernel<KeyT, false>),
              2u * (uint32_t)c->sm_count, kThreads, sizeof(OrderedShared), ep);
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

// The steady-state frame of a flat few-bucket scene whose buckets depend on the Sprite only: one
// kernel (velocity system + key statistics + rank + compose + emit) placed by the tile offsets the
// last full frame left in the scratch buffer. Returns true in *ok when the frame's own statistics
// confirm the cached plan; otherwise the caller falls back to the full frame (the velocity system
// has already run exactly once either way).
// List frames (tb_frame_sharded): `remap` shifts the shard's bucket positions to the global list's; with `flags` the
// tiles go out in `nchunks` launches, each followed by this rank's per-bucket completion flags (list_signal_buckets_kernel).
tb_status fast_frame(tb_ctx* c, bool integrate, float4* d_inst, float4* d_vtx, uint32_t* d_order, bool* ok, const uint32_t* remap = nullptr,
                     uint32_t nchunks = 1, unsigned long long* flags = nullptr, uint32_t seq = 0, uint32_t nslots = 0) {
    *ok = false;
    const uint32_t n = (uint32_t)c->n;
    const FramePlan& plan = c->plan;
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrStats, 0, 6 * 4, c->stream));
    EmitParams<uint32_t> ep{};
    ep.pass.n = n;
    ep.pass.shift = plan.shift[0];
    ep.pass.bits = plan.bits[0];
    ep.pass.tile_off = c->scratch + c->static_offs_off;
    ep.pass.ntiles = c->static_tiles;
    ep.mask_t = mask_or_null(c, TB_TRANSFORM);
    ep.mask_s = mask_or_null(c, TB_SPRITE);
    ep.rows = c->rows;
    ep.plan = plan;
    ep.instances = d_inst;
    ep.vertices = d_vtx;
    ep.order = d_order;
    ep.global_first = (uint32_t)c->global_first;
    ep.staged = c->staged_emit;
    ep.remap = remap;  // list frames: shard-local bucket positions -> positions of the global list
    if (d_inst == nullptr) {
        ep.world_rows_out = 1u;
        ep.staged = 1u;
    }
    ep.vel = c->store[TB_VELOCITY] ? c->vel : nullptr;
    ep.mask_v = mask_or_null(c, TB_VELOCITY);
    ep.dt = c->pending_dt;
    ep.integrate = (integrate && ep.vel != nullptr) ? c->pending_ticks : 0u;
    ep.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    set_camera(c, &ep);
    const uint32_t per = std::max<uint32_t>(div_up(c->static_tiles, nchunks), 1u);
    for (uint32_t k = 0; k < nchunks; ++k) {
        const uint32_t t0 = std::min(c->static_tiles, k * per), t1 = k + 1 == nchunks ? c->static_tiles : std::min(c->static_tiles, (k + 1) * per);
        if (t0 < t1) {
            ep.tile_first = t0;
            ep.tile_end = nchunks == 1 ? 0u : t1;
            const uint32_t grid = std::min<uint32_t>(div_up(t1 - t0, kWarps), 2u * (uint32_t)c->sm_count);
            const uint64_t bytes = (uint64_t)(t1 - t0) * 64 * (64 + (d_inst ? 80 : 0) + (d_vtx ? 64 : 0) + 4 + (ep.integrate ? 12 + 16 : 0));
            if (ep.world_rows_out)
                TB_LAUNCH(c, "emit", bytes, (emit_small_kernel<uint32_t, true, true>), grid, kThreads, sizeof(SmallShared), ep);
            else
                TB_LAUNCH(c, "emit", bytes, (emit_small_kernel<uint32_t, true>), grid, kThreads, sizeof(SmallShared), ep);
        }
        if (flags != nullptr)
            TB_LAUNCH(c, "list_signal", 8ull * nslots, list_signal_buckets_kernel, 1, nslots, 0, flags + (size_t)k * kListSlots, seq,
                      (const uint32_t*)ep.pass.tile_off, c->static_tiles, t1, remap, (const uint32_t*)c->shard_hist);
    }
    tb_status s = read_header(c);
    if (s) return s;
    *ok = c->h_hdr[2] == c->static_count && plan_covers(plan, c->h_hdr[0], ~c->h_hdr[1]);
    return TB_OK;
}

struct MathArg {
    const void* in;   // host input (nullptr: this slot is an output)
    void* out;        // host output
    size_t elem;      // bytes per element
};
// Runs `launch(device pointers, first element, count)` over n elements in chunks: inputs host -> device, outputs
// device -> host, all on the context's stream.
template <typename Launch>
tb_status math_batch(tb_ctx* c, uint64_t n, const MathArg* args, int nargs, Launch launch) {
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    for (int k = 0; k < nargs; ++k)
        if (!args[k].in && !args[k].out) return fail(c, TB_ERR_INVALID, "null pointer");
    const uint64_t per = 1u << 20;
    const uint64_t m0 = std::min<uint64_t>(n, per);
    DeviceTemp mem[8];
    void* d[8] = {};
    for (int k = 0; k < nargs; ++k) {
        TB_CUDA(c, mem[k].alloc(m0 * args[k].elem));
        d[k] = mem[k].p;
    }
    for (uint64_t off = 0; off < n; off += per) {
        const uint64_t m = std::min<uint64_t>(per, n - off);
        for (int k = 0; k < nargs; ++k)
            if (args[k].in)
                TB_CUDA(c, cudaMemcpyAsync(d[k], (const char*)args[k].in + off * args[k].elem, m * args[k].elem, cudaMemcpyHostToDevice, c->stream));
        launch(d, (uint32_t)m);
        for (int k = 0; k < nargs; ++k)
            if (!args[k].in)
                TB_CUDA(c, cudaMemcpyAsync((char*)args[k].out + off * args[k].elem, d[k], m * args[k].elem, cudaMemcpyDeviceToHost, c->stream));
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "tuber-math batch");
    return TB_OK;
}
}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================

extern "C" {

static tb_status shard_frame(tb_ctx* c, tb_frame_out* out);
static tb_status shard_finalize_impl(tb_ctx* c, bool collective);

const char* tb_version(void) { return "tuber_b200 0.1 (sm_100a)"; }

const char* tb_status_string(tb_status s) {
    switch (s) {
        case TB_OK: return "ok";
        case TB_ERR_INVALID: return "invalid argument";
        case TB_ERR_CUDA: return "CUDA error";
        case TB_ERR_OOM: return "out of memory";
        case TB_ERR_RANGE: return "component value out of range";
        case TB_ERR_NO_DEVICE: return "no sm_100 device (there is no CPU fallback)";
        case TB_ERR_CAPACITY: return "capacity exceeded";
        case TB_ERR_CYCLE: return "cycle in Parent links";
        case TB_ERR_COMM: return "communication failure";
        case TB_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

tb_status tb_ctx_create(int device, uint64_t max_entities, tb_ctx** out) {
    if (!out) return TB_ERR_INVALID;
    *out = nullptr;
    if (max_entities == 0 || max_entities >= (1ull << 30)) return TB_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return TB_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TB_ERR_NO_DEVICE;
    if (prop.major != 10) return TB_ERR_NO_DEVICE;  // the only code in this library is sm_100a
    DeviceGuard device_guard__(device);  // the caller's current device is restored on return
    {
        int now = -1;
        if (cudaGetDevice(&now) != cudaSuccess || now != device) return TB_ERR_NO_DEVICE;
    }
    tb_ctx* c = new (std::nothrow) tb_ctx();
    if (!c) return TB_ERR_OOM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cap = max_entities;
    const char* env = getenv("TB_INTERLEAVED_ROWS");
    if (env) {
        c->interleaved = atoi(env);
        c->layout_auto = false;
    }
    env = getenv("TB_STAGED_EMIT");
    if (env) c->staged_emit = atoi(env);
    cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return TB_ERR_CUDA;
    }
    c->stream = c->own_stream;
    {
        // the reference's sin / cos are the host libm's; glibc picks its FMA build exactly when the CPU has FMA and
        // AVX2 (sysdeps/x86_64/fpu/multiarch/ifunc-fma.h) and the device evaluates the same build (libm_sincosf.h)
        __builtin_cpu_init();
        int fma = (__builtin_cpu_supports("fma") && __builtin_cpu_supports("avx2")) ? 1 : 0;
        if (const char* v = getenv("TB_LIBM_FMA")) fma = atoi(v) != 0;
        c->libm_fma = fma;
        if (cudaMemcpyToSymbol(c_libm_fma, &fma, sizeof(int)) != cudaSuccess) {
            tb_ctx_destroy(c);
            return TB_ERR_CUDA;
        }
    }
    const size_t words = (max_entities + 63) / 64;
    bool ok = cudaMalloc((void**)&c->rows_mem, max_entities * 64) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&c->yaw_mem, max_entities * 4) == cudaSuccess;
    if (ok) cudaMemsetAsync(c->yaw_mem, 0, max_entities * 4, c->stream);
    for (int k = 0; ok && k < TB_COMPONENT_COUNT; ++k) {
        // +2 words: the emit kernel copies presence words two at a time (16-byte cp.async)
        ok = cudaMalloc((void**)&c->mask[k], (words + 2) * 8) == cudaSuccess;
        if (ok) cudaMemsetAsync(c->mask[k], 0, (words + 2) * 8, c->stream);
    }
    ok = ok && cudaMalloc((void**)&c->gen, max_entities * 4) == cudaSuccess;
    if (ok) cudaMemsetAsync(c->gen, 0, max_entities * 4, c->stream);
    ok = ok && cudaMallocHost((void**)&c->h_hdr, 64) == cudaSuccess;
    ok = ok && cudaMallocHost(&c->h_peek, kPeekBytes) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->copied, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&c->zset, (kZSetSlots + 1) * 8) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&c->h_zset, (kZSetSlots + 1) * 8) == cudaSuccess;
    if (ok) ok = ensure_scratch(c, kHdrWords) == TB_OK;
    if (!ok) {
        cudaGetLastError();
        tb_ctx_destroy(c);
        return TB_ERR_OOM;
    }
    cudaMemsetAsync(c->rows_mem, 0, max_entities * 64, c->stream);
    c->rows = make_row_view(c->rows_mem, max_entities, c->interleaved != 0, c->yaw_mem);
    cudaFuncSetAttribute(emit_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitShared<uint32_t>));
    cudaFuncSetAttribute(emit_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitShared<uint64_t>));
    cudaFuncSetAttribute(radix_scatter_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterShared<uint32_t>));
    cudaFuncSetAttribute(radix_scatter_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterShared<uint64_t>));
    cudaFuncSetAttribute(emit_sequential_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint32_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint64_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint32_t, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_sequential_kernel<uint64_t, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SeqShared));
    cudaFuncSetAttribute(emit_ordered_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OrderedShared));
    cudaFuncSetAttribute(emit_ordered_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OrderedShared));
    cudaFuncSetAttribute(emit_ordered_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OrderedShared));
    cudaFuncSetAttribute(emit_ordered_kernel<uint64_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OrderedShared));
    cudaFuncSetAttribute(emit_small_kernel<uint32_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
    cudaFuncSetAttribute(emit_small_kernel<uint64_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
    cudaFuncSetAttribute(emit_small_kernel<uint32_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
    cudaFuncSetAttribute(emit_small_kernel<uint32_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmallShared));
    cudaFuncSetAttribute(emit_list_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ListShared));
    cudaFuncSetAttribute(emit_list_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ListShared));
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
        tb_ctx_destroy(c);
        return TB_ERR_CUDA;
    }
    *out = c;
    return TB_OK;
}

void tb_ctx_destroy(tb_ctx* c) {
    if (!c) return;
    TB_ENTER(c);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    shard_finalize_impl(c, false);
    prof_clear(c);
    cudaFree(c->rows_mem);
    cudaFree(c->yaw_mem);
    for (int k = 0; k < TB_COMPONENT_COUNT; ++k) cudaFree(c->mask[k]);
    cudaFree(c->vel);
    cudaFree(c->spin);
    cudaFree(c->gen);
    for (int k = 0; k < TB_SYSTEM_COUNT; ++k) {
        if (c->side_stream[k]) cudaStreamDestroy(c->side_stream[k]);
        if (c->join_event[k]) cudaEventDestroy(c->join_event[k]);
    }
    if (c->fork_event) cudaEventDestroy(c->fork_event);
    if (c->bundle_graph) cudaGraphExecDestroy(c->bundle_graph);
    cudaFree(c->parent);
    cudaFree(c->stage);
    cudaFree(c->scratch);
    for (int k = 0; k < 2; ++k) {
        cudaFree(c->keys[k]);
        cudaFree(c->idx[k]);
    }
    cudaFree(c->ekeys);
    cudaFree(c->final_idx);
    cudaFree(c->final_keys);
    cudaFree(c->final_dst);
    cudaFree(c->seq_rows);
    cudaFree(c->seq_vel);
    cudaFree(c->seq_par);
    cudaFree(c->texlayer);
    cudaFree(c->world);
    cudaFree(c->inst);
    cudaFree(c->vtx);
    cudaFree(c->order);
    cudaFree(c->batches);
    cudaFree(c->level_list);
    cudaFree(c->sane_parent);
    cudaFree(c->remap);
    cudaFree(c->lt_global);
    cudaFree(c->place_tmp);
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    if (c->h_peek) cudaFreeHost(c->h_peek);
    if (c->copied) cudaEventDestroy(c->copied);
    if (c->frame_graph) cudaGraphExecDestroy(c->frame_graph);
    cudaFree(c->zset);
    if (c->h_zset) cudaFreeHost(c->h_zset);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* tb_last_error(tb_ctx* c) { return c ? c->err.c_str() : "null context"; }

tb_status tb_ctx_set_stream(tb_ctx* c, void* cuda_stream) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    cudaStreamSynchronize(c->stream);
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return TB_OK;
}

tb_status tb_ctx_synchronize(tb_ctx* c) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

tb_status tb_ctx_set_max_batches(tb_ctx* c, uint64_t max_batches) {
    if (!c || max_batches == 0 || max_batches >= (1ull << 30)) return TB_ERR_INVALID;
    c->max_batches = max_batches;
    return TB_OK;
}

tb_status tb_ctx_set_option(tb_ctx* c, const char* name, int64_t value) {
    TB_ENTER(c);
    if (!c || !name) return TB_ERR_INVALID;
    if (tb_status sy = ensure_rows_current(c)) return sy;
    std::string k(name);
    if (k == "interleaved_rows") {
        c->layout_auto = value < 0;
        if (value >= 0) {
            tb_status s = relayout_rows(c, value != 0);
            if (s) return s;
        }
    } else if (k == "staged_emit") {
        c->staged_emit = value != 0;
    } else if (k == "static_buckets") {
        c->static_buckets = value != 0;
        drop_kept_frames(c);
    } else if (k == "small_emit") {
        drop_kept_frames(c);
        c->small_emit = value != 0;
    } else if (k == "digit_bits") {
        if (value < 1 || value > kMaxDigitBits) return fail(c, TB_ERR_INVALID, "digit_bits must be 1..8");
        c->digit_bits = (int)value;
        c->plan_valid = false;
        drop_kept_frames(c);
    } else if (k == "sorted_cache") {
        c->sorted_cache = value != 0;
        c->sorted_valid = c->list_kept = c->final_valid = c->seq_valid = false;
    } else if (k == "chunked_sort") {
        c->chunked_sort = value != 0;
        c->plan_valid = false;
        drop_kept_frames(c);
        c->list_kept = c->final_valid = c->seq_valid = false;
    } else if (k == "hier_fused") {
        c->hier_fused = value != 0;
    } else if (k == "chunk_l2_hint") {
        c->chunk_l2_hint = value != 0;
    } else if (k == "chunk_shift") {
        if (value != 0 && (value < 11 || value > 24)) return fail(c, TB_ERR_INVALID, "chunk_shift must be 0 (automatic) or 11..24");
        c->chunk_shift_opt = (int)value;
        c->plan_valid = false;
        drop_kept_frames(c);
        c->list_kept = c->final_valid = c->seq_valid = false;
    } else if (k == "cuda_graphs") {
        c->use_graphs = value != 0;
        c->frame_graph_valid = false;
    } else if (k == "z_dictionary") {
        c->z_dictionary = value != 0;
        c->plan_valid = false;
        drop_kept_frames(c);
    } else if (k == "lt_limit") {
        if (value < 0 || value > (int64_t)kMaxLtBits) return fail(c, TB_ERR_INVALID, "lt_limit must be 0..16");
        c->lt_limit = (int)value;
        c->plan_valid = false;
        drop_kept_frames(c);
    } else {
        return fail(c, TB_ERR_INVALID, "unknown option " + k);
    }
    // a captured frame graph bakes in what the options selected (which kernels, which optional outputs)
    c->frame_graph_valid = c->frame_graph_seen = false;
    return TB_OK;
}

tb_status tb_host_alloc(uint64_t bytes, void** out) {
    if (!out || bytes == 0) return TB_ERR_INVALID;
    *out = nullptr;
    if (cudaMallocHost(out, bytes) != cudaSuccess) {
        cudaGetLastError();
        return TB_ERR_OOM;
    }
    return TB_OK;
}

void tb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ---- storage ------------------------------------------------------------------------------------

tb_status tb_set_entity_count(tb_ctx* c, uint64_t n) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (n > c->cap) return fail(c, TB_ERR_CAPACITY, "entity_count beyond max_entities");
    tb_status s = flush_integrate(c);
    if (s) return s;
    if (n < c->n) {
        // shrinking drops the components of the removed indices
        for (int k = 0; k < TB_COMPONENT_COUNT; ++k) mark_presence(c, k, n, c->n - n, 0);
    }
    c->n = n;
    c->plan_valid = false;
    drop_kept_frames(c);
    c->levels_dirty = true;
    c->world_valid = false;
    c->frame_valid = false;
    return TB_OK;
}

tb_status tb_entity_count(tb_ctx* c, uint64_t* n) {
    if (!c || !n) return TB_ERR_INVALID;
    *n = c->n;
    return TB_OK;
}

static tb_status upload_chunks(tb_ctx* c, const void* src, uint64_t n, size_t elem,
                               void (*launch)(tb_ctx*, const void*, uint64_t, uint64_t, void*), uint64_t first, void* user) {
    const uint64_t per = kStageChunk / elem;
    tb_status s = ensure_stage(c, (size_t)std::min<uint64_t>(n, per) * elem);
    if (s) return s;
    LinkTurn turn(c, LinkTurn::kUp, n * elem);
    for (uint64_t off = 0; off < n; off += per) {
        const uint64_t m = std::min<uint64_t>(per, n - off);
        TB_CUDA(c, cudaMemcpyAsync(c->stage, (const char*)src + off * elem, m * elem, cudaMemcpyHostToDevice, c->stream));
        if (off + per >= n) TB_CUDA(c, finish_upload_turn(c, turn));
        launch(c, c->stage, first + off, m, user);
        // the staging buffer is reused by the next chunk: same stream, so ordering is implicit
    }
    return TB_OK;
}

tb_status tb_upload_transforms(tb_ctx* c, uint64_t first, uint64_t n, const tb_transform* src) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, src);
    if (s || n == 0) return s;
    if ((s = flush_integrate(c))) return s;
    c->rows_touched = true;
    s = upload_chunks(
        c, src, n, sizeof(tb_transform),
        [](tb_ctx* c, const void* d, uint64_t f, uint64_t m, void*) {
            TB_LAUNCH(c, "convert_transforms", m * 96, convert_transforms_kernel, div_up(m, kThreads), kThreads, 0,
                      (const float*)d, c->rows, (uint32_t)f, (uint32_t)m);
        },
        first, nullptr);
    if (s) return s;
    c->store[TB_TRANSFORM] = true;
    mark_presence(c, TB_TRANSFORM, first, n, 1);
    // Component values changed. The cached plan stays (every frame validates it against its own statistics) and
    // so does the kept sort (the key pass compares every entity's key with the kept one); the frames that run
    // WITHOUT a key pass need one clean prepare-based frame first.
    c->static_valid = false;
    c->fast_clean = c->seq_valid = false;
    c->levels_dirty = true;
    c->world_valid = false;
    c->frame_valid = false;
    return TB_OK;
}

tb_status tb_upload_sprites(tb_ctx* c, uint64_t first, uint64_t n, const tb_sprite* src) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, src);
    if (s || n == 0) return s;
    if ((s = ensure_rows_current(c))) return s;
    c->rows_touched = true;
    s = upload_chunks(
        c, src, n, sizeof(tb_sprite),
        [](tb_ctx* c, const void* d, uint64_t f, uint64_t m, void*) {
            TB_LAUNCH(c, "convert_sprites", m * 32, convert_sprites_kernel, div_up(m, kThreads), kThreads, 0,
                      (const uint32_t*)d, c->rows, (uint32_t)f, (uint32_t)m, c->scratch + kHdrRangeErr);
        },
        first, nullptr);
    if (s) return s;
    c->store[TB_SPRITE] = true;
    mark_presence(c, TB_SPRITE, first, n, 1);
    c->static_valid = false;  // as for transforms: plan and kept sort are validated by the next frame's key pass
    c->fast_clean = c->seq_valid = false;
    c->world_valid = false;
    c->frame_valid = false;
    // range check of layer / texture (SPEC: layer <= 65534, texture <= 65535)
    uint32_t bad = 0;
    TB_CUDA(c, peek_sync(c, &bad, c->scratch + kHdrRangeErr, 4));
    if (bad) {
        cudaMemsetAsync(c->scratch + kHdrRangeErr, 0, 4, c->stream);
        return fail(c, TB_ERR_RANGE, "sprite layer > 65534 or texture > 65535");
    }
    return TB_OK;
}

tb_status tb_upload_velocities(tb_ctx* c, uint64_t first, uint64_t n, const tb_velocity* src) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, src);
    if (s || n == 0) return s;
    if ((s = flush_integrate(c))) return s;
    if (!c->vel) {
        // +16 bytes: tiles of velocities are fetched in 16-byte multiples
        TB_CUDA(c, cudaMalloc((void**)&c->vel, c->cap * sizeof(tb_velocity) + 16));
        TB_CUDA(c, cudaMemsetAsync(c->vel, 0, c->cap * sizeof(tb_velocity) + 16, c->stream));
    }
    {
        LinkTurn turn(c, LinkTurn::kUp, n * sizeof(tb_velocity));
        TB_CUDA(c, cudaMemcpyAsync(c->vel + first * 3, src, n * sizeof(tb_velocity), cudaMemcpyHostToDevice, c->stream));
        TB_CUDA(c, finish_upload_turn(c, turn));
    }
    c->store[TB_VELOCITY] = true;
    mark_presence(c, TB_VELOCITY, first, n, 1);
    c->seq_valid = false;  // the draw-ordered copy carries the velocities
    return TB_OK;
}

tb_status tb_upload_parents(tb_ctx* c, uint64_t first, uint64_t n, const uint32_t* src) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, src);
    if (s || n == 0) return s;
    if (!c->parent) {
        TB_CUDA(c, cudaMalloc((void**)&c->parent, c->cap * 4));
        TB_CUDA(c, cudaMemsetAsync(c->parent, 0xFF, c->cap * 4, c->stream));
    }
    {
        LinkTurn turn(c, LinkTurn::kUp, n * 4);
        TB_CUDA(c, cudaMemcpyAsync(c->parent + first, src, n * 4, cudaMemcpyHostToDevice, c->stream));
        TB_CUDA(c, finish_upload_turn(c, turn));
    }
    c->store[TB_PARENT] = true;
    // presence == "has a Parent component": every uploaded index whose value is not TB_NO_PARENT
    {
        const uint64_t nw = ((first + n - 1) >> 6) - (first >> 6) + 1;
        TB_LAUNCH(c, "mark_parent_presence", n * 4, mark_parent_presence_kernel, div_up(nw, kThreads), kThreads, 0,
                  c->mask[TB_PARENT], c->parent, first, n);
    }
    c->levels_dirty = true;
    c->plan_valid = false;
    drop_kept_frames(c);
    c->world_valid = false;
    c->frame_valid = false;
    return TB_OK;
}

tb_status tb_upload_presence(tb_ctx* c, int comp, uint64_t first_word, uint64_t nwords, const uint64_t* words) {
    TB_ENTER(c);
    if (!c || comp < 0 || comp >= TB_COMPONENT_COUNT) return TB_ERR_INVALID;
    if (nwords == 0) return TB_OK;
    if (!words) return fail(c, TB_ERR_INVALID, "null pointer");
    if (first_word + nwords > (c->cap + 63) / 64) return fail(c, TB_ERR_INVALID, "presence words outside max_entities");
    tb_status s = flush_integrate(c);
    if (s) return s;
    TB_CUDA(c, cudaMemcpyAsync(c->mask[comp] + first_word, words, nwords * 8, cudaMemcpyHostToDevice, c->stream));
    c->store[comp] = true;
    c->plan_valid = false;
    drop_kept_frames(c);
    c->seq_valid = false;
    c->levels_dirty = true;
    c->world_valid = false;
    c->frame_valid = false;
    return TB_OK;
}

tb_status tb_download_presence(tb_ctx* c, int comp, uint64_t first_word, uint64_t nwords, uint64_t* words) {
    TB_ENTER(c);
    if (!c || comp < 0 || comp >= TB_COMPONENT_COUNT) return TB_ERR_INVALID;
    if (nwords == 0) return TB_OK;
    if (!words) return fail(c, TB_ERR_INVALID, "null pointer");
    if (first_word + nwords > (c->cap + 63) / 64) return fail(c, TB_ERR_INVALID, "presence words outside max_entities");
    TB_CUDA(c, cudaMemcpyAsync(words, c->mask[comp] + first_word, nwords * 8, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

tb_status tb_download_transforms(tb_ctx* c, uint64_t first, uint64_t n, tb_transform* dst) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, dst);
    if (s || n == 0) return s;
    if ((s = flush_integrate(c))) return s;
    const uint64_t per = kStageChunk / sizeof(tb_transform);
    if ((s = ensure_stage(c, (size_t)std::min<uint64_t>(n, per) * sizeof(tb_transform)))) return s;
    for (uint64_t off = 0; off < n; off += per) {
        const uint64_t m = std::min<uint64_t>(per, n - off);
        TB_LAUNCH(c, "unconvert_transforms", m * 112, unconvert_transforms_kernel, div_up(m, kThreads), kThreads, 0, c->rows,
                  (uint32_t)(first + off), (uint32_t)m, (float*)c->stage);
        TB_CUDA(c, cudaMemcpyAsync(dst + off, c->stage, m * sizeof(tb_transform), cudaMemcpyDeviceToHost, c->stream));
    }
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

// ---- structural changes on the device ---------------------------------------------------------------------

namespace {
// what every change of the renderable set invalidates (as tb_upload_presence)
void structure_changed(tb_ctx* c) {
    c->plan_valid = false;
    drop_kept_frames(c);
    c->list_kept = c->final_valid = c->seq_valid = c->fast_clean = false;
    c->levels_dirty = true;
    c->world_valid = false;
    c->frame_valid = false;
}
AllMasks all_masks(tb_ctx* c) {
    AllMasks m{};
    m.count = TB_COMPONENT_COUNT;
    for (int k = 0; k < TB_COMPONENT_COUNT; ++k) m.m[k] = c->mask[k];
    return m;
}
}  // namespace

tb_status tb_upload_spins(tb_ctx* c, uint64_t first, uint64_t n, const float* src) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, src);
    if (s || n == 0) return s;
    if (!c->spin) {
        TB_CUDA(c, cudaMalloc((void**)&c->spin, c->cap * 4));
        TB_CUDA(c, cudaMemsetAsync(c->spin, 0, c->cap * 4, c->stream));
    }
    TB_CUDA(c, cudaMemcpyAsync(c->spin + first, src, n * 4, cudaMemcpyHostToDevice, c->stream));
    c->store[TB_SPIN] = true;
    mark_presence(c, TB_SPIN, first, n, 1);
    return TB_OK;
}

tb_status tb_insert(tb_ctx* c, const tb_transform* t, const tb_sprite* sp, const tb_velocity* v, const uint32_t* parent, uint64_t* index) {
    TB_ENTER(c);
    if (!c || !index) return TB_ERR_INVALID;
    const uint64_t e = c->n;
    tb_status s = tb_set_entity_count(c, e + 1);
    if (s) return s;
    if (t && (s = tb_upload_transforms(c, e, 1, t))) return s;
    if (sp && (s = tb_upload_sprites(c, e, 1, sp))) return s;
    if (v && (s = tb_upload_velocities(c, e, 1, v))) return s;
    if (parent && (s = tb_upload_parents(c, e, 1, parent))) return s;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));  // the sources may be temporaries of the caller
    *index = e;
    return TB_OK;
}

tb_status tb_delete_by_ids(tb_ctx* c, const uint32_t* ids, uint64_t n) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    if (!ids) return fail(c, TB_ERR_INVALID, "null pointer");
    tb_status s = flush_integrate(c);
    if (s) return s;
    if ((s = ensure_stage(c, n * 4))) return s;
    TB_CUDA(c, cudaMemcpyAsync(c->stage, ids, n * 4, cudaMemcpyHostToDevice, c->stream));
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrPending, 0, 4, c->stream));
    TB_LAUNCH(c, "delete_ids", n * 4, delete_ids_kernel, div_up(n, kThreads), kThreads, 0, (const uint32_t*)c->stage, (uint32_t)n, (uint32_t)c->n,
              all_masks(c), c->gen, c->scratch + kHdrPending);
    uint32_t bad = 0;
    TB_CUDA(c, peek_sync(c, &bad, c->scratch + kHdrPending, 4));
    structure_changed(c);
    if (bad) return fail(c, TB_ERR_INVALID, "delete_by_ids: an entity index outside entity_count");
    return TB_OK;
}

tb_status tb_delete_by_query(tb_ctx* c, uint32_t required, uint64_t* deleted) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (deleted) *deleted = 0;
    if (required >> TB_COMPONENT_COUNT) return fail(c, TB_ERR_INVALID, "unknown component bit");
    for (int k = 0; k < TB_COMPONENT_COUNT; ++k)
        if (((required >> k) & 1u) && !c->store[k]) return TB_OK;  // query.rs:110
    if (c->n == 0) return TB_OK;
    tb_status s = flush_integrate(c);
    if (s) return s;
    QueryParams q{};
    for (int k = 0; k < TB_COMPONENT_COUNT; ++k)
        if ((required >> k) & 1u) q.masks[q.nmasks++] = c->mask[k];
    q.nwords = div_up(c->n, 64);
    q.n = (uint32_t)c->n;
    if ((s = ensure_stage(c, 16))) return s;
    unsigned long long* d_cnt = static_cast<unsigned long long*>(c->stage);
    TB_CUDA(c, cudaMemsetAsync(d_cnt, 0, 8, c->stream));
    TB_LAUNCH(c, "delete_query", q.nwords * 8ull * (q.nmasks + TB_COMPONENT_COUNT), delete_query_kernel, div_up(q.nwords, kThreads), kThreads, 0, q,
              all_masks(c), c->gen, d_cnt);
    unsigned long long cnt = 0;
    TB_CUDA(c, peek_sync(c, &cnt, d_cnt, 8));
    if (cnt) structure_changed(c);
    if (deleted) *deleted = cnt;
    return TB_OK;
}

tb_status tb_add_component(tb_ctx* c, int comp, uint64_t index, const void* value) {
    TB_ENTER(c);
    if (!c || comp < 0 || comp >= TB_COMPONENT_COUNT) return TB_ERR_INVALID;
    if (!value) return fail(c, TB_ERR_INVALID, "null pointer");
    if (index >= c->n) return fail(c, TB_ERR_INVALID, "add_component: no such entity index");
    if (!c->store[comp]) return TB_OK;  // ecs.rs:114-118: no store of that type yet -> nothing happens
    tb_status s = TB_OK;
    switch (comp) {
        case TB_TRANSFORM: s = tb_upload_transforms(c, index, 1, static_cast<const tb_transform*>(value)); break;
        case TB_SPRITE: s = tb_upload_sprites(c, index, 1, static_cast<const tb_sprite*>(value)); break;
        case TB_VELOCITY: s = tb_upload_velocities(c, index, 1, static_cast<const tb_velocity*>(value)); break;
        case TB_PARENT: s = tb_upload_parents(c, index, 1, static_cast<const uint32_t*>(value)); break;
        case TB_SPIN: s = tb_upload_spins(c, index, 1, static_cast<const float*>(value)); break;
    }
    if (s) return s;
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    structure_changed(c);
    return TB_OK;
}

tb_status tb_remove_component(tb_ctx* c, int comp, uint64_t index) {
    TB_ENTER(c);
    if (!c || comp < 0 || comp >= TB_COMPONENT_COUNT) return TB_ERR_INVALID;
    if (index >= c->n) return fail(c, TB_ERR_INVALID, "remove_component: no such entity index");
    if (!c->store[comp]) return TB_OK;  // ecs.rs:120-124
    tb_status s = flush_integrate(c);
    if (s) return s;
    mark_presence(c, comp, index, 1, 0);
    structure_changed(c);
    return TB_OK;
}

tb_status tb_entity_generations(tb_ctx* c, uint64_t first, uint64_t n, uint32_t* generations) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, generations);
    if (s || n == 0) return s;
    TB_CUDA(c, cudaMemcpyAsync(generations, c->gen + first, n * 4, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

tb_status tb_entities_alive(tb_ctx* c, const uint32_t* ids, const uint32_t* generations, uint64_t n, uint8_t* alive) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    if (!ids || !generations || !alive) return fail(c, TB_ERR_INVALID, "null pointer");
    tb_status s = ensure_stage(c, n * 9 + 16);
    if (s) return s;
    uint32_t* d_ids = static_cast<uint32_t*>(c->stage);
    uint32_t* d_gens = d_ids + n;
    uint8_t* d_alive = reinterpret_cast<uint8_t*>(d_gens + n);
    TB_CUDA(c, cudaMemcpyAsync(d_ids, ids, n * 4, cudaMemcpyHostToDevice, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(d_gens, generations, n * 4, cudaMemcpyHostToDevice, c->stream));
    TB_LAUNCH(c, "handles_alive", n * 13, handles_alive_kernel, div_up(n, kThreads), kThreads, 0, (const uint32_t*)d_ids, (const uint32_t*)d_gens,
              (uint32_t)n, (uint32_t)c->n, (const uint32_t*)c->gen, d_alive);
    TB_CUDA(c, cudaMemcpyAsync(alive, d_alive, n, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

// ---- query ----------------------------------------------------------------------------------------

tb_status tb_query(tb_ctx* c, uint32_t required, uint32_t* out_indices, uint64_t cap, uint64_t* count) {
    TB_ENTER(c);
    if (!c || !count) return TB_ERR_INVALID;
    *count = 0;
    if (required >> TB_COMPONENT_COUNT) return fail(c, TB_ERR_INVALID, "unknown component bit");
    // query.rs:110 — a required component without a store matches nothing
    for (int k = 0; k < TB_COMPONENT_COUNT; ++k)
        if (((required >> k) & 1u) && !c->store[k]) return TB_OK;
    if (c->n == 0) return TB_OK;
    QueryParams q{};
    for (int k = 0; k < TB_COMPONENT_COUNT; ++k)
        if ((required >> k) & 1u) q.masks[q.nmasks++] = c->mask[k];
    q.nwords = div_up(c->n, 64);
    q.n = (uint32_t)c->n;
    const uint32_t nblocks = div_up(q.nwords, 1024);
    tb_status s = ensure_stage(c, (size_t)nblocks * 4 + 16 + (out_indices ? std::min<uint64_t>(cap, c->n) * 4 : 0));
    if (s) return s;
    uint32_t* d_counts = (uint32_t*)c->stage;
    uint32_t* d_total = d_counts + nblocks;
    uint32_t* d_out = d_counts + nblocks + 4;
    TB_LAUNCH(c, "query_count", q.nwords * 8ull * q.nmasks, query_count_kernel, nblocks, kThreads, 0, q, d_counts);
    TB_LAUNCH(c, "query_scan", 0, scan_blocks_kernel, 1, 1024, 0, d_counts, nblocks, d_total);
    const uint64_t wcap = out_indices ? std::min<uint64_t>(cap, c->n) : 0;
    if (wcap) TB_LAUNCH(c, "query_write", q.nwords * 8ull * q.nmasks, query_write_kernel, nblocks, kThreads, 0, q, d_counts, d_out, (uint32_t)wcap);
    uint32_t total = 0;
    TB_CUDA(c, peek_sync(c, &total, d_total, 4));
    if (wcap) {
        TB_CUDA(c, cudaMemcpyAsync(out_indices, d_out, std::min<uint64_t>(wcap, total) * 4, cudaMemcpyDeviceToHost, c->stream));
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    *count = total;
    return TB_OK;
}

// ---- compose ----------------------------------------------------------------------------------------

tb_status tb_as_matrix4(tb_ctx* c, const tb_transform* src, uint64_t n, float* dst16) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    if (!src || !dst16) return fail(c, TB_ERR_INVALID, "null pointer");
    // temporary rows + world rows, chunked
    const uint64_t per = 1u << 20;
    const uint64_t m0 = std::min<uint64_t>(n, per);
    DeviceTemp rows_mem, yaw_mem, world_mem, in_mem, out_mem;  // freed on every return path
    TB_CUDA(c, rows_mem.alloc(m0 * 64));
    TB_CUDA(c, yaw_mem.alloc(m0 * 4));
    TB_CUDA(c, world_mem.alloc(m0 * 64));
    TB_CUDA(c, in_mem.alloc(m0 * 48));
    TB_CUDA(c, out_mem.alloc(m0 * 64));
    WorldRow* d_world = static_cast<WorldRow*>(world_mem.p);
    float* d_in = static_cast<float*>(in_mem.p);
    float* d_out = static_cast<float*>(out_mem.p);
    const RowView rv = make_row_view(static_cast<float4*>(rows_mem.p), m0, false, static_cast<float*>(yaw_mem.p));
    for (uint64_t off = 0; off < n; off += per) {
        const uint64_t m = std::min<uint64_t>(per, n - off);
        TB_CUDA(c, cudaMemcpyAsync(d_in, src + off, m * 48, cudaMemcpyHostToDevice, c->stream));
        TB_LAUNCH(c, "convert_transforms", m * 96, convert_transforms_kernel, div_up(m, kThreads), kThreads, 0, d_in, rv, 0u, (uint32_t)m);
        TB_LAUNCH(c, "compose_world", m * 128, compose_world_kernel, div_up(m, kThreads), kThreads, 0, rv, (const uint64_t*)nullptr, (uint32_t)m, d_world);
        TB_LAUNCH(c, "world_to_mat4", m * 128, world_to_mat4_kernel, div_up(m, kThreads), kThreads, 0, d_world, (const uint64_t*)nullptr, 0u, (uint32_t)m, d_out);
        TB_CUDA(c, cudaMemcpyAsync(dst16 + off * 16, d_out, m * 64, cudaMemcpyDeviceToHost, c->stream));
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return TB_OK;
}

// ---- the rest of tuber-math, batched (mathops.cuh) -------------------------------------------------------

tb_status tb_quat_from_euler(tb_ctx* c, const float* euler_xyz, uint64_t n, float* wxyz) {
    TB_ENTER(c);
    const MathArg a[] = {{euler_xyz, nullptr, 12}, {nullptr, wxyz, 16}};
    return math_batch(c, n, a, 2, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "quat_from_euler", m * 28ull, quat_from_euler_kernel, div_up(m, kThreads), kThreads, 0, (const float*)d[0], m, (float4*)d[1]);
    });
}

tb_status tb_quat_from_axis_angle(tb_ctx* c, const float* axis_angle4, uint64_t n, float* wxyz) {
    TB_ENTER(c);
    const MathArg a[] = {{axis_angle4, nullptr, 16}, {nullptr, wxyz, 16}};
    return math_batch(c, n, a, 2, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "quat_from_axis_angle", m * 32ull, quat_from_axis_angle_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], m, (float4*)d[1]);
    });
}

tb_status tb_quat_rotation_matrix(tb_ctx* c, const float* wxyz, uint64_t n, float* m16) {
    TB_ENTER(c);
    const MathArg a[] = {{wxyz, nullptr, 16}, {nullptr, m16, 64}};
    return math_batch(c, n, a, 2, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "quat_rotation_matrix", m * 80ull, quat_rotation_matrix_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], m, (float4*)d[1]);
    });
}

tb_status tb_quat_mul(tb_ctx* c, const float* a_wxyz, const float* b_wxyz, uint64_t n, float* out_wxyz) {
    TB_ENTER(c);
    const MathArg a[] = {{a_wxyz, nullptr, 16}, {b_wxyz, nullptr, 16}, {nullptr, out_wxyz, 16}};
    return math_batch(c, n, a, 3, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "quat_mul", m * 48ull, quat_mul_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], (const float4*)d[1], m, (float4*)d[2]);
    });
}

tb_status tb_quat_normalize(tb_ctx* c, const float* wxyz, uint64_t n, float* out_wxyz) {
    TB_ENTER(c);
    const MathArg a[] = {{wxyz, nullptr, 16}, {nullptr, out_wxyz, 16}};
    return math_batch(c, n, a, 2, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "quat_normalize", m * 32ull, quat_normalize_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], m, (float4*)d[1]);
    });
}

tb_status tb_mat4_mul(tb_ctx* c, const float* a16, const float* b16, uint64_t n, float* out16) {
    TB_ENTER(c);
    const MathArg a[] = {{a16, nullptr, 64}, {b16, nullptr, 64}, {nullptr, out16, 64}};
    return math_batch(c, n, a, 3, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "mat4_mul", m * 192ull, mat4_mul_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], (const float4*)d[1], m, (float4*)d[2]);
    });
}

tb_status tb_mat4_transform_vec3(tb_ctx* c, const float* m16, const float* v3, uint64_t n, float* out3) {
    TB_ENTER(c);
    const MathArg a[] = {{m16, nullptr, 64}, {v3, nullptr, 12}, {nullptr, out3, 12}};
    return math_batch(c, n, a, 3, [&](void** d, uint32_t m) {
        TB_LAUNCH(c, "mat4_transform_vec3", m * 88ull, mat4_transform_vec3_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], (const float*)d[1], m, (float*)d[2]);
    });
}

tb_status tb_mat4_try_inverse(tb_ctx* c, const float* m16, uint64_t n, float* out16, uint8_t* some) {
    TB_ENTER(c);
    const MathArg a[] = {{m16, nullptr, 64}, {nullptr, out16, 64}, {nullptr, some, 1}};
    return math_batch(c, n, a, 3, [&](void** d, uint32_t m) {
        cudaMemsetAsync(d[1], 0, (size_t)m * 64, c->stream);  // None: the output matrix reads as zeros
        TB_LAUNCH(c, "mat4_try_inverse", m * 129ull, mat4_try_inverse_kernel, div_up(m, kThreads), kThreads, 0, (const float4*)d[0], m, (float4*)d[1], (uint8_t*)d[2]);
    });
}

// ---- systems ------------------------------------------------------------------------------------------

tb_status tb_system_integrate(tb_ctx* c, float dt) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    // several ticks of the same dt between two frames (a 100 Hz update loop under a 60 Hz render loop,
    // tuber-winit/src/lib.rs:117-141) run as ONE pass over the rows, each tick rounded on its own
    if (c->pending && memcmp(&dt, &c->pending_dt, sizeof(float)) == 0 && c->pending_ticks < 4096) {
        c->pending_ticks += 1;
        if (c->seq_world) c->seq_valid = false;
        return TB_OK;
    }
    if (c->pending) {  // earlier ticks of another dt that no frame consumed
        tb_status s = flush_integrate(c);
        if (s) return s;
    }
    c->pending = true;
    c->pending_dt = dt;
    c->pending_ticks = 1;
    // a flat scene's draw-ordered copy is working storage: the next frame runs this tick on it; a scene graph's copy
    // holds world rows, which the tick invalidates
    if (c->seq_world || !c->seq_has_vel) c->seq_valid = false;
    c->world_valid = false;
    c->frame_valid = false;
    return TB_OK;
}

namespace {
// The spin system as a launch on `stream` (the rows' sin / cos slots and the yaw array change).
void launch_spin(tb_ctx* c, float dt, cudaStream_t stream) {
    prof_begin(c, "spin", c->n * 24ull);
    spin_kernel<<<div_up(c->n, kThreads), kThreads, 0, stream>>>(c->rows, mask_or_null(c, TB_TRANSFORM), mask_or_null(c, TB_SPIN), c->spin, dt,
                                                               (uint32_t)c->n);
    prof_end(c);
    c->launches++;
    c->frame_launches++;
}
void rows_changed_outside_a_frame(tb_ctx* c) {
    c->world_valid = false;
    c->frame_valid = false;
    c->fast_clean = c->seq_valid = false;  // frames without a key pass of their own need one first
}
struct SystemAccess {
    uint32_t reads, writes;
};
constexpr SystemAccess kSystemAccess[TB_SYSTEM_COUNT] = {
    {(1u << TB_COL_VELOCITY) | (1u << TB_COL_TRANSLATION) | (1u << TB_COL_PRESENCE), 1u << TB_COL_TRANSLATION},
    {(1u << TB_COL_SPIN) | (1u << TB_COL_YAW) | (1u << TB_COL_PRESENCE), 1u << TB_COL_YAW},
};
bool systems_conflict(int a, int b) {
    const SystemAccess &x = kSystemAccess[a], &y = kSystemAccess[b];
    return (x.writes & (y.reads | y.writes)) != 0 || (y.writes & (x.reads | x.writes)) != 0;
}
}  // namespace

tb_status tb_system_spin(tb_ctx* c, float dt) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (c->n == 0 || !c->store[TB_TRANSFORM] || !c->store[TB_SPIN]) return TB_OK;
    tb_status s = ensure_rows_current(c);
    if (s) return s;
    launch_spin(c, dt, c->stream);
    rows_changed_outside_a_frame(c);
    return TB_OK;
}

tb_status tb_system_access(int system, uint32_t* reads, uint32_t* writes) {
    if (system < 0 || system >= TB_SYSTEM_COUNT || !reads || !writes) return TB_ERR_INVALID;
    *reads = kSystemAccess[system].reads;
    *writes = kSystemAccess[system].writes;
    return TB_OK;
}

tb_status tb_bundle_step(tb_ctx* c, const int* systems, uint32_t n, float dt, uint32_t* stages) {
    TB_ENTER(c);
    NvtxRange nvtx_call__("tb_bundle_step");
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    if (!systems) return fail(c, TB_ERR_INVALID, "null pointer");
    for (uint32_t k = 0; k < n; ++k)
        if (systems[k] < 0 || systems[k] >= TB_SYSTEM_COUNT) return fail(c, TB_ERR_INVALID, "unknown system id");
    // list scheduling: a system goes one stage after the latest earlier system it conflicts with
    std::vector<uint32_t> stage(n, 0);
    uint32_t nstages = 0;
    for (uint32_t k = 0; k < n; ++k) {
        for (uint32_t j = 0; j < k; ++j)
            if (systems_conflict(systems[j], systems[k])) stage[k] = std::max(stage[k], stage[j] + 1);
        nstages = std::max(nstages, stage[k] + 1);
        if (stages) stages[k] = stage[k];
    }
    if (n == 1 && systems[0] == TB_SYSTEM_VELOCITY) return tb_system_integrate(c, dt);  // fused into the next frame
    tb_status s = flush_integrate(c);  // ticks queued earlier come first
    if (s) return s;
    if (!c->fork_event) {
        TB_CUDA(c, cudaEventCreateWithFlags(&c->fork_event, cudaEventDisableTiming));
        for (int k = 0; k < TB_SYSTEM_COUNT; ++k) {
            TB_CUDA(c, cudaStreamCreateWithFlags(&c->side_stream[k], cudaStreamNonBlocking));
            TB_CUDA(c, cudaEventCreateWithFlags(&c->join_event[k], cudaEventDisableTiming));
        }
    }
    // the same bundle as last time over the same storage: replay the captured graph
    uint32_t dt_bits;
    memcpy(&dt_bits, &dt, 4);
    const uint64_t sig[8] = {c->n, dt_bits, (uint64_t)(uintptr_t)c->rows.t, (uint64_t)(uintptr_t)c->rows.b, (uint64_t)(uintptr_t)c->vel,
                             (uint64_t)(uintptr_t)c->spin, (uint64_t)(uintptr_t)c->stream,
                             (uint64_t)((c->store[TB_VELOCITY] ? 1 : 0) | (c->store[TB_SPIN] ? 2 : 0) | (c->store[TB_TRANSFORM] ? 4 : 0) | (c->rows.st << 8))};
    const std::vector<int> sys_list(systems, systems + n);
    if (c->use_graphs && !c->prof && c->bundle_graph_valid && sys_list == c->bundle_sig_systems && memcmp(sig, c->bundle_sig, sizeof(sig)) == 0) {
        TB_CUDA(c, cudaGraphLaunch(c->bundle_graph, c->stream));
        c->launches += n;
        rows_changed_outside_a_frame(c);
        c->static_valid = false;
        return TB_OK;
    }
    const bool capture = c->use_graphs && !c->prof && cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (!capture) cudaGetLastError();
    struct CaptureGuard {  // an early error return must not leave the stream capturing
        cudaStream_t stream;
        bool active;
        ~CaptureGuard() {
            if (!active) return;
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(stream, &g);
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
        }
    } capture_guard{c->stream, capture};
    for (uint32_t st = 0; st < nstages; ++st) {
        std::vector<uint32_t> members;
        for (uint32_t k = 0; k < n; ++k)
            if (stage[k] == st) members.push_back(k);
        // the first member runs on the context's stream, the others on side streams forked from it
        if (members.size() > 1) TB_CUDA(c, cudaEventRecord(c->fork_event, c->stream));
        for (size_t m = 0; m < members.size(); ++m) {
            const int sys = systems[members[m]];
            cudaStream_t stream = c->stream;
            if (m > 0) {
                stream = c->side_stream[m % TB_SYSTEM_COUNT];
                TB_CUDA(c, cudaStreamWaitEvent(stream, c->fork_event, 0));
            }
            if (sys == TB_SYSTEM_VELOCITY) {
                if (c->n && c->store[TB_TRANSFORM] && c->store[TB_VELOCITY]) {
                    prof_begin(c, "integrate", c->n * 36ull);
                    integrate_kernel<<<div_up(c->n, kThreads), kThreads, 0, stream>>>(c->rows, mask_or_null(c, TB_TRANSFORM),
                                                                                    mask_or_null(c, TB_VELOCITY), c->vel, dt, 1u, (uint32_t)c->n);
                    prof_end(c);
                    c->launches++;
                }
            } else if (sys == TB_SYSTEM_SPIN) {
                if (c->n && c->store[TB_TRANSFORM] && c->store[TB_SPIN]) launch_spin(c, dt, stream);
            }
            if (m > 0) {
                TB_CUDA(c, cudaEventRecord(c->join_event[m % TB_SYSTEM_COUNT], stream));
                TB_CUDA(c, cudaStreamWaitEvent(c->stream, c->join_event[m % TB_SYSTEM_COUNT], 0));
            }
        }
    }
    if (capture) {
        capture_guard.active = false;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(c->stream, &g);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaStreamEndCapture(bundle)");
        if (c->bundle_graph) cudaGraphExecDestroy(c->bundle_graph);
        c->bundle_graph = nullptr;
        c->bundle_graph_valid = false;
        e = cudaGraphInstantiate(&c->bundle_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaGraphInstantiate(bundle)");
        c->bundle_graph_valid = true;
        c->bundle_sig_systems = sys_list;
        memcpy(c->bundle_sig, sig, sizeof(sig));
        TB_CUDA(c, cudaGraphLaunch(c->bundle_graph, c->stream));  // the capture recorded the work: this runs it
    }
    rows_changed_outside_a_frame(c);
    c->static_valid = false;
    return TB_OK;
}

// ---- frame ----------------------------------------------------------------------------------------------

tb_status tb_frame_set_view_projection(tb_ctx* c, const float* m16) {
    if (!c) return TB_ERR_INVALID;
    c->has_vp = m16 != nullptr;
    if (m16) memcpy(c->vp, m16, sizeof(c->vp));
    c->frame_valid = false;
    return TB_OK;
}

// Matrix4::new_orthographic (matrix.rs:41-58), f32, row-major.
tb_status tb_orthographic(float left, float right, float bottom, float top, float near_, float far_, float* out16) {
    if (!out16) return TB_ERR_INVALID;
    for (int k = 0; k < 16; ++k) out16[k] = 0.0f;
    out16[0] = 2.0f / (right - left);
    out16[3] = -((right + left) / (right - left));
    out16[5] = 2.0f / (top - bottom);
    out16[7] = -((top + bottom) / (top - bottom));
    out16[10] = -2.0f / (far_ - near_);
    out16[11] = -((far_ + near_) / (far_ - near_));
    out16[15] = 1.0f;
    return TB_OK;
}

tb_status tb_frame_force_path(tb_ctx* c, uint32_t path) {
    TB_ENTER(c);
    if (!c || path > TB_PATH_SORT) return TB_ERR_INVALID;
    if (tb_status sy = ensure_rows_current(c)) return sy;
    c->force_path = path;
    c->plan_valid = false;
    drop_kept_frames(c);
    return TB_OK;
}

tb_status tb_frame_transform_batch(tb_ctx* c, tb_frame_out* out) {
    TB_ENTER(c);
    NvtxRange nvtx_call__("tb_frame_transform_batch");
    if (!c || !out) return TB_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    c->frame_launches = 0;
    c->frame_valid = false;
    c->last_instances = c->last_batches = 0;
    if (c->shard_state == 4) return shard_frame(c, out);
    c->shard_state = 0;
    c->list_kept = false;  // this frame reuses the kept-order buffers for the shard's own list
    const uint64_t n = c->n;
    // query::<(&Transform, &Sprite)>: no store for either => no matches (query.rs:110)
    if (n == 0 || !c->store[TB_TRANSFORM] || !c->store[TB_SPRITE]) {
        tb_status s = flush_integrate(c);
        if (s) return s;
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        c->frame_valid = true;
        out->kernel_launches = c->frame_launches;
        return TB_OK;
    }
    tb_status s;
    if (c->levels_dirty && (s = plan_levels(c))) return s;
    FrameMode fm{};
    fm.hier = c->nlevels > 1;
    if ((s = ensure_outputs(c, n))) return s;
    if (fm.hier && (s = ensure_world(c, n))) return s;
    float4* d_inst = c->ext_inst ? (float4*)c->ext_inst : c->inst;
    float4* d_vtx = c->ext_inst ? (float4*)c->ext_vtx : c->vtx;
    uint32_t* d_order = c->ext_order ? (uint32_t*)c->ext_order : c->order;

    bool integrate = c->pending;
    c->pending = false;
    uint32_t replanned = 0;
    uint64_t count = 0, num_batches = 0;
    bool fast_done = false;
    if (!c->static_valid && c->sorted_valid && c->final_valid && c->fast_clean && c->static_buckets && c->sorted_cache &&
        c->plan_valid && !fm.hier && c->shard_state != 4) {
        // flat sorted scene with its sort and draw order kept: one kernel emits and validates — or, when nothing at
        // all has been touched for a few frames, streams from the draw-ordered copy of the rows
        const bool untouched = !integrate;
        if (c->chunked) {
            // chunked sort: the kept chunk-major list serves moving and static frames alike (window-local gathers)
            s = chunk_fast_frame(c, integrate, d_inst, d_vtx, d_order, &fast_done);
        } else if (c->seq_valid && !c->seq_world) {
            // rows live in draw order: stream them (a pending tick runs on the copy, with every key re-checked)
            const bool validate = !untouched;
            if (c->plan.key64) s = sequential_frame<uint64_t>(c, validate, integrate, d_inst, d_vtx, d_order, &fast_done);
            else s = sequential_frame<uint32_t>(c, validate, integrate, d_inst, d_vtx, d_order, &fast_done);
        } else {
            if (c->plan.key64) s = sorted_fast_frame<uint64_t>(c, integrate, d_inst, d_vtx, d_order, &fast_done);
            else s = sorted_fast_frame<uint32_t>(c, integrate, d_inst, d_vtx, d_order, &fast_done);
        }
        if (s) return s;
        integrate = false;  // the velocity system has run (as its own pass)
        if (fast_done) {
            count = c->sorted_count;
            num_batches = c->sorted_batches;
            c->clean_streak += 1;
            if (!c->chunked && !c->seq_valid && c->clean_streak >= 2 && (s = build_sequential_rows(c, false))) return s;
        } else {
            // some key changed: the entity-order rows take over again and the full frame below re-sorts
            if ((s = ensure_rows_current(c))) return s;
            c->seq_valid = false;
            c->clean_streak = 0;
            c->fast_backoff = 2;
        }
        // else: some key changed; the full frame below re-sorts (its key pass finds the same difference)
    }
    if (fm.hier && !integrate && c->world_valid && c->sorted_valid && c->final_valid && c->fast_clean && c->static_buckets &&
        c->sorted_cache && c->plan_valid && c->shard_state != 4) {
        // a scene graph nothing has touched since the frame that left its world rows, sort and draw order (no
        // upload, no tick): the frame is the emit over those, e.g. under a new camera
        bool seq_ok = false;
        if (c->chunked) {
            s = chunk_replay_hierarchy(c, d_inst, d_vtx, d_order);
        } else if (c->seq_valid && c->seq_world) {
            if (c->plan.key64) s = sequential_frame<uint64_t>(c, false, false, d_inst, d_vtx, d_order, &seq_ok);
            else s = sequential_frame<uint32_t>(c, false, false, d_inst, d_vtx, d_order, &seq_ok);
        } else {
            if (c->plan.key64) s = replay_hierarchy_frame<uint64_t>(c, d_inst, d_vtx, d_order);
            else s = replay_hierarchy_frame<uint32_t>(c, d_inst, d_vtx, d_order);
        }
        if (s) return s;
        fast_done = true;
        count = c->sorted_count;
        num_batches = c->sorted_batches;
        c->clean_streak += 1;
        if (!c->chunked && !c->seq_valid && c->clean_streak >= 2 && (s = build_sequential_rows(c, true))) return s;
    }
    if (fm.hier && !fast_done && c->hier_fused && c->seq_valid && c->seq_graph && c->level_ranges && c->nlevels > 1 && c->sorted_valid &&
        c->final_valid && c->fast_clean && c->static_buckets && c->sorted_cache && c->plan_valid && !c->chunked && c->shard_state != 4 &&
        !c->levels_dirty && c->seq_leaf_first == c->level_first[c->nlevels - 1]) {
        // scene graph with a kept draw order on the draw-ordered working storage (see hierarchy_seq_frame)
        const ScratchLayout HL = layout_scratch(c->plan, n, true, false, false);
        if (HL.alloc_words <= c->scratch_words) {
            if (c->plan.key64) s = hierarchy_seq_frame<uint64_t>(c, integrate, HL, d_inst, d_vtx, d_order, &fast_done);
            else s = hierarchy_seq_frame<uint32_t>(c, integrate, HL, d_inst, d_vtx, d_order, &fast_done);
            if (s) return s;
            integrate = false;  // the velocity system has run
            if (fast_done) {
                count = c->sorted_count;
                num_batches = c->sorted_batches;
            } else {
                if ((s = ensure_rows_current(c))) return s;
                c->seq_valid = false;
                c->hier_streak = 0;
                c->fast_backoff = 2;
            }
        }
    }
    if (!fast_done && c->seq_live) {
        // whatever kept this frame off the draw-ordered working storage: the entity-order rows take over again
        if ((s = ensure_rows_current(c))) return s;
        c->seq_valid = false;
    }
    if (c->static_valid && c->static_buckets && c->plan_valid && !fm.hier) {
        if ((s = fast_frame(c, integrate, d_inst, d_vtx, d_order, &fast_done))) return s;
        integrate = false;
        if (fast_done) {
            count = c->static_count;
            num_batches = c->static_batches;
        } else {
            drop_kept_frames(c);  // e.g. a velocity moved z: fall back to the full frame below
            replanned = 1;
            c->fast_backoff = 2;
        }
    }
    if (!fast_done && integrate) c->seq_valid = false;  // the full frame ticks the entity-order rows: a draw-ordered copy goes stale
    for (int attempt = 0; attempt < 4 && !fast_done; ++attempt) {
        const bool cached_plan = c->plan_valid;
        if (!c->plan_valid) {
            // discovery: key statistics only (and the velocity system, exactly once)
            TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
            c->zset_active = c->z_dictionary != 0;
            if (c->zset_active) TB_CUDA(c, cudaMemsetAsync(c->zset, 0, (kZSetSlots + 1) * 8, c->stream));
            s = run_prepare(c, false, integrate, fm, nullptr);
            const bool have_zset = c->zset_active;
            c->zset_active = false;
            if (s) return s;
            integrate = false;
            if (have_zset)
                peek_async(c, c->h_zset, c->zset, (kZSetSlots + 1) * 8);
            if ((s = read_header(c))) return s;
            count = c->h_hdr[2];
            if (count == 0) break;
            // few distinct depths: sort them by rank instead of by their varying float bits
            uint32_t zvals[kZSetSlots];
            uint32_t nz = 0;
            if (have_zset && (uint32_t)c->h_zset[kZSetSlots] == 0) {
                for (int k = 0; k < kZSetSlots; ++k)
                    if (c->h_zset[k] >> 32) zvals[nz++] = (uint32_t)c->h_zset[k];
                if (nz > (uint32_t)kMaxZDict) nz = 0;
                std::sort(zvals, zvals + nz);
            }
            make_plan(c, c->h_hdr[0], ~c->h_hdr[1], zvals, nz);
        }
        const FramePlan& plan = c->plan;
        if (c->force_path == TB_PATH_BUCKET && plan.npasses > 1)
            return fail(c, TB_ERR_UNSUPPORTED, "TB_PATH_BUCKET needs at most 256 distinct key buckets");
        fm.lt_alias = plan.npasses == 1 && plan.zbits == 0 && (int)plan.ltbits <= c->lt_limit;
        fm.lt_hist = !fm.lt_alias && (int)plan.ltbits <= c->lt_limit;
        fm.lt_generic = !fm.lt_alias && !fm.lt_hist;
        // Chunked sort (chunked.cuh): short keys of 4..16 bits. The id range is cut into chunks whose rows fit the L2
        // several times over; a (chunk, key) run then holds 2^shift / 2^bits entities on average, so wider keys take
        // bigger chunks (32 entities per run: whole-line stores), within 2^17..2^20 entities (8..64 MB of rows).
        const bool was_chunked = c->chunked;
        c->chunked = c->chunked_sort != 0 && !plan.key64 && plan.pext.nbits > kSmallMaxBits && plan.pext.nbits <= 16 &&
                     !fm.lt_generic && c->shard_state != 4 && n < (1ull << 31);
        if (c->chunked) {
            uint32_t shift = c->chunk_shift_opt ? (uint32_t)c->chunk_shift_opt : std::min(20u, std::max(17u, plan.pext.nbits + 5u));
            while (shift > 11 && (1ull << (shift - 1)) >= n) --shift;  // small scenes: one chunk
            const ChunkGeom cg{shift, div_up(n, 1ull << shift), plan.pext.nbits};
            if (!was_chunked || cg.shift != c->cg.shift || cg.nchunks != c->cg.nchunks || cg.nbits != c->cg.nbits)
                c->sorted_valid = c->list_kept = c->final_valid = false;  // the kept list has another geometry
            c->cg = cg;
            fm.lt_alias = false;  // the batch bins come from the per-chunk histogram
            fm.lt_hist = true;
        } else if (was_chunked) {
            c->sorted_valid = c->list_kept = c->final_valid = false;
        }
        // Frames that gather rows in draw order from anywhere in the storage read whole 64-byte rows: one DRAM atom
        // per entity instead of three partial ones. Chunked frames gather inside an L2-sized window, where every
        // sector of the split arrays gets used, and keep the split layout (dense key pass and write-back).
        if (c->layout_auto && !c->interleaved && plan.npasses > 1 && !fm.hier && !c->chunked && (s = relayout_rows(c, true))) return s;
        const bool prepare_made_counts = !fm.hier;  // flat frames: pass A counts pass 0's tile digits
        const ScratchLayout L = layout_scratch(plan, n, fm.lt_hist, prepare_made_counts, use_small_emit(c, plan),
                                               c->chunked ? &c->cg : nullptr);
        if ((s = ensure_scratch(c, L.alloc_words))) return s;
        if (c->chunked) {
            c->chunk_cnt_off = L.cnt_off;
            c->chunk_remap_off = L.remap_off;
        }
        if ((plan.npasses > 1 || fm.hier || c->chunked) &&
            (s = ensure_sort_buffers(c, c->chunked ? ((uint64_t)c->cg.nchunks << c->cg.shift) : n, plan.key64 ? 8 : 4)))
            return s;
        if (fm.lt_generic) {
            size_t cap = c->texlayer ? c->texlayer_cap : 0;
            if ((s = grow(c, &c->texlayer, &cap, (size_t)n, 4))) return s;
            c->texlayer_cap = cap;
        }
        // Steady state (cached plan, nothing changed since the last frame): replay the captured graph.
        // A frame whose launches would differ in any parameter has a different signature.
        const bool armed_at_start = c->sorted_valid;
        const bool graphable = c->use_graphs && cached_plan && !c->prof && !fm.lt_generic;
        FrameSig sig;
        if (graphable) frame_signature(c, fm, integrate, d_inst, d_vtx, d_order, &sig);
        if (graphable && c->frame_graph_valid && sig == c->frame_sig) {
            TB_CUDA(c, cudaGraphLaunch(c->frame_graph, c->stream));
            c->launches += c->frame_graph_launches;
            c->frame_launches += c->frame_graph_launches;
            if (fm.hier) c->world_valid = true;
        } else {
            bool capture = graphable && !c->frame_graph_valid && c->frame_graph_seen && sig == c->frame_sig;
            const uint32_t launches_before = c->frame_launches;
            if (capture && cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();  // e.g. the caller handed us the legacy default stream: run eagerly from now on
                c->use_graphs = 0;
                capture = false;
            }
            s = record_full_frame(c, fm, L, integrate, prepare_made_counts, d_inst, d_vtx, d_order);
            if (capture) {
                cudaGraph_t g = nullptr;
                cudaError_t e = cudaStreamEndCapture(c->stream, &g);
                if (s == TB_OK && e != cudaSuccess) s = cuda_fail(c, e, "cudaStreamEndCapture");
                if (s == TB_OK) {
                    if (c->frame_graph) cudaGraphExecDestroy(c->frame_graph);
                    c->frame_graph = nullptr;
                    e = cudaGraphInstantiate(&c->frame_graph, g, 0);
                    if (e != cudaSuccess) s = cuda_fail(c, e, "cudaGraphInstantiate");
                }
                if (g) cudaGraphDestroy(g);
                if (s == TB_OK) {
                    c->frame_graph_valid = true;
                    c->frame_graph_launches = c->frame_launches - launches_before;
                    TB_CUDA(c, cudaGraphLaunch(c->frame_graph, c->stream));
                }
            }
            if (s) return s;
            if (graphable && !capture) {
                // first steady frame with this signature runs eagerly; the next one captures
                c->frame_sig = sig;
                c->frame_graph_seen = true;
                c->frame_graph_valid = false;
            }
        }
        integrate = false;
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        count = c->h_hdr[2];
        const uint64_t k_or = c->h_hdr[0], k_and = ~c->h_hdr[1];
        if (count == 0) break;
        if (!plan_covers(plan, k_or, k_and) || c->h_hdr[(kHdrZMiss - kHdrFrame) / 2] != 0) {
            // the scene's keys moved outside the cached plan: re-plan through discovery (which collects
            // the depths for the z dictionary again), or from this frame's statistics without dictionaries
            if (c->z_dictionary) c->plan_valid = false;
            else make_plan(c, k_or, k_and);
            replanned = 1;
            continue;
        }
        if (fm.lt_generic) {
            const uint32_t nb = div_up(count, 1024);
            if ((s = ensure_stage(c, (size_t)nb * 4))) return s;
            uint32_t* d_counts = (uint32_t*)c->stage;
            TB_LAUNCH(c, "batch_heads_count", count * 4, batch_heads_count_kernel, nb, kThreads, 0, c->texlayer, (uint32_t)count, d_counts);
            TB_LAUNCH(c, "batch_heads_scan", 0, scan_blocks_kernel, 1, 1024, 0, d_counts, nb, c->scratch + kHdrNumBatches);
            TB_LAUNCH(c, "batch_heads_write", count * 4, batch_heads_write_kernel, nb, kThreads, 0, c->texlayer, (uint32_t)count,
                      d_counts, c->batches, (uint32_t)c->max_batches);
            TB_LAUNCH(c, "batch_counts", 0, batch_counts_kernel, div_up(c->max_batches, kThreads), kThreads, 0, c->batches,
                      c->scratch + kHdrNumBatches, (uint32_t)c->max_batches, (uint32_t)count);
            if ((s = read_header(c))) return s;
        }
        num_batches = ((uint32_t*)c->h_hdr)[kHdrNumBatches - kHdrFrame];
        // buckets that depend on the Sprite only (no varying z bit): the tile offsets stay valid until
        // a component changes, so the next frames can run as one kernel
        const bool backing_off = c->fast_backoff > 0;
        if (backing_off) c->fast_backoff -= 1;
        if (c->static_buckets && L.small && !fm.hier && !fm.lt_generic && plan.zbits == 0 && !plan.key64 &&
            num_batches <= c->max_batches && !backing_off) {
            c->static_valid = true;
            c->static_count = count;
            c->static_batches = num_batches;
            c->static_offs_off = L.offs_off[0];
            c->static_tiles = L.tiles[0];
        }
        // sorted frames: this frame's sorted (key, id) arrays, tile offsets and entity-order keys stay where
        // they are; the next frame's key pass compares every key with them and its sort passes only run
        // if one changed (a u32 short key of all ones would collide with the "not renderable" mark)
        if (c->sorted_cache && (plan.npasses > 1 || fm.hier || c->chunked) && !fm.lt_generic && (plan.key64 || plan.pext.nbits < 32)) {
            const bool resorted = !armed_at_start || c->h_hdr[(kHdrKeysChanged - kHdrFrame) / 2] != 0;
            c->sorted_valid = true;
            // a frame that ranked has left its draw order and short keys in final_idx / final_keys
            if (resorted) c->final_valid = plan.npasses > 1 || c->chunked;  // written by emit_kernel / the last chunk scatter
            if (resorted) c->seq_valid = false;
            c->clean_streak = 0;
            c->sorted_count = count;
            c->sorted_batches = num_batches;
            c->fast_clean = !backing_off;
            // scene graphs whose entities move keep coming through here (level passes every frame): after two frames that
            // kept their sort, the drawn entities' local rows, velocities and parent ids are copied into draw order and
            // the next frames stream (hierarchy_seq_frame)
            c->hier_streak = (fm.hier && !resorted) ? c->hier_streak + 1 : 0;
            if (fm.hier && c->hier_fused && c->static_buckets && c->hier_streak >= 2 && !(c->seq_valid && c->seq_graph) && c->final_valid &&
                c->level_ranges && c->nlevels > 1 && !c->chunked && (int)plan.ltbits <= c->lt_limit &&
                c->level_first[c->nlevels - 1] + c->level_count[c->nlevels - 1] == n && c->store[TB_VELOCITY]) {
                if ((s = build_sequential_rows(c, false, true))) return s;
                TB_CUDA(c, cudaStreamSynchronize(c->stream));
            }
        }
        break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "frame");
    if (c->prof) prof_harvest(c);
    if (num_batches > c->max_batches) return fail(c, TB_ERR_CAPACITY, "more batches than tb_ctx_set_max_batches allows");
    c->last_instances = count;
    c->last_batches = num_batches;
    c->frame_valid = true;
    out->num_instances = count;
    out->num_batches = num_batches;
    out->d_instances = (const tb_instance*)d_inst;
    out->d_vertices = (const tb_quad_vertex*)d_vtx;
    out->d_batches = (const tb_batch*)c->batches;
    out->d_order = d_order;
    out->path = count == 0 ? TB_PATH_NONE : (c->plan.npasses == 1 ? TB_PATH_BUCKET : TB_PATH_SORT);
    out->key_bits = count == 0 ? 0 : c->plan.pext.nbits;
    out->sort_passes = count == 0 ? 0 : c->plan.npasses;
    out->kernel_launches = c->frame_launches;
    out->hierarchy_levels = fm.hier ? c->nlevels : 0;
    out->replanned = replanned;
    return TB_OK;
}

// ---- multi-GPU frame -------------------------------------------------------------------------------------

tb_status tb_shard_key_stats(tb_ctx* c, uint64_t stats[3]) {
    TB_ENTER(c);
    if (!c || !stats) return TB_ERR_INVALID;
    stats[0] = 0;
    stats[1] = ~0ull;
    stats[2] = 0;
    if (tb_status sy = ensure_rows_current(c)) return sy;
    c->seq_valid = false;
    c->shard_state = 0;
    c->frame_valid = false;
    c->list_kept = c->sorted_valid = false;  // the multi-GPU protocol re-plans every frame from the global statistics
    tb_status s;
    c->shard_count = 0;
    if (c->n == 0 || !c->store[TB_TRANSFORM] || !c->store[TB_SPRITE]) {
        if ((s = flush_integrate(c))) return s;
        stats[1] = 0;  // neutral element of the OR-reduction of ~key
        c->shard_state = 1;
        return TB_OK;
    }
    // Parent links must resolve inside this context: a shard also holds (as entities without a
    // Sprite, after its own id range) the ancestors of its entities that live on other ranks, and
    // recomputes their world matrices itself — see shard.ancestor_closure / DESIGN.md §7.
    if (c->levels_dirty && (s = plan_levels(c))) return s;
    FrameMode fm{};
    fm.hier = c->nlevels > 1;
    if (fm.hier && (s = ensure_world(c, c->n))) return s;
    const bool integrate = c->pending;
    c->pending = false;
    TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
    if ((s = run_prepare(c, false, integrate, fm, nullptr))) return s;
    if ((s = read_header(c))) return s;
    stats[0] = c->h_hdr[0];
    stats[1] = c->h_hdr[1];  // OR of ~key: reduces over ranks with OR, like word 0
    stats[2] = c->h_hdr[2];
    c->shard_count = c->h_hdr[2];
    c->shard_state = 1;
    return TB_OK;
}

tb_status tb_shard_plan(tb_ctx* c, const uint64_t g[3], uint64_t* nbins) {
    if (!c || !g || !nbins) return TB_ERR_INVALID;
    *nbins = 0;
    if (c->shard_state != 1) return fail(c, TB_ERR_INVALID, "tb_shard_plan follows tb_shard_key_stats");
    c->shard_global_count = g[2];
    if (g[2] == 0) {
        c->shard_state = 2;
        *nbins = 1;
        build_frame_plan(0, ~0ull, &c->plan);
        return TB_OK;
    }
    build_frame_plan(g[0], ~g[1], &c->plan);
    c->plan_valid = false;
    drop_kept_frames(c);  // a later single-GPU frame re-plans for itself
    if (c->plan.pext.nbits > TB_SHARD_MAX_KEY_BITS)
        return fail(c, TB_ERR_UNSUPPORTED, "sort key too wide for histogram placement across GPUs");
    if ((int)c->plan.ltbits > (int)kMaxLtBits) return fail(c, TB_ERR_UNSUPPORTED, "too many (layer, texture) bins");
    *nbins = 1ull << c->plan.pext.nbits;
    c->shard_state = 2;
    return TB_OK;
}

tb_status tb_shard_key_histogram(tb_ctx* c, void* d_hist) {
    TB_ENTER(c);
    if (!c || !d_hist) return TB_ERR_INVALID;
    if (c->shard_state != 2) return fail(c, TB_ERR_INVALID, "tb_shard_key_histogram follows tb_shard_plan");
    const FramePlan& plan = c->plan;
    const uint64_t nbins = 1ull << plan.pext.nbits;
    TB_CUDA(c, cudaMemsetAsync(d_hist, 0, nbins * 4, c->stream));
    if (c->shard_count > 0) {
        tb_status s;
        const uint64_t n = c->n;
        FrameMode fm{};  // the batch table comes from the global bins, not from pass A
        fm.hier = c->nlevels > 1;
        const bool prepare_made_counts = !fm.hier;  // flat frames: pass A counts pass 0's tile digits
        const ScratchLayout L = layout_scratch(plan, n, false, prepare_made_counts, use_small_emit(c, plan));
        if ((s = ensure_scratch(c, L.alloc_words))) return s;
        if ((plan.npasses > 1 || fm.hier) && (s = ensure_sort_buffers(c, n, plan.key64 ? 8 : 4))) return s;
        TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, (L.total_words - kHdrFrame) * 4, c->stream));
        if ((s = run_prepare(c, true, false, fm, &L, (uint32_t*)d_hist))) return s;
    }
    TB_CUDA(c, cudaStreamSynchronize(c->stream));  // the caller's collective may run on another stream
    c->shard_state = 3;
    return TB_OK;
}

tb_status tb_shard_place(tb_ctx* c, const void* d_all_hists, uint64_t* global_instances) {
    TB_ENTER(c);
    if (!c || !d_all_hists) return TB_ERR_INVALID;
    if (c->shard_state != 3) return fail(c, TB_ERR_INVALID, "tb_shard_place follows tb_shard_key_histogram");
    const FramePlan& plan = c->plan;
    const uint32_t nbins = 1u << plan.pext.nbits, nlt = 1u << plan.ltbits;
    const uint32_t nblocks = div_up(nbins, 1024);
    tb_status s;
    if ((s = grow(c, &c->remap, &c->remap_cap, (size_t)nbins, 4))) return s;
    if ((s = grow(c, &c->lt_global, &c->lt_global_cap, (size_t)nlt, 4))) return s;
    if ((s = grow(c, &c->place_tmp, &c->place_tmp_cap, (size_t)nblocks * 2 + 2, 4))) return s;
    uint32_t* bt = c->place_tmp;
    uint32_t* bl = c->place_tmp + nblocks;
    uint32_t* totals = c->place_tmp + 2 * nblocks;
    TB_CUDA(c, cudaMemsetAsync(c->lt_global, 0, (size_t)nlt * 4, c->stream));
    const uint32_t* H = (const uint32_t*)d_all_hists;
    TB_LAUNCH(c, "place_sums", 0, place_sums_kernel, nblocks, kThreads, 0, H, (uint32_t)c->nranks, (uint32_t)c->rank, nbins, bt, bl);
    TB_LAUNCH(c, "place_scan", 0, scan_blocks_kernel, 1, 1024, 0, bt, nblocks, totals);
    TB_LAUNCH(c, "place_scan", 0, scan_blocks_kernel, 1, 1024, 0, bl, nblocks, totals + 1);
    TB_LAUNCH(c, "place_write", 0, place_write_kernel, nblocks, kThreads, 0, H, (uint32_t)c->nranks, (uint32_t)c->rank, nbins,
              (const uint32_t*)bt, (const uint32_t*)bl, plan.zbits, c->remap, c->lt_global);
    uint32_t h[2] = {0, 0};
    TB_CUDA(c, peek_sync(c, h, totals, 8));
    if (h[1] != c->shard_count) return fail(c, TB_ERR_INVALID, "histograms do not match this shard's renderable count");
    c->shard_global_count = h[0];
    if (global_instances) *global_instances = h[0];
    c->shard_state = 4;
    return TB_OK;
}

// The rest of a placed multi-GPU frame: batch table from the global bins, radix passes, emit at
// global positions.
static tb_status shard_frame(tb_ctx* c, tb_frame_out* out) {
    const uint64_t n = c->n;
    const FramePlan& plan = c->plan;
    tb_status s;
    if ((s = ensure_outputs(c, std::max<uint64_t>(c->ext_inst ? 0 : c->shard_global_count, 1)))) return s;
    if (c->ext_inst && c->ext_cap < c->shard_global_count)
        return fail(c, TB_ERR_CAPACITY, "output buffers smaller than the global instance count");
    float4* d_inst = c->ext_inst ? (float4*)c->ext_inst : c->inst;
    float4* d_vtx = c->ext_inst ? (float4*)c->ext_vtx : c->vtx;
    uint32_t* d_order = c->ext_order ? (uint32_t*)c->ext_order : c->order;
    uint64_t num_batches = 0;
    if (c->shard_global_count > 0) {
        TB_LAUNCH(c, "build_batches", 0, build_batches_kernel, 1, 1024, 0, (const uint32_t*)c->lt_global, 1u << plan.ltbits, plan,
                  c->batches, (uint32_t)c->max_batches, c->scratch + kHdrNumBatches);
    }
    bool small_plan = false;
    if (c->shard_count > 0) {
        FrameMode fm{};
        fm.hier = c->nlevels > 1;
        const bool prepare_made_counts = !fm.hier;  // flat frames: pass A counts pass 0's tile digits
        const ScratchLayout L = layout_scratch(plan, n, false, prepare_made_counts, use_small_emit(c, plan));
        small_plan = L.small && plan.npasses == 1 && plan.zbits == 0 && !plan.key64 && !fm.hier;
        c->static_offs_off = L.offs_off[0];
        c->static_tiles = L.tiles[0];
        if (plan.key64) s = run_sort_and_emit<uint64_t>(c, fm, L, prepare_made_counts, d_inst, d_vtx, d_order);
        else s = run_sort_and_emit<uint32_t>(c, fm, L, prepare_made_counts, d_inst, d_vtx, d_order);
        if (s) return s;
    }
    if ((s = read_header(c))) return s;
    if (c->shard_global_count > 0) num_batches = ((uint32_t*)c->h_hdr)[kHdrNumBatches - kHdrFrame];
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "frame");
    if (c->prof) prof_harvest(c);
    c->shard_state = 0;
    if (num_batches > c->max_batches) return fail(c, TB_ERR_CAPACITY, "more batches than tb_ctx_set_max_batches allows");
    c->last_instances = c->shard_global_count;  // the output buffers hold the global order
    c->last_batches = num_batches;
    c->frame_valid = true;
    // The ranking emit left this shard's draw order and short keys behind (final_idx / final_keys, shard-local positions):
    // with `remap` they are this rank's part of the global list, which tb_frame_sharded's steady frames keep.
    // Plans of one pass (a few buckets that depend on the Sprite only): what stays is the per-tile bucket offsets of the
    // single-kernel frame (fast_frame) — the shard's steady frame is that kernel with `remap` added.
    if (plan.npasses > 1) c->list_kept = c->sorted_cache && c->nlevels <= 1 && c->shard_global_count > 0;
    else c->list_kept = c->static_buckets && (small_plan || c->shard_count == 0) && c->nlevels <= 1 && c->shard_global_count > 0;
    if (c->list_kept) {
        c->sorted_count = c->static_count = c->shard_count;
        c->list_batches = num_batches;
        c->fast_clean = true;  // a complete key pass has just seen every entity
    }
    out->num_instances = c->shard_global_count;
    out->num_batches = num_batches;
    out->d_instances = (const tb_instance*)d_inst;
    out->d_vertices = (const tb_quad_vertex*)d_vtx;
    out->d_batches = (const tb_batch*)c->batches;
    out->d_order = d_order;
    out->path = c->shard_count == 0 ? TB_PATH_NONE : (plan.npasses == 1 ? TB_PATH_BUCKET : TB_PATH_SORT);
    out->key_bits = plan.pext.nbits;
    out->sort_passes = plan.npasses;
    out->kernel_launches = c->frame_launches;
    out->hierarchy_levels = c->nlevels > 1 ? c->nlevels : 0;
    return TB_OK;
}

// ---- library-owned multi-GPU frame ---------------------------------------------------------------------------------------

tb_status tb_shard_unique_id(uint8_t id[TB_SHARD_ID_BYTES]) {
    static_assert(TB_SHARD_ID_BYTES == kCommIdBytes, "id size");
    if (!id) return TB_ERR_INVALID;
    NcclApi& api = NcclApi::get();
    if (!api.ok()) return TB_ERR_COMM;
    ncclUniqueId uid;
    if (api.GetUniqueId(&uid) != ncclSuccess) return TB_ERR_COMM;
    memcpy(id, &uid, sizeof(uid));
    return TB_OK;
}

tb_status tb_shard_local_id(uint8_t id[TB_SHARD_ID_BYTES]) {
    if (!id) return TB_ERR_INVALID;
    static std::mutex m;
    static uint64_t counter = 0;
    std::lock_guard<std::mutex> lk(m);
    memset(id, 0, TB_SHARD_ID_BYTES);
    memcpy(id, kLocalIdMagic, sizeof(kLocalIdMagic));
    const uint64_t v[2] = {(uint64_t)getpid(), ++counter};
    memcpy(id + 8, v, sizeof(v));
    return TB_OK;
}

namespace {
// A steady list frame goes out as this many launches per rank, each followed by a completion flag (list_signal_kernel).
constexpr uint32_t kListChunks = 16;      // flag slots per rank
constexpr uint32_t kListChunksUsed = 16;  // launches per rank and frame
// the completion flags live behind the order array of the list (one allocation, one IPC mapping)
inline size_t list_flags_offset(uint64_t total) { return ((size_t)total * 4 + 255) & ~(size_t)255; }
struct ListHandles {
    cudaIpcMemHandle_t inst, order, rows;
    void* p_inst;
    void* p_order;
    void* p_rows;
    uint64_t cap;
};
}  // namespace

tb_status tb_shard_init(tb_ctx* c, int rank, int nranks, const uint8_t id[TB_SHARD_ID_BYTES], uint64_t global_first, int root) {
    TB_ENTER(c);
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks || root < 0 || root >= nranks) return TB_ERR_INVALID;
    tb_status s = tb_shard_finalize(c);
    if (s) return s;
    if ((s = tb_shard_set(c, rank, nranks, global_first))) return s;
    if (memcmp(id, kLocalIdMagic, sizeof(kLocalIdMagic)) == 0) {
        std::unique_ptr<LocalComm> lc(new LocalComm());
        if (!lc->init(rank, nranks, id, c->device)) return fail(c, TB_ERR_COMM, lc->err);
        c->comm = std::move(lc);
    } else {
        std::unique_ptr<NcclComm> nc(new NcclComm());
        if (!nc->init(rank, nranks, id)) return fail(c, TB_ERR_COMM, nc->err);
        c->comm = std::move(nc);
    }
    c->list_root = root;
    TB_CUDA(c, cudaMalloc((void**)&c->d_xchg, (size_t)(4 + 4 * nranks) * 8 + sizeof(ListHandles)));
    TB_CUDA(c, cudaMallocHost((void**)&c->h_xchg, (size_t)(4 + 4 * nranks) * 8 + sizeof(ListHandles)));
    // capacity of the global draw list = sum of the ranks' capacities
    c->h_xchg[0] = c->cap;
    TB_CUDA(c, cudaMemcpyAsync(c->d_xchg, c->h_xchg, 8, cudaMemcpyHostToDevice, c->stream));
    if (!c->comm->all_gather(c->d_xchg, c->d_xchg + 4, 8, c->stream)) return fail(c, TB_ERR_COMM, c->comm->err);
    peek_async(c, c->h_xchg + 4, c->d_xchg + 4, (size_t)nranks * 8);
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint64_t total = 0;
    for (int r = 0; r < nranks; ++r) total += c->h_xchg[4 + r];
    if (total >= (1ull << 32)) return fail(c, TB_ERR_CAPACITY, "the global draw list must stay below 2^32 instances");
    // the presenting rank owns the list; everyone else maps it
    ListHandles* hh = reinterpret_cast<ListHandles*>(c->h_xchg + 4 + 4 * nranks);
    ListHandles* dh = reinterpret_cast<ListHandles*>(c->d_xchg + 4 + 4 * nranks);
    memset(hh, 0, sizeof(*hh));
    if (rank == root) {
        TB_CUDA(c, cudaMalloc(&c->list_inst, total * 80));
        TB_CUDA(c, cudaMalloc(&c->list_vtx, total * 64));
        const size_t flag_bytes = (size_t)nranks * kListChunks * kListSlots * 8;
        TB_CUDA(c, cudaMalloc(&c->list_order, list_flags_offset(total) + flag_bytes));
        TB_CUDA(c, cudaMemsetAsync((char*)c->list_order + list_flags_offset(total), 0, flag_bytes, c->stream));
        TB_CUDA(c, cudaMalloc((void**)&c->list_state, (kListStateWords + kListMaxStreams) * sizeof(uint32_t)));
        TB_CUDA(c, cudaMemsetAsync(c->list_state, 0, (kListStateWords + kListMaxStreams) * sizeof(uint32_t), c->stream));
        TB_CUDA(c, cudaMalloc(&c->list_rows, total * 64));
        hh->p_inst = c->list_inst;
        hh->p_order = c->list_order;
        hh->p_rows = c->list_rows;
        hh->cap = total;
        if (!c->comm->same_process() && nranks > 1) {
            TB_CUDA(c, cudaIpcGetMemHandle(&hh->inst, c->list_inst));
            TB_CUDA(c, cudaIpcGetMemHandle(&hh->order, c->list_order));
            TB_CUDA(c, cudaIpcGetMemHandle(&hh->rows, c->list_rows));
        }
        TB_CUDA(c, cudaMemcpyAsync(dh, hh, sizeof(*hh), cudaMemcpyHostToDevice, c->stream));
    }
    if (!c->comm->broadcast(dh, sizeof(ListHandles), root, c->stream)) return fail(c, TB_ERR_COMM, c->comm->err);
    if (rank != root) {
        TB_CUDA(c, cudaMemcpyAsync(hh, dh, sizeof(*hh), cudaMemcpyDeviceToHost, c->stream));
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->comm->same_process()) {
            c->list_inst = hh->p_inst;
            c->list_order = hh->p_order;
            c->list_rows = hh->p_rows;
        } else {
            if (cudaIpcOpenMemHandle(&c->list_inst, hh->inst, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&c->list_order, hh->order, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle(&c->list_rows, hh->rows, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c->list_inst = c->list_order = c->list_rows = nullptr;
                return fail(c, TB_ERR_COMM, "cannot map the presenting rank's draw list (CUDA IPC)");
            }
            c->list_ipc = true;
        }
    }
    c->list_cap = total;
    c->list_flags = reinterpret_cast<unsigned long long*>((char*)c->list_order + list_flags_offset(total));
    c->list_kept = false;
    c->list_backoff = 0;
    c->list_seq = 0;
    // every rank has mapped the list before the owner could ever free it
    if (!c->comm->barrier(c->stream)) return fail(c, TB_ERR_COMM, c->comm->err);
    return TB_OK;
}

tb_status tb_shard_finalize(tb_ctx* c) { return shard_finalize_impl(c, true); }
// collective == false (tb_ctx_destroy): no exchange with the other ranks, which may be gone already
static tb_status shard_finalize_impl(tb_ctx* c, bool collective) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (!c->comm) return TB_OK;
    cudaStreamSynchronize(c->stream);
    if (c->ext_inst == c->list_inst) {
        c->ext_inst = c->ext_vtx = c->ext_order = nullptr;
        c->ext_cap = 0;
    }
    const bool owner = c->rank == c->list_root;
    if (!owner && c->list_ipc) {
        if (c->list_inst) cudaIpcCloseMemHandle(c->list_inst);
        if (c->list_order) cudaIpcCloseMemHandle(c->list_order);
        if (c->list_rows) cudaIpcCloseMemHandle(c->list_rows);
    }
    if (collective) c->comm->barrier(c->stream);  // nobody stores into the list any more
    if (owner) {
        cudaFree(c->list_inst);
        cudaFree(c->list_vtx);
        cudaFree(c->list_order);
        cudaFree(c->list_rows);
        cudaFree(c->list_state);
    }
    c->list_inst = c->list_vtx = c->list_order = c->list_rows = nullptr;
    c->list_flags = nullptr;
    c->list_state = nullptr;
    c->list_kept = false;
    c->list_cap = 0;
    c->list_ipc = false;
    c->list_root = -1;
    cudaFree(c->d_xchg);
    if (c->h_xchg) cudaFreeHost(c->h_xchg);
    c->d_xchg = c->h_xchg = nullptr;
    cudaFree(c->shard_hist);
    cudaFree(c->shard_all_hists);
    cudaFree(c->merge_keys);
    cudaFree(c->merge_all);
    cudaFree(c->merge_batches);
    c->merge_keys = c->merge_all = nullptr;
    c->merge_batches = nullptr;
    c->merge_keys_cap = c->merge_all_cap = c->merge_batches_cap = 0;
    c->shard_hist = c->shard_all_hists = nullptr;
    c->shard_hist_cap = c->shard_all_cap = 0;
    c->comm.reset();
    c->frame_valid = false;
    return TB_OK;
}

// tb_frame_sharded for short keys wider than TB_SHARD_MAX_KEY_BITS (continuous z): every rank sorts and frames its own
// shard, the ranks all-gather their sorted full keys and co-rank (shard.cuh), and a second emit over (entity,
// destination) pairs stores the instances at their global positions in the presenter's list. Slower than histogram
// placement (an extra all-gather of 8 B per element to every rank, binary searches, two emits) but exact.
static tb_status sharded_merge_frame(tb_ctx* c, tb_frame_out* out) {
    Comm& comm = *c->comm;
    const uint32_t nranks = (uint32_t)comm.nranks;
    tb_status s;
    // 1. the shard's own frame into its own buffers (the tick was consumed by tb_shard_key_stats)
    c->shard_state = 0;
    c->ext_inst = c->ext_vtx = c->ext_order = nullptr;
    c->ext_cap = 0;
    tb_frame_out local{};
    if ((s = tb_frame_transform_batch(c, &local))) return s;
    const uint64_t count = local.num_instances, nb = local.num_batches;
    // 2. counts and batch-table sizes of every rank
    uint64_t counts[64], nbs[64];
    if (nranks > 64) return fail(c, TB_ERR_UNSUPPORTED, "more than 64 ranks");
    for (int round = 0; round < 2; ++round) {
        c->h_xchg[0] = round == 0 ? count : nb;
        TB_CUDA(c, cudaMemcpyAsync(c->d_xchg, c->h_xchg, 8, cudaMemcpyHostToDevice, c->stream));
        if (!comm.all_gather(c->d_xchg, c->d_xchg + 4 + (round ? nranks : 0), 8, c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    }
    peek_async(c, c->h_xchg + 4, c->d_xchg + 4, (size_t)nranks * 16);
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint64_t total = 0, maxcount = 1, maxnb = 1;
    for (uint32_t r = 0; r < nranks; ++r) {
        counts[r] = c->h_xchg[4 + r];
        nbs[r] = c->h_xchg[4 + nranks + r];
        total += counts[r];
        maxcount = std::max(maxcount, counts[r]);
        maxnb = std::max(maxnb, nbs[r]);
    }
    if (total > c->list_cap) return fail(c, TB_ERR_CAPACITY, "more instances than the draw list was sized for");
    // 3. sorted full keys of every rank
    if ((s = grow(c, &c->merge_keys, &c->merge_keys_cap, (size_t)maxcount, 8))) return s;
    if ((s = grow(c, &c->merge_all, &c->merge_all_cap, (size_t)maxcount * nranks, 8))) return s;
    const bool hier = c->nlevels > 1;
    TB_LAUNCH(c, "draw_keys", maxcount * 8 + count * 68, draw_keys_kernel, div_up(maxcount, kThreads), kThreads, 0, (const uint32_t*)c->order,
              (uint32_t)c->global_first, (uint32_t)count, c->rows, hier ? (const WorldRow*)c->world : (const WorldRow*)nullptr, c->merge_keys,
              (uint32_t)maxcount);
    if (!comm.all_gather(c->merge_keys, c->merge_all, (size_t)maxcount * 8, c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    // 4. global position of every local draw position
    if ((s = ensure_sort_buffers(c, std::max<uint64_t>(c->n, 1), c->sort_keybytes ? c->sort_keybytes : 4))) return s;
    c->sorted_valid = c->list_kept = c->final_valid = c->seq_valid = false;  // final_idx / final_dst are reused below
    if (count) {
        TB_CUDA(c, cudaMemcpyAsync(c->final_idx, c->order, count * 4, cudaMemcpyDeviceToDevice, c->stream));
        TB_LAUNCH(c, "corank", count * 8ull * nranks, corank_kernel, div_up(count, kThreads), kThreads, 0, (const unsigned long long*)c->merge_all,
                  (const unsigned long long*)(c->d_xchg + 4), (uint32_t)maxcount, nranks, (uint32_t)c->rank, (uint32_t)count, c->final_dst);
    }
    // 5. the global batch table: the ranks' tables merged by (layer, texture) on the host (a few thousand entries)
    if ((s = grow(c, &c->merge_batches, &c->merge_batches_cap, (size_t)maxnb * (nranks + 1), sizeof(BatchOut)))) return s;
    BatchOut* send = c->merge_batches + (size_t)maxnb * nranks;
    TB_CUDA(c, cudaMemsetAsync(send, 0, maxnb * sizeof(BatchOut), c->stream));
    if (nb) TB_CUDA(c, cudaMemcpyAsync(send, c->batches, nb * sizeof(BatchOut), cudaMemcpyDeviceToDevice, c->stream));
    if (!comm.all_gather(send, c->merge_batches, (size_t)maxnb * sizeof(BatchOut), c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    std::vector<BatchOut> tables((size_t)maxnb * nranks);
    TB_CUDA(c, cudaMemcpyAsync(tables.data(), c->merge_batches, tables.size() * sizeof(BatchOut), cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<std::pair<uint64_t, uint64_t>> lt;  // ((layer << 32) | texture, count)
    for (uint32_t r = 0; r < nranks; ++r)
        for (uint64_t k = 0; k < nbs[r]; ++k) {
            const BatchOut& b = tables[(size_t)r * maxnb + k];
            lt.emplace_back(((uint64_t)b.layer << 32) | b.texture, b.count);
        }
    std::sort(lt.begin(), lt.end());
    std::vector<BatchOut> merged;
    uint64_t first = 0;
    for (size_t k = 0; k < lt.size(); ++k) {
        if (merged.empty() || (((uint64_t)merged.back().layer << 32) | merged.back().texture) != lt[k].first)
            merged.push_back(BatchOut{(uint32_t)(lt[k].first >> 32), (uint32_t)lt[k].first, (uint32_t)first, 0});
        merged.back().count += (uint32_t)lt[k].second;
        first += lt[k].second;
    }
    if (merged.size() > c->max_batches) return fail(c, TB_ERR_CAPACITY, "more batches than tb_ctx_set_max_batches allows");
    if (!merged.empty())
        TB_CUDA(c, cudaMemcpyAsync(c->batches, merged.data(), merged.size() * sizeof(BatchOut), cudaMemcpyHostToDevice, c->stream));
    // 6. emit at the global positions into the presenter's list
    if (count) {
        uint32_t shift = 6;
        while ((1ull << shift) < count) ++shift;
        c->h_xchg[0] = count;  // low word: the single chunk's element count
        TB_CUDA(c, cudaMemcpyAsync(c->scratch + kHdrPending, c->h_xchg, 4, cudaMemcpyHostToDevice, c->stream));
        ListParams lp{};
        lp.ids = c->final_idx;
        lp.dst = c->final_dst;
        lp.chunk_cnt = c->scratch + kHdrPending;
        lp.chunk_shift = shift;
        lp.npad = 1u << shift;
        lp.rows = c->rows;
        lp.world = hier ? c->world : nullptr;
        lp.instances = (float4*)c->list_inst;
        lp.vertices = nullptr;
        lp.order = (uint32_t*)c->list_order;
        lp.global_first = (uint32_t)c->global_first;
        lp.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
        lp.plan = c->plan;
        set_camera(c, &lp);
        TB_LAUNCH(c, "emit_list", count * (64 + 80 + 4 + 8), emit_list_kernel<false>, 2u * (uint32_t)c->sm_count, kThreads, sizeof(ListShared), lp);
    }
    TB_CUDA(c, cudaStreamSynchronize(c->stream));  // `merged` (pageable) and the chunk count have been consumed
    if (!comm.barrier(c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    c->ext_inst = c->list_inst;
    c->ext_vtx = nullptr;
    c->ext_order = c->list_order;
    c->ext_cap = c->list_cap;
    *out = local;
    out->num_instances = total;
    out->num_batches = merged.size();
    out->d_instances = (const tb_instance*)c->list_inst;
    out->d_order = (const uint32_t*)c->list_order;
    out->d_vertices = nullptr;
    out->kernel_launches = c->frame_launches;
    if (c->rank == c->list_root && total > 0) {
        if ((s = tb_vertices_from_instances(c, c->list_inst, 0, total, c->list_vtx))) return s;
        out->d_vertices = (const tb_quad_vertex*)c->list_vtx;
    }
    c->last_instances = total;
    c->last_batches = merged.size();
    c->frame_valid = true;
    return TB_OK;
}

// One frame of a scene sharded over the ranks of tb_shard_init, as ONE draw list in the exact global order on the
// presenting rank: key statistics -> (all-gather) -> common plan -> short-key histogram -> (all-gather) -> placement
// -> every rank's emit kernel stores its instances and order at their global positions in the presenter's list ->
// (barrier) -> the presenter regenerates the quad vertices of the whole list.
}  // extern "C"

namespace {
// The steady-state list frame: every rank's keys, order and placement are the previous frame's (`list_kept` everywhere, agreed
// by the caller). No exchange before the emit: each rank streams its draw-ordered working storage (velocity system fused,
// every key recomputed and compared with the kept one) and stores a WORLD ROW (64 bytes, not the 80-byte instance) and the
// entity id at position + remap[key] in the presenting rank's memory — kListChunks launches over consecutive ranges of its
// draw order, each followed by a completion flag. The presenting rank expands chunk k of the rows into instances and quad
// vertices as soon as every rank has passed it (list_wait_kernel + expand_world_rows_kernel), so the expansion overlaps
// the ingress instead of following the barrier. The closing sum carries
// every rank's verdict: *ok_all only if all of them confirmed their keys; otherwise the caller redoes the frame in full.
CameraParams list_camera(const tb_ctx* c) {
    CameraParams cam{};
    set_camera(c, &cam);
    return cam;
}
// the whole list at once (ranks that share a GPU: after the closing barrier)
void expand_whole_list(tb_ctx* c, uint32_t total) {
    TB_LAUNCH(c, "expand_world_rows", (uint64_t)total * 208, expand_world_rows_kernel, std::min<uint32_t>(div_up(total, 32 * kWarps), 8u * (uint32_t)c->sm_count),
              kThreads, 0, (const float4*)c->list_rows, 0u, total, (float4*)c->list_inst, (float4*)c->list_vtx, (const uint32_t*)nullptr, list_camera(c));
}

// The presenting rank's side of the completion flags: per chunk, wait for every rank's flag(s) and regenerate the quad
// vertices of what has arrived since the previous chunk (list_wait_kernel writes the range table on the device).
// *silent: a rank's flags stayed away past the wait kernel's patience. Not an error by itself: the caller votes against the
// frame in the closing sum (which it must still join — the other ranks are on their way there) and everyone redoes it in full.
tb_status expand_list_behind_the_ranks(tb_ctx* c, uint32_t nchunks, uint32_t seq, uint32_t nslots, const uint32_t* d_starts, bool* silent) {
    const long long patience = 4000000000ll;  // about two seconds of SM clocks
    const uint32_t total = (uint32_t)c->shard_global_count;
    TB_CUDA(c, cudaMemsetAsync(c->list_state, 0, 4, c->stream));
    for (uint32_t k = 0; k < nchunks; ++k) {
        TB_LAUNCH(c, "list_wait", (uint64_t)c->comm->nranks * nslots * 8, list_wait_kernel, 1, 1, 0, (const unsigned long long*)c->list_flags,
                  (uint32_t)c->comm->nranks, kListChunks, k, seq, nslots, d_starts, c->list_state, patience);
        TB_LAUNCH(c, "expand_world_rows", (uint64_t)total / nchunks * 208, expand_world_rows_kernel, 8u * (uint32_t)c->sm_count, kThreads, 0,
                  (const float4*)c->list_rows, 0u, 0u, (float4*)c->list_inst, (float4*)c->list_vtx, (const uint32_t*)(c->list_state + kListTable),
                  list_camera(c));
    }
    uint32_t gave_up = 0;
    TB_CUDA(c, peek_sync(c, &gave_up, c->list_state, 4));
    *silent = gave_up != 0;
    return TB_OK;
}

template <typename KeyT>
tb_status list_fast_frame(tb_ctx* c, tb_frame_out* out, bool* ok_all) {
    *ok_all = false;
    Comm& comm = *c->comm;
    tb_status s;
    const bool root = c->rank == c->list_root;
    const uint32_t count = (uint32_t)c->sorted_count, total = (uint32_t)c->shard_global_count;
    c->frame_launches = 0;
    // this rank can serve the frame from its kept order (the caller built the working storage; no memory for it: this
    // rank votes for the full frame below)
    const bool mine = count == 0 || c->seq_valid;
    const bool integrate = mine && c->pending;
    if (mine) c->pending = false;
    EmitParams<KeyT> ep{};
    ep.pass.n = count;
    ep.idx_in = c->final_idx;
    ep.keys_in = (const KeyT*)c->final_keys;
    ep.seq_rows = c->seq_rows;
    ep.seq_rows_rw = c->seq_rows;
    ep.seq_split = c->seq_split;
    ep.seq_vel = c->seq_has_vel ? c->seq_vel : nullptr;
    ep.plan = c->plan;
    ep.remap = c->remap;
    ep.instances = nullptr;  // world rows cross the link (68 instead of 84 bytes per element); the presenter expands them
    ep.vertices = (float4*)c->list_rows;
    ep.world_rows_out = 1u;
    ep.order = (uint32_t*)c->list_order;
    ep.global_first = (uint32_t)c->global_first;
    ep.staged = c->staged_emit;
    ep.dt = c->pending_dt;
    ep.integrate = (integrate && c->seq_has_vel) ? c->pending_ticks : 0u;
    ep.stats = reinterpret_cast<unsigned long long*>(c->scratch + kHdrStats);
    set_camera(c, &ep);
    const bool emit = mine && count > 0;
    if (emit) {
        TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
        if (ep.integrate && c->hidden_movers > 0) tick_hidden_movers(c);
        if (ep.integrate) c->seq_live = true;
    }
    const uint32_t seq = ++c->list_seq;
    static const uint32_t nchunks = [] {  // experiments: TB_LIST_CHUNKS (the same on every rank)
        const char* v = getenv("TB_LIST_CHUNKS");
        return v ? std::min<uint32_t>(std::max(atoi(v), 1), kListChunks) : kListChunksUsed;
    }();
    const uint32_t ntiles = div_up(count, 64), per = std::max<uint32_t>(div_up(ntiles, nchunks), 1u);
    unsigned long long* my_flags = c->list_flags + (size_t)c->rank * kListChunks * kListSlots;
    for (uint32_t k = 0; k < nchunks; ++k) {
        const uint32_t t0 = std::min(ntiles, k * per), t1 = k + 1 == nchunks ? ntiles : std::min(ntiles, (k + 1) * per);
        if (emit && t0 < t1) {
            ep.tile_first = t0;
            ep.tile_end = t1;
            TB_LAUNCH(c, "emit_sequential", (uint64_t)(t1 - t0) * 64 * (64 + 80 + 4 + 4 + sizeof(KeyT) + (ep.integrate ? 28 : 0)),
                      (emit_sequential_kernel<KeyT, true, false, true>), std::min<uint32_t>(2u * (uint32_t)c->sm_count, div_up(t1 - t0, kWarps)), kThreads,
                      sizeof(SeqShared), ep);
        }
        // a rank that emits nothing still reports, so that the presenter's waits end; its vote below voids the frame
        TB_LAUNCH(c, "list_signal", 8, list_signal_kernel<KeyT>, 1, 1, 0, my_flags + (size_t)k * kListSlots, seq, (const KeyT*)c->final_keys, (const uint32_t*)c->remap,
                  (uint32_t)std::min<uint64_t>((uint64_t)t1 * 64, count), count, total);
    }
    bool confirmed = mine;
    // ranks that share a GPU: none of their kernels may wait for another rank's (see Comm::shares_device); the presenter
    // then expands the whole list after the closing barrier
    const bool overlap = !comm.shares_device();
    if (root && total > 0 && overlap) {
        bool silent = false;
        if ((s = expand_list_behind_the_ranks(c, nchunks, seq, 1, nullptr, &silent))) return s;
        uint32_t reached = 0;
        TB_CUDA(c, peek_sync(c, &reached, c->list_state + kListDone, 4));
        if (silent || reached != total) confirmed = false;
    }
    if (emit) {
        if ((s = read_header(c))) return s;
        confirmed = confirmed && c->h_hdr[(kHdrKeysChanged - kHdrFrame) / 2] == 0 && c->h_hdr[(kHdrZMiss - kHdrFrame) / 2] == 0 &&
                    c->h_hdr[2] == c->sorted_count && plan_covers(c->plan, c->h_hdr[0], ~c->h_hdr[1]);
    }
    int unconfirmed = 0;
    if (!comm.sum(c->stream, confirmed ? 0 : 1, &unconfirmed)) return fail(c, TB_ERR_COMM, comm.err);
    if (unconfirmed == 0 && root && total > 0 && !overlap) {
        expand_whole_list(c, total);
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (c->prof) prof_harvest(c);
    if (unconfirmed != 0) return TB_OK;
    *ok_all = true;
    c->ext_inst = c->list_inst;
    c->ext_vtx = nullptr;
    c->ext_order = c->list_order;
    c->ext_cap = c->list_cap;
    c->last_instances = total;
    c->last_batches = c->list_batches;
    c->frame_valid = true;
    c->shard_state = 0;
    out->num_instances = total;
    out->num_batches = c->list_batches;
    out->d_instances = (const tb_instance*)c->list_inst;
    out->d_vertices = root ? (const tb_quad_vertex*)c->list_vtx : nullptr;
    out->d_batches = (const tb_batch*)c->batches;
    out->d_order = (const uint32_t*)c->list_order;
    out->path = count == 0 ? TB_PATH_NONE : TB_PATH_SORT;
    out->key_bits = c->plan.pext.nbits;
    out->sort_passes = 0;
    out->kernel_launches = c->frame_launches;
    return TB_OK;
}

// ... and for plans of one pass: the single-kernel frame (velocity system + key statistics + rank + emit from the cached
// per-tile bucket offsets) storing world rows at bucket position + remap[bucket]. It walks the shard in ENTITY order, so
// progress is monotone per bucket, not per list: with few buckets every chunk of tiles is followed by one flag per bucket
// and the presenter expands ranks x buckets ranges behind them; otherwise it expands the list after the closing barrier.
tb_status list_static_frame(tb_ctx* c, tb_frame_out* out, bool* ok_all) {
    *ok_all = false;
    Comm& comm = *c->comm;
    tb_status s;
    const bool root = c->rank == c->list_root;
    const uint32_t total = (uint32_t)c->shard_global_count;
    c->frame_launches = 0;
    bool confirmed = true;
    // few buckets: every rank reports, per chunk of its tiles, how far each of its bucket segments has come, and the
    // presenter expands the vertices behind them (ranks x buckets streams); otherwise after the closing barrier
    const uint32_t nb = 1u << c->plan.pext.nbits;
    const bool flagged = nb <= kListSlots && (uint32_t)comm.nranks * nb <= kListMaxStreams;
    const bool overlap = flagged && !comm.shares_device();
    const uint32_t seq = ++c->list_seq;
    unsigned long long* my_flags = c->list_flags + (size_t)c->rank * kListChunks * kListSlots;
    if (c->static_count > 0) {
        const bool integrate = c->pending;
        c->pending = false;
        // (nullptr, list_rows): world rows out, see EmitParams::world_rows_out
        if (flagged) s = fast_frame(c, integrate, nullptr, (float4*)c->list_rows, (uint32_t*)c->list_order, &confirmed, c->remap, kListChunksUsed, my_flags, seq, nb);
        else s = fast_frame(c, integrate, nullptr, (float4*)c->list_rows, (uint32_t*)c->list_order, &confirmed, c->remap);
        if (s) return s;
    } else if (flagged) {
        for (uint32_t k = 0; k < kListChunksUsed; ++k)  // nothing to emit, but the presenter's waits must end
            TB_LAUNCH(c, "list_signal", 8ull * nb, list_signal_buckets_kernel, 1, nb, 0, my_flags + (size_t)k * kListSlots, seq,
                      (const uint32_t*)nullptr, 0u, 0u, (const uint32_t*)c->remap, (const uint32_t*)c->shard_hist);
    }
    if (root && total > 0 && overlap) {
        uint32_t* d_starts = c->list_state + kListStateWords;
        TB_LAUNCH(c, "list_starts", (uint64_t)comm.nranks * nb * 4, list_starts_kernel, 1, 1, 0, (const uint32_t*)c->shard_all_hists, (uint32_t)comm.nranks, nb,
                  d_starts);
        bool silent = false;
        if ((s = expand_list_behind_the_ranks(c, kListChunksUsed, seq, nb, d_starts, &silent))) return s;
        if (silent) confirmed = false;
    }
    int unconfirmed = 0;
    if (!comm.sum(c->stream, confirmed ? 0 : 1, &unconfirmed)) return fail(c, TB_ERR_COMM, comm.err);
    if (unconfirmed == 0 && root && total > 0 && !overlap) {
        expand_whole_list(c, total);
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (c->prof) prof_harvest(c);
    if (unconfirmed != 0) return TB_OK;
    *ok_all = true;
    c->ext_inst = c->list_inst;
    c->ext_vtx = nullptr;
    c->ext_order = c->list_order;
    c->ext_cap = c->list_cap;
    c->last_instances = total;
    c->last_batches = c->list_batches;
    c->frame_valid = true;
    c->shard_state = 0;
    out->num_instances = total;
    out->num_batches = c->list_batches;
    out->d_instances = (const tb_instance*)c->list_inst;
    out->d_vertices = root ? (const tb_quad_vertex*)c->list_vtx : nullptr;
    out->d_batches = (const tb_batch*)c->batches;
    out->d_order = (const uint32_t*)c->list_order;
    out->path = c->static_count == 0 ? TB_PATH_NONE : TB_PATH_BUCKET;
    out->key_bits = c->plan.pext.nbits;
    out->sort_passes = 0;
    out->kernel_launches = c->frame_launches;
    return TB_OK;
}
}  // namespace

extern "C" {

tb_status tb_frame_sharded(tb_ctx* c, tb_frame_out* out) {
    TB_ENTER(c);
    NvtxRange nvtx_call__("tb_frame_sharded");
    if (!c || !out) return TB_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    if (!c->comm) return fail(c, TB_ERR_INVALID, "tb_frame_sharded follows tb_shard_init");
    Comm& comm = *c->comm;
    const int nranks = comm.nranks;
    tb_status s;
    {
        // Steady frames: one word of agreement (does EVERY rank still hold last frame's order, untouched by uploads or
        // structural changes?), then the frame without any further exchange before its closing barrier.
        const bool want = c->list_kept && c->fast_clean && c->list_backoff == 0 && c->nlevels <= 1 && !c->levels_dirty;
        const bool one_pass = c->plan.npasses == 1;  // the plan is the global one: the same kind of steady frame on every rank
        // the draw-ordered working storage is (re)built BEFORE the agreement: allocation synchronises the device, which must
        // not happen once some rank's kernels may be waiting for this one's
        if (want && !one_pass && c->sorted_count > 0 && !c->seq_valid) {
            if ((s = ensure_rows_current(c))) return s;
            if ((s = build_sequential_rows(c, false))) return s;
        }
        int cannot = 0;
        if (!comm.sum(c->stream, want ? 0 : 1, &cannot)) return fail(c, TB_ERR_COMM, comm.err);
        if (cannot == 0) {
            bool done = false;
            if (one_pass) s = list_static_frame(c, out, &done);
            else if (c->plan.key64) s = list_fast_frame<uint64_t>(c, out, &done);
            else s = list_fast_frame<uint32_t>(c, out, &done);
            if (s) return s;
            if (done) return TB_OK;
            // some rank's keys moved: the frame is redone in full by everyone (this rank's tick, if it ran, is in the
            // working storage and goes back to the entity-order rows first)
            if ((s = ensure_rows_current(c))) return s;
            c->seq_valid = false;
            c->list_kept = false;
            c->list_backoff = 2;
            memset(out, 0, sizeof(*out));
        } else if (c->list_backoff > 0) {
            c->list_backoff -= 1;
        }
    }
    uint64_t st[3];
    if ((s = tb_shard_key_stats(c, st))) return s;
    memcpy(c->h_xchg, st, sizeof(st));
    TB_CUDA(c, cudaMemcpyAsync(c->d_xchg, c->h_xchg, 24, cudaMemcpyHostToDevice, c->stream));
    if (!comm.all_gather(c->d_xchg, c->d_xchg + 4, 24, c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    peek_async(c, c->h_xchg + 4, c->d_xchg + 4, (size_t)nranks * 24);
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint64_t g[3] = {0, 0, 0};
    for (int r = 0; r < nranks; ++r) {
        g[0] |= c->h_xchg[4 + 3 * r];
        g[1] |= c->h_xchg[4 + 3 * r + 1];
        g[2] += c->h_xchg[4 + 3 * r + 2];
    }
    {
        // identical inputs on every rank: they take the same branch. Wide keys (continuous z): co-ranking instead of
        // histogram placement
        FramePlan gp;
        build_frame_plan(g[0], ~g[1], &gp);
        if (g[2] > 0 && (gp.pext.nbits > TB_SHARD_MAX_KEY_BITS || (int)gp.ltbits > (int)kMaxLtBits)) return sharded_merge_frame(c, out);
    }
    uint64_t nbins = 0;
    if ((s = tb_shard_plan(c, g, &nbins))) return s;
    if ((s = grow(c, &c->shard_hist, &c->shard_hist_cap, (size_t)nbins, 4))) return s;
    if ((s = grow(c, &c->shard_all_hists, &c->shard_all_cap, (size_t)nbins * nranks, 4))) return s;
    if ((s = tb_shard_key_histogram(c, c->shard_hist))) return s;
    if (!comm.all_gather(c->shard_hist, c->shard_all_hists, (size_t)nbins * 4, c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    uint64_t total = 0;
    if ((s = tb_shard_place(c, c->shard_all_hists, &total))) return s;
    if (total > c->list_cap) return fail(c, TB_ERR_CAPACITY, "more instances than the draw list was sized for");
    c->ext_inst = c->list_inst;
    c->ext_vtx = nullptr;  // instances and order cross the link; the presenter regenerates the vertices
    c->ext_order = c->list_order;
    c->ext_cap = c->list_cap;
    if ((s = tb_frame_transform_batch(c, out))) return s;
    if (!comm.barrier(c->stream)) return fail(c, TB_ERR_COMM, comm.err);
    if (c->rank == c->list_root && total > 0) {
        if ((s = tb_vertices_from_instances(c, c->list_inst, 0, total, c->list_vtx))) return s;
        out->d_vertices = (const tb_quad_vertex*)c->list_vtx;
        out->kernel_launches += 1;
    }
    return TB_OK;
}

tb_status tb_device_alloc(int device, uint64_t bytes, void** d_ptr) {
    if (!d_ptr || bytes == 0) return TB_ERR_INVALID;
    *d_ptr = nullptr;
    DeviceGuard device_guard__(device);
    if (cudaMalloc(d_ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        return TB_ERR_OOM;
    }
    return TB_OK;
}
tb_status tb_device_free(void* d_ptr) { return cudaFree(d_ptr) == cudaSuccess ? TB_OK : TB_ERR_CUDA; }
tb_status tb_ipc_export(void* d_ptr, uint8_t handle[TB_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == TB_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d_ptr || !handle) return TB_ERR_INVALID;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, d_ptr) != cudaSuccess) {
        cudaGetLastError();
        return TB_ERR_COMM;
    }
    memcpy(handle, &h, sizeof(h));
    return TB_OK;
}
tb_status tb_ipc_open(int device, const uint8_t handle[TB_IPC_HANDLE_BYTES], void** d_ptr) {
    if (!d_ptr || !handle) return TB_ERR_INVALID;
    *d_ptr = nullptr;
    DeviceGuard device_guard__(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    if (cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return TB_ERR_COMM;
    }
    return TB_OK;
}
tb_status tb_ipc_close(void* d_ptr) { return cudaIpcCloseMemHandle(d_ptr) == cudaSuccess ? TB_OK : TB_ERR_COMM; }
tb_status tb_device_read(tb_ctx* c, const void* d_src, uint64_t bytes, void* dst) {
    TB_ENTER(c);
    if (!c || !d_src || !dst) return TB_ERR_INVALID;
    TB_CUDA(c, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

static tb_status download(tb_ctx* c, const void* d_src, uint64_t first, uint64_t n, uint64_t limit, size_t elem, void* dst) {
    if (!c) return TB_ERR_INVALID;
    if (n == 0) return TB_OK;
    if (!dst) return fail(c, TB_ERR_INVALID, "null pointer");
    if (!c->frame_valid) return fail(c, TB_ERR_INVALID, "no frame has been run since the last change");
    if (first + n > limit) return fail(c, TB_ERR_INVALID, "range outside the last frame's output");
    if (n * elem <= kPeekBytes && (n * elem) % 4 == 0 && (first * elem) % 4 == 0) {  // e.g. the batch table
        peek_async(c, c->h_peek, (const char*)d_src + first * elem, n * elem);
        TB_CUDA(c, cudaStreamSynchronize(c->stream));
        memcpy(dst, c->h_peek, n * elem);
        return TB_OK;
    }
    TB_CUDA(c, cudaStreamSynchronize(c->stream));  // the frame's kernels are not part of the link turn
    LinkTurn turn(c, LinkTurn::kDown, n * elem);
    TB_CUDA(c, cudaMemcpyAsync(dst, (const char*)d_src + first * elem, n * elem, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

tb_status tb_download_instances(tb_ctx* c, uint64_t first, uint64_t n, tb_instance* dst) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    return download(c, c->ext_inst ? c->ext_inst : (void*)c->inst, first, n, c->last_instances, sizeof(tb_instance), dst);
}
tb_status tb_download_vertices(tb_ctx* c, uint64_t first, uint64_t n, tb_quad_vertex* dst) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (c->comm && c->list_vtx && c->ext_inst == c->list_inst)  // the presenting rank of tb_frame_sharded
        return download(c, c->list_vtx, first, n, c->last_instances, 4 * sizeof(tb_quad_vertex), dst);
    if (c->ext_inst && !c->ext_vtx) return fail(c, TB_ERR_INVALID, "this context does not emit vertices");
    return download(c, c->ext_inst ? c->ext_vtx : (void*)c->vtx, first, n, c->last_instances, 4 * sizeof(tb_quad_vertex), dst);
}
tb_status tb_download_batches(tb_ctx* c, uint64_t first, uint64_t n, tb_batch* dst) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    return download(c, c->batches, first, n, c->last_batches, sizeof(tb_batch), dst);
}
tb_status tb_download_order(tb_ctx* c, uint64_t first, uint64_t n, uint32_t* dst) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    return download(c, c->ext_order ? c->ext_order : (void*)c->order, first, n, c->last_instances, 4, dst);
}

tb_status tb_download_world_matrices(tb_ctx* c, uint64_t first, uint64_t n, float* dst16) {
    TB_ENTER(c);
    tb_status s = check_range(c, first, n, dst16);
    if (s || n == 0) return s;
    if ((s = flush_integrate(c))) return s;
    if (c->levels_dirty && (s = plan_levels(c))) return s;
    if (!c->world_valid) {
        if ((s = ensure_world(c, c->n))) return s;
        if (c->nlevels > 1) {
            FrameMode fm{};
            fm.hier = true;
            if ((s = ensure_scratch(c, kHdrWords))) return s;
            TB_CUDA(c, cudaMemsetAsync(c->scratch + kHdrFrame, 0, kHdrReadWords * 4, c->stream));
            if ((s = run_prepare(c, false, false, fm, nullptr))) return s;
        } else {
            TB_LAUNCH(c, "compose_world", c->n * 128, compose_world_kernel, div_up(c->n, kThreads), kThreads, 0, c->rows,
                      mask_or_null(c, TB_TRANSFORM), (uint32_t)c->n, c->world);
        }
        c->world_valid = true;
    }
    const uint64_t per = kStageChunk / 64;
    if ((s = ensure_stage(c, (size_t)std::min<uint64_t>(n, per) * 64))) return s;
    for (uint64_t off = 0; off < n; off += per) {
        const uint64_t m = std::min<uint64_t>(per, n - off);
        TB_LAUNCH(c, "world_to_mat4", m * 128, world_to_mat4_kernel, div_up(m, kThreads), kThreads, 0, c->world,
                  c->store[TB_TRANSFORM] ? mask_or_null(c, TB_TRANSFORM) : c->mask[TB_TRANSFORM], (uint32_t)(first + off),
                  (uint32_t)m, (float*)c->stage);
        TB_CUDA(c, cudaMemcpyAsync(dst16 + off * 16, c->stage, m * 64, cudaMemcpyDeviceToHost, c->stream));
    }
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

// ---- headless presenter ----------------------------------------------------------------------------------------------

tb_status tb_present_headless(tb_ctx* c, uint32_t width, uint32_t height, uint32_t* ids, uint32_t* rgba, uint64_t* drawn,
                              uint64_t* culled) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (drawn) *drawn = 0;
    if (culled) *culled = 0;
    if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 28)) return fail(c, TB_ERR_INVALID, "framebuffer size");
    if (!c->frame_valid) return fail(c, TB_ERR_INVALID, "no frame has been run since the last change");
    const bool list_root = c->comm && c->list_vtx && c->ext_inst == c->list_inst;
    const float4* vtx = list_root ? (const float4*)c->list_vtx : c->ext_inst ? (const float4*)c->ext_vtx : c->vtx;
    const float4* inst = c->ext_inst ? (const float4*)c->ext_inst : c->inst;
    if (c->last_instances > 0 && vtx == nullptr) return fail(c, TB_ERR_INVALID, "this context does not emit vertices");
    const uint32_t npix = width * height;
    DeviceTemp fb;  // ids | rgba | 2 counters
    TB_CUDA(c, fb.alloc((size_t)npix * 8 + 16));
    uint32_t* d_ids = static_cast<uint32_t*>(fb.p);
    uint32_t* d_rgba = d_ids + npix;
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_rgba + npix);
    TB_CUDA(c, cudaMemsetAsync(fb.p, 0, (size_t)npix * 8 + 16, c->stream));
    if (c->last_instances > 0) {
        RasterParams rp{};
        rp.vertices = vtx;
        rp.count = (uint32_t)c->last_instances;
        rp.width = width;
        rp.height = height;
        rp.ids = d_ids;
        rp.counters = d_cnt;
        const uint32_t grid = std::min<uint32_t>(div_up(rp.count, kWarps), 16u * (uint32_t)c->sm_count);
        TB_LAUNCH(c, "raster_quads", c->last_instances * 64, raster_quads_kernel, grid, kThreads, 0, rp);
        TB_LAUNCH(c, "raster_resolve", (uint64_t)npix * 8, raster_resolve_kernel, div_up(npix, kThreads), kThreads, 0, (const uint32_t*)d_ids, inst,
                  npix, d_rgba);
    }
    unsigned long long h_cnt[2] = {0, 0};
    if (ids) TB_CUDA(c, cudaMemcpyAsync(ids, d_ids, (size_t)npix * 4, cudaMemcpyDeviceToHost, c->stream));
    if (rgba) TB_CUDA(c, cudaMemcpyAsync(rgba, d_rgba, (size_t)npix * 4, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaMemcpyAsync(h_cnt, d_cnt, 16, cudaMemcpyDeviceToHost, c->stream));
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (drawn) *drawn = h_cnt[0];
    if (culled) *culled = h_cnt[1];
    return TB_OK;
}

// ---- instrumentation ----------------------------------------------------------------------------------------

tb_status tb_profile_enable(tb_ctx* c, int on) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    cudaStreamSynchronize(c->stream);
    prof_clear(c);
    c->prof = on != 0;
    return TB_OK;
}

tb_status tb_profile_count(tb_ctx* c, uint32_t* n) {
    TB_ENTER(c);
    if (!c || !n) return TB_ERR_INVALID;
    cudaStreamSynchronize(c->stream);
    prof_harvest(c);
    *n = (uint32_t)c->prof_totals.size();
    return TB_OK;
}

tb_status tb_profile_get(tb_ctx* c, uint32_t i, const char** name, double* ms, uint64_t* bytes, uint32_t* launches) {
    if (!c || i >= c->prof_totals.size()) return TB_ERR_INVALID;
    if (name) *name = c->prof_totals[i].name;
    if (ms) *ms = c->prof_totals[i].ms;
    if (bytes) *bytes = c->prof_totals[i].bytes;
    if (launches) *launches = c->prof_totals[i].launches;
    return TB_OK;
}

tb_status tb_launch_count(tb_ctx* c, uint64_t* n) {
    if (!c || !n) return TB_ERR_INVALID;
    *n = c->launches;
    return TB_OK;
}

// ---- sharding ------------------------------------------------------------------------------------------------

tb_status tb_shard_set(tb_ctx* c, int rank, int nranks, uint64_t global_first) {
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return TB_ERR_INVALID;
    if (global_first + c->cap > (1ull << 32)) return fail(c, TB_ERR_INVALID, "global entity ids must fit 32 bits");
    c->rank = rank;
    c->nranks = nranks;
    if (tb_status sy = ensure_rows_current(c)) return sy;
    if (global_first != c->global_first) {
        // final_idx / level lists / captured launches hold ids relative to the old range
        drop_kept_frames(c);
        c->list_kept = c->final_valid = c->seq_valid = c->fast_clean = false;
        c->plan_valid = false;
        c->frame_graph_valid = c->frame_graph_seen = false;
        c->world_valid = false;
    }
    c->global_first = global_first;
    c->levels_dirty = true;
    c->frame_valid = false;
    return TB_OK;
}

tb_status tb_frame_set_output(tb_ctx* c, void* d_instances, void* d_vertices, void* d_order, uint64_t capacity) {
    if (!c) return TB_ERR_INVALID;
    if (!d_instances && !d_vertices && !d_order) {
        c->ext_inst = c->ext_vtx = c->ext_order = nullptr;
        c->ext_cap = 0;
    } else {
        if (!d_instances || !d_order) return fail(c, TB_ERR_INVALID, "instance and order buffers are required");
        c->ext_inst = d_instances;
        c->ext_vtx = d_vertices;  // NULL: this context does not emit vertices (see tb_vertices_from_instances)
        c->ext_order = d_order;
        c->ext_cap = capacity;
    }
    c->frame_valid = false;
    return TB_OK;
}

tb_status tb_vertices_from_instances(tb_ctx* c, const void* d_instances, uint64_t first, uint64_t count, void* d_vertices) {
    TB_ENTER(c);
    if (!c) return TB_ERR_INVALID;
    if (count == 0) return TB_OK;
    if (!d_instances || !d_vertices) return fail(c, TB_ERR_INVALID, "null pointer");
    if (first + count >= (1ull << 32)) return fail(c, TB_ERR_INVALID, "range beyond 2^32 instances");
    const uint32_t grid = std::min<uint32_t>(div_up(count, 32 * kWarps), 8u * (uint32_t)c->sm_count);
    TB_LAUNCH(c, "vertices_from_instances", count * 144, vertices_from_instances_kernel, grid, kThreads, 0, (const float4*)d_instances,
              (uint32_t)first, (uint32_t)count, (float4*)d_vertices, (const uint32_t*)nullptr);
    TB_CUDA(c, cudaStreamSynchronize(c->stream));
    return TB_OK;
}

}  // extern "C"
