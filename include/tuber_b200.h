/* tuber_b200.h — C ABI of the B200-native tuber hot path.
 *
 * One shared library (libtuber_b200.so, sm_100a only, no CPU fallback) that a host ECS binds
 * over FFI. It replaces, for the per-frame data-parallel path only:
 *
 *   tuber-ecs    dense component storage + query   crates/tuber-ecs/src/ecs.rs:16-53,93-129
 *                                                   crates/tuber-ecs/src/query.rs:94-141
 *   tuber-core   Transform::as_matrix4             crates/tuber-core/src/transform.rs:28-38
 *   tuber-math   Matrix4::mul / transform_vec3     crates/tuber-math/src/matrix.rs:160-194
 *   tuber-graphics  GraphicsAPI::render_scene      crates/tuber-graphics/src/lib.rs:34-36,134-156
 *   tuber-engine    system step / render           crates/tuber-engine/src/state.rs:81-87,
 *                                                   crates/tuber-engine/src/lib.rs:105-111
 *
 * Conventions: every call returns tb_status (0 == ok) and never aborts the process; the
 * caller owns every host pointer, the library owns all device memory; one tb_ctx per GPU,
 * used from one host thread at a time (the reference ECS is single-threaded, ecs.rs:12,17).
 * All pointers are plain C; nothing in this header needs CUDA or PyTorch headers.
 * The reference-side binding (Rust `extern "C"` block) is shown in INTEGRATION.md.
 */
#ifndef TUBER_B200_H
#define TUBER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB_API __attribute__((visibility("default")))

typedef struct tb_ctx tb_ctx;
typedef int32_t tb_status;

enum {
    TB_OK = 0,
    TB_ERR_INVALID = 1,     /* bad argument (null pointer, range outside entity_count, ...) */
    TB_ERR_CUDA = 2,        /* a CUDA runtime call failed; see tb_last_error */
    TB_ERR_OOM = 3,         /* device or host allocation failed */
    TB_ERR_RANGE = 4,       /* a component value is outside the SPEC range (layer/texture) */
    TB_ERR_NO_DEVICE = 5,   /* no usable sm_100 device: there is NO CPU fallback */
    TB_ERR_CAPACITY = 6,    /* more entities / batches than the context was created for */
    TB_ERR_CYCLE = 7,       /* the Parent links contain a cycle */
    TB_ERR_COMM = 8,        /* NCCL / peer communication failure */
    TB_ERR_UNSUPPORTED = 9
};

/* ---- component layouts ------------------------------------------------------------------ */

/* tuber_core::transform::Transform (transform.rs:5-11): 4 x Vector3<f32>, this order. */
typedef struct tb_transform {
    float translation[3];
    float angle[3];           /* Euler radians: x roll, y pitch, z yaw (quaternion.rs:39-42) */
    float rotation_center[3];
    float scale[3];
} tb_transform;               /* 48 bytes */

/* [SPEC] the renderable component (absent from the reference snapshot). layer <= 65534,
 * texture <= 65535 (the sort key keeps 16 bits of each). */
typedef struct tb_sprite {
    float size[2];
    uint32_t texture;
    uint32_t layer;
} tb_sprite;                  /* 16 bytes */

/* [SPEC] */
typedef struct tb_velocity {
    float v[3];
} tb_velocity;                /* 12 bytes */

/* tuber_ecs::Parent(EntityIndex) (tuber-ecs/src/lib.rs:18-19), narrowed to u32 on device. */
#define TB_NO_PARENT 0xFFFFFFFFu

/* ---- outputs ([SPEC]) ----------------------------------------------------- */

typedef struct tb_instance {
    float model[16];          /* column-major world matrix (matrix.rs:219-251) */
    float size[2];
    uint32_t texture;
    uint32_t layer;
} tb_instance;                /* 80 bytes */

typedef struct tb_quad_vertex {
    float x, y, z;            /* transform_vec3(world, corner) (matrix.rs:168-171) */
    uint32_t uv;              /* corner k: 0 -> 0x00000000, 1 -> 0x00000001, 2 -> 0x00010001, 3 -> 0x00010000 */
} tb_quad_vertex;             /* 16 bytes, 4 per instance */

typedef struct tb_batch {
    uint32_t layer;
    uint32_t texture;
    uint32_t first_instance;
    uint32_t count;
} tb_batch;                   /* 16 bytes: one per maximal run of equal (layer, texture) */

/* [SPEC] angular velocity about z in radians per second (one f32): the component of the second device system. */
enum tb_component { TB_TRANSFORM = 0, TB_SPRITE = 1, TB_VELOCITY = 2, TB_PARENT = 3, TB_SPIN = 4, TB_COMPONENT_COUNT = 5 };

/* Which batching path a frame took (both give identical results). */
enum tb_path {
    TB_PATH_NONE = 0,
    TB_PATH_BUCKET = 1,   /* <= 256 distinct key buckets: histogram -> scan -> fused compose+scatter emit */
    TB_PATH_SORT = 2      /* general: key-gen -> LSD radix sort of (key, index) pairs -> gather emit */
};

typedef struct tb_frame_out {
    uint64_t num_instances;            /* entities that have Transform and Sprite */
    uint64_t num_batches;
    const tb_instance* d_instances;    /* DEVICE pointers, valid until the next frame / destroy */
    const tb_quad_vertex* d_vertices;
    const tb_batch* d_batches;
    const uint32_t* d_order;           /* entity index per sorted position */
    uint32_t path;                     /* enum tb_path */
    uint32_t key_bits;                 /* sort-key bits that actually vary this frame */
    uint32_t sort_passes;              /* radix passes run (TB_PATH_SORT) */
    uint32_t kernel_launches;          /* kernels this frame launched (its result read-back, a one-warp store, not counted) */
    uint32_t hierarchy_levels;         /* 0 when no entity has a Parent */
    uint32_t replanned;                /* 1 when the cached plan was stale and the frame re-ran batching */
} tb_frame_out;

/* ---- context ------------------------------------------------------------------------------ */

TB_API const char* tb_version(void);
TB_API const char* tb_status_string(tb_status s);
/* Creates a context on CUDA device `device` with room for `max_entities` entity indices.
 * Fails with TB_ERR_NO_DEVICE when there is no sm_100 GPU. Every entry point makes the context's
 * device current for the duration of the call and restores the calling thread's current device on
 * return, so one process may hold a context per GPU. */
TB_API tb_status tb_ctx_create(int device, uint64_t max_entities, tb_ctx** out);
TB_API void tb_ctx_destroy(tb_ctx* ctx);
TB_API const char* tb_last_error(tb_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (a cudaStream_t passed as void*),
 * e.g. the host framework's current stream. NULL restores the context's own stream. */
TB_API tb_status tb_ctx_set_stream(tb_ctx* ctx, void* cuda_stream);
TB_API tb_status tb_ctx_synchronize(tb_ctx* ctx);
/* Cap on the batch table (default 65536 entries). */
TB_API tb_status tb_ctx_set_max_batches(tb_ctx* ctx, uint64_t max_batches);
/* Tuning / test knobs; results never depend on them. Names:
 *   "interleaved_rows" 0|1|-1  one 64-byte row per entity (1) or three dense arrays (0: translation chunk,
 *                           key chunk, rest); -1, the default, starts with the arrays and switches to 64-byte
 *                           rows at the first frame that gathers rows in sorted order (multi-pass, no
 *                           hierarchy). Changing the layout of uploaded rows costs one conversion pass.
 *   "staged_emit"      0|1  stage emitted instances/vertices through shared memory (default 1)
 *   "static_buckets"   0|1  single-kernel steady-state frames (default 1): flat few-bucket scenes whose buckets
 *                           depend on the Sprite only reuse the per-tile offsets of the last full frame, flat
 *                           sorted scenes reuse its draw order (every key re-checked while emitting) and, once
 *                           that order has held for two frames, LIVE in draw order (rows and velocities copied
 *                           once; the velocity system runs on the copy; entity-order rows are brought up to date
 *                           lazily before any call that reads or replaces them); all of it is dropped when a
 *                           component is uploaded. A frame whose validation fails is redone the long way and the
 *                           next two frames skip the single-kernel attempt
 *   "small_emit"       0|1  single-pass frames with <= 8 buckets use the warp-autonomous emit (default 1)
 *   "digit_bits"       1..8 widest radix digit (default 8)
 *   "sorted_cache"     0|1  sorted (multi-pass / hierarchy) frames keep the previous frame's sorted arrays; when the
 *                           key pass finds every entity's sort key unchanged, the sort passes are skipped on the
 *                           device (default 1)
 *   "cuda_graphs"      0|1  steady-state multi-kernel frames (same plan, same buffers as the frame before)
 *                           replay a captured CUDA graph instead of launching kernel by kernel (default 1)
 *   "z_dictionary"     0|1  scenes with at most 64 distinct depths sort by the rank of z instead of by
 *                           its varying float bits: fewer sort passes (default 1)
 *   "lt_limit"         0..16 widest (layer, texture) histogram kept as direct bins (default 16);
 *                           wider scenes find batch boundaries by run-head detection
 *   "hier_fused"       0|1  scene graphs whose entities move and whose draw order is kept run on the draw-ordered
 *                           working storage: level passes for the inner levels only, the leaves are ticked, composed
 *                           (world[parent] * local) and key-checked inside one streaming emit (default 1)
 *   "chunked_sort"     0|1  sorted frames with 4..16 key bits sort per chunk of entity ids and emit chunk by chunk
 *                           (window-local row gathers; default 0: measured slower than the kept-order paths on B200)
 *   "chunk_shift"      0|11..24  log2 of the chunk size in entities (0: chosen from the key width)
 *   "chunk_l2_hint"    0|1  the chunked emit's row gathers carry an L2 evict_last policy (default 1) */
TB_API tb_status tb_ctx_set_option(tb_ctx* ctx, const char* name, int64_t value);

/* Page-locked host memory for the caller's component / output arrays (uploads and downloads
 * from pageable memory work too, at the driver's staging speed). */
TB_API tb_status tb_host_alloc(uint64_t bytes, void** out);
TB_API void tb_host_free(void* p);

/* ---- storage: replaces ComponentStore / Ecs::insert (ecs.rs:16-53, 93-98) ---------------- */

/* Ecs::entity_count (ecs.rs:178-180): indices [0, n) exist. Growing keeps existing rows. */
TB_API tb_status tb_set_entity_count(tb_ctx* ctx, uint64_t n);
TB_API tb_status tb_entity_count(tb_ctx* ctx, uint64_t* n);
/* Column uploads for entity indices [first, first + n). Uploading a column marks the
 * component present on that range (ComponentStore::add_to_entity, ecs.rs:40-43).
 * SOURCE LIFETIME: uploads are enqueued on the context's stream and may return before the copy has
 * read `src` (always so for page-locked memory from tb_host_alloc; pageable memory is staged by the
 * driver before the call returns). The caller must not modify or free `src` until the next call that
 * synchronises the context: tb_ctx_synchronize, tb_frame_transform_batch, any tb_download_*, tb_query,
 * tb_upload_sprites (it reads back the range check). The same holds for tb_upload_presence.
 * LINK TURNS: host transfers of 8 MB or more take turns per device and direction across all contexts of the
 * process (a streaming host keeps two or three contexts in flight so that one step's download overlaps the next
 * step's upload; same-direction copies sharing PCIe are slower than copies taking turns). Such an upload returns
 * once its copy has finished, a download always did. Calls never nest turns, so there is no ordering to respect. */
TB_API tb_status tb_upload_transforms(tb_ctx* ctx, uint64_t first, uint64_t n, const tb_transform* src);
TB_API tb_status tb_upload_sprites(tb_ctx* ctx, uint64_t first, uint64_t n, const tb_sprite* src);
TB_API tb_status tb_upload_velocities(tb_ctx* ctx, uint64_t first, uint64_t n, const tb_velocity* src);
TB_API tb_status tb_upload_parents(tb_ctx* ctx, uint64_t first, uint64_t n, const uint32_t* src);
TB_API tb_status tb_upload_spins(tb_ctx* ctx, uint64_t first, uint64_t n, const float* src);
/* Presence bitset words of one component, same bit layout as bitset.rs:15-36
 * (entity e -> word e/64, bit e%64). Clearing a bit is remove_from_entity (ecs.rs:35-38). */
TB_API tb_status tb_upload_presence(tb_ctx* ctx, int component, uint64_t first_word, uint64_t nwords,
                                    const uint64_t* words);
TB_API tb_status tb_download_presence(tb_ctx* ctx, int component, uint64_t first_word, uint64_t nwords,
                                      uint64_t* words);
TB_API tb_status tb_download_transforms(tb_ctx* ctx, uint64_t first, uint64_t n, tb_transform* dst);

/* ---- structural changes on the device (ecs.rs:93-124) -------------------------------------------
 *
 * Incremental: a few presence words and rows change, nothing is re-uploaded. Entity indices are never
 * reused (ecs.rs:94-97), so an index stays valid for ever; what a delete invalidates is a HANDLE: every
 * slot has a generation on the device, bumped by tb_delete_*, and tb_entities_alive resolves
 * (index, generation) pairs against it (the north star's generation-checked entity indices). */
/* Ecs::insert (ecs.rs:93-98): one new entity with the given components (NULL = not part of the entity);
 * *index receives its EntityIndex (the old entity_count). */
TB_API tb_status tb_insert(tb_ctx* ctx, const tb_transform* t, const tb_sprite* s, const tb_velocity* v, const uint32_t* parent,
                           uint64_t* index);
/* Ecs::delete_by_ids (ecs.rs:106-112): every component of the listed entities is removed. An id outside
 * [0, entity_count) is TB_ERR_INVALID (the reference panics) and nothing is changed for it. */
TB_API tb_status tb_delete_by_ids(tb_ctx* ctx, const uint32_t* ids, uint64_t n);
/* Ecs::delete_by_query (ecs.rs:100-104): the same for every entity that has all `required_components`
 * (bit c == enum tb_component c); a required component without a store matches nothing. */
TB_API tb_status tb_delete_by_query(tb_ctx* ctx, uint32_t required_components, uint64_t* deleted);
/* Ecs::add_component / remove_component (ecs.rs:114-124): as in the reference, a component type that has no store
 * yet is left alone (no-op). `value`: tb_transform / tb_sprite / tb_velocity / uint32_t parent / float spin. */
TB_API tb_status tb_add_component(tb_ctx* ctx, int component, uint64_t index, const void* value);
TB_API tb_status tb_remove_component(tb_ctx* ctx, int component, uint64_t index);
TB_API tb_status tb_entity_generations(tb_ctx* ctx, uint64_t first, uint64_t n, uint32_t* generations);
TB_API tb_status tb_entities_alive(tb_ctx* ctx, const uint32_t* ids, const uint32_t* generations, uint64_t n, uint8_t* alive);

/* ---- query: replaces QueryIterator::new (query.rs:94-129) -------------------------------- */

/* AND of the presence bitsets of the components whose bit is set in `required_components`
 * (bit c == enum tb_component c) over [0, entity_count); writes the matching entity
 * indices in ASCENDING order (QueryIterator pops them from the back, query.rs:138) to
 * `out_indices` (host, capacity `cap`, may be NULL) and the match count to `count`. */
TB_API tb_status tb_query(tb_ctx* ctx, uint32_t required_components, uint32_t* out_indices, uint64_t cap,
                          uint64_t* count);

/* ---- compose: replaces AsMatrix4::as_matrix4 (transform.rs:24-38) ------------------------ */

/* n host Transforms -> n row-major Matrix4f (16 f32 each, matrix.rs:9-12) on the host. */
TB_API tb_status tb_as_matrix4(tb_ctx* ctx, const tb_transform* src, uint64_t n, float* dst16);

/* ---- the rest of tuber-math on the device, batched -------------------------------
 *
 * Host arrays in, host arrays out, n elements each; every operation individually rounded in the
 * reference's own order, so the results are the reference's bit for bit (sin / cos: the host libm's
 * sinf / cosf reproduced on the device). Quaternions are (w, x, y, z); matrices row-major 16 f32
 * (matrix.rs:9-12). */
/* Quaternion::from_euler (quaternion.rs:39-59): (roll, pitch, yaw) per element. */
TB_API tb_status tb_quat_from_euler(tb_ctx* ctx, const float* euler_xyz, uint64_t n, float* wxyz);
/* Quaternion::from_axis_angle (quaternion.rs:28-37): (axis.x, axis.y, axis.z, angle) per element. */
TB_API tb_status tb_quat_from_axis_angle(tb_ctx* ctx, const float* axis_angle4, uint64_t n, float* wxyz);
/* Quaternion::rotation_matrix (quaternion.rs:63-90); the quaternion is not normalised first. */
TB_API tb_status tb_quat_rotation_matrix(tb_ctx* ctx, const float* wxyz, uint64_t n, float* m16);
/* impl Mul for Quaternion (quaternion.rs:132-158). */
TB_API tb_status tb_quat_mul(tb_ctx* ctx, const float* a_wxyz, const float* b_wxyz, uint64_t n, float* out_wxyz);
/* Quaternion::normalized (quaternion.rs:92-120). */
TB_API tb_status tb_quat_normalize(tb_ctx* ctx, const float* wxyz, uint64_t n, float* out_wxyz);
/* impl Mul for Matrix4 (matrix.rs:174-194). */
TB_API tb_status tb_mat4_mul(tb_ctx* ctx, const float* a16, const float* b16, uint64_t n, float* out16);
/* Matrix4::transform_vec3 (matrix.rs:160-171): one point (3 f32) per matrix. */
TB_API tb_status tb_mat4_transform_vec3(tb_ctx* ctx, const float* m16, const float* v3, uint64_t n, float* out3);
/* Matrix4::try_inverse (matrix.rs:98-149): some[i] = 1 and out16[i] = inverse, or some[i] = 0 (None:
 * |det| < 1e-8, number_traits.rs:80-84) and out16[i] all zero. */
TB_API tb_status tb_mat4_try_inverse(tb_ctx* ctx, const float* m16, uint64_t n, float* out16, uint8_t* some);

/* ---- systems: what SystemBundle::step (system.rs:17-23) runs per tick -------------------- */

/* [SPEC] velocity integration for every entity with Transform and Velocity:
 * translation += velocity * dt, dt = DeltaTime.0 as f32 (state.rs:81). Queued on the
 * context's stream; when a frame follows it is fused into that frame's first pass. Several
 * ticks of the same dt before one frame (the fixed-timestep loop of tuber-winit/src/lib.rs:117-141
 * runs 1-2 updates per render) run as one pass, each tick rounded as if run alone. */
TB_API tb_status tb_system_integrate(tb_ctx* ctx, float dt);

/* [SPEC] a second device system: angle.z += spin * dt for every entity with Transform and Spin,
 * then sin / cos of the new half angle (bit-exact host libm values) into the row. Runs at once on the
 * context's stream. */
TB_API tb_status tb_system_spin(tb_ctx* ctx, float dt);

/* What a device system reads and writes, as bit sets over enum tb_column: the declared access sets a
 * scheduler needs. Two systems may run concurrently iff neither writes a column the other reads or writes. */
enum tb_system_id { TB_SYSTEM_VELOCITY = 0, TB_SYSTEM_SPIN = 1, TB_SYSTEM_COUNT = 2 };
enum tb_column { TB_COL_TRANSLATION = 0, TB_COL_YAW = 1, TB_COL_VELOCITY = 2, TB_COL_SPIN = 3, TB_COL_PRESENCE = 4 };
TB_API tb_status tb_system_access(int system, uint32_t* reads, uint32_t* writes);
/* SystemBundle::step (system.rs:17-23) for device systems: the state afterwards is the one running
 * systems[0], systems[1], ... one after the other leaves (insertion order, as the reference runs them).
 * Systems are list-scheduled into stages by their access sets; the systems of one stage touch disjoint
 * columns and run concurrently on separate streams, forked from and joined to the context's stream by
 * events (capturable into a CUDA graph by a caller that captures the context's stream). stages[k], when
 * not NULL, receives the stage system k was placed in. A bundle that is just the velocity system is
 * queued like tb_system_integrate (fused into the next frame). */
TB_API tb_status tb_bundle_step(tb_ctx* ctx, const int* systems, uint32_t n, float dt, uint32_t* stages);

/* ---- frame: what a batching render_scene(&Ecs) does (tuber-graphics/src/lib.rs:134-156) -- */

/* query (&Transform, &Sprite, Opt<&Parent>) -> world matrices (with parent propagation)
 * -> sort key (layer, texture, world z, entity) -> instances + quad vertices + batch table.
 * Synchronises the stream before returning (the counts in `out` are host values). */
TB_API tb_status tb_frame_transform_batch(tb_ctx* ctx, tb_frame_out* out);
/* [SPEC] camera: a row-major 4x4 view-projection matrix folded into the emitted
 * geometry: instance model = view_proj * world (Matrix4::mul, matrix.rs:174-194), quad corners
 * transformed by that product (transform_vec3, matrix.rs:160-171). The draw order stays by world z.
 * NULL removes the camera (the default). */
TB_API tb_status tb_frame_set_view_projection(tb_ctx* ctx, const float* view_proj16);
/* Matrix4::new_orthographic (matrix.rs:41-58) in f32, row-major, into out16. */
TB_API tb_status tb_orthographic(float left, float right, float bottom, float top, float near_plane, float far_plane,
                                 float* out16);
/* Force one path (enum tb_path; TB_PATH_NONE == automatic). Results never depend on it. */
TB_API tb_status tb_frame_force_path(tb_ctx* ctx, uint32_t path);

/* Copies of the last frame's outputs for parity checks / CPU presenters. */
TB_API tb_status tb_download_instances(tb_ctx* ctx, uint64_t first, uint64_t n, tb_instance* dst);
TB_API tb_status tb_download_vertices(tb_ctx* ctx, uint64_t first_instance, uint64_t n_instances, tb_quad_vertex* dst);
TB_API tb_status tb_download_batches(tb_ctx* ctx, uint64_t first, uint64_t n, tb_batch* dst);
TB_API tb_status tb_download_order(tb_ctx* ctx, uint64_t first, uint64_t n, uint32_t* dst);
/* World matrices (row-major 4x4) of entities [first, first+n) as of the last frame; entities
 * without a Transform get the identity. Runs the compose/propagate kernels if needed. */
TB_API tb_status tb_download_world_matrices(tb_ctx* ctx, uint64_t first, uint64_t n, float* dst16);

/* ---- headless presenter ([SPEC]) ------------------------------------------------------
 *
 * What sits behind render_scene after the frame (tuber-graphics/src/lib.rs:134-156 submits an empty encoder
 * and presents): consumes the last frame's draw list on the device. The quad vertices, taken as clip-space
 * coordinates (set a camera with tb_frame_set_view_projection, e.g. tb_orthographic), are rasterised in draw
 * order with the painter's rule into a width x height buffer, row 0 at clip y = -1:
 *   ids[p]  = 1 + draw position of the quad on top of pixel p, 0 where nothing was drawn
 *   rgba[p] = an opaque colour derived from that instance's (texture, layer), 0 for the background
 * Either host buffer may be NULL. Quads entirely outside the clip square are culled here (per quad, on the
 * presenter side: the frame itself never culls, so what it keeps between frames survives camera moves);
 * `drawn` / `culled` count them. Coverage is decided at pixel centres by edge functions in individually
 * rounded f32, so a CPU rasterisation of the same list reproduces the id buffer exactly. */
TB_API tb_status tb_present_headless(tb_ctx* ctx, uint32_t width, uint32_t height, uint32_t* ids, uint32_t* rgba,
                                     uint64_t* drawn, uint64_t* culled);

/* ---- instrumentation ----------------------------------------------------------------------- */

/* Per-kernel device times, measured with CUDA events on the launching stream and summed per
 * kernel name since tb_profile_enable(ctx, 1) (which also resets the totals). `bytes` is the
 * memory traffic the host code expects of those launches. Names are static strings. */
TB_API tb_status tb_profile_enable(tb_ctx* ctx, int on);
TB_API tb_status tb_profile_count(tb_ctx* ctx, uint32_t* n);
TB_API tb_status tb_profile_get(tb_ctx* ctx, uint32_t i, const char** name, double* total_ms, uint64_t* total_bytes,
                                uint32_t* launches);
/* Total kernels launched by this context since creation. Small device -> host reads (a frame's result header, counters,
 * a batch table) are one-warp kernels storing into pinned memory rather than copy-engine transfers (which would queue
 * behind another context's bulk download): they count here, not in tb_frame_out.kernel_launches. */
TB_API tb_status tb_launch_count(tb_ctx* ctx, uint64_t* n);

/* ---- sharding across GPUs ------------------------------------------------- */

/* This context holds entity indices [global_first, global_first + entity_count) of a
 * scene partitioned by contiguous ID range over `nranks` contexts. d_order and parents
 * are then global indices. */
TB_API tb_status tb_shard_set(tb_ctx* ctx, int rank, int nranks, uint64_t global_first);
/* Emit the next frames' instances / vertices / order into caller-provided DEVICE buffers
 * (e.g. a slice of a gather buffer, or memory shared with the presenter). All NULL restores
 * the context's own buffers. d_vertices alone may be NULL: the context then emits no vertices
 * (ranks that ship instances to a presenter; see tb_vertices_from_instances). Capacities are in
 * elements; buffers must be 16-byte aligned. */
TB_API tb_status tb_frame_set_output(tb_ctx* ctx, void* d_instances, void* d_vertices, void* d_order,
                                     uint64_t capacity_instances);
/* Quad vertices of instances [first, first+count) of a device instance buffer, written to the same
 * positions of a device vertex buffer: bit-identical to what the emit kernels write (the vertices
 * are a pure function of the instance). Lets only instances cross NVLink to the presenting rank. */
TB_API tb_status tb_vertices_from_instances(tb_ctx* ctx, const void* d_instances, uint64_t first, uint64_t count,
                                            void* d_vertices);

/* ---- one scene over several GPUs: exact global draw order --------------------
 *
 * Every rank holds a contiguous id range (tb_shard_set). A multi-GPU frame is five calls per rank,
 * with two small exchanges done by the caller (NCCL / any transport) in between:
 *
 *   tb_shard_key_stats        -> 3 words; caller ORs words 0,1 and sums word 2 over the ranks
 *   tb_shard_plan             <- the reduced words: every rank now sorts by the same short key
 *   tb_shard_key_histogram    -> this rank's histogram over the short key (caller-provided device memory)
 *   tb_shard_place            <- all ranks' histograms, rank-major (an all-gather of the above)
 *   tb_frame_transform_batch  emits every element at its position in the GLOBAL order, into the
 *                             buffers given to tb_frame_set_output (e.g. the presenting rank's
 *                             buffers mapped with tb_ipc_open: the emit kernel then stores straight
 *                             over NVLink), and builds the global batch table.
 *
 * Histogram placement is exact for short keys up to TB_SHARD_MAX_KEY_BITS bits; tb_shard_plan returns
 * TB_ERR_UNSUPPORTED beyond (e.g. continuous z). tb_frame_sharded (below) serves such scenes too: the
 * ranks all-gather their sorted full keys and co-rank them.
 * Parent links must resolve inside the context: a shard additionally stores, after its own id
 * range and WITHOUT a Sprite, the ancestors of its entities that live on other ranks, and uploads
 * Parent values as global_first + local index. Their world matrices are then recomputed on this
 * rank (a thin slice of a BFS-ordered tree) and no parent matrix ever crosses the link. */
#define TB_SHARD_MAX_KEY_BITS 22
TB_API tb_status tb_shard_key_stats(tb_ctx* ctx, uint64_t stats[3]);
TB_API tb_status tb_shard_plan(tb_ctx* ctx, const uint64_t global_stats[3], uint64_t* nbins);
TB_API tb_status tb_shard_key_histogram(tb_ctx* ctx, void* d_hist /* nbins x u32, device */);
TB_API tb_status tb_shard_place(tb_ctx* ctx, const void* d_all_hists /* nranks x nbins x u32, device */,
                                uint64_t* global_instances);

/* ---- the same, with the exchanges inside the library ------------------------------
 *
 * A host binds two calls: tb_shard_init once, tb_frame_sharded per frame. The collectives run inside the
 * library on the context's stream, over one of two transports chosen by the id:
 *   tb_shard_unique_id  NCCL (libnccl.so.2 is opened at run time; a copy the process already loaded is
 *                       shared): one process per GPU. Rank 0 creates the id and hands the 128 bytes to
 *                       the other ranks by any means (it is an ncclUniqueId).
 *   tb_shard_local_id   several contexts of ONE process, one host thread per context (one or several
 *                       GPUs): rendezvous in host memory, peer copies.
 * tb_shard_init is collective: every rank calls it with the same id, nranks and root. It also does what
 * tb_shard_set does, allocates the global draw list on `root` (sum of the ranks' max_entities) and maps it
 * into every other rank (CUDA IPC between processes). tb_frame_sharded is collective too: every rank's
 * emit kernel stores its instances and order at their positions of the exact global draw order straight
 * into root's list; on root `out` describes the whole list (instances, regenerated quad vertices, global
 * batch table, global order), elsewhere only the counts are meaningful. A rank that fails before an
 * exchange leaves the others waiting, as with any collective. tb_download_* on root read the global list.
 * STEADY FRAMES keep the list: once a frame has been placed (keys sorted in more than one pass, no scene graph) and no
 * rank has uploaded or changed structure since, the next frame costs one word of agreement, then every rank streams its
 * shard in its kept draw order (queued tb_system_integrate ticks fused, every key recomputed and compared) straight to
 * its kept positions in root's memory — as 64-byte world rows, not 80-byte instances —, in 16 launches each followed by a
 * completion flag; root expands chunk after chunk of the rows into instances and quad vertices (same arithmetic, same
 * bits) as soon as every rank has passed it, and a closing sum carries every rank's verdict. If
 * any rank saw a key change, all of them redo the frame in full inside the same call (out->sort_passes == 0 tells a
 * steady frame from a placed one). Root's device-side wait for a rank's flags gives up after about two seconds: root then
 * votes against the frame like a rank whose keys moved, so a slow rank costs a full frame, never a hung device. */
#define TB_SHARD_ID_BYTES 128
TB_API tb_status tb_shard_unique_id(uint8_t id[TB_SHARD_ID_BYTES]);
TB_API tb_status tb_shard_local_id(uint8_t id[TB_SHARD_ID_BYTES]);
TB_API tb_status tb_shard_init(tb_ctx* ctx, int rank, int nranks, const uint8_t id[TB_SHARD_ID_BYTES],
                               uint64_t global_first, int root);
TB_API tb_status tb_frame_sharded(tb_ctx* ctx, tb_frame_out* out);
TB_API tb_status tb_shard_finalize(tb_ctx* ctx);

/* Raw device buffers that can be shared with other processes of the node (CUDA IPC). */
#define TB_IPC_HANDLE_BYTES 64
TB_API tb_status tb_device_alloc(int device, uint64_t bytes, void** d_ptr);
TB_API tb_status tb_device_free(void* d_ptr);
TB_API tb_status tb_ipc_export(void* d_ptr, uint8_t handle[TB_IPC_HANDLE_BYTES]);
TB_API tb_status tb_ipc_open(int device, const uint8_t handle[TB_IPC_HANDLE_BYTES], void** d_ptr);
TB_API tb_status tb_ipc_close(void* d_ptr);
/* Device-to-host copy of n bytes from any device pointer (peer-mapped ones included). */
TB_API tb_status tb_device_read(tb_ctx* ctx, const void* d_src, uint64_t bytes, void* dst);

#ifdef __cplusplus
}
#endif
#endif /* TUBER_B200_H */
