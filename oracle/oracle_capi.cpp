// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// extern "C" surface over the restated reference (ref_math.hpp, ref_ecs.hpp, ref_frame.hpp)
// so tests/ (ctypes), __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs can drive it. Nothing under tuber_legacy_b200/ may link or load this library.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ref_ecs.hpp"
#include "ref_frame.hpp"
#include "ref_math.hpp"

using namespace ref;

extern "C" {

// ---------------------------------------------------------------------------------------------
// math known-answer surface
// ---------------------------------------------------------------------------------------------
void orc_mat4_mul_i32(const int32_t* a, const int32_t* b, int32_t* out) {
    Mat4<int32_t> A, B;
    std::memcpy(A.v, a, 64);
    std::memcpy(B.v, b, 64);
    Mat4<int32_t> R = mul(A, B);
    std::memcpy(out, R.v, 64);
}
void orc_mat4_mul_f32(const float* a, const float* b, float* out) {
    Mat4f A, B;
    std::memcpy(A.v, a, 64);
    std::memcpy(B.v, b, 64);
    Mat4f R = mul(A, B);
    std::memcpy(out, R.v, 64);
}
void orc_mat4_identity_i32(int32_t* out) {
    Mat4<int32_t> I = identity<int32_t>();
    std::memcpy(out, I.v, 64);
}
int orc_mat4_try_inverse(const float* a, float* out) {
    Mat4f A, R;
    std::memcpy(A.v, a, 64);
    if (!try_inverse(A, &R)) return 0;
    std::memcpy(out, R.v, 64);
    return 1;
}
void orc_quat_from_euler(const float* angles3, float* wxyz) {
    Quat q = from_euler(Vec3{angles3[0], angles3[1], angles3[2]});
    wxyz[0] = q.w; wxyz[1] = q.x; wxyz[2] = q.y; wxyz[3] = q.z;
}
void orc_quat_rotation_matrix(const float* wxyz, float* out16) {
    Mat4f m = rotation_matrix(Quat{wxyz[0], wxyz[1], wxyz[2], wxyz[3]});
    std::memcpy(out16, m.v, 64);
}
void orc_quat_mul(const float* a, const float* b, float* out) {
    Quat r = quat_mul(Quat{a[0], a[1], a[2], a[3]}, Quat{b[0], b[1], b[2], b[3]});
    out[0] = r.w; out[1] = r.x; out[2] = r.y; out[3] = r.z;
}
void orc_new_translation(const float* t3, float* out16) {
    Mat4f m = new_translation(Vec3{t3[0], t3[1], t3[2]});
    std::memcpy(out16, m.v, 64);
}
void orc_new_scale(const float* s3, float* out16) {
    Mat4f m = new_scale(Vec3{s3[0], s3[1], s3[2]});
    std::memcpy(out16, m.v, 64);
}
// n transforms (12 f32 each) -> n row-major matrices (16 f32 each)
void orc_as_matrix4(const float* transforms, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        Transform t;
        std::memcpy(&t, transforms + 12 * i, 48);
        Mat4f m = as_matrix4(t);
        std::memcpy(out + 16 * i, m.v, 64);
    }
}
void orc_as_matrix4_closed(const float* transforms, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        Transform t;
        std::memcpy(&t, transforms + 12 * i, 48);
        Mat4f m = as_matrix4_closed(t);
        std::memcpy(out + 16 * i, m.v, 64);
    }
}
void orc_transform_default(float* out12) {
    Transform t = transform_default();
    std::memcpy(out12, &t, 48);
}
void orc_transform_vec3(const float* m16, const float* p3, float* out3) {
    Mat4f m;
    std::memcpy(m.v, m16, 64);
    Vec3 r = transform_vec3(m, Vec3{p3[0], p3[1], p3[2]});
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
// ---- batched forms (n elements) for the device-vs-oracle tests of the tuber-math entry points ----
void orc_batch_quat_from_euler(const float* e3, uint64_t n, float* wxyz) {
    for (uint64_t i = 0; i < n; ++i) orc_quat_from_euler(e3 + 3 * i, wxyz + 4 * i);
}
void orc_batch_quat_from_axis_angle(const float* aa4, uint64_t n, float* wxyz) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* a = aa4 + 4 * i;
        Quat q = from_axis_angle(Vec3{a[0], a[1], a[2]}, a[3]);
        float* o = wxyz + 4 * i;
        o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    }
}
void orc_batch_quat_rotation_matrix(const float* wxyz, uint64_t n, float* out16) {
    for (uint64_t i = 0; i < n; ++i) orc_quat_rotation_matrix(wxyz + 4 * i, out16 + 16 * i);
}
void orc_batch_quat_mul(const float* a, const float* b, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) orc_quat_mul(a + 4 * i, b + 4 * i, out + 4 * i);
}
void orc_batch_quat_normalize(const float* q, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* a = q + 4 * i;
        Quat r = quat_normalized(Quat{a[0], a[1], a[2], a[3]});
        float* o = out + 4 * i;
        o[0] = r.w; o[1] = r.x; o[2] = r.y; o[3] = r.z;
    }
}
void orc_batch_mat4_mul(const float* a, const float* b, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) orc_mat4_mul_f32(a + 16 * i, b + 16 * i, out + 16 * i);
}
void orc_batch_mat4_try_inverse(const float* a, uint64_t n, float* out, uint8_t* some) {
    for (uint64_t i = 0; i < n; ++i) {
        for (int k = 0; k < 16; ++k) out[16 * i + k] = 0.0f;
        some[i] = (uint8_t)orc_mat4_try_inverse(a + 16 * i, out + 16 * i);
    }
}
void orc_batch_mat4_transform_vec3(const float* m16, const float* p3, uint64_t n, float* out3) {
    for (uint64_t i = 0; i < n; ++i) orc_transform_vec3(m16 + 16 * i, p3 + 3 * i, out3 + 3 * i);
}
void orc_to_column_major(const float* m16, float* out16) {
    Mat4f m;
    std::memcpy(m.v, m16, 64);
    to_column_major(m, out16);
}
uint64_t orc_sort_key(uint32_t layer, uint32_t texture, float z) { return sort_key(layer, texture, z); }

// ---------------------------------------------------------------------------------------------
// bitset known-answer surface (bitset.rs)
// ---------------------------------------------------------------------------------------------
uint64_t orc_u64_set_bit(uint64_t w, uint64_t bit) { u64_set_bit(w, bit); return w; }
uint64_t orc_u64_unset_bit(uint64_t w, uint64_t bit) { u64_unset_bit(w, bit); return w; }
int orc_u64_bit(uint64_t w, uint64_t bit) { return u64_bit(w, bit) ? 1 : 0; }
// returns 0 ok, -1 out of range (the reference panics)
int orc_words_set_bit(uint64_t* words, uint64_t nwords, uint64_t bit) {
    BitSet b(0);
    b.words.assign(words, words + nwords);
    try { b.set_bit(bit); } catch (const std::out_of_range&) { return -1; }
    std::memcpy(words, b.words.data(), nwords * 8);
    return 0;
}
int orc_words_unset_bit(uint64_t* words, uint64_t nwords, uint64_t bit) {
    BitSet b(0);
    b.words.assign(words, words + nwords);
    try { b.unset_bit(bit); } catch (const std::out_of_range&) { return -1; }
    std::memcpy(words, b.words.data(), nwords * 8);
    return 0;
}
int orc_words_bit(const uint64_t* words, uint64_t nwords, uint64_t bit) {
    BitSet b(0);
    b.words.assign(words, words + nwords);
    try { return b.bit(bit) ? 1 : 0; } catch (const std::out_of_range&) { return -1; }
}

// ---------------------------------------------------------------------------------------------
// generic ECS surface (component types are caller-chosen 64-bit ids, payloads are bytes)
// ---------------------------------------------------------------------------------------------
struct OrcEcs {
    Ecs ecs;
    std::string last_error;
    explicit OrcEcs(size_t words) : ecs(words) {}
};

void* orc_ecs_new(uint64_t bitset_words) { return new OrcEcs(bitset_words ? bitset_words : 1024); }
void orc_ecs_free(void* h) { delete static_cast<OrcEcs*>(h); }
const char* orc_ecs_last_error(void* h) { return static_cast<OrcEcs*>(h)->last_error.c_str(); }
uint64_t orc_ecs_entity_count(void* h) { return static_cast<OrcEcs*>(h)->ecs.entity_count(); }

// Ecs::insert (ecs.rs:93-98). Returns the new index, or -1 when the reference would panic
// (entity index beyond the presence bitset).
int64_t orc_ecs_insert(void* h, uint32_t ncomp, const uint64_t* types, const void* const* payloads,
                       const uint64_t* sizes) {
    auto* o = static_cast<OrcEcs*>(h);
    std::vector<ComponentValue> def;
    for (uint32_t i = 0; i < ncomp; ++i) def.push_back(ComponentValue{types[i], payloads[i], size_t(sizes[i])});
    try {
        return int64_t(o->ecs.insert(def));
    } catch (const std::out_of_range&) {
        o->last_error = "entity index out of presence-bitset range (reference panics)";
        return -1;
    }
}

static std::vector<AccessorSpec> make_accessors(uint32_t n, const uint64_t* types, const uint8_t* required,
                                                const uint8_t* exclusive) {
    std::vector<AccessorSpec> a;
    for (uint32_t i = 0; i < n; ++i) a.push_back(AccessorSpec{types[i], required[i] != 0, exclusive[i] != 0});
    return a;
}

// Ecs::query (ecs.rs:127-129): writes the yielded entity indices IN ITERATION ORDER and, per
// yielded row and accessor, 1 when the accessor produced a value (0 == Opt yielded None).
// Returns the number of rows.
uint64_t orc_ecs_query(void* h, uint32_t nacc, const uint64_t* types, const uint8_t* required,
                       const uint8_t* exclusive, uint64_t* out_indices, uint8_t* out_present, uint64_t cap) {
    auto* o = static_cast<OrcEcs*>(h);
    QueryIterator it = o->ecs.query(make_accessors(nacc, types, required, exclusive));
    std::vector<Guard> row;
    uint64_t count = 0;
    while (it.next(row)) {
        if (count < cap) {
            if (out_indices) out_indices[count] = it.index;
            if (out_present)
                for (uint32_t k = 0; k < nacc; ++k) out_present[count * nacc + k] = row[k].some() ? 1 : 0;
        }
        ++count;
    }
    return count;
}

// Ecs::query_by_ids (ecs.rs:132-134)
uint64_t orc_ecs_query_by_ids(void* h, uint32_t nacc, const uint64_t* types, const uint8_t* required,
                              const uint8_t* exclusive, const uint64_t* ids, uint64_t nids, uint64_t* out_indices,
                              uint64_t cap) {
    auto* o = static_cast<OrcEcs*>(h);
    std::unordered_set<EntityIndex> set(ids, ids + nids);
    QueryIteratorByIds it(o->ecs.query(make_accessors(nacc, types, required, exclusive)), std::move(set));
    std::vector<Guard> row;
    uint64_t count = 0;
    while (it.next(row)) {
        if (count < cap && out_indices) out_indices[count] = it.inner.index;
        ++count;
    }
    return count;
}

// Ecs::query_one (ecs.rs:137-169): returns the index or -1 for None.
int64_t orc_ecs_query_one(void* h, uint32_t nacc, const uint64_t* types, const uint8_t* required,
                          const uint8_t* exclusive, uint8_t* out_present) {
    auto* o = static_cast<OrcEcs*>(h);
    EntityIndex idx = 0;
    std::vector<Guard> row;
    if (!o->ecs.query_one(make_accessors(nacc, types, required, exclusive), &idx, row)) return -1;
    if (out_present)
        for (uint32_t k = 0; k < nacc; ++k) out_present[k] = row[k].some() ? 1 : 0;
    return int64_t(idx);
}

// query_one_by_id::<(&T,)> (ecs.rs:172-174) then copy the payload out. Returns the payload
// size, or -1 for None.
int64_t orc_ecs_get(void* h, uint64_t type, uint64_t index, void* out, uint64_t cap) {
    auto* o = static_cast<OrcEcs*>(h);
    std::vector<Guard> row;
    if (index >= o->ecs.entity_count()) return -1;
    if (!o->ecs.query_one_by_id({{type, true, false}}, index, row)) return -1;
    size_t sz = row[0].cell->size;
    std::memcpy(out, row[0].cell->payload, sz < cap ? sz : cap);
    return int64_t(sz);
}
// query_one_by_id::<(&mut T,)> then overwrite the payload (what `*guard = value` does).
int orc_ecs_set(void* h, uint64_t type, uint64_t index, const void* bytes, uint64_t size) {
    auto* o = static_cast<OrcEcs*>(h);
    std::vector<Guard> row;
    if (index >= o->ecs.entity_count()) return -1;
    if (!o->ecs.query_one_by_id({{type, true, true}}, index, row)) return -1;
    if (row[0].cell->size != size) return -2;
    std::memcpy(row[0].cell->payload, bytes, size);
    return 0;
}
void orc_ecs_delete_by_ids(void* h, const uint64_t* ids, uint64_t n) {
    std::vector<EntityIndex> v(ids, ids + n);
    static_cast<OrcEcs*>(h)->ecs.delete_by_ids(v);
}
void orc_ecs_delete_by_query(void* h, uint32_t nacc, const uint64_t* types, const uint8_t* required,
                             const uint8_t* exclusive) {
    static_cast<OrcEcs*>(h)->ecs.delete_by_query(make_accessors(nacc, types, required, exclusive));
}
void orc_ecs_add_component(void* h, uint64_t type, const void* bytes, uint64_t size, uint64_t index) {
    static_cast<OrcEcs*>(h)->ecs.add_component(type, bytes, size, index);
}
void orc_ecs_remove_component(void* h, uint64_t type, uint64_t index) {
    static_cast<OrcEcs*>(h)->ecs.remove_component(type, index);
}
void orc_ecs_insert_shared_resource(void* h, uint64_t type, const void* bytes, uint64_t size) {
    static_cast<OrcEcs*>(h)->ecs.insert_shared_resource(type, bytes, size);
}
int64_t orc_ecs_shared_resource(void* h, uint64_t type, void* out, uint64_t cap) {
    Guard g = static_cast<OrcEcs*>(h)->ecs.shared_resource(type, false);
    if (!g.some()) return -1;
    size_t sz = g.cell->size;
    std::memcpy(out, g.cell->payload, sz < cap ? sz : cap);
    return int64_t(sz);
}
// Two simultaneous exclusive borrows of the same component: the reference panics
// (RefCell). Returns 1 when the restatement raises the same conflict.
int orc_ecs_double_borrow_conflicts(void* h, uint64_t type, uint64_t index) {
    auto* o = static_cast<OrcEcs*>(h);
    std::vector<Guard> a, b;
    if (!o->ecs.query_one_by_id({{type, true, true}}, index, a)) return -1;
    try {
        o->ecs.query_one_by_id({{type, true, true}}, index, b);
    } catch (const BorrowError&) {
        return 1;
    }
    return 0;
}

// SystemBundle (system.rs:8-24). A system is a C callback returning 0 for Ok(()) and
// non-zero for Err(_).
typedef int (*orc_system_fn)(void* ecs, void* additional_data);
struct OrcBundle {
    SystemBundle<void*> bundle;
};
void* orc_bundle_new() { return new OrcBundle(); }
void orc_bundle_free(void* b) { delete static_cast<OrcBundle*>(b); }
uint64_t orc_bundle_len(void* b) { return static_cast<OrcBundle*>(b)->bundle.systems.size(); }
void orc_bundle_add_system(void* b, orc_system_fn fn, void* ecs_handle) {
    static_cast<OrcBundle*>(b)->bundle.add_system([fn, ecs_handle](Ecs&, void*& ad) -> std::string {
        int rc = fn(ecs_handle, ad);
        return rc == 0 ? std::string() : ("system error " + std::to_string(rc));
    });
}
// returns 0 for Ok(()), 1 for the first Err (remaining systems are not run)
int orc_bundle_step(void* b, void* ecs_handle, void* additional_data) {
    auto* o = static_cast<OrcEcs*>(ecs_handle);
    std::string err = static_cast<OrcBundle*>(b)->bundle.step(o->ecs, additional_data);
    if (!err.empty()) { o->last_error = err; return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// scene + frame surface (SPEC)
// ---------------------------------------------------------------------------------------------
static inline bool mask_bit(const uint64_t* mask, uint64_t i) {
    return mask == nullptr || ((mask[i / 64] >> (i % 64)) & 1u);
}

// Builds a restated-reference Ecs holding n entities. A component array that is NULL is
// absent from every entity; a non-NULL array with a NULL mask is present on every entity.
// parents[i] == 0xFFFFFFFF means entity i has no Parent component.
void* orc_scene_create(uint64_t n, const float* transforms, const uint64_t* mask_transform, const uint32_t* sprites,
                       const uint64_t* mask_sprite, const float* velocities, const uint64_t* mask_velocity,
                       const uint32_t* parents, uint64_t bitset_words, const float* spins, const uint64_t* mask_spin) {
    uint64_t need = (n + 63) / 64;
    if (bitset_words == 0) bitset_words = need > 1024 ? need : 1024;
    auto* o = new OrcEcs(bitset_words);
    std::vector<ComponentValue> def;
    try {
        for (uint64_t i = 0; i < n; ++i) {
            def.clear();
            Parent p;
            if (transforms && mask_bit(mask_transform, i)) def.push_back({TY_TRANSFORM, transforms + 12 * i, 48});
            if (sprites && mask_bit(mask_sprite, i)) def.push_back({TY_SPRITE, sprites + 4 * i, 16});
            if (velocities && mask_bit(mask_velocity, i)) def.push_back({TY_VELOCITY, velocities + 3 * i, 12});
            if (parents && parents[i] != 0xFFFFFFFFu) {
                p.index = parents[i];
                def.push_back({TY_PARENT, &p, sizeof(Parent)});
            }
            if (spins && mask_bit(mask_spin, i)) def.push_back({TY_SPIN, spins + i, sizeof(Spin)});
            o->ecs.insert(def, /*faithful_insert_cost=*/n <= 4096);
        }
    } catch (const std::out_of_range&) {
        delete o;
        return nullptr;  // the reference panics: entity index beyond its presence bitset
    }
    return o;
}

// One engine tick of the velocity system: insert DeltaTime (state.rs:81), run the bundle
// (state.rs:85-87). Returns elapsed seconds, or -1 on system error.
double orc_scene_step(void* h, double dt) {
    auto* o = static_cast<OrcEcs*>(h);
    auto t0 = std::chrono::steady_clock::now();
    DeltaTime d{dt};
    o->ecs.insert_shared_resource(TY_DELTA_TIME, &d, sizeof(d));
    SystemBundle<int> bundle;
    bundle.add_system([](Ecs& ecs, int&) { return velocity_system(ecs); });
    int ad = 0;
    std::string err = bundle.step(o->ecs, ad);
    if (!err.empty()) { o->last_error = err; return -1.0; }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// One engine tick running a bundle of the built-in systems in the given order (0: velocity, 1: spin) — sequentially,
// in insertion order, as SystemBundle::step does (system.rs:17-23). Returns 0, or -1 on a system error.
int orc_scene_step_systems(void* h, double dt, const int* systems, uint32_t n) {
    auto* o = static_cast<OrcEcs*>(h);
    DeltaTime d{dt};
    o->ecs.insert_shared_resource(TY_DELTA_TIME, &d, sizeof(d));
    SystemBundle<int> bundle;
    for (uint32_t k = 0; k < n; ++k) {
        if (systems[k] == 0) bundle.add_system([](Ecs& ecs, int&) { return velocity_system(ecs); });
        else bundle.add_system([](Ecs& ecs, int&) { return spin_system(ecs); });
    }
    int ad = 0;
    std::string err = bundle.step(o->ecs, ad);
    if (!err.empty()) { o->last_error = err; return -1; }
    return 0;
}

// Copies every entity's Transform out (zeros where absent).
void orc_scene_read_transforms(void* h, float* out12n) {
    auto* o = static_cast<OrcEcs*>(h);
    uint64_t n = o->ecs.entity_count();
    std::memset(out12n, 0, n * 48);
    QueryIterator it = o->ecs.query({{TY_TRANSFORM, true, false}});
    std::vector<Guard> row;
    while (it.next(row)) std::memcpy(out12n + 12 * it.index, row[0].get<Transform>(), 48);
}

struct OrcFrame {
    FrameResult r;
};
// render_scene over the Ecs. Returns a frame handle; read it with the accessors below.
void* orc_frame_run(void* h) {
    auto* o = static_cast<OrcEcs*>(h);
    auto* f = new OrcFrame();
    render_scene(o->ecs, f->r);
    return f;
}
// The same with a view-projection matrix (row-major 16 f32) folded into the emitted geometry.
void* orc_frame_run_camera(void* h, const float* view_proj16) {
    auto* o = static_cast<OrcEcs*>(h);
    auto* f = new OrcFrame();
    Mat4f vp;
    for (int k = 0; k < 16; ++k) vp.v[k] = view_proj16[k];
    render_scene(o->ecs, f->r, &vp);
    return f;
}
void orc_orthographic(float l, float r, float b, float t, float n, float fa, float* out16) {
    const Mat4f m = new_orthographic(l, r, b, t, n, fa);
    for (int k = 0; k < 16; ++k) out16[k] = m.v[k];
}
// Forward bounds of the frame's model matrices (ref_frame.hpp forward_bounds): 16 doubles per instance.
void orc_frame_bounds(void* h, void* f, const float* view_proj16, double* out) {
    auto* o = static_cast<OrcEcs*>(h);
    auto* fr = static_cast<OrcFrame*>(f);
    Mat4f vp;
    if (view_proj16) for (int k = 0; k < 16; ++k) vp.v[k] = view_proj16[k];
    forward_bounds(o->ecs, fr->r.order, view_proj16 ? &vp : nullptr, out);
}
void orc_frame_free(void* f) { delete static_cast<OrcFrame*>(f); }
double orc_frame_seconds(void* f) { return static_cast<OrcFrame*>(f)->r.seconds; }
uint64_t orc_frame_num_instances(void* f) { return static_cast<OrcFrame*>(f)->r.instances.size(); }
uint64_t orc_frame_num_batches(void* f) { return static_cast<OrcFrame*>(f)->r.batches.size(); }
uint64_t orc_frame_num_entities(void* f) { return static_cast<OrcFrame*>(f)->r.world.size(); }
void orc_frame_read_world(void* f, float* out16n, uint8_t* has_world) {
    auto& r = static_cast<OrcFrame*>(f)->r;
    if (out16n) std::memcpy(out16n, r.world.data(), r.world.size() * 64);
    if (has_world) std::memcpy(has_world, r.has_world.data(), r.has_world.size());
}
void orc_frame_read_order(void* f, uint32_t* out) {
    auto& r = static_cast<OrcFrame*>(f)->r;
    std::memcpy(out, r.order.data(), r.order.size() * 4);
}
void orc_frame_read_keys(void* f, uint64_t* out) {
    auto& r = static_cast<OrcFrame*>(f)->r;
    std::memcpy(out, r.keys.data(), r.keys.size() * 8);
}
void orc_frame_read_instances(void* f, void* out) {
    auto& r = static_cast<OrcFrame*>(f)->r;
    std::memcpy(out, r.instances.data(), r.instances.size() * sizeof(Instance));
}
void orc_frame_read_vertices(void* f, void* out) {
    auto& r = static_cast<OrcFrame*>(f)->r;
    std::memcpy(out, r.vertices.data(), r.vertices.size() * sizeof(QuadVertex));
}
void orc_frame_read_batches(void* f, void* out) {
    auto& r = static_cast<OrcFrame*>(f)->r;
    std::memcpy(out, r.batches.data(), r.batches.size() * sizeof(Batch));
}

}  // extern "C"
