// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of one frame of the hot path, driven through the restated reference ECS
// (ref_ecs.hpp) and math (ref_math.hpp):
//
//   velocity system  ->  Transform::as_matrix4  ->  parent->child propagation  ->
//   sort key / stable order  ->  instance + quad-vertex emission  ->  batch table
//
// PARITY UNPINNED for everything in this file except the calls into ref_math/ref_ecs:
// the reference snapshot has no Sprite, no batcher, no sort, no instance or vertex format,
// no velocity system and no hierarchy propagation (`render_scene` ignores the Ecs,
// crates/tuber-graphics/src/lib.rs:134-156; `Parent` is declared and never used,
// crates/tuber-ecs/src/lib.rs:18-19). Those parts follow the builder's SPEC in DESIGN.md §3
// and are built only from reference primitives: Matrix4::mul (matrix.rs:174-194),
// transform_vec3 (matrix.rs:168-171), the column-major conversion (matrix.rs:219-251) and
// query iteration (query.rs:94-141).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <vector>

#include "ref_ecs.hpp"
#include "ref_math.hpp"

namespace ref {

// Component type ids of the built-in device-visible components.
enum : TypeId { TY_TRANSFORM = 0x7472616e73666f72ull, TY_SPRITE = 0x7370726974650000ull,
                TY_VELOCITY = 0x76656c6f63697479ull, TY_PARENT = 0x706172656e740000ull,
                TY_DELTA_TIME = 0x64656c746174696dull, TY_SPIN = 0x7370696e00000000ull };

// SPEC §3.1 component layouts.
struct Sprite {
    float size[2];
    uint32_t texture;
    uint32_t layer;
};
static_assert(sizeof(Sprite) == 16, "Sprite is 16 bytes");
struct Velocity {
    float v[3];
};
struct Parent {
    size_t index;  // tuber-ecs/src/lib.rs:18-19: Parent(pub EntityIndex)
};
struct Spin {
    float w;  // [SPEC] angular velocity about z, radians per second
};
struct DeltaTime {
    double dt;  // tuber-core/src/lib.rs:13
};

// SPEC §3.3 outputs.
struct Instance {
    float model[16];  // column-major (matrix.rs:219-251)
    float size[2];
    uint32_t texture;
    uint32_t layer;
};
static_assert(sizeof(Instance) == 80, "Instance is 80 bytes");
struct QuadVertex {
    float x, y, z;
    uint32_t uv;
};
static_assert(sizeof(QuadVertex) == 16, "QuadVertex is 16 bytes");
struct Batch {
    uint32_t layer, texture, first_instance, count;
};

// SPEC §3.2 sort key: (layer, texture, orderable bits of world z) — ties broken by entity
// index. `z + 0.0f` folds -0 into +0 so equal floats give equal keys.
inline uint64_t sort_key(uint32_t layer, uint32_t texture, float world_z) {
    float z = world_z + 0.0f;
    uint32_t u;
    std::memcpy(&u, &z, 4);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
    return (uint64_t(layer & 0xFFFFu) << 48) | (uint64_t(texture & 0xFFFFu) << 32) | uint64_t(u);
}

static const uint32_t kCornerUV[4] = {0x00000000u, 0x00000001u, 0x00010001u, 0x00010000u};

// The velocity-integration system, written the way a tuber user system is written
// (system.rs:132-137): dt comes from the DeltaTime resource the engine inserts every tick
// (tuber-engine/src/state.rs:81), cast to f32 once; translation += velocity * dt with the
// multiply and the add individually rounded.
inline std::string velocity_system(Ecs& ecs) {
    float dt;
    {
        Guard g = ecs.shared_resource(TY_DELTA_TIME, false);
        if (!g.some()) return "DeltaTime resource missing";
        dt = float(g.get<DeltaTime>()->dt);
    }
    QueryIterator it = ecs.query({{TY_TRANSFORM, true, true}, {TY_VELOCITY, true, false}});
    std::vector<Guard> row;
    while (it.next(row)) {
        Transform* t = row[0].get<Transform>();
        const Velocity* v = row[1].get<Velocity>();
        t->translation.x = t->translation.x + v->v[0] * dt;
        t->translation.y = t->translation.y + v->v[1] * dt;
        t->translation.z = t->translation.z + v->v[2] * dt;
    }
    return std::string();
}

// [SPEC] the spin system, a second user system of the same shape: angle.z += spin * dt.
inline std::string spin_system(Ecs& ecs) {
    float dt;
    {
        Guard g = ecs.shared_resource(TY_DELTA_TIME, false);
        if (!g.some()) return "DeltaTime resource missing";
        dt = float(g.get<DeltaTime>()->dt);
    }
    QueryIterator it = ecs.query({{TY_TRANSFORM, true, true}, {TY_SPIN, true, false}});
    std::vector<Guard> row;
    while (it.next(row)) {
        Transform* t = row[0].get<Transform>();
        const Spin* s = row[1].get<Spin>();
        t->angle.z = t->angle.z + s->w * dt;
    }
    return std::string();
}

struct FrameResult {
    std::vector<Mat4f> world;        // per entity index (identity-filled where no Transform)
    std::vector<uint8_t> has_world;  // 1 where world[e] is defined
    std::vector<uint32_t> order;     // entity index per sorted position
    std::vector<uint64_t> keys;      // sort key per sorted position
    std::vector<Instance> instances;
    std::vector<QuadVertex> vertices;  // 4 per instance
    std::vector<Batch> batches;
    double seconds = 0;
};

// What a batching `render_scene(&Ecs)` does with the Ecs it is handed
// (tuber-graphics/src/lib.rs:134-156 is the slot; contents per SPEC).
// `view_proj` ([SPEC]): when given, instances carry view_proj * world and the quad
// corners are transformed by it (Matrix4::mul then transform_vec3); the draw order stays by world z.
inline void render_scene(const Ecs& ecs, FrameResult& out, const Mat4f* view_proj = nullptr) {
    auto t0 = std::chrono::steady_clock::now();
    const size_t n = ecs.entity_count();
    out.world.assign(n, identity<float>());
    out.has_world.assign(n, 0);

    // 1. local matrices + parent links: query::<(&Transform, Opt<&Parent>)>()
    std::vector<size_t> parent(n, SIZE_MAX);
    {
        QueryIterator it = ecs.query({{TY_TRANSFORM, true, false}, {TY_PARENT, false, false}});
        std::vector<Guard> row;
        while (it.next(row)) {
            size_t e = it.index;
            out.world[e] = as_matrix4(*row[0].get<Transform>());
            out.has_world[e] = 1;
            if (row[1].some()) parent[e] = row[1].get<Parent>()->index;
        }
    }
    // 2. propagation: world = world[parent] * local (Matrix4::mul), parents before children.
    //    A parent without a Transform (or out of range) is treated as no parent.
    {
        std::vector<uint8_t> done(n, 0);
        std::vector<size_t> chain;
        for (size_t e = 0; e < n; ++e) {
            if (!out.has_world[e] || done[e]) continue;
            chain.clear();
            size_t cur = e;
            while (true) {
                chain.push_back(cur);
                size_t p = parent[cur];
                if (p == SIZE_MAX || p >= n || !out.has_world[p]) break;
                if (done[p]) break;
                if (chain.size() > n) break;  // cycle guard; SPEC forbids cycles
                cur = p;
            }
            for (size_t k = chain.size(); k-- > 0;) {
                size_t c = chain[k];
                size_t p = parent[c];
                if (p != SIZE_MAX && p < n && out.has_world[p] && done[p]) out.world[c] = mul(out.world[p], out.world[c]);
                done[c] = 1;
            }
        }
    }
    // 3. renderables: query::<(&Transform, &Sprite)>() -> key; order by (key, entity).
    struct Item {
        uint64_t key;
        uint32_t entity;
        Sprite sprite;
    };
    std::vector<Item> items;
    {
        QueryIterator it = ecs.query({{TY_TRANSFORM, true, false}, {TY_SPRITE, true, false}});
        std::vector<Guard> row;
        while (it.next(row)) {
            const Sprite* s = row[1].get<Sprite>();
            size_t e = it.index;
            items.push_back(Item{sort_key(s->layer, s->texture, out.world[e].v[11]), uint32_t(e), *s});
        }
    }
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
        return a.key != b.key ? a.key < b.key : a.entity < b.entity;
    });
    // 4. emit instances, quad vertices, batches in sorted order.
    const size_t m = items.size();
    out.order.resize(m);
    out.keys.resize(m);
    out.instances.resize(m);
    out.vertices.resize(4 * m);
    out.batches.clear();
    for (size_t pos = 0; pos < m; ++pos) {
        const Item& it = items[pos];
        const Mat4f w = view_proj ? mul(*view_proj, out.world[it.entity]) : out.world[it.entity];
        out.order[pos] = it.entity;
        out.keys[pos] = it.key;
        Instance& inst = out.instances[pos];
        to_column_major(w, inst.model);
        inst.size[0] = it.sprite.size[0];
        inst.size[1] = it.sprite.size[1];
        inst.texture = it.sprite.texture;
        inst.layer = it.sprite.layer;
        const float cw = it.sprite.size[0], ch = it.sprite.size[1];
        const Vec3 corners[4] = {{0, 0, 0}, {cw, 0, 0}, {cw, ch, 0}, {0, ch, 0}};
        for (int k = 0; k < 4; ++k) {
            Vec3 p = transform_vec3(w, corners[k]);
            out.vertices[4 * pos + k] = QuadVertex{p.x, p.y, p.z, kCornerUV[k]};
        }
        if (out.batches.empty() || out.batches.back().layer != it.sprite.layer ||
            out.batches.back().texture != it.sprite.texture) {
            out.batches.push_back(Batch{it.sprite.layer, it.sprite.texture, uint32_t(pos), 0});
        }
        out.batches.back().count += 1;
    }
    out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- forward error bounds (the parity metric's denominator; not part of the frame) ----------------------
//
// For every emitted instance, the entry-wise forward bound of its model matrix: the same chain of products
// as render_scene, evaluated on ABSOLUTE values in double — i.e. for every entry the sum of the magnitudes of
// all the terms that are added up to produce it (an entry that is a cancelling sum, such as a
// world translation near 0, is judged against the size of its terms, not against its own tiny value).
// out: 16 doubles per instance, row-major, in draw order (`order`).
struct AbsMat {
    double v[16];
};
inline AbsMat abs_of(const Mat4f& m) {
    AbsMat r;
    for (int k = 0; k < 16; ++k) r.v[k] = std::fabs(double(m.v[k]));
    return r;
}
inline AbsMat abs_mul(const AbsMat& a, const AbsMat& b) {
    AbsMat r;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a.v[j * 4 + k] * b.v[k * 4 + i];
            r.v[j * 4 + i] = s;
        }
    return r;
}
inline void forward_bounds(const Ecs& ecs, const std::vector<uint32_t>& order, const Mat4f* view_proj, double* out) {
    const size_t n = ecs.entity_count();
    std::vector<AbsMat> bound(n);
    std::vector<uint8_t> has(n, 0), done(n, 0);
    std::vector<size_t> parent(n, SIZE_MAX);
    {
        QueryIterator it = ecs.query({{TY_TRANSFORM, true, false}, {TY_PARENT, false, false}});
        std::vector<Guard> row;
        while (it.next(row)) {
            const size_t e = it.index;
            const Transform& t = *row[0].get<Transform>();
            AbsMat m = abs_mul(abs_of(new_scale(t.scale)), abs_of(new_translation(t.translation)));
            m = abs_mul(m, abs_of(new_translation(t.rotation_center)));
            m = abs_mul(m, abs_of(rotation_matrix(from_euler(t.angle))));
            m = abs_mul(m, abs_of(new_translation(neg(t.rotation_center))));
            bound[e] = m;
            has[e] = 1;
            if (row[1].some()) parent[e] = row[1].get<Parent>()->index;
        }
    }
    std::vector<size_t> chain;
    for (size_t e = 0; e < n; ++e) {
        if (!has[e] || done[e]) continue;
        chain.clear();
        size_t cur = e;
        while (true) {
            chain.push_back(cur);
            const size_t p = parent[cur];
            if (p == SIZE_MAX || p >= n || !has[p] || done[p] || chain.size() > n) break;
            cur = p;
        }
        for (size_t k = chain.size(); k-- > 0;) {
            const size_t c = chain[k], p = parent[c];
            if (p != SIZE_MAX && p < n && has[p] && done[p]) bound[c] = abs_mul(bound[p], bound[c]);
            done[c] = 1;
        }
    }
    for (size_t pos = 0; pos < order.size(); ++pos) {
        const AbsMat b = view_proj ? abs_mul(abs_of(*view_proj), bound[order[pos]]) : bound[order[pos]];
        for (int k = 0; k < 16; ++k) out[16 * pos + k] = b.v[k];
    }
}

}  // namespace ref
