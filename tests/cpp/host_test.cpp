// C++ host-side tests: the reference's own ECS / system unit tests (crates/tuber-ecs/src/ecs.rs:217-310,
// system.rs:78-176) replayed against include/tuber_host.hpp (same entities, same assertions), then
// render_scene parity of the C++ host path against the CPU oracle (oracle/ref_frame.hpp).
// Needs a B200. Prints "ok <name>" per test; exit code 0 iff all passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>

#include "../../include/tuber_host.hpp"
#include "../../oracle/ref_frame.hpp"

using namespace tuber;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            ++failures;                                                      \
        }                                                                    \
    } while (0)

struct Position { float x, y; bool operator==(const Position& o) const { return x == o.x && y == o.y; } };
struct Vel2 { float x, y; };  // the reference tests' own host-only `Velocity { x, y }`

static void ecs_new_and_insert() {  // ecs.rs:235-248
    Ecs ecs(1024);
    CHECK(ecs.entity_count() == 0);
    ecs.insert(Position{0, 1}, Vel2{2, 3});
    CHECK(ecs.entity_count() == 1);
    ecs.insert(Position{4, 5}, Vel2{6, 7});
    CHECK(ecs.entity_count() == 2);
    std::puts("ok ecs_new_and_insert");
}

static void ecs_query() {  // ecs.rs:250-267
    Ecs ecs(1024);
    ecs.insert(Position{12, 1}, Vel2{2, 3});
    ecs.insert(Position{4, 5}, Vel2{6, 7});
    ecs.insert(Position{4, 5});
    for (EntityIndex e : ecs.query<Vel2>()) ecs.get_mut<Vel2>(e)->x = 0.0f;
    for (EntityIndex e : ecs.query<Vel2>()) CHECK(std::fabs(ecs.get<Vel2>(e)->x) <= 0.01f);
    CHECK(ecs.query<Position>().size() == 3);
    CHECK(ecs.query<Vel2>().size() == 2);
    std::puts("ok ecs_query");
}

static void ecs_query_one_and_optional() {  // ecs.rs:269-309
    Ecs ecs(1024);
    ecs.insert(Position{12, 1}, Vel2{2, 3});
    CHECK(ecs.query_one<Position>().value() == 0);
    CHECK(*ecs.get<Position>(*ecs.query_one<Position>()) == (Position{12, 1}));
    ecs.insert(Position{15, 8});
    CHECK((ecs.query<Position, Opt<Vel2>>().size() == 2));  // Opt never filters
    Ecs e2(1024);
    e2.insert(Position{12, 1});
    auto one = e2.query_one<Position, Opt<Vel2>>();
    CHECK(one.has_value() && *e2.get<Position>(*one) == (Position{12, 1}) && e2.get<Vel2>(*one) == nullptr);
    std::puts("ok ecs_query_one_and_optional");
}

static void systems() {  // system.rs:96-175
    struct Value { int v; };
    struct OtherComponent {};
    Ecs ecs(1024);
    ecs.insert(Value{12});
    ecs.insert(Value{18}, OtherComponent{});
    SystemBundle<int> bundle;
    bundle.add_system(std::function<SystemResult(Ecs&)>([](Ecs& e) {
        for (EntityIndex i : e.query<Value>()) e.get_mut<Value>(i)->v += 35;
        return Ok();
    }));
    CHECK(bundle.size() == 1);  // system_bundle_add
    bundle.add_system(std::function<SystemResult(Ecs&)>([](Ecs& e) {
        for (EntityIndex i : e.query<Value>()) e.get_mut<Value>(i)->v -= 6;
        return Ok();
    }));
    int ad = 0;
    CHECK(!bundle.step(ecs, ad).has_value());
    std::set<int> vals;
    for (EntityIndex i : ecs.query<Value>()) vals.insert(ecs.get<Value>(i)->v);
    CHECK(vals.count(41) && vals.count(47));  // system_bundle_step
    // failing_system: Err propagates and stops the bundle
    SystemBundle<int> failing;
    bool ran_after = false;
    failing.add_system(std::function<SystemResult(Ecs&)>([](Ecs&) { return Err("ATROCIOUS ERROR"); }));
    failing.add_system(std::function<void(Ecs&)>([&](Ecs&) { ran_after = true; }));
    auto r = failing.step(ecs, ad);
    CHECK(r.has_value() && *r == "ATROCIOUS ERROR" && !ran_after);
    // system_bundle_with_additional_data
    SystemBundle<int> with_ad;
    with_ad.add_system(std::function<void(Ecs&, int&)>([](Ecs&, int& d) { d += 1; }));
    with_ad.step(ecs, ad);
    with_ad.step(ecs, ad);
    CHECK(ad == 2);
    std::puts("ok systems");
}

static void source_only_semantics() {
    Ecs ecs(128);
    for (int k = 0; k < 10; ++k) {
        if (k % 3) ecs.insert(Transform{});
        else ecs.insert(Transform{}, Sprite{1, 1, 0, 0});
    }
    // descending iteration (query.rs:138), on the device path (both types are device-visible)
    auto t = ecs.query<Transform>();
    CHECK(t.size() == 10 && t.front() == 9 && t.back() == 0);
    auto ts = ecs.query<Transform, Sprite>();
    CHECK((ts == std::vector<EntityIndex>{9, 6, 3, 0}));
    CHECK(ecs.query_one<Sprite>().value() == 0);
    CHECK(ecs.query<Velocity>().empty());                       // query.rs:110: no store -> nothing
    ecs.add_component(Velocity{1, 2, 3}, 4);                    // ecs.rs:120-124: no store yet -> no-op
    CHECK(ecs.get<Velocity>(4) == nullptr);
    ecs.remove_component<Sprite>(3);
    CHECK((ecs.query<Transform, Sprite>() == std::vector<EntityIndex>{9, 6, 0}));
    const tuber::Entity h9 = ecs.handle(9), h6 = ecs.handle(6);
    CHECK(ecs.alive(h9) && ecs.resolve(h9).value() == 9 && ecs.get<Transform>(h9) != nullptr);
    ecs.delete_by_ids({9});
    CHECK(!ecs.alive(h9) && !ecs.resolve(h9).has_value() && ecs.get<Transform>(h9) == nullptr);  // stale handle
    CHECK(ecs.alive(h6) && ecs.handle(9).generation == h9.generation + 1);
    CHECK((ecs.query<Transform, Sprite>() == std::vector<EntityIndex>{6, 0}));
    CHECK(ecs.entity_count() == 10);                            // indices are never reused
    CHECK((ecs.query_by_ids<Transform>({0, 1, 9, 42}) == std::vector<EntityIndex>{1, 0}));
    ecs.delete_by_query<Sprite>();
    CHECK(ecs.query<Sprite>().empty() && ecs.query<Transform>().size() == 7);
    // the presence-bitset cap (ecs.rs:14): the reference panics, the host mirror throws
    bool threw = false;
    try {
        for (int k = 0; k < 200; ++k) ecs.insert(Transform{});
    } catch (const Error& e) { threw = e.status == TB_ERR_CAPACITY; }
    CHECK(threw && ecs.entity_count() == 128);
    std::puts("ok source_only_semantics");
}

// Query items as the reference yields them (query.rs:27, 131-141) and the RefCell borrow rules (query.rs:170-199).
static void query_items_and_borrows() {
    Ecs ecs(64);
    for (int k = 0; k < 6; ++k) {
        Transform t;
        t.translation = {float(k), 0, 0};
        if (k % 2) ecs.insert(t, Velocity{1, 2, 0});
        else ecs.insert(t, Velocity{1, 2, 0}, Sprite{4, 4, 0, 0});
    }
    // (EntityIndex, (RefMut<Transform>, Ref<Velocity>, Option<Ref<Sprite>>)), descending
    std::vector<EntityIndex> seen;
    int with_sprite = 0;
    for (auto it = ecs.query_items<tuber::Mut<Transform>, Velocity, tuber::Opt<Sprite>>(); auto item = it.next();) {
        auto& [e, row] = *item;
        auto& [t, v, s] = row;
        t->translation.x += v->x;  // the idiom of system.rs:132-137
        t->translation.y += v->y;
        with_sprite += s.has_value();
        seen.push_back(e);
    }
    CHECK((seen == std::vector<EntityIndex>{5, 4, 3, 2, 1, 0}) && with_sprite == 3);
    CHECK(ecs.get<Transform>(2)->translation.x == 3.0f && ecs.get<Transform>(2)->translation.y == 2.0f);
    // the written Transforms reach the device with the next device call: a frame sees them
    tuber::Graphics gfx;
    CHECK(!gfx.render_scene(ecs).has_value());
    auto inst = gfx.instances();
    auto order = gfx.order();
    for (size_t p = 0; p < order.size(); ++p) CHECK(inst[p].model[12] == float(order[p]) + 1.0f);  // column-major translation.x
    // two shared borrows coexist; an exclusive one conflicts with either kind (RefCell: "already borrowed")
    {
        auto a = ecs.query_one_by_id<Transform>(3);
        auto b = ecs.query_one_by_id<Transform>(3);
        CHECK(a.has_value() && b.has_value());
        bool threw = false;
        try {
            auto c = ecs.query_one_by_id<tuber::Mut<Transform>>(3);
        } catch (const tuber::BorrowError&) { threw = true; }
        CHECK(threw);
    }
    {
        auto a = ecs.query_one_by_id<tuber::Mut<Transform>>(3);  // released guards: exclusive borrow succeeds
        CHECK(a.has_value());
        bool threw = false;
        try {
            auto b = ecs.query_one_by_id<Transform>(3);
        } catch (const tuber::BorrowError&) { threw = true; }
        CHECK(threw);
        auto other = ecs.query_one_by_id<tuber::Mut<Transform>>(4);  // another entity's slot is another RefCell
        CHECK(other.has_value());
    }
    CHECK(!ecs.query_one_by_id<Sprite>(1).has_value());  // entity 1 has no Sprite
    std::puts("ok query_items_and_borrows");
}

// A second device system and the bundle scheduler through the host API.
static void device_bundle() {
    Ecs ecs(64);
    for (int k = 0; k < 8; ++k) {
        Transform t;
        t.angle = {0, 0, 0.25f * k};
        ecs.insert(t, Sprite{4, 4, 0, 0}, Velocity{1, 0, 0}, tuber::Spin{2.0f});
    }
    ecs.insert_shared_resource(tuber::DeltaTime{0.5});
    std::vector<uint32_t> stages;
    CHECK(!tuber::device_bundle_step(ecs, {TB_SYSTEM_VELOCITY, TB_SYSTEM_SPIN}, &stages).has_value());
    CHECK((stages == std::vector<uint32_t>{0, 0}));  // disjoint access sets: one stage
    CHECK(ecs.get<Transform>(3)->translation.x == 0.5f);
    CHECK(ecs.get<Transform>(3)->angle.z == 0.25f * 3 + 2.0f * 0.5f);
    tuber::SystemBundle<int> bundle;
    bundle.add_system(std::function<tuber::SystemResult(Ecs&)>(tuber::spin_system));
    bundle.add_system(std::function<tuber::SystemResult(Ecs&)>(tuber::velocity_system));
    int ad = 0;
    CHECK(!bundle.step(ecs, ad).has_value());
    tuber::Graphics gfx;
    CHECK(!gfx.render_scene(ecs).has_value());  // the queued velocity tick is consumed by the frame
    CHECK(ecs.get<Transform>(3)->translation.x == 1.0f && ecs.get<Transform>(3)->angle.z == 0.25f * 3 + 2.0f);
    std::puts("ok device_bundle");
}

static bool close(float g, float r, float scale) { return std::fabs(g - r) <= 1e-5f * std::max(std::fabs(r), scale); }

struct Gen {
    Transform t;
    Sprite s;
    Velocity v;
    bool has_sprite, has_vel, has_parent;
    int parent_k;
};

static Gen gen(std::mt19937& rng, int i, bool hierarchy) {
    std::uniform_real_distribution<float> U(0.f, 1.f);
    Gen g;
    g.t.translation = {U(rng) * 200 - 100, U(rng) * 200 - 100, float(int(U(rng) * 4))};
    g.t.angle = {0, 0, U(rng) * 6.28f};
    g.t.scale = {0.9f + 0.2f * U(rng), 0.9f + 0.2f * U(rng), 1};
    g.s = Sprite{8 + 56 * U(rng), 8 + 56 * U(rng), uint32_t(U(rng) * 8), uint32_t(U(rng) * 3)};
    g.t.rotation_center = {g.s.width * 0.5f, g.s.height * 0.5f, 0};
    g.v = Velocity{U(rng) * 20 - 10, U(rng) * 20 - 10, 0};
    g.has_sprite = i % 7 != 3;
    g.has_vel = i % 2 == 0;
    g.has_parent = hierarchy && i >= 50;
    g.parent_k = (i - 1) / 4;
    return g;
}

// One engine tick + render through the C++ host API vs the oracle, entity by entity.
// add_component is a no-op until a store exists (ecs.rs:120-124), so entity 1 is a throw-away entity
// whose insert tuple creates the four stores and which is deleted again: a hole, in both worlds.
static void frame_parity(bool hierarchy) {
    const int n = 20000;
    auto id_of = [](int k) { return size_t(k == 0 ? 0 : k + 1); };  // ids shift by one after the hole
    Ecs ecs2(1 << 16);
    ref::Ecs oracle2((n + 1 + 63) / 64 + 1);
    std::mt19937 rng(hierarchy ? 7 : 5);
    for (int i = 0; i < n; ++i) {
        Gen g = gen(rng, i, hierarchy);
        Parent p{id_of(g.parent_k)};
        ref::Parent rp{id_of(g.parent_k)};
        EntityIndex e = ecs2.insert(g.t);
        CHECK(e == id_of(i));
        std::vector<ref::ComponentValue> def{{ref::TY_TRANSFORM, &g.t, 48}};
        if (g.has_sprite) def.push_back({ref::TY_SPRITE, &g.s, 16});
        if (g.has_vel) def.push_back({ref::TY_VELOCITY, &g.v, 12});
        if (g.has_parent) def.push_back({ref::TY_PARENT, &rp, sizeof(rp)});
        oracle2.insert(def, false);
        if (i == 0) {
            ecs2.insert(g.t, g.s, g.v, p);
            ecs2.delete_by_ids({1});
            oracle2.insert({}, false);
        }
        if (g.has_sprite) ecs2.add_component(g.s, e);
        if (g.has_vel) ecs2.add_component(g.v, e);
        if (g.has_parent) ecs2.add_component(p, e);
    }
    // tick: DeltaTime resource then the bundle (state.rs:81-87)
    ecs2.insert_shared_resource(DeltaTime{0.01});
    SystemBundle<int> bundle;
    bundle.add_system(std::function<SystemResult(Ecs&)>(velocity_system));
    int ad = 0;
    CHECK(!bundle.step(ecs2, ad).has_value());
    ref::DeltaTime d{0.01};
    oracle2.insert_shared_resource(ref::TY_DELTA_TIME, &d, sizeof(d));
    CHECK(ref::velocity_system(oracle2).empty());
    // render
    Graphics gfx;
    auto res = gfx.render_scene(ecs2);
    CHECK(!res.has_value());
    ref::FrameResult want;
    ref::render_scene(oracle2, want);
    auto order = gfx.order();
    auto inst = gfx.instances();
    auto vtx = gfx.vertices();
    auto batches = gfx.batches();
    CHECK(order.size() == want.order.size());
    CHECK(batches.size() == want.batches.size());
    CHECK(gfx.frame().hierarchy_levels == (hierarchy ? 8u : 0u) || hierarchy);
    bool ints_ok = order.size() == want.order.size(), floats_ok = true;
    // Hierarchy: a world translation is a sum of ancestors' terms that can cancel to ~0, so element
    // errors are judged against the scene extent (largest world-matrix entry), not the element alone.
    float extent = 0;
    if (hierarchy)
        for (const auto& in : want.instances)
            for (int k = 0; k < 16; ++k) extent = std::max(extent, std::fabs(in.model[k]));
    for (size_t p = 0; ints_ok && p < order.size(); ++p) {
        ints_ok = order[p] == want.order[p] && inst[p].texture == want.instances[p].texture &&
                  inst[p].layer == want.instances[p].layer;
        float scale = extent;
        for (int k = 0; k < 16; ++k) scale = std::max(scale, std::fabs(want.instances[p].model[k]));
        for (int k = 0; k < 16; ++k) {
            const bool ok = close(inst[p].model[k], want.instances[p].model[k], scale);
            if (!ok && floats_ok)
                std::printf("first float mismatch: position %zu entity %u model[%d] gpu %.9g ref %.9g (matrix scale %.6g)\n", p,
                            order[p], k, inst[p].model[k], want.instances[p].model[k], scale);
            floats_ok = floats_ok && ok;
        }
        for (int k = 0; k < 4; ++k) {
            const auto& g = vtx[4 * p + k];
            const auto& r = want.vertices[4 * p + k];
            floats_ok = floats_ok && close(g.x, r.x, scale) && close(g.y, r.y, scale) && close(g.z, r.z, scale) && g.uv == r.uv;
        }
    }
    CHECK(ints_ok);
    CHECK(floats_ok);
    for (size_t b = 0; b < batches.size() && b < want.batches.size(); ++b)
        CHECK(std::memcmp(&batches[b], &want.batches[b], sizeof(tb_batch)) == 0);
    // the host mirror sees the device system's writes (lazy refresh), bit for bit
    ref::QueryIterator it = oracle2.query({{ref::TY_TRANSFORM, true, false}});
    std::vector<ref::Guard> row;
    bool same = true;
    while (it.next(row)) same = same && std::memcmp(ecs2.get<Transform>(it.index), row[0].get<ref::Transform>(), 48) == 0;
    CHECK(same);
    // camera: Graphics::set_orthographic folds Matrix4::new_orthographic (matrix.rs:41-58) into the next frame
    {
        gfx.set_orthographic(-1200.0f, 1200.0f, -900.0f, 900.0f, -100.0f, 100.0f);
        CHECK(!gfx.render_scene(ecs2).has_value());
        const ref::Mat4f vp = ref::new_orthographic(-1200.0f, 1200.0f, -900.0f, 900.0f, -100.0f, 100.0f);
        ref::FrameResult cam;
        ref::render_scene(oracle2, cam, &vp);
        auto order2 = gfx.order();
        auto inst2 = gfx.instances();
        bool ok = order2.size() == cam.order.size();
        for (size_t p = 0; ok && p < order2.size(); ++p) {
            ok = order2[p] == cam.order[p];
            for (int k = 0; ok && k < 16; ++k) ok = close(inst2[p].model[k], cam.instances[p].model[k], 1.0f);  // clip space: the matrix holds a 1
        }
        CHECK(ok);
        gfx.clear_camera();
    }
    std::printf("ok frame_parity(%s): %zu instances, %zu batches\n", hierarchy ? "hierarchy" : "flat", order.size(), batches.size());
}

int main() {
    try {
        ecs_new_and_insert();
        ecs_query();
        ecs_query_one_and_optional();
        systems();
        source_only_semantics();
        query_items_and_borrows();
        device_bundle();
        frame_parity(false);
        frame_parity(true);
    } catch (const std::exception& e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf("%s (%d failures)\n", failures ? "FAILED" : "ALL OK", failures);
    return failures ? 1 : 0;
}
