// tuber_host.hpp — C++17 host side above the C ABI (include/tuber_b200.h).
//
// The reference's host code is Rust; no Rust toolchain exists in this image, so this header is
// the host-language mirror of the reference interface for the hot path: the same names, argument
// meaning and error behaviour as
//
//   tuber_ecs::ecs::Ecs                    crates/tuber-ecs/src/ecs.rs:56-181
//   tuber_ecs::query (iteration order, Opt) crates/tuber-ecs/src/query.rs:94-141, 226-245
//   tuber_ecs::system::SystemBundle        crates/tuber-ecs/src/system.rs:8-24
//   tuber_core::transform::Transform       crates/tuber-core/src/transform.rs:5-38
//   tuber_graphics::GraphicsAPI            crates/tuber-graphics/src/lib.rs:34-36
//
// Any `'static` type can be a component (type-erased host columns, like the reference). Four
// component types are *device-visible* — Transform, Sprite, Velocity, Parent — and are mirrored
// into the tb_ctx columns: structural changes and writes mark index ranges dirty, and the next
// device call (a device system, render_scene, a device-side query) pushes exactly those ranges
// with tb_upload_* / tb_upload_presence. After a device system has moved Transforms the host copy
// is refreshed lazily from the device the next time it is read.
//
// Header-only; link with -ltuber_b200. No CPU fallback: Ecs's constructor throws without a B200.
#ifndef TUBER_HOST_HPP
#define TUBER_HOST_HPP

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <typeindex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "tuber_b200.h"

namespace tuber {

using EntityIndex = std::size_t;  // tuber-ecs/src/lib.rs:16

// Generation-checked entity handle. The reference hands out bare indices and never reuses one
// (ecs.rs:94-97), so an index cannot come to name ANOTHER entity, but it can outlive its own
// (delete_by_ids / delete_by_query). A handle pairs the index with the generation it was issued at;
// deleting an entity bumps its slot's generation, so stale handles stop resolving. Plain indices keep
// working everywhere, exactly as in the reference.
struct Entity {
    EntityIndex index;
    uint32_t generation;
    bool operator==(const Entity& o) const { return index == o.index && generation == o.generation; }
};

struct Vector3f {
    float x = 0, y = 0, z = 0;
};

// tuber_core::transform::Transform (transform.rs:5-22): Default is identity.
struct Transform {
    Vector3f translation{0, 0, 0};
    Vector3f angle{0, 0, 0};
    Vector3f rotation_center{0, 0, 0};
    Vector3f scale{1, 1, 1};
};
static_assert(sizeof(Transform) == sizeof(tb_transform), "Transform is the 48-byte reference layout");

struct Sprite {  // [SPEC]
    float width = 0, height = 0;
    uint32_t texture = 0, layer = 0;
};
static_assert(sizeof(Sprite) == sizeof(tb_sprite), "Sprite is 16 bytes");

struct Velocity {  // [SPEC]
    float x = 0, y = 0, z = 0;
};
static_assert(sizeof(Velocity) == sizeof(tb_velocity), "Velocity is 12 bytes");

struct Parent {  // tuber-ecs/src/lib.rs:18-19
    EntityIndex index = 0;
};

struct DeltaTime {  // tuber-core/src/lib.rs:13
    double dt = 0;
};

// Opt<T> accessor marker (query.rs:226-245): never filters.
template <class T>
struct Opt {};
// `&mut T` accessor marker (query.rs:181-208); a bare T is `&T` (query.rs:159-179).
template <class T>
struct Mut {};

struct Spin {  // [SPEC] angular velocity about z, radians per second
    float w = 0;
};

// The reference keeps every component in a RefCell (ecs.rs:17): a second exclusive borrow, or a shared one while an
// exclusive one is alive, PANICS ("already borrowed", query.rs:170-199). Here it throws.
struct BorrowError : std::logic_error {
    using std::logic_error::logic_error;
};

// Ref<T> / RefMut<T>: what the accessors of a query hand out (query.rs:27, 166-208). They hold the slot's borrow flag
// like std::cell::Ref / RefMut and release it when they go out of scope.
template <class T>
class Ref {
public:
    Ref(const T* p, int32_t* flag) : p_(p), flag_(flag) {
        if (*flag_ < 0) throw BorrowError("already mutably borrowed");
        ++*flag_;
    }
    Ref(Ref&& o) noexcept : p_(o.p_), flag_(o.flag_) { o.flag_ = nullptr; }
    Ref(const Ref&) = delete;
    Ref& operator=(const Ref&) = delete;
    ~Ref() {
        if (flag_) --*flag_;
    }
    const T& operator*() const { return *p_; }
    const T* operator->() const { return p_; }

private:
    const T* p_;
    int32_t* flag_;
};
template <class T>
class RefMut {
public:
    RefMut(T* p, int32_t* flag, std::function<void()> on_drop) : p_(p), flag_(flag), on_drop_(std::move(on_drop)) {
        if (*flag_ != 0) throw BorrowError("already borrowed");
        *flag_ = -1;
    }
    RefMut(RefMut&& o) noexcept : p_(o.p_), flag_(o.flag_), on_drop_(std::move(o.on_drop_)) { o.flag_ = nullptr; }
    RefMut(const RefMut&) = delete;
    RefMut& operator=(const RefMut&) = delete;
    ~RefMut() {
        if (flag_) {
            *flag_ = 0;
            if (on_drop_) on_drop_();  // a written device-visible component is pushed by the next device call
        }
    }
    T& operator*() const { return *p_; }
    T* operator->() const { return p_; }

private:
    T* p_;
    int32_t* flag_;
    std::function<void()> on_drop_;
};

struct Error : std::runtime_error {
    tb_status status;
    Error(tb_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};

namespace detail {

template <class T> struct device_component { static constexpr int id = -1; };
template <> struct device_component<Transform> { static constexpr int id = TB_TRANSFORM; };
template <> struct device_component<Sprite> { static constexpr int id = TB_SPRITE; };
template <> struct device_component<Velocity> { static constexpr int id = TB_VELOCITY; };
template <> struct device_component<Parent> { static constexpr int id = TB_PARENT; };
template <> struct device_component<Spin> { static constexpr int id = TB_SPIN; };

template <class T> struct accessor { using type = T; static constexpr bool required = true; static constexpr bool mut = false; using guard = Ref<T>; };
template <class T> struct accessor<Mut<T>> { using type = T; static constexpr bool required = true; static constexpr bool mut = true; using guard = RefMut<T>; };
template <class T> struct accessor<Opt<T>> { using type = T; static constexpr bool required = false; static constexpr bool mut = false; using guard = std::optional<Ref<T>>; };
template <class T> struct accessor<Opt<Mut<T>>> { using type = T; static constexpr bool required = false; static constexpr bool mut = true; using guard = std::optional<RefMut<T>>; };

// One type-erased component column: ComponentStore (ecs.rs:16-53).
struct Store {
    std::vector<std::shared_ptr<void>> data;  // nullptr == None
    std::vector<uint64_t> bits;               // entities_bitset, bit layout of bitset.rs:15-36
    mutable std::vector<int32_t> borrow;      // RefCell flag per slot: n > 0 shared borrows, -1 one exclusive borrow
    int32_t* flag(EntityIndex e) const {
        if (e >= borrow.size()) borrow.resize(e + 1, 0);
        return &borrow[e];
    }
    bool test(EntityIndex e) const { return e / 64 < bits.size() && ((bits[e / 64] >> (e % 64)) & 1u); }
    void set(EntityIndex e) {
        if (e / 64 >= bits.size()) bits.resize(e / 64 + 1, 0);
        bits[e / 64] |= uint64_t(1) << (e % 64);
    }
    void unset(EntityIndex e) {
        if (e / 64 < bits.size()) bits[e / 64] &= ~(uint64_t(1) << (e % 64));
    }
};

struct DirtyRange {
    EntityIndex lo = SIZE_MAX, hi = 0;  // [lo, hi)
    void mark(EntityIndex e) {
        lo = std::min(lo, e);
        hi = std::max(hi, e + 1);
    }
    bool any() const { return lo < hi; }
    void clear() {
        lo = SIZE_MAX;
        hi = 0;
    }
};

}  // namespace detail

class Ecs {
public:
    // The reference's presence bitset is a fixed [u64; 1024] (ecs.rs:14): 65 536 entities, beyond
    // which insert panics. `max_entities` is that cap here (insert throws TB_ERR_CAPACITY).
    explicit Ecs(uint64_t max_entities = 65536, int device = 0) : cap_(max_entities) {
        tb_status s = tb_ctx_create(device, max_entities, &ctx_);
        if (s != TB_OK) throw Error(s, std::string("tb_ctx_create: ") + tb_status_string(s));
    }
    ~Ecs() { tb_ctx_destroy(ctx_); }
    Ecs(const Ecs&) = delete;
    Ecs& operator=(const Ecs&) = delete;

    // ecs.rs:93-98 — returns the new entity's index; indices are never reused.
    template <class... Cs>
    EntityIndex insert(Cs... components) {
        if (next_ >= cap_) throw Error(TB_ERR_CAPACITY, "entity index beyond the presence bitset (the reference panics)");
        const EntityIndex e = next_++;
        generation_.push_back(0);
        for (auto& kv : stores_) kv.second.data.push_back(nullptr);  // ecs.rs:194-196
        (put(e, std::move(components)), ...);
        count_dirty_ = true;
        return e;
    }
    std::size_t entity_count() const { return next_; }  // ecs.rs:178-180

    // ecs.rs:120-124 — silently does nothing when no store for the type exists yet.
    template <class C>
    void add_component(C c, EntityIndex e) {
        auto it = stores_.find(std::type_index(typeid(C)));
        if (it == stores_.end() || e >= next_) return;
        write(it->second, e, std::move(c));
    }
    // ecs.rs:114-118
    template <class C>
    void remove_component(EntityIndex e) {
        auto it = stores_.find(std::type_index(typeid(C)));
        if (it == stores_.end() || e >= next_) return;
        erase(it->first, it->second, e);
    }
    // ecs.rs:105-112
    void delete_by_ids(const std::vector<EntityIndex>& ids) {
        for (EntityIndex e : ids) {
            if (e >= next_) continue;
            for (auto& kv : stores_) erase(kv.first, kv.second, e);
            generation_[e] += 1;  // handles issued before the delete no longer resolve
        }
    }

    // Generation-checked handles over the never-reused indices.
    Entity handle(EntityIndex e) const {
        if (e >= next_) throw Error(TB_ERR_INVALID, "handle(): no such entity index");
        return Entity{e, generation_[e]};
    }
    bool alive(Entity h) const { return h.index < next_ && generation_[h.index] == h.generation; }
    std::optional<EntityIndex> resolve(Entity h) const {
        return alive(h) ? std::optional<EntityIndex>(h.index) : std::nullopt;
    }
    template <class C>
    const C* get(Entity h) {
        return alive(h) ? get<C>(h.index) : nullptr;
    }
    // ecs.rs:100-103
    template <class... As>
    void delete_by_query() { delete_by_ids(query<As...>()); }

    // ecs.rs:127-129 + query.rs:94-141 — the matching entity indices IN THE REFERENCE'S ITERATION
    // ORDER (descending: QueryIterator pops matches from the back). Opt<T> accessors never filter;
    // a required component without a store matches nothing. When every accessor is device-visible
    // the presence AND and the compaction run on the GPU (tb_query).
    template <class... As>
    std::vector<EntityIndex> query() {
        std::vector<const detail::Store*> req;
        uint32_t device_bits = 0;
        bool all_device = true, missing = false;
        (collect<As>(req, device_bits, all_device, missing), ...);
        std::vector<EntityIndex> out;
        if (missing) return out;
        if (all_device && next_ > 0) {
            flush();
            std::vector<uint32_t> idx(next_);
            uint64_t count = 0;
            check(tb_query(ctx_, device_bits, idx.data(), idx.size(), &count), "tb_query");
            out.resize(count);
            for (uint64_t k = 0; k < count; ++k) out[k] = idx[count - 1 - k];
            return out;
        }
        for (EntityIndex e = next_; e-- > 0;) {
            bool all = true;
            for (auto* s : req) all = all && s->test(e);
            if (all) out.push_back(e);
        }
        return out;
    }
    // ecs.rs:137-169 — the LOWEST matching index.
    template <class... As>
    std::optional<EntityIndex> query_one() {
        auto ids = query<As...>();
        if (ids.empty()) return std::nullopt;
        return ids.back();
    }
    // ecs.rs:132-134 — matches whose index is in `ids`, still descending.
    template <class... As>
    std::vector<EntityIndex> query_by_ids(const std::unordered_set<EntityIndex>& ids) {
        std::vector<EntityIndex> out;
        for (EntityIndex e : query<As...>())
            if (ids.count(e)) out.push_back(e);
        return out;
    }

    // The query as the reference yields it (query.rs:27, 131-141): items `(EntityIndex, (A::RefType, ...))` in descending
    // index order, every accessor a Ref / RefMut / optional guard with RefCell borrow semantics:
    //     for (auto it = ecs.query_items<Mut<Transform>, Velocity, Opt<Sprite>>(); auto item = it.next();) {
    //         auto& [e, row] = *item;  auto& [t, v, s] = row;  t->translation.x += v->x; ...
    // The guards must not outlive the iterator's Ecs; slots of a column must not be added while guards of it are alive.
    template <class... As>
    class QueryIterator {
    public:
        using Item = std::tuple<EntityIndex, std::tuple<typename detail::accessor<As>::guard...>>;
        QueryIterator(Ecs* ecs, std::vector<EntityIndex> matches) : ecs_(ecs), matches_(std::move(matches)) {}
        std::optional<Item> next() {  // query.rs:131-141
            if (pos_ >= matches_.size()) return std::nullopt;
            const EntityIndex e = matches_[pos_++];
            return Item(e, std::tuple<typename detail::accessor<As>::guard...>(ecs_->template fetch<As>(e)...));
        }

    private:
        Ecs* ecs_;
        std::vector<EntityIndex> matches_;  // already in iteration (descending) order
        std::size_t pos_ = 0;
    };
    template <class... As>
    QueryIterator<As...> query_items() {
        return QueryIterator<As...>(this, query<As...>());
    }
    // ecs.rs:172-174
    template <class... As>
    std::optional<typename QueryIterator<As...>::Item> query_one_by_id(EntityIndex id) {
        for (EntityIndex e : query<As...>())
            if (e == id) return typename QueryIterator<As...>::Item(e, std::tuple<typename detail::accessor<As>::guard...>(fetch<As>(e)...));
        return std::nullopt;
    }
    // Accessor::fetch (query.rs:170-177, 192-199, 230-232)
    template <class A>
    typename detail::accessor<A>::guard fetch(EntityIndex e) {
        using Acc = detail::accessor<A>;
        using T = typename Acc::type;
        if constexpr (std::is_same_v<T, Transform>) refresh_transforms();
        auto it = stores_.find(std::type_index(typeid(T)));
        T* p = nullptr;
        if (it != stores_.end() && e < it->second.data.size()) p = static_cast<T*>(it->second.data[e].get());
        if constexpr (!Acc::required) {
            if (!p) return std::nullopt;
            if constexpr (Acc::mut) return RefMut<T>(p, it->second.flag(e), [this, e] { mark_dirty<T>(e); });
            else return Ref<T>(p, it->second.flag(e));
        } else {
            if (!p) throw Error(TB_ERR_INVALID, "fetch: required component missing");  // cannot happen for matches
            if constexpr (Acc::mut) return RefMut<T>(p, it->second.flag(e), [this, e] { mark_dirty<T>(e); });
            else return Ref<T>(p, it->second.flag(e));
        }
    }

    // Component access (what the Ref / RefMut guards of query.rs:166-208 give). nullptr == None.
    template <class C>
    const C* get(EntityIndex e) {
        if constexpr (std::is_same_v<C, Transform>) refresh_transforms();
        auto it = stores_.find(std::type_index(typeid(C)));
        if (it == stores_.end() || e >= it->second.data.size()) return nullptr;
        return static_cast<const C*>(it->second.data[e].get());
    }
    template <class C>
    C* get_mut(EntityIndex e) {
        C* p = const_cast<C*>(get<C>(e));
        if (p) mark_dirty<C>(e);
        return p;
    }

    // ecs.rs:64-86
    template <class T>
    void insert_shared_resource(T value) { resources_[std::type_index(typeid(T))] = std::make_shared<T>(std::move(value)); }
    template <class T>
    T* shared_resource() {
        auto it = resources_.find(std::type_index(typeid(T)));
        return it == resources_.end() ? nullptr : static_cast<T*>(it->second.get());
    }

    // ---- device side -------------------------------------------------------------------------
    tb_ctx* ctx() const { return ctx_; }

    // Pushes every dirty range of the device-visible columns. const: render_scene takes `&Ecs`
    // (tuber-graphics/src/lib.rs:34-36); what changes is the device mirror's bookkeeping, not the Ecs.
    void flush() const {
        if (count_dirty_) {
            check(tb_set_entity_count(ctx_, next_), "tb_set_entity_count");
            count_dirty_ = false;
        }
        push<Transform>();
        push<Sprite>();
        push<Velocity>();
        push<Parent>();
        push<Spin>();
    }
    // A device system has written Transforms: host copies are stale until read again.
    void device_wrote_transforms() { transforms_stale_ = true; }

    void check(tb_status s, const char* what) const {
        if (s != TB_OK) throw Error(s, std::string(what) + ": " + tb_last_error(ctx_));
    }

private:
    template <class C>
    void put(EntityIndex e, C c) {
        const std::type_index ty(typeid(C));
        auto it = stores_.find(ty);
        if (it == stores_.end()) {  // ecs.rs:199: ComponentStore::with_size(index)
            it = stores_.emplace(ty, detail::Store{}).first;
            it->second.data.assign(e + 1, nullptr);
        }
        write(it->second, e, std::move(c));
    }
    template <class C>
    void write(detail::Store& s, EntityIndex e, C c) {
        if (e >= s.data.size()) s.data.resize(e + 1, nullptr);
        s.data[e] = std::make_shared<C>(std::move(c));
        s.set(e);
        mark_dirty<C>(e);
    }
    void erase(std::type_index ty, detail::Store& s, EntityIndex e) {
        if (e < s.data.size()) s.data[e] = nullptr;
        s.unset(e);
        for (int k = 0; k < TB_COMPONENT_COUNT; ++k)
            if (device_type_[k] && *device_type_[k] == ty) presence_dirty_[k].mark(e);
    }
    template <class C>
    void mark_dirty(EntityIndex e) {
        constexpr int id = detail::device_component<C>::id;
        if constexpr (id >= 0) {
            if (!device_type_[id]) device_type_[id] = std::make_unique<std::type_index>(typeid(C));
            value_dirty_[id].mark(e);
            presence_dirty_[id].mark(e);
        }
    }
    template <class A>
    void collect(std::vector<const detail::Store*>& req, uint32_t& bits, bool& all_device, bool& missing) {
        using T = typename detail::accessor<A>::type;
        if (!detail::accessor<A>::required) return;  // query.rs:105
        auto it = stores_.find(std::type_index(typeid(T)));
        if (it == stores_.end()) {
            missing = true;  // query.rs:110
            return;
        }
        req.push_back(&it->second);
        constexpr int id = detail::device_component<T>::id;
        if constexpr (id >= 0) bits |= 1u << id;
        else all_device = false;
    }
    template <class C>
    void push() const {
        constexpr int id = detail::device_component<C>::id;
        auto it = stores_.find(std::type_index(typeid(C)));
        if (it == stores_.end()) return;
        const detail::Store& s = it->second;
        if (value_dirty_[id].any()) {
            const EntityIndex lo = value_dirty_[id].lo, hi = std::min<EntityIndex>(value_dirty_[id].hi, next_);
            if constexpr (std::is_same_v<C, Parent>) {
                std::vector<uint32_t> buf(hi - lo, TB_NO_PARENT);
                for (EntityIndex e = lo; e < hi; ++e)
                    if (auto* p = static_cast<Parent*>(s.data[e].get())) buf[e - lo] = (uint32_t)p->index;
                check(tb_upload_parents(ctx_, lo, hi - lo, buf.data()), "tb_upload_parents");
            } else {
                std::vector<C> buf(hi - lo);
                for (EntityIndex e = lo; e < hi; ++e)
                    if (auto* p = static_cast<C*>(s.data[e].get())) buf[e - lo] = *p;
                if constexpr (std::is_same_v<C, Transform>)
                    check(tb_upload_transforms(ctx_, lo, hi - lo, reinterpret_cast<const tb_transform*>(buf.data())), "tb_upload_transforms");
                else if constexpr (std::is_same_v<C, Sprite>)
                    check(tb_upload_sprites(ctx_, lo, hi - lo, reinterpret_cast<const tb_sprite*>(buf.data())), "tb_upload_sprites");
                else if constexpr (std::is_same_v<C, Spin>)
                    check(tb_upload_spins(ctx_, lo, hi - lo, reinterpret_cast<const float*>(buf.data())), "tb_upload_spins");
                else
                    check(tb_upload_velocities(ctx_, lo, hi - lo, reinterpret_cast<const tb_velocity*>(buf.data())), "tb_upload_velocities");
            }
            value_dirty_[id].clear();
        }
        if (presence_dirty_[id].any()) {
            // uploads mark whole ranges present: restore the store's own bitset words over the range
            const uint64_t w0 = presence_dirty_[id].lo / 64, w1 = (std::min<EntityIndex>(presence_dirty_[id].hi, next_) + 63) / 64;
            if (w1 > w0) {
                std::vector<uint64_t> words(w1 - w0, 0);
                for (uint64_t w = w0; w < w1 && w < s.bits.size(); ++w) words[w - w0] = s.bits[w];
                check(tb_upload_presence(ctx_, id, w0, w1 - w0, words.data()), "tb_upload_presence");
            }
            presence_dirty_[id].clear();
        }
    }
    void refresh_transforms() {
        if (!transforms_stale_) return;
        transforms_stale_ = false;
        auto it = stores_.find(std::type_index(typeid(Transform)));
        if (it == stores_.end() || next_ == 0) return;
        std::vector<Transform> buf(next_);
        check(tb_download_transforms(ctx_, 0, next_, reinterpret_cast<tb_transform*>(buf.data())), "tb_download_transforms");
        for (EntityIndex e = 0; e < next_ && e < it->second.data.size(); ++e)
            if (auto* p = static_cast<Transform*>(it->second.data[e].get())) *p = buf[e];
    }

    tb_ctx* ctx_ = nullptr;
    uint64_t cap_;
    EntityIndex next_ = 0;
    std::vector<uint32_t> generation_;  // per entity slot
    std::unordered_map<std::type_index, detail::Store> stores_;             // Components (ecs.rs:11)
    std::unordered_map<std::type_index, std::shared_ptr<void>> resources_;  // Resources (ecs.rs:12)
    std::unique_ptr<std::type_index> device_type_[TB_COMPONENT_COUNT];
    // the device mirror's bookkeeping (changed by const calls: flush() from render_scene(&Ecs))
    mutable detail::DirtyRange value_dirty_[TB_COMPONENT_COUNT], presence_dirty_[TB_COMPONENT_COUNT];
    mutable bool count_dirty_ = false;
    bool transforms_stale_ = false;
};

// system.rs:5-6 — Ok(()) is an empty optional, Err(e) carries the message.
using SystemResult = std::optional<std::string>;
inline SystemResult Ok() { return std::nullopt; }
inline SystemResult Err(std::string what) { return what; }

// system.rs:8-24 — boxed systems run in insertion order; the first Err stops the bundle.
template <class AD>
class SystemBundle {
public:
    using System = std::function<SystemResult(Ecs&, AD&)>;
    // the four closure shapes IntoSystem accepts (system.rs:32-76)
    void add_system(System s) { systems_.push_back(std::move(s)); }
    void add_system(std::function<SystemResult(Ecs&)> s) {
        systems_.push_back([s](Ecs& e, AD&) { return s(e); });
    }
    void add_system(std::function<void(Ecs&, AD&)> s) {
        systems_.push_back([s](Ecs& e, AD& a) { s(e, a); return Ok(); });
    }
    void add_system(std::function<void(Ecs&)> s) {
        systems_.push_back([s](Ecs& e, AD&) { s(e); return Ok(); });
    }
    SystemResult step(Ecs& ecs, AD& additional_data) {
        for (auto& s : systems_)
            if (auto err = s(ecs, additional_data)) return err;
        return Ok();
    }
    std::size_t size() const { return systems_.size(); }

private:
    std::vector<System> systems_;
};

// [SPEC] the velocity system as a device system: translation += velocity * dt with
// dt = DeltaTime.0 as f32 (tuber-engine/src/state.rs:81). Queued; fused into the next frame.
inline SystemResult velocity_system(Ecs& ecs) {
    DeltaTime* dt = ecs.shared_resource<DeltaTime>();
    if (!dt) return Err("DeltaTime resource missing");
    try {
        ecs.flush();
        ecs.check(tb_system_integrate(ecs.ctx(), (float)dt->dt), "tb_system_integrate");
        ecs.device_wrote_transforms();
    } catch (const Error& e) {
        return Err(e.what());
    }
    return Ok();
}

// [SPEC] the spin system as a device system: angle.z += spin * dt.
inline SystemResult spin_system(Ecs& ecs) {
    DeltaTime* dt = ecs.shared_resource<DeltaTime>();
    if (!dt) return Err("DeltaTime resource missing");
    try {
        ecs.flush();
        ecs.check(tb_system_spin(ecs.ctx(), (float)dt->dt), "tb_system_spin");
        ecs.device_wrote_transforms();
    } catch (const Error& e) {
        return Err(e.what());
    }
    return Ok();
}

// A bundle of built-in device systems stepped with ONE library call (SystemBundle::step order, system.rs:17-23):
// systems with disjoint access sets run concurrently on forked streams, the state afterwards is the sequential one.
inline SystemResult device_bundle_step(Ecs& ecs, const std::vector<int>& systems, std::vector<uint32_t>* stages = nullptr) {
    DeltaTime* dt = ecs.shared_resource<DeltaTime>();
    if (!dt) return Err("DeltaTime resource missing");
    try {
        ecs.flush();
        std::vector<uint32_t> st(systems.size());
        ecs.check(tb_bundle_step(ecs.ctx(), systems.data(), (uint32_t)systems.size(), (float)dt->dt, st.data()), "tb_bundle_step");
        if (stages) *stages = st;
        ecs.device_wrote_transforms();
    } catch (const Error& e) {
        return Err(e.what());
    }
    return Ok();
}

// tuber-graphics/src/lib.rs:25-36
struct GraphicsError {
    std::string what;
};
using GraphicsResult = std::optional<GraphicsError>;  // nullopt == Ok(())

class GraphicsAPI {
public:
    virtual ~GraphicsAPI() = default;
    virtual GraphicsResult render_scene(const Ecs& ecs) = 0;  // tuber-graphics/src/lib.rs:34-36: `&Ecs`
};

// The batching renderer behind render_scene: leaves instances / vertices / batches on the device
// in draw order (tb_frame_out) for the presenter.
class Graphics : public GraphicsAPI {
public:
    // Camera ([SPEC]): an orthographic view-projection (Matrix4::new_orthographic, matrix.rs:41-58) folded
    // into the emitted instances and vertices; clear_camera() goes back to world-space output.
    void set_orthographic(float left, float right, float bottom, float top, float near_plane, float far_plane) {
        tb_orthographic(left, right, bottom, top, near_plane, far_plane, view_proj_);
        has_camera_ = true;
    }
    void set_view_projection(const float (&row_major)[16]) {
        std::memcpy(view_proj_, row_major, sizeof(view_proj_));
        has_camera_ = true;
    }
    void clear_camera() { has_camera_ = false; }

    GraphicsResult render_scene(const Ecs& ecs) override {
        try {
            ecs.flush();
            ecs.check(tb_frame_set_view_projection(ecs.ctx(), has_camera_ ? view_proj_ : nullptr), "tb_frame_set_view_projection");
            ecs.check(tb_frame_transform_batch(ecs.ctx(), &frame_), "tb_frame_transform_batch");
        } catch (const Error& e) {
            return GraphicsError{e.what()};
        }
        ecs_ = &ecs;
        return std::nullopt;
    }
    const tb_frame_out& frame() const { return frame_; }
    std::vector<tb_instance> instances() const {
        std::vector<tb_instance> v(frame_.num_instances);
        if (!v.empty()) ecs_->check(tb_download_instances(ecs_->ctx(), 0, v.size(), v.data()), "tb_download_instances");
        return v;
    }
    std::vector<tb_quad_vertex> vertices() const {
        std::vector<tb_quad_vertex> v(frame_.num_instances * 4);
        if (!v.empty()) ecs_->check(tb_download_vertices(ecs_->ctx(), 0, frame_.num_instances, v.data()), "tb_download_vertices");
        return v;
    }
    std::vector<tb_batch> batches() const {
        std::vector<tb_batch> v(frame_.num_batches);
        if (!v.empty()) ecs_->check(tb_download_batches(ecs_->ctx(), 0, v.size(), v.data()), "tb_download_batches");
        return v;
    }
    std::vector<uint32_t> order() const {
        std::vector<uint32_t> v(frame_.num_instances);
        if (!v.empty()) ecs_->check(tb_download_order(ecs_->ctx(), 0, v.size(), v.data()), "tb_download_order");
        return v;
    }

private:
    tb_frame_out frame_{};
    const Ecs* ecs_ = nullptr;
    float view_proj_[16] = {};
    bool has_camera_ = false;
};

}  // namespace tuber

#endif  // TUBER_HOST_HPP
