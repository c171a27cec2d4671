// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's tuber-ecs storage, query and system semantics,
// including its cost model: one heap box per component per entity, a hash-map lookup per
// accessor per fetched entity (SipHash-1-3, Rust's default hasher), a runtime borrow flag,
// and DESCENDING entity-index iteration (QueryIterator pops matches from the back).
//
// Follows: crates/tuber-ecs/src/ecs.rs:11-215, query.rs:12-258, bitset.rs:1-54,
// system.rs:5-76, lib.rs:16-19.
//
// Deliberate, documented deviation: the reference's presence bitset is a fixed [u64; 1024]
// (ecs.rs:14) and panics for entity index >= 65 536 (bitset.rs:16-20). `bitset_words` is a
// constructor parameter here: 1024 reproduces the reference (including the out-of-range
// failure, surfaced as std::out_of_range); larger values are used for BASELINE configs 2-5.
//
// Parity status: PINNED by replaying the reference's own tests (ecs.rs:235-309,
// system.rs:96-175, bitset.rs:60-108) in tests/test_oracle_ecs.py. Iteration order,
// query_by_ids, delete_by_*, add/remove_component follow the source text only.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace ref {

using EntityIndex = size_t;  // tuber-ecs/src/lib.rs:16

// ---- bitset.rs:15-36 (impl BitSet for [u64]) ------------------------------------------
struct BitSet {
    std::vector<uint64_t> words;
    explicit BitSet(size_t nwords = 1024) : words(nwords, 0) {}
    void set_bit(size_t bit) { words.at(bit / 64) |= (uint64_t(1) << (bit % 64)); }
    void unset_bit(size_t bit) { words.at(bit / 64) &= ~(uint64_t(1) << (bit % 64)); }
    bool bit(size_t bit) const { return (words.at(bit / 64) & (uint64_t(1) << (bit % 64))) != 0; }
    size_t bit_count() const { return words.size() * 64; }
};
// bitset.rs:38-54 (impl BitSet for u64)
inline void u64_set_bit(uint64_t& w, size_t bit) { w |= (uint64_t(1) << bit); }
inline void u64_unset_bit(uint64_t& w, size_t bit) { w &= ~(uint64_t(1) << bit); }
inline bool u64_bit(uint64_t w, size_t bit) { return (w & (uint64_t(1) << bit)) != 0; }

// ---- TypeId + Rust's default hasher ------------------------------------------------------
// Rust's HashMap<TypeId, _> (ecs.rs:11) hashes the TypeId with SipHash-1-3 under a
// per-process random key. The algorithm is restated from the SipHash paper so the baseline
// pays the same per-lookup cost; the key is fixed (its value is irrelevant to results).
using TypeId = uint64_t;

struct SipHasher13 {
    static inline uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
    size_t operator()(TypeId id) const {
        const uint64_t k0 = 0x0706050403020100ull, k1 = 0x0f0e0d0c0b0a0908ull;
        uint64_t v0 = k0 ^ 0x736f6d6570736575ull, v1 = k1 ^ 0x646f72616e646f6dull;
        uint64_t v2 = k0 ^ 0x6c7967656e657261ull, v3 = k1 ^ 0x7465646279746573ull;
        auto round = [&]() {
            v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
            v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
            v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
            v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
        };
        uint64_t m = id;  // one 8-byte block
        v3 ^= m; round(); v0 ^= m;
        uint64_t b = uint64_t(8) << 56;  // length byte, no tail bytes
        v3 ^= b; round(); v0 ^= b;
        v2 ^= 0xff; round(); round(); round();
        return size_t(v0 ^ v1 ^ v2 ^ v3);
    }
};

// ---- RefCell<Box<dyn Any>> ------------------------------------------------------------------
// ecs.rs:17 — each present component is its own heap allocation behind a runtime borrow
// flag. borrow >= 0: that many shared borrows; borrow == -1: one exclusive borrow.
struct BoxedAny {
    TypeId type;
    void* payload;  // malloc'd, `size` bytes
    size_t size;
    int borrow;
};

struct BorrowError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Ref<'a, T> / RefMut<'a, T> guards (query.rs:168,190): release the flag on drop.
struct Guard {
    BoxedAny* cell = nullptr;
    bool exclusive = false;
    Guard() = default;
    Guard(BoxedAny* c, bool excl) : cell(c), exclusive(excl) {
        if (excl) {
            if (c->borrow != 0) throw BorrowError("already borrowed");  // RefCell::borrow_mut panics
            c->borrow = -1;
        } else {
            if (c->borrow < 0) throw BorrowError("already mutably borrowed");  // RefCell::borrow panics
            c->borrow += 1;
        }
    }
    Guard(const Guard&) = delete;
    Guard& operator=(const Guard&) = delete;
    Guard(Guard&& o) noexcept : cell(o.cell), exclusive(o.exclusive) { o.cell = nullptr; }
    Guard& operator=(Guard&& o) noexcept {
        release();
        cell = o.cell;
        exclusive = o.exclusive;
        o.cell = nullptr;
        return *this;
    }
    ~Guard() { release(); }
    void release() {
        if (!cell) return;
        if (exclusive) cell->borrow = 0; else cell->borrow -= 1;
        cell = nullptr;
    }
    bool some() const { return cell != nullptr; }
    template <typename T> T* get() const { return static_cast<T*>(cell->payload); }
};

// ---- ComponentStore (ecs.rs:16-53) -----------------------------------------------------------
struct ComponentStore {
    std::vector<BoxedAny*> component_data;  // nullptr == None
    BitSet entities_bitset;

    // ecs.rs:23-33 — `vec![None]` then `size` more pushes: size + 1 slots.
    ComponentStore(size_t size, size_t bitset_words) : component_data(size + 1, nullptr), entities_bitset(bitset_words) {}
    ComponentStore(const ComponentStore&) = delete;
    ComponentStore& operator=(const ComponentStore&) = delete;
    ~ComponentStore() { for (auto* b : component_data) drop(b); }

    static void drop(BoxedAny* b) {
        if (!b) return;
        std::free(b->payload);
        delete b;
    }
    static BoxedAny* box(TypeId type, const void* bytes, size_t size) {
        auto* b = new BoxedAny{type, std::malloc(size ? size : 1), size, 0};
        if (size) std::memcpy(b->payload, bytes, size);
        return b;
    }
    // ecs.rs:35-38
    void remove_from_entity(EntityIndex e) {
        entities_bitset.unset_bit(e);
        drop(component_data.at(e));
        component_data.at(e) = nullptr;
    }
    // ecs.rs:40-43
    void add_to_entity(TypeId type, const void* bytes, size_t size, EntityIndex e) {
        entities_bitset.set_bit(e);
        drop(component_data.at(e));
        component_data.at(e) = box(type, bytes, size);
    }
};

using Components = std::unordered_map<TypeId, std::unique_ptr<ComponentStore>, SipHasher13>;  // ecs.rs:11

// One member of an EntityDefinition tuple (ecs.rs:184-215).
struct ComponentValue {
    TypeId type;
    const void* bytes;
    size_t size;
};

// One accessor of a Query tuple (query.rs:157-245): &T, &mut T or Opt<&T>/Opt<&mut T>.
struct AccessorSpec {
    TypeId type;
    bool required;   // false == Opt<_>  (query.rs:226-245)
    bool exclusive;  // true == &mut T   (query.rs:188-208)
};

struct Ecs;

// ---- QueryIterator (query.rs:87-141) ----------------------------------------------------------
struct QueryIterator {
    const Components* components;
    std::vector<AccessorSpec> accessors;
    std::vector<EntityIndex> matching_entities;
    EntityIndex index = 0;

    // query.rs:94-129
    QueryIterator(size_t entity_count, const Components* comps, std::vector<AccessorSpec> accs)
        : components(comps), accessors(std::move(accs)) {
        std::vector<BitSet> bitsets;  // copied by value, like `[u64;1024]: Copy` (query.rs:102)
        size_t required = 0;
        for (const auto& a : accessors) {
            if (!a.required) continue;  // query.rs:105
            ++required;
            auto it = components->find(a.type);
            if (it != components->end()) bitsets.push_back(it->second->entities_bitset);
        }
        if (bitsets.size() == required) {  // query.rs:110
            for (size_t i = 0; i < entity_count; ++i) {
                bool all = true;
                for (const auto& b : bitsets) {
                    if (!b.bit(i)) { all = false; break; }
                }
                if (all) matching_entities.push_back(i);
            }
        }
    }

    // Accessor::fetch (query.rs:170-177, 192-199, 230-232). Returns false when a required
    // accessor cannot be fetched (the tuple fetch's `?`, query.rs:29-31).
    static bool fetch(const Components* components, const std::vector<AccessorSpec>& accessors, EntityIndex e,
                      std::vector<Guard>& out) {
        out.clear();
        out.reserve(accessors.size());
        for (const auto& a : accessors) {
            BoxedAny* cell = nullptr;
            auto it = components->find(a.type);  // one hash lookup per accessor per entity
            if (it != components->end()) cell = it->second->component_data.at(e);
            if (cell && cell->type != a.type) cell = nullptr;  // downcast_ref (always succeeds here)
            if (!cell) {
                if (a.required) { out.clear(); return false; }
                out.emplace_back();  // Some(None)
                continue;
            }
            out.emplace_back(cell, a.exclusive);
        }
        return true;
    }

    // query.rs:137-140 — pop the LAST match, so iteration runs in descending index order.
    bool next(std::vector<Guard>& out) {
        if (matching_entities.empty()) return false;
        index = matching_entities.back();
        matching_entities.pop_back();
        return fetch(components, accessors, index, out);
    }
};

// ---- Ecs (ecs.rs:56-181) ----------------------------------------------------------------------
struct Ecs {
    Components components;
    std::unordered_map<TypeId, BoxedAny*, SipHasher13> shared_resources;  // ecs.rs:12
    EntityIndex next_index = 0;
    size_t bitset_words;

    explicit Ecs(size_t words = 1024) : bitset_words(words) {}
    Ecs(const Ecs&) = delete;
    Ecs& operator=(const Ecs&) = delete;
    ~Ecs() { for (auto& kv : shared_resources) ComponentStore::drop(kv.second); }

    // ecs.rs:64-67
    void insert_shared_resource(TypeId type, const void* bytes, size_t size) {
        auto it = shared_resources.find(type);
        if (it != shared_resources.end()) { ComponentStore::drop(it->second); shared_resources.erase(it); }
        shared_resources.emplace(type, ComponentStore::box(type, bytes, size));
    }
    // ecs.rs:70-86
    Guard shared_resource(TypeId type, bool exclusive) const {
        auto it = shared_resources.find(type);
        if (it == shared_resources.end()) return Guard();
        return Guard(it->second, exclusive);
    }

    // ecs.rs:93-98 with store_components (ecs.rs:191-203). `faithful_insert_cost` keeps the
    // reference's eager `ComponentStore::with_size(index)` argument evaluation (an
    // (index+1)-slot Vec plus a bitset allocated and dropped per tuple member per insert,
    // i.e. O(N^2) bulk loading); bulk scene construction for the large configs turns it off —
    // results are identical, only setup time changes, and setup is never timed.
    EntityIndex insert(const std::vector<ComponentValue>& def, bool faithful_insert_cost = true) {
        EntityIndex index = next_index;
        for (auto& kv : components) kv.second->component_data.push_back(nullptr);  // ecs.rs:194-196
        for (const auto& c : def) {
            std::unique_ptr<ComponentStore> fresh;
            auto it = components.find(c.type);
            if (faithful_insert_cost || it == components.end()) fresh.reset(new ComponentStore(index, bitset_words));
            if (it == components.end()) it = components.emplace(c.type, std::move(fresh)).first;  // ecs.rs:199
            ComponentStore& store = *it->second;
            ComponentStore::drop(store.component_data.back());
            store.component_data.back() = ComponentStore::box(c.type, c.bytes, c.size);  // ecs.rs:200
            store.entities_bitset.set_bit(index);                                        // ecs.rs:201
        }
        next_index += 1;
        return index;
    }

    size_t entity_count() const { return next_index; }  // ecs.rs:178-180

    // ecs.rs:105-112
    void delete_by_ids(const std::vector<EntityIndex>& ids) {
        for (EntityIndex e : ids)
            for (auto& kv : components) kv.second->remove_from_entity(e);
    }
    // ecs.rs:114-118
    void remove_component(TypeId type, EntityIndex e) {
        auto it = components.find(type);
        if (it != components.end()) it->second->remove_from_entity(e);
    }
    // ecs.rs:120-124 — silently does nothing when no store for the type exists yet.
    void add_component(TypeId type, const void* bytes, size_t size, EntityIndex e) {
        auto it = components.find(type);
        if (it != components.end()) it->second->add_to_entity(type, bytes, size, e);
    }
    // ecs.rs:127-129
    QueryIterator query(std::vector<AccessorSpec> accessors) const {
        return QueryIterator(entity_count(), &components, std::move(accessors));
    }
    // Query::matching_ids (query.rs:34-38) with Accessor::matching_ids (query.rs:210-224,
    // 234-236): set intersection; Opt<_> matches every index below entity_count.
    std::unordered_set<EntityIndex> matching_ids(const std::vector<AccessorSpec>& accessors) const {
        std::unordered_set<EntityIndex> result;
        bool first = true;
        for (const auto& a : accessors) {
            std::unordered_set<EntityIndex> ids;
            if (!a.required) {
                for (size_t i = 0; i < entity_count(); ++i) ids.insert(i);
            } else {
                auto it = components.find(a.type);
                if (it != components.end()) {
                    const BitSet& b = it->second->entities_bitset;
                    // query.rs:216 iterates to max(entity_count, bit_count); bits at or above
                    // entity_count are never set, so scanning bit_count alone is equivalent.
                    size_t upper = entity_count() > b.bit_count() ? b.bit_count() : entity_count();
                    for (size_t i = 0; i < upper; ++i)
                        if (b.bit(i)) ids.insert(i);
                }
            }
            if (first) { result = std::move(ids); first = false; }
            else {
                std::unordered_set<EntityIndex> inter;
                for (auto e : result) if (ids.count(e)) inter.insert(e);
                result = std::move(inter);
            }
        }
        return result;
    }
    // ecs.rs:100-103
    void delete_by_query(const std::vector<AccessorSpec>& accessors) {
        auto ids = matching_ids(accessors);
        delete_by_ids(std::vector<EntityIndex>(ids.begin(), ids.end()));
    }
    // ecs.rs:137-169 — the LOWEST matching index; None when a required store is missing.
    bool query_one(const std::vector<AccessorSpec>& accessors, EntityIndex* index, std::vector<Guard>& out) const {
        std::vector<const BitSet*> bitsets;
        size_t required = 0;
        for (const auto& a : accessors) {
            if (!a.required) continue;
            ++required;
            auto it = components.find(a.type);
            if (it != components.end()) bitsets.push_back(&it->second->entities_bitset);
        }
        if (bitsets.size() != required) return false;
        bool found = false;
        EntityIndex idx = 0;
        for (size_t i = 0; i < entity_count(); ++i) {
            bool all = true;
            for (auto* b : bitsets) if (!b->bit(i)) { all = false; break; }
            if (all) { idx = i; found = true; break; }
        }
        if (!found) return false;
        *index = idx;
        return QueryIterator::fetch(&components, accessors, idx, out);
    }
    // ecs.rs:172-174
    bool query_one_by_id(const std::vector<AccessorSpec>& accessors, EntityIndex id, std::vector<Guard>& out) const {
        return QueryIterator::fetch(&components, accessors, id, out);
    }
};

// QueryIteratorByIds (query.rs:56-85): the inner iterator filtered by an id set. The
// reference tests `ids.contains(inner.index)` AFTER advancing, so it yields exactly the
// matches whose index is in `ids`, still in descending order.
struct QueryIteratorByIds {
    QueryIterator inner;
    std::unordered_set<EntityIndex> ids;
    QueryIteratorByIds(QueryIterator it, std::unordered_set<EntityIndex> s) : inner(std::move(it)), ids(std::move(s)) {}
    bool next(std::vector<Guard>& out) {
        bool ok = inner.next(out);
        while (!ids.count(inner.index) && ok) ok = inner.next(out);
        return ok;
    }
};

// ---- SystemBundle (system.rs:8-24) ---------------------------------------------------------------
// Boxed systems run in insertion order; the first Err stops the bundle and is returned.
// A system returns an empty string for Ok(()) and the error text for Err(_).
template <typename AD>
struct SystemBundle {
    using System = std::function<std::string(Ecs&, AD&)>;
    std::vector<System> systems;
    void add_system(System s) { systems.push_back(std::move(s)); }  // system.rs:13-15
    std::string step(Ecs& ecs, AD& additional_data) {               // system.rs:17-23
        for (auto& s : systems) {
            std::string err = s(ecs, additional_data);
            if (!err.empty()) return err;
        }
        return std::string();
    }
};

}  // namespace ref
