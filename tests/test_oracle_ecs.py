"""Pins oracle/ref_ecs.hpp by replaying the reference's own unit tests for the ECS on the hot path:
crates/tuber-ecs/src/bitset.rs:56-109, ecs.rs:217-310, system.rs:78-176 (same entities, same
assertions), plus the source-only semantics SURVEY.md §7 lists (descending iteration, Opt,
add_component no-op, 65 536-entity cap)."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import oracle

POSITION, VELOCITY, VALUE, OTHER, COMP_A, COMP_B = 101, 102, 103, 104, 105, 106


def _p(a):
    return C.c_void_p(a.ctypes.data)


class Ecs:
    def __init__(self, words=1024):
        self.L = oracle.lib()
        self.h = self.L.orc_ecs_new(words)

    def __del__(self):
        self.L.orc_ecs_free(self.h)

    def insert(self, *components):
        types = np.array([t for t, _ in components], np.uint64)
        bufs = [C.create_string_buffer(b, len(b)) for _, b in components]
        ptrs = (C.c_void_p * len(bufs))(*[C.cast(b, C.c_void_p) for b in bufs])
        sizes = np.array([len(b) for _, b in components], np.uint64)
        return self.L.orc_ecs_insert(self.h, len(components), _p(types), ptrs, _p(sizes))

    def query(self, accessors):
        """accessors: [(type, required, exclusive)] -> (indices in iteration order, present flags)"""
        n = self.count()
        types = np.array([a[0] for a in accessors], np.uint64)
        req = np.array([a[1] for a in accessors], np.uint8)
        exc = np.array([a[2] for a in accessors], np.uint8)
        out = np.zeros(max(n, 1), np.uint64)
        present = np.zeros((max(n, 1), len(accessors)), np.uint8)
        c = self.L.orc_ecs_query(self.h, len(accessors), _p(types), _p(req), _p(exc), _p(out), _p(present), n)
        return out[:c].tolist(), present[:c]

    def query_one(self, accessors):
        types = np.array([a[0] for a in accessors], np.uint64)
        req = np.array([a[1] for a in accessors], np.uint8)
        exc = np.array([a[2] for a in accessors], np.uint8)
        present = np.zeros(len(accessors), np.uint8)
        i = self.L.orc_ecs_query_one(self.h, len(accessors), _p(types), _p(req), _p(exc), _p(present))
        return i, present

    def get(self, type_, index, size):
        buf = C.create_string_buffer(size)
        r = self.L.orc_ecs_get(self.h, type_, index, buf, size)
        return None if r < 0 else buf.raw

    def set(self, type_, index, data):
        return self.L.orc_ecs_set(self.h, type_, index, C.create_string_buffer(data, len(data)), len(data))

    def count(self):
        return self.L.orc_ecs_entity_count(self.h)


def pos(x, y):
    return struct.pack("<ff", x, y)


# ---- bitset.rs:56-109 -------------------------------------------------------------------------------

def test_bitset_u64(oracle_lib):
    assert oracle_lib.orc_u64_set_bit(0, 1) == 2                      # set_bit_u64
    assert oracle_lib.orc_u64_unset_bit(3, 1) == 1                    # unset_bit_u64
    b = oracle_lib.orc_u64_set_bit(oracle_lib.orc_u64_set_bit(0, 0), 2)
    assert oracle_lib.orc_u64_bit(b, 2) == 1 and b == 5               # bit_u64


def test_bitset_u64_array(oracle_lib):
    w = np.zeros(1024, np.uint64)
    assert oracle_lib.orc_words_set_bit(_p(w), 1024, 64) == 0 and w[1] == 1      # set_bit_u64_array
    assert oracle_lib.orc_words_unset_bit(_p(w), 1024, 64) == 0 and w[1] == 0    # unset_bit_u64_array
    oracle_lib.orc_words_set_bit(_p(w), 1024, 66)
    oracle_lib.orc_words_set_bit(_p(w), 1024, 2)
    assert oracle_lib.orc_words_bit(_p(w), 1024, 66) == 1 and oracle_lib.orc_words_bit(_p(w), 1024, 2) == 1
    # bitset.rs:16-20 indexes self[bit/64]: beyond [u64;1024] the reference panics
    assert oracle_lib.orc_words_set_bit(_p(w), 1024, 65536) == -1


# ---- ecs.rs:217-310 ---------------------------------------------------------------------------------

def test_ecs_new_and_insert():
    e = Ecs()
    assert e.count() == 0                                                          # ecs_new
    assert e.insert((POSITION, pos(0, 1)), (VELOCITY, pos(2, 3))) == 0
    assert e.count() == 1
    assert e.insert((POSITION, pos(4, 5)), (VELOCITY, pos(6, 7))) == 1
    assert e.count() == 2                                                          # ecs_insert


def test_ecs_query():
    e = Ecs()
    e.insert((POSITION, pos(12, 1)), (VELOCITY, pos(2, 3)))
    e.insert((POSITION, pos(4, 5)), (VELOCITY, pos(6, 7)))
    e.insert((POSITION, pos(4, 5)))
    ids, _ = e.query([(VELOCITY, 1, 1)])           # for (_, (mut velocity,)) in query::<(&mut Velocity,)>()
    for i in ids:
        x, y = struct.unpack("<ff", e.get(VELOCITY, i, 8))
        assert e.set(VELOCITY, i, pos(0.0, y)) == 0
    ids, _ = e.query([(VELOCITY, 1, 0)])
    for i in ids:
        assert abs(struct.unpack("<ff", e.get(VELOCITY, i, 8))[0]) <= 0.01
    assert len(e.query([(POSITION, 1, 0)])[0]) == 3
    assert len(e.query([(VELOCITY, 1, 0)])[0]) == 2


def test_ecs_query_one():
    e = Ecs()
    e.insert((POSITION, pos(12, 1)), (VELOCITY, pos(2, 3)))
    i, _ = e.query_one([(POSITION, 1, 0)])
    assert i == 0 and e.get(POSITION, 0, 8) == pos(12, 1)


def test_ecs_query_optional():
    e = Ecs()
    e.insert((POSITION, pos(12, 1)), (VELOCITY, pos(2, 3)))
    e.insert((POSITION, pos(15, 8)))
    ids, present = e.query([(POSITION, 1, 0), (VELOCITY, 0, 0)])   # (&Position, Opt<&Velocity>)
    assert len(ids) == 2                                            # Opt never filters
    assert present[ids.index(1)].tolist() == [1, 0] and present[ids.index(0)].tolist() == [1, 1]


def test_ecs_query_one_optional_and_implicit_read():
    e = Ecs()
    e.insert((POSITION, pos(12, 1)))
    i, present = e.query_one([(POSITION, 1, 0), (VELOCITY, 0, 0)])
    assert i == 0 and present.tolist() == [1, 0]                    # velocity.is_none()
    i, _ = e.query_one([(POSITION, 1, 0)])
    assert e.get(POSITION, i, 8) == pos(12, 1)


# ---- system.rs:78-176 ---------------------------------------------------------------------------------

SYSTEM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


def test_failing_system_and_bundle_order(oracle_lib):
    L = oracle_lib
    L.orc_bundle_add_system.argtypes = [C.c_void_p, SYSTEM_FN, C.c_void_p]
    e = Ecs()
    e.insert((VALUE, struct.pack("<i", 12)))
    e.insert((VALUE, struct.pack("<i", 18)), (OTHER, b"\0"))

    def add(delta):
        def system(ecs_handle, _ad):
            for i in e.query([(VALUE, 1, 1)])[0]:
                v = struct.unpack("<i", e.get(VALUE, i, 4))[0]
                e.set(VALUE, i, struct.pack("<i", v + delta))
            return 0
        return SYSTEM_FN(system)

    s1, s2 = add(35), add(-6)
    b = L.orc_bundle_new()
    L.orc_bundle_add_system(b, s1, e.h)
    assert L.orc_bundle_len(b) == 1                                     # system_bundle_add
    L.orc_bundle_add_system(b, s2, e.h)
    assert L.orc_bundle_step(b, e.h, None) == 0                         # system_bundle_step
    vals = {struct.unpack("<i", e.get(VALUE, i, 4))[0] for i in (0, 1)}
    assert vals == {41, 47}
    L.orc_bundle_free(b)
    # failing_system: the first Err stops the bundle; later systems do not run
    ran = []
    fail = SYSTEM_FN(lambda h, ad: 7)
    after = SYSTEM_FN(lambda h, ad: ran.append(1) or 0)
    b = L.orc_bundle_new()
    L.orc_bundle_add_system(b, fail, e.h)
    L.orc_bundle_add_system(b, after, e.h)
    assert L.orc_bundle_step(b, e.h, None) == 1 and ran == []
    L.orc_bundle_free(b)


def test_system_bundle_with_additional_data(oracle_lib):
    L = oracle_lib
    L.orc_bundle_add_system.argtypes = [C.c_void_p, SYSTEM_FN, C.c_void_p]
    e = Ecs()
    e.insert((COMP_A, b"\0"), (COMP_B, b"\0"))
    e.insert((COMP_B, b"\0"))
    data = C.c_int(0)

    def system(h, ad):
        C.cast(ad, C.POINTER(C.c_int))[0] += 1
        return 0
    fn = SYSTEM_FN(system)
    b = L.orc_bundle_new()
    L.orc_bundle_add_system(b, fn, e.h)
    L.orc_bundle_step(b, e.h, C.byref(data))
    L.orc_bundle_step(b, e.h, C.byref(data))
    assert data.value == 2
    L.orc_bundle_free(b)


# ---- source-only semantics (SURVEY.md §7 "reference quirks to preserve") ----------------------------

def test_iteration_is_descending_and_query_one_is_lowest():
    e = Ecs()
    for k in range(10):
        e.insert((POSITION, pos(k, 0))) if k % 3 else e.insert((POSITION, pos(k, 0)), (VELOCITY, pos(0, 0)))
    ids, _ = e.query([(POSITION, 1, 0)])
    assert ids == list(range(9, -1, -1))                      # query.rs:138 pops from the back
    ids, _ = e.query([(POSITION, 1, 0), (VELOCITY, 1, 0)])
    assert ids == [9, 6, 3, 0]
    assert e.query_one([(VELOCITY, 1, 0)])[0] == 0            # ecs.rs:137-169 lowest index
    assert e.query([(OTHER, 1, 0)])[0] == []                  # query.rs:110 missing store -> empty
    assert e.query_one([(OTHER, 1, 0)])[0] == -1


def test_add_remove_delete(oracle_lib):
    L = oracle_lib
    e = Ecs()
    e.insert((POSITION, pos(1, 1)))
    e.insert((POSITION, pos(2, 2)), (VELOCITY, pos(3, 3)))
    buf = C.create_string_buffer(pos(9, 9), 8)
    L.orc_ecs_add_component(e.h, OTHER, buf, 8, 0)            # ecs.rs:120-124: no store yet -> silently nothing
    assert e.get(OTHER, 0, 8) is None
    L.orc_ecs_add_component(e.h, VELOCITY, buf, 8, 0)
    assert e.get(VELOCITY, 0, 8) == pos(9, 9)
    L.orc_ecs_remove_component(e.h, VELOCITY, 1)
    assert e.query([(VELOCITY, 1, 0)])[0] == [0]
    ids = np.array([0], np.uint64)
    L.orc_ecs_delete_by_ids(e.h, _p(ids), 1)
    assert e.query([(POSITION, 1, 0)])[0] == [1] and e.count() == 2   # indices are never reused


def test_double_mutable_borrow_conflicts(oracle_lib):
    e = Ecs()
    e.insert((POSITION, pos(1, 1)))
    assert oracle_lib.orc_ecs_double_borrow_conflicts(e.h, POSITION, 0) == 1   # RefCell panics


def test_reference_bitset_caps_entities_at_65536():
    """ecs.rs:14: [u64; 1024]. The oracle reproduces the cap at the reference width."""
    e = Ecs(words=2)  # 128 entities: same mechanism, small enough to test quickly
    for k in range(128):
        assert e.insert((POSITION, pos(k, 0))) == k
    assert e.insert((POSITION, pos(0, 0))) == -1
    t = np.zeros((65537, 12), np.float32)
    with pytest.raises(OverflowError):
        oracle.Scene(65537, transforms=t, bitset_words=1024)


def test_scene_frame_small_known_answer():
    """[SPEC] frame on a hand-checkable scene: order, batches, instance layout, vertices."""
    T = np.zeros((4, 12), np.float32)
    T[:, 9:12] = 1
    T[:, 0] = [10, 20, 30, 40]
    T[:, 2] = [0.5, 0.25, 0.25, 0.5]
    S = np.zeros(4, oracle.np.dtype([("size", "<f4", 2), ("texture", "<u4"), ("layer", "<u4")]))
    S["size"] = [[2, 3]] * 4
    S["texture"] = [1, 0, 0, 1]
    S["layer"] = [0, 1, 0, 0]
    f = oracle.Scene(4, T, S).frame()
    assert f.order.tolist() == [2, 0, 3, 1]          # (l0,t0,z.25,e2) < (l0,t1,z.5,e0) < (l0,t1,z.5,e3: tie -> index) < (l1,...)
    f_order_keys = [(int(S["layer"][e]), int(S["texture"][e]), float(T[e, 2]), e) for e in f.order]
    assert f_order_keys == sorted(f_order_keys)
    assert [(b["layer"], b["texture"], b["first_instance"], b["count"]) for b in f.batches] == [(0, 0, 0, 1), (0, 1, 1, 2), (1, 0, 3, 1)]
    m = f.instances["model"][0]                       # entity 2: translation (30, 0, .25), column-major
    assert m.tolist() == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 30, 0, 0.25, 1]
    v = f.vertices[0]
    assert [(q["x"], q["y"], q["z"], q["uv"]) for q in v] == [(30, 0, .25, 0), (32, 0, .25, 1), (32, 3, .25, 0x10001), (30, 3, .25, 0x10000)]
