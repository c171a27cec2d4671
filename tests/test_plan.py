"""CPU tests of the frame planner (tuber_legacy_b200/csrc/plan.h): which key bits are sorted, how
they are packed into the short key and split into radix digits. The property that matters: for any
set of keys, ordering by the short key equals ordering by the full 64-bit key."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


class PextPlan(C.Structure):
    _fields_ = [("nruns", C.c_uint32), ("nbits", C.c_uint32), ("src_shift", C.c_uint8 * 16),
                ("width", C.c_uint8 * 16), ("dst_shift", C.c_uint8 * 16)]


class FramePlan(C.Structure):
    _fields_ = [("pext", PextPlan), ("varying", C.c_uint64), ("constant", C.c_uint64), ("npasses", C.c_uint32),
                ("shift", C.c_uint32 * 8), ("bits", C.c_uint32 * 8), ("ltbits", C.c_uint32), ("zbits", C.c_uint32),
                ("key64", C.c_uint32), ("zdict_n", C.c_uint32), ("zdict", C.c_uint32 * 64)]


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "helpers", "plan_capi.cpp")
    out = os.path.join(HERE, "helpers", "_plan_capi.so")
    hdr = os.path.join(HERE, "..", "tuber_legacy_b200", "csrc", "plan.h")
    if not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.plan_build.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(FramePlan)]
    L.plan_pext.restype = C.c_uint64
    L.plan_pext.argtypes = [C.POINTER(FramePlan), C.c_uint64]
    L.plan_pdep.restype = C.c_uint64
    L.plan_pdep.argtypes = [C.POINTER(FramePlan), C.c_uint64]
    L.plan_covers.argtypes = [C.POINTER(FramePlan), C.c_uint64, C.c_uint64]
    L.plan_build_dict.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(FramePlan)]
    L.plan_equal.argtypes = [C.POINTER(FramePlan), C.POINTER(FramePlan)]
    L.plan_short_key.restype = C.c_uint64
    L.plan_short_key.argtypes = [C.POINTER(FramePlan), C.c_uint64, C.POINTER(C.c_int)]
    assert L.plan_sizeof() == C.sizeof(FramePlan)
    return L


def build(lib, keys):
    k_or = int(np.bitwise_or.reduce(keys))
    k_and = int(np.bitwise_and.reduce(keys))
    f = FramePlan()
    lib.plan_build(k_or, k_and, C.byref(f))
    return f, k_or, k_and


def check_plan(lib, keys):
    f, k_or, k_and = build(lib, keys)
    short = np.array([lib.plan_pext(C.byref(f), int(k)) for k in keys], np.uint64)
    # order preservation, including equality
    o_full = np.argsort(keys, kind="stable")
    o_short = np.argsort(short, kind="stable")
    assert np.array_equal(o_full, o_short)
    assert np.array_equal(np.diff(keys[o_full]) == 0, np.diff(short[o_short]) == 0)
    # pdep inverts pext on top of the constant bits
    for k, s in zip(keys[:50], short[:50]):
        assert (lib.plan_pdep(C.byref(f), int(s)) | f.constant) == int(k)
    # digits tile the short key
    assert f.pext.nbits == f.ltbits + f.zbits
    assert sum(f.bits[p] for p in range(f.npasses)) == f.pext.nbits
    assert all(f.bits[p] <= 8 for p in range(f.npasses))
    assert [f.shift[p] for p in range(f.npasses)] == list(np.cumsum([0] + [f.bits[p] for p in range(f.npasses - 1)]))
    assert f.key64 == (1 if f.pext.nbits > 32 else 0)
    assert lib.plan_covers(C.byref(f), k_or, k_and) == 1
    return f


def test_single_key_has_no_bits(lib):
    f = check_plan(lib, np.array([0x0003000780000000] * 5, np.uint64))
    assert f.pext.nbits == 0 and f.npasses == 1 and f.bits[0] == 0


def test_particles_four_textures(lib):
    keys = (np.arange(1000, dtype=np.uint64) % 4) << np.uint64(32) | np.uint64(0x80000000)
    f = check_plan(lib, keys)
    assert (f.pext.nbits, f.ltbits, f.zbits, f.npasses) == (2, 2, 0, 1)


def test_config2_layers_textures_z_steps(lib):
    rng = np.random.default_rng(0)
    n = 5000
    z = rng.integers(0, 16, n).astype(np.float32).view(np.uint32).astype(np.uint64) ^ np.uint64(0x80000000)
    keys = (rng.integers(0, 16, n).astype(np.uint64) << np.uint64(48)) | (rng.integers(0, 64, n).astype(np.uint64) << np.uint64(32)) | z
    f = check_plan(lib, keys)
    assert f.ltbits == 10 and f.npasses == (f.pext.nbits + 7) // 8 and not f.key64


def test_random_float_z_needs_64bit_short_keys(lib):
    rng = np.random.default_rng(1)
    n = 5000
    zf = rng.uniform(-1, 1, n).astype(np.float32)
    u = zf.view(np.uint32).astype(np.uint64)
    u = np.where(u >> np.uint64(31), u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))
    keys = (rng.integers(0, 16, n).astype(np.uint64) << np.uint64(48)) | (rng.integers(0, 64, n).astype(np.uint64) << np.uint64(32)) | u
    f = check_plan(lib, keys)
    assert f.key64 == 1 and f.npasses >= 5


def test_more_than_16_runs_fills_gaps(lib):
    # alternating bits: 32 runs -> gaps are filled until 16 remain; ordering must survive
    rng = np.random.default_rng(2)
    mask = np.uint64(0x5555555555555555)
    keys = rng.integers(0, 2**63, 4000).astype(np.uint64) & mask
    f = check_plan(lib, keys)
    assert f.pext.nruns <= 16 and f.pext.nbits > 32


def test_runs_never_straddle_bit_32(lib):
    keys = np.array([0, 0x00000001FFFFFFFF, 0x0000000180000000, 0x000000007FFFFFFF], np.uint64)
    f = check_plan(lib, keys)
    assert f.ltbits == 1 and f.zbits == 32
    for r in range(f.pext.nruns):
        lo, hi = f.pext.src_shift[r], f.pext.src_shift[r] + f.pext.width[r]
        assert hi <= 32 or lo >= 32


def test_plan_covers_detects_stale_plans(lib):
    keys = (np.arange(100, dtype=np.uint64) % 4) << np.uint64(32)
    f, _, _ = build(lib, keys)
    # fewer varying bits: still covered
    sub = keys[keys < (np.uint64(2) << np.uint64(32))]
    assert lib.plan_covers(C.byref(f), int(np.bitwise_or.reduce(sub)), int(np.bitwise_and.reduce(sub))) == 1
    # a new varying bit, or a changed constant bit: not covered
    more = np.append(keys, np.uint64(8) << np.uint64(32))
    assert lib.plan_covers(C.byref(f), int(np.bitwise_or.reduce(more)), int(np.bitwise_and.reduce(more))) == 0
    moved = keys | np.uint64(1)
    assert lib.plan_covers(C.byref(f), int(np.bitwise_or.reduce(moved)), int(np.bitwise_and.reduce(moved))) == 0


def orderable(zf):
    u = np.asarray(zf, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(u >> np.uint64(31), u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))


def build_dict(lib, keys):
    k_or = int(np.bitwise_or.reduce(keys))
    k_and = int(np.bitwise_and.reduce(keys))
    zvals = np.unique((keys & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    f = FramePlan()
    lib.plan_build_dict(k_or, k_and, zvals.ctypes.data_as(C.POINTER(C.c_uint32)), len(zvals), C.byref(f))
    return f, k_or, k_and


def short_keys(lib, f, keys):
    miss = C.c_int(0)
    out = []
    for k in keys:
        out.append(lib.plan_short_key(C.byref(f), int(k), C.byref(miss)))
        assert miss.value == 0
    return np.array(out, np.uint64)


@pytest.mark.parametrize("nz", [2, 3, 5, 17, 64])
def test_z_dictionary_ranks_preserve_order(lib, nz):
    # depths 0.0, 1.0, 2.0 ... differ in many float bits but are few values: the short key holds their rank
    rng = np.random.default_rng(nz)
    n = 3000
    z = orderable(rng.integers(0, nz, n).astype(np.float32) - 1.0)
    z[:nz] = orderable(np.arange(nz, dtype=np.float32) - 1.0)
    keys = (rng.integers(0, 3, n).astype(np.uint64) << np.uint64(48)) | (rng.integers(0, 5, n).astype(np.uint64) << np.uint64(32)) | z
    plain, _, _ = build(lib, keys)
    f, k_or, k_and = build_dict(lib, keys)
    rank_bits = int(np.ceil(np.log2(nz)))
    assert f.zdict_n == nz and f.zbits == rank_bits and f.zbits < plain.zbits
    assert f.ltbits == plain.ltbits and f.pext.nbits == f.ltbits + f.zbits
    assert sum(f.bits[p] for p in range(f.npasses)) == f.pext.nbits and f.npasses <= plain.npasses
    short = short_keys(lib, f, keys)
    assert int(short.max()) < (1 << f.pext.nbits)
    o_full = np.argsort(keys, kind="stable")
    o_short = np.argsort(short, kind="stable")
    assert np.array_equal(o_full, o_short)
    assert np.array_equal(np.diff(keys[o_full]) == 0, np.diff(short[o_short]) == 0)
    # the (layer, texture) bin of a short key expands back to the key's high half
    for k, s in zip(keys[:50], short[:50]):
        hi = lib.plan_pdep(C.byref(f), int(s) >> f.zbits) | (f.constant & 0xFFFFFFFF00000000)
        assert hi == int(k) & 0xFFFFFFFF00000000
    assert lib.plan_covers(C.byref(f), k_or, k_and) == 1


def test_z_dictionary_not_taken_when_it_does_not_help(lib):
    # two depths one bit apart: the varying-bit key is already 1 bit wide
    z = orderable(np.array([1.0, np.nextafter(np.float32(1.0), np.float32(2.0))] * 10, np.float32))
    keys = z | (np.uint64(1) << np.uint64(32))
    f, _, _ = build_dict(lib, keys)
    assert f.zdict_n == 0 and f.zbits == 1
    # more than 64 depths
    z = orderable(np.arange(65, dtype=np.float32))
    f, _, _ = build_dict(lib, z)
    assert f.zdict_n == 0


def test_z_dictionary_miss_and_stale_high_half(lib):
    z = orderable(np.array([0.0, 1.0, 2.0, 3.0], np.float32))
    keys = np.concatenate([z | (np.uint64(t) << np.uint64(32)) for t in range(4)])
    f, k_or, k_and = build_dict(lib, keys)
    assert f.zdict_n == 4 and f.zbits == 2 and f.ltbits == 2
    miss = C.c_int(0)
    lib.plan_short_key(C.byref(f), int(orderable(np.float32(1.5))), C.byref(miss))
    assert miss.value == 1
    # low-half changes are the dictionary's business; a new texture bit is not covered
    assert lib.plan_covers(C.byref(f), k_or | 0x7FFFFFFF, k_and & ~0x7FFFFFFF) == 1
    assert lib.plan_covers(C.byref(f), k_or | (8 << 32), k_and) == 0


def test_plan_equal_ignores_padding_and_unused_dictionary_slots(lib):
    """the frame signature (which steady frames may replay a captured graph) compares plans field by field"""
    z = orderable(np.array([0.0, 1.0, 2.0, 3.0], np.float32))
    keys = np.concatenate([z | (np.uint64(t) << np.uint64(32)) for t in range(4)])
    a, _, _ = build_dict(lib, keys)
    b, _, _ = build_dict(lib, keys)
    for k in range(a.zdict_n, 64):
        b.zdict[k] = 0xDEADBEEF  # slots beyond zdict_n do not count
    assert lib.plan_equal(C.byref(a), C.byref(b)) == 1
    b.zdict[1] += 1
    assert lib.plan_equal(C.byref(a), C.byref(b)) == 0
    c, _, _ = build(lib, keys)  # the same keys without a dictionary: another plan
    assert lib.plan_equal(C.byref(a), C.byref(c)) == 0
    d, _, _ = build(lib, keys)
    assert lib.plan_equal(C.byref(c), C.byref(d)) == 1
    d.bits[0] += 1
    assert lib.plan_equal(C.byref(c), C.byref(d)) == 0
