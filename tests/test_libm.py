"""sinf / cosf as the reference computes them (f32::sin / f32::cos -> the platform libm, number_traits.rs:137-143):
the restatement the device evaluates (csrc/libm_sincosf.h) against this box's real libm, bit for bit — on the CPU
(host build of the same header) and on the GPU (through Quaternion::from_euler with roll = pitch = 0, whose w and z
are exactly cos and sin of yaw / 2)."""
import ctypes

import numpy as np
import pytest

import util


def test_selected_build_matches_libm_on_a_spread_of_all_exponents():
    L = util.libm_check_lib()
    fma = L.libm_uses_fma()
    bs, bc, first = util.libm_compare(0, 509, 1 << 23, fma)  # 8 M bit patterns over the whole float range
    assert (bs, bc) == (0, 0), f"restated sinf/cosf (fma={fma}) differ from libm, first at bits {first:#x}"


@pytest.mark.parametrize("lo,hi", [(0.0, 0.79), (0.78, 7.0), (6.0, 120.5), (119.0, 1.0e6), (1e6, 3.0e38)])
def test_dense_ranges(lo, hi):
    # consecutive bit patterns from lo upwards (positive and negative): rounding boundaries of every branch
    L = util.libm_check_lib()
    fma = L.libm_uses_fma()
    first = int(np.float32(lo).view(np.uint32))
    last = int(np.float32(hi).view(np.uint32))
    stride = max(1, (last - first) // (1 << 21))
    for sign in (0, 0x80000000):
        bs, bc, fb = util.libm_compare(first | sign, stride, (last - first) // stride, fma)
        assert (bs, bc) == (0, 0), f"[{lo}, {hi}) sign {sign:#x}: first mismatch at bits {fb:#x}"


def test_special_values():
    L = util.libm_check_lib()
    y = np.array([0.0, -0.0, 1e-45, -1e-45, 2.0 ** -12, 2.0 ** -13, np.pi / 4, np.nextafter(np.float32(np.pi / 4), 0),
                  119.99999, 120.0, 1e9, -1e9, 3.4e38], np.float32)
    outs = [np.zeros_like(y) for _ in range(4)]
    L.libm_eval(y.ctypes.data, len(y), L.libm_uses_fma(), *[o.ctypes.data for o in outs])
    assert np.array_equal(outs[0].view(np.uint32), outs[2].view(np.uint32))
    assert np.array_equal(outs[1].view(np.uint32), outs[3].view(np.uint32))


def test_the_other_build_is_a_different_function():
    """The FMA and the generic build differ on a few inputs (34 of 2^32, found by an exhaustive run): the device has
    to evaluate the build the host libm uses, which is why tb_ctx_create selects it like glibc does."""
    a = util.libm_compare(0x418a3adb, 1, 1, 1)
    b = util.libm_compare(0x418a3adb, 1, 1, 0)
    assert a != b


@pytest.mark.gpu
def test_device_sincos_is_the_host_libm_bit_for_bit():
    from tuber_legacy_b200 import capi
    rng = np.random.default_rng(5)
    parts = [rng.uniform(-7, 7, 1 << 20), rng.uniform(-130, 130, 1 << 19), rng.uniform(-1e-3, 1e-3, 1 << 16),
             rng.uniform(-1e7, 1e7, 1 << 16), (rng.integers(0, 0x7f000000, 1 << 19, dtype=np.uint32)).view(np.float32)]
    half = np.concatenate([np.asarray(p, np.float32) for p in parts])
    yaw = (half * np.float32(2.0)).astype(np.float32)
    half = (yaw * np.float32(0.5)).astype(np.float32)
    ok = np.isfinite(yaw)
    yaw, half = yaw[ok], half[ok]
    e = np.zeros((len(yaw), 3), np.float32)
    e[:, 2] = yaw
    with capi.Context(16) as ctx:
        q = ctx.quat_from_euler(e)
    L = util.libm_check_lib()
    outs = [np.zeros_like(half) for _ in range(4)]
    L.libm_eval(half.ctypes.data, len(half), L.libm_uses_fma(), *[o.ctypes.data for o in outs])
    s_libm, c_libm = outs[2], outs[3]
    # roll = pitch = 0: w = 1*1*cy + 0*0*sy = cy, z = 1*1*sy - 0*0*cy = sy (x1, x0 and +-0 are exact)
    assert np.array_equal(q[:, 0].view(np.uint32), c_libm.view(np.uint32)), "device cos differs from libm cosf"
    bad = q[:, 3] != s_libm  # value comparison: sy - 0 keeps -0, but compare as floats
    assert not bad.any(), f"device sin differs from libm sinf at {half[bad][:4]}"
