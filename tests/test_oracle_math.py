"""Pins oracle/ref_math.hpp against every known-answer test the reference holds for the hot
path's arithmetic: crates/tuber-math/src/matrix.rs:272-408,
quaternion.rs:160-260. Values and tolerances are the reference's own. as_matrix4 has no
reference test; it is pinned by the survey's float32 emulation vectors (tests/golden) and by
structural checks against the reference primitives."""
import ctypes as C
import json
import os

import numpy as np

from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_matrix_identity_index(oracle_lib):  # matrix.rs:276-304
    m = np.zeros(16, np.int32)
    oracle_lib.orc_mat4_identity_i32(_p(m))
    assert m.reshape(4, 4).tolist() == np.eye(4, dtype=np.int32).tolist()


def test_matrix_mul_exact_i32(oracle_lib):  # matrix.rs:307-377 (mul and mul_assign share the product)
    a = np.array([1, 2, 3, 4, 5, 6, 7, 8, 9, 39, 11, 12, 13, 14, 15, 16], np.int32)
    b = np.arange(17, 33, dtype=np.int32)
    out = np.zeros(16, np.int32)
    oracle_lib.orc_mat4_mul_i32(_p(a), _p(b), _p(out))
    assert out.tolist() == [250, 260, 270, 280, 618, 644, 670, 696, 1595, 1666, 1737, 1808, 1354, 1412, 1470, 1528]


def test_matrix_try_inverse(oracle_lib):  # matrix.rs:379-407, tolerance 0.1
    a = np.array([1, 0, 0, 1, 0, 2, 1, 2, 2, 1, 0, 1, 2, 0, 1, 4], np.float32)
    out = np.zeros(16, np.float32)
    assert oracle_lib.orc_mat4_try_inverse(_p(a), _p(out)) == 1
    want = [-2.0, -0.5, 1.0, 0.5, 1.0, 0.5, 0.0, -0.5, -8.0, -1.0, 2.0, 2.0, 3.0, 0.5, -1.0, -0.5]
    assert np.all(np.abs(out - np.array(want, np.float32)) <= 0.1)
    singular = np.zeros(16, np.float32)
    assert oracle_lib.orc_mat4_try_inverse(_p(singular), _p(out)) == 0


def test_quaternion_mul(oracle_lib):  # quaternion.rs:166-177, tolerance 0.01
    a = np.array([12.4, 1.1, 2.0, 4.4], np.float32)
    b = np.array([4.0, 0.3, 45.0, 5.0], np.float32)
    out = np.zeros(4, np.float32)
    oracle_lib.orc_quat_mul(_p(a), _p(b), _p(out))
    assert np.all(np.abs(out - np.array([-62.73, -179.88, 561.82, 128.5], np.float32)) <= 0.01)


def test_quaternion_rotation_matrix(oracle_lib):  # quaternion.rs:179-201, tolerance 0.02
    q = np.array([0.56, 0.77, -0.31, 0.0], np.float32)
    out = np.zeros(16, np.float32)
    oracle_lib.orc_quat_rotation_matrix(_p(q), _p(out))
    want = [0.80, -0.47, -0.34, 0.0, -0.47, -0.18, -0.86, 0.0, 0.34, 0.86, -0.37, 0.0, 0.0, 0.0, 0.0, 1.0]
    assert np.all(np.abs(out - np.array(want, np.float32)) <= 0.02)


def test_quaternion_from_euler(oracle_lib):  # quaternion.rs:249-259, tolerance 0.01
    angles = np.array([0.4, 1.3, 5.0], np.float32)
    out = np.zeros(4, np.float32)
    oracle_lib.orc_quat_from_euler(_p(angles), _p(out))
    assert np.all(np.abs(out - np.array([-0.55, -0.48, -0.38, 0.56], np.float32)) <= 0.01)
    # the survey's float32 emulation of the same call
    assert np.allclose(out, [-0.5531089, -0.48167437, -0.38052386, 0.5632601], atol=2e-7)


def test_constructors_and_layout(oracle_lib):
    t = np.array([3, 4, 5], np.float32)
    out = np.zeros(16, np.float32)
    oracle_lib.orc_new_translation(_p(t), _p(out))  # matrix.rs:61-71: elements 3, 7, 11
    assert out.tolist() == [1, 0, 0, 3, 0, 1, 0, 4, 0, 0, 1, 5, 0, 0, 0, 1]
    oracle_lib.orc_new_scale(_p(t), _p(out))  # matrix.rs:80-90
    assert out.tolist() == [3, 0, 0, 0, 0, 4, 0, 0, 0, 0, 5, 0, 0, 0, 0, 1]
    m = np.arange(16, dtype=np.float32)
    oracle_lib.orc_to_column_major(_p(m), _p(out))  # matrix.rs:219-251
    assert out.tolist() == [0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15]
    p = np.array([1, 2, 3], np.float32)
    o3 = np.zeros(3, np.float32)
    oracle_lib.orc_transform_vec3(_p(m), _p(p), _p(o3))  # matrix.rs:160-171: M.(v,1)
    assert o3.tolist() == [0 * 1 + 1 * 2 + 2 * 3 + 3, 4 * 1 + 5 * 2 + 6 * 3 + 7, 8 * 1 + 9 * 2 + 10 * 3 + 11]
    d = np.zeros(12, np.float32)
    oracle_lib.orc_transform_default(_p(d))  # transform.rs:13-22
    assert d.tolist() == [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1]


def test_as_matrix4_golden_vectors():
    """tests/golden/as_matrix4.json: the survey's float32 emulation of transform.rs:28-38 with
    this image's glibc sinf/cosf + Transform::default()."""
    gold = json.load(open(os.path.join(HERE, "golden", "as_matrix4.json")))
    for case in gold["cases"]:
        t = np.array([case["transform"]], np.float32)
        got = oracle.as_matrix4(t)[0]
        want = np.array([int(h, 16) for h in case["matrix_hex"]], np.uint32).view(np.float32)
        assert np.array_equal(got, want), (case["name"], got.tolist(), want.tolist())


def test_as_matrix4_is_the_four_product_chain(oracle_lib):
    """S.T(t).T(c).R.T(-c), left-associative, built from the pinned primitives"""
    rng = np.random.default_rng(1)
    for _ in range(200):
        t = rng.uniform(-5, 5, 12).astype(np.float32)
        S, T, Tc, Tn, R = (np.zeros(16, np.float32) for _ in range(5))
        q = np.zeros(4, np.float32)
        oracle_lib.orc_new_scale(_p(t[9:12].copy()), _p(S))
        oracle_lib.orc_new_translation(_p(t[0:3].copy()), _p(T))
        oracle_lib.orc_new_translation(_p(t[6:9].copy()), _p(Tc))
        oracle_lib.orc_new_translation(_p((-t[6:9]).copy()), _p(Tn))
        oracle_lib.orc_quat_from_euler(_p(t[3:6].copy()), _p(q))
        oracle_lib.orc_quat_rotation_matrix(_p(q), _p(R))
        m = S
        for rhs in (T, Tc, R, Tn):
            out = np.zeros(16, np.float32)
            oracle_lib.orc_mat4_mul_f32(_p(m), _p(rhs), _p(out))
            m = out
        assert np.array_equal(m, oracle.as_matrix4(t[None])[0])


def test_closed_form_equals_the_chain():
    """The device kernels evaluate as_matrix4 in closed form. It must be
    value-identical (==, so +-0 are equal) to the four-product chain."""
    rng = np.random.default_rng(2)
    n = 200_000
    t = np.zeros((n, 12), np.float32)
    t[:, 0:3] = rng.uniform(-1000, 1000, (n, 3))
    t[:, 3:6] = rng.uniform(-7, 7, (n, 3))
    t[:, 6:9] = rng.uniform(-64, 64, (n, 3))
    t[:, 9:12] = rng.uniform(-3, 3, (n, 3))
    t[: n // 4, 3:5] = 0          # planar rotations
    t[n // 4: n // 2, 6:9] = 0    # no rotation centre
    t[n // 2: n // 2 + 1000, 9:12] = 0  # zero scale
    a = oracle.as_matrix4(t)
    b = oracle.as_matrix4_closed(t)
    assert np.array_equal(a, b)


def test_sort_key_orders_like_floats():
    zs = np.array([-np.inf, -1e30, -2.5, -1.0, -1e-30, -0.0, 0.0, 1e-30, 1.0, 2.5, 1e30, np.inf], np.float32)
    keys = [oracle.sort_key(3, 7, z) for z in zs]
    assert keys[5] == keys[6]  # -0 and +0 fold
    assert all(a <= b for a, b in zip(keys, keys[1:]))
    assert oracle.sort_key(1, 0, 5.0) > oracle.sort_key(0, 65535, 1e30)  # layer dominates
    assert oracle.sort_key(0, 2, -5.0) > oracle.sort_key(0, 1, 5.0)       # then texture
    assert keys[8] >> 32 == (3 << 16 | 7)


def test_orthographic_matches_matrix_rs_and_the_library():
    """Matrix4::new_orthographic (matrix.rs:41-58): entries by hand, and the product library's host helper
    (tb_orthographic, no GPU needed) is bit-identical to the oracle's restatement."""
    from tuber_legacy_b200 import capi
    m = oracle.orthographic(0.0, 800.0, 0.0, 600.0, -1.0, 1.0)
    f = np.float32
    expect = np.array([[f(2) / f(800), 0, 0, -(f(800) / f(800))],
                       [0, f(2) / f(600), 0, -(f(600) / f(600))],
                       [0, 0, f(-2) / f(2), -(f(0) / f(2))],
                       [0, 0, 0, 1]], np.float32)
    np.testing.assert_array_equal(m, expect)
    rng = np.random.default_rng(5)
    for _ in range(200):
        l, b, n = rng.uniform(-1000, 0, 3)
        r, t, fa = rng.uniform(1, 1000, 3)
        a = oracle.orthographic(l, r, b, t, n, fa)
        g = capi.orthographic(l, r, b, t, n, fa)
        np.testing.assert_array_equal(a.view(np.uint32), g.view(np.uint32))
