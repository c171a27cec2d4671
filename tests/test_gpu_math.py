"""The rest of tuber-math on the device: tb_quat_* / tb_mat4_* against the oracle's restatement of
crates/tuber-math/src/quaternion.rs:28-158 and matrix.rs:98-194 — bit for bit — and against the reference's own
known-answer tests."""
import numpy as np
import pytest

from oracle import oracle
from tuber_legacy_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(16)
    yield c
    c.close()


def same(a, b):
    """== as floats (+0 == -0), NaN equal to NaN."""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_quat_from_euler_and_rotation_matrix(ctx):
    rng = np.random.default_rng(1)
    e = rng.uniform(-7, 7, (200_000, 3)).astype(np.float32)
    e[:1000, 0] = 0
    e[500:1500, 1] = 0
    q = ctx.quat_from_euler(e)
    assert same(q, oracle.quat_from_euler(e))
    assert same(ctx.quat_rotation_matrix(q), oracle.quat_rotation_matrix(q))
    # quaternion.rs:250-259: from_euler(0.1, 0.2, 0.3) within 0.01
    k = ctx.quat_from_euler([[0.1, 0.2, 0.3]])[0]
    assert np.allclose(k, [0.983347, 0.0342708, 0.106021, 0.143572], atol=0.01)


def test_quat_from_axis_angle(ctx):
    rng = np.random.default_rng(2)
    a = rng.uniform(-3, 3, (100_000, 4)).astype(np.float32)
    assert same(ctx.quat_from_axis_angle(a), oracle.quat_from_axis_angle(a))


def test_quat_mul_and_normalize(ctx):
    rng = np.random.default_rng(3)
    a = rng.normal(size=(200_000, 4)).astype(np.float32)
    b = rng.normal(size=(200_000, 4)).astype(np.float32)
    assert same(ctx.quat_mul(a, b), oracle.quat_mul(a, b))
    assert same(ctx.quat_normalize(a), oracle.quat_normalize(a))
    # quaternion.rs:166-177: (1 + 2i + 3j + 4k)(5 + 6i + 7j + 8k) = -60 + 12i + 30j + 24k
    assert same(ctx.quat_mul([[1, 2, 3, 4]], [[5, 6, 7, 8]]), np.array([[-60, 12, 30, 24]], np.float32))
    n = ctx.quat_normalize(a)
    assert np.allclose((n.astype(np.float64) ** 2).sum(axis=1), 1.0, atol=1e-6)


def test_mat4_mul_and_transform(ctx):
    rng = np.random.default_rng(4)
    a = rng.uniform(-10, 10, (100_000, 16)).astype(np.float32)
    b = rng.uniform(-10, 10, (100_000, 16)).astype(np.float32)
    v = rng.uniform(-100, 100, (100_000, 3)).astype(np.float32)
    assert same(ctx.mat4_mul(a, b), oracle.mat4_mul(a, b))
    assert same(ctx.mat4_transform_vec3(a, v), oracle.mat4_transform_vec3(a, v))
    # matrix.rs:309-377: the exact integer product pins the row-major convention
    A = np.arange(1, 17, dtype=np.float32).reshape(1, 16)
    B = np.arange(17, 33, dtype=np.float32).reshape(1, 16)
    want = (A.reshape(4, 4).astype(np.float64) @ B.reshape(4, 4).astype(np.float64)).reshape(1, 16)
    assert same(ctx.mat4_mul(A, B), want.astype(np.float32))
    assert want[0, 0] == 250 and want[0, 15] == 1528


def test_mat4_try_inverse(ctx):
    rng = np.random.default_rng(5)
    m = rng.uniform(-4, 4, (200_000, 16)).astype(np.float32)
    m[:100] = 0                       # singular: None
    m[100:200, 4:8] = m[100:200, 0:4]  # two equal rows: det ~ 0
    m[200:300] *= np.float32(1e-3)    # |det| < 1e-8 although invertible: the reference still says None
    got, some = ctx.mat4_try_inverse(m)
    want, wsome = oracle.mat4_try_inverse(m)
    assert np.array_equal(some, wsome)
    assert some[:100].sum() == 0 and some[200:300].sum() == 0
    assert same(got, want)
    # matrix.rs:381-407: inverse of a translation * scale is its algebraic inverse; product with the input is I
    k = np.array([[2, 0, 0, 3, 0, 4, 0, 5, 0, 0, 8, 7, 0, 0, 0, 1]], np.float32)
    inv, ok = ctx.mat4_try_inverse(k)
    assert ok[0] == 1
    assert np.allclose(ctx.mat4_mul(k, inv).reshape(4, 4), np.eye(4), atol=1e-6)


def test_orthographic_inverse_maps_clip_back_to_world(ctx):
    """screen -> world, the use try_inverse exists for: unproject the clip-space corners of new_orthographic."""
    vp = capi.orthographic(-400, 400, -300, 300, -1, 1).reshape(1, 16)
    inv, ok = ctx.mat4_try_inverse(vp)
    assert ok[0] == 1
    corners = np.array([[-1, -1, 0], [1, 1, 0]], np.float32)
    w = ctx.mat4_transform_vec3(np.repeat(inv, 2, axis=0), corners)
    assert np.allclose(w[:, :2], [[-400, -300], [400, 300]], atol=1e-3)
