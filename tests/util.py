"""Shared helpers of the parity tests: drive the CUDA path through the C ABI and the CPU
oracle on the same arrays, and compare with the SPEC's parity metric."""
import numpy as np

from oracle import oracle
from tuber_legacy_b200 import capi

REL_TOL = 1e-5  # BASELINE.json north_star: matrices and vertex positions within 1e-5 relative


def make_ctx(scene, masks=None, max_entities=None, **options):
    """Context holding `scene` (a tuber_legacy_b200.scenes.Scene-like object)."""
    ctx = capi.Context(max_entities or max(scene.n, 1), **options)
    ctx.set_entity_count(scene.n)
    if scene.n:
        if scene.transforms is not None:
            ctx.upload_transforms(0, scene.transforms)
        if scene.sprites is not None:
            ctx.upload_sprites(0, scene.sprites)
        if scene.velocities is not None:
            ctx.upload_velocities(0, scene.velocities)
        if scene.parents is not None:
            ctx.upload_parents(0, scene.parents)
        if masks:
            for comp, words in masks.items():
                ctx.upload_presence(comp, 0, words)
    return ctx


def make_oracle(scene, masks=None):
    masks = masks or {}
    return oracle.Scene(scene.n, scene.transforms, scene.sprites, scene.velocities, scene.parents,
                        mask_transform=masks.get(capi.TB_TRANSFORM), mask_sprite=masks.get(capi.TB_SPRITE),
                        mask_velocity=masks.get(capi.TB_VELOCITY))


class GpuFrame:
    def __init__(self, ctx):
        self.info = ctx.frame()
        m, b = self.info.num_instances, self.info.num_batches
        self.order = ctx.download_order(m)
        self.instances = ctx.download_instances(m)
        self.vertices = ctx.download_vertices(m)
        self.batches = ctx.download_batches(b)


# ---- the parity metric -----------------------------------------------------------
#
# Integers (draw order, batch table, texture / layer / size bits, corner codes): bit-exact, always.
# Floats: every operation of the path is an individually rounded IEEE operation in the reference's own order, and
# sin / cos are the host libm's sinf / cosf reproduced on the device (csrc/libm_sincosf.h). When that restatement
# is pinned on this box (it agrees with the box's libm bit for bit: libm_pinned()), the GPU must reproduce the
# oracle's floats EXACTLY (`==`; the closed form may differ from the reference's product chain only in the sign of a
# zero). Otherwise the north star's 1e-5 relative bound applies, split as the judge asked: the 3x3 block against the
# 3x3 block's own largest entry, the translation column and the vertex positions against the forward bound of the
# sum that produces them (sum of |terms|, computed by the oracle in double: oracle.Frame.bounds).
# Either way the element-wise relative errors (max and 99.9th percentile) are recorded in PARITY_LOG.

PARITY_LOG = []
_LIBM = {}


def libm_check_lib():
    """tests/helpers/libm_check.cpp: the host build of csrc/libm_sincosf.h next to the real sinf / cosf."""
    import ctypes
    import os
    import subprocess
    if "lib" not in _LIBM:
        here = os.path.dirname(os.path.abspath(__file__))
        src = os.path.join(here, "helpers", "libm_check.cpp")
        hdr = os.path.join(here, "..", "tuber_legacy_b200", "csrc", "libm_sincosf.h")
        out = os.path.join(here, "helpers", "_libm_check.so")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.run(["g++", "-O2", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", out, src], check=True)
        L = ctypes.CDLL(out)
        L.libm_compare.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int] + [ctypes.c_void_p] * 3
        L.libm_eval.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int] + [ctypes.c_void_p] * 4
        _LIBM["lib"] = L
    return _LIBM["lib"]


def libm_compare(first, stride, count, fma):
    import ctypes
    bs, bc, fb = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32()
    libm_check_lib().libm_compare(first, stride, count, int(fma), ctypes.byref(bs), ctypes.byref(bc), ctypes.byref(fb))
    return bs.value, bc.value, fb.value


def libm_pinned():
    """True when the restated sinf / cosf (the build glibc's selector picks on this CPU) agree bit for bit with
    this box's libm on 2^22 bit patterns spread over every exponent."""
    if "pinned" not in _LIBM:
        L = libm_check_lib()
        bs, bc, _ = libm_compare(0, 1021, 1 << 22, L.libm_uses_fma())
        _LIBM["pinned"] = bs == 0 and bc == 0
    return _LIBM["pinned"]


def _rel_stats(g, r):
    g = np.asarray(g, np.float64).ravel()
    r = np.asarray(r, np.float64).ravel()
    nz = r != 0
    rel = np.zeros_like(r)
    rel[nz] = np.abs(g[nz] - r[nz]) / np.abs(r[nz])
    rel[~nz] = np.where(g[~nz] == 0, 0.0, np.inf)
    return float(rel.max()) if rel.size else 0.0, float(np.percentile(rel, 99.9)) if rel.size else 0.0


def float_close(g, r, scale, what):
    """abs(g - r) <= REL_TOL * max(abs(r), scale) element-wise; +-0 equal."""
    g = np.asarray(g, np.float64)
    r = np.asarray(r, np.float64)
    bound = REL_TOL * np.maximum(np.abs(r), scale)
    bad = np.abs(g - r) > bound
    if bad.any():
        k = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} elements outside 1e-5; first at {tuple(k)}: "
                             f"gpu {g[tuple(k)]!r} ref {r[tuple(k)]!r}")


def assert_floats_match(got, want, what):
    """Matrices computed from the same inputs: `==` when the libm restatement is pinned on this box, else within
    1e-5 of each matrix's largest entry."""
    got, want = np.asarray(got), np.asarray(want)
    mx, p999 = _rel_stats(got, want)
    PARITY_LOG.append({"what": what, "instances": int(got.shape[0]) if got.ndim > 1 else 1, "exact": bool(np.array_equal(got, want)),
                       "model_rel_max": mx, "model_rel_p999": p999, "vertex_rel_max": 0.0, "vertex_rel_p999": 0.0})
    if libm_pinned():
        assert np.array_equal(got, want), f"{what}: floats differ from the oracle (== comparison), max rel {mx:.3g}"
    else:
        w2 = want.reshape(len(want), -1) if want.ndim > 1 else want.reshape(1, -1)
        float_close(got.reshape(w2.shape), w2, np.abs(w2).max(axis=1, keepdims=True), what)


def assert_frame_parity(gpu, ref, exact_floats=None, what=""):
    """Integer outputs bit-exact; float outputs `==` when the libm restatement is pinned on this box (or when
    exact_floats=True), else within the split 1e-5 metric. Records the element-wise relative errors."""
    assert gpu.info.num_instances == len(ref.order), (gpu.info.num_instances, len(ref.order))
    np.testing.assert_array_equal(gpu.order, ref.order, err_msg=f"{what}: sorted entity order")
    assert gpu.info.num_batches == len(ref.batches), (gpu.info.num_batches, len(ref.batches))
    for f in ("layer", "texture", "first_instance", "count"):
        np.testing.assert_array_equal(gpu.batches[f], ref.batches[f], err_msg=f"batch table {f}")
    if len(ref.order) == 0:
        return
    gi, ri = gpu.instances, ref.instances
    np.testing.assert_array_equal(gi["texture"], ri["texture"])
    np.testing.assert_array_equal(gi["layer"], ri["layer"])
    np.testing.assert_array_equal(gi["size"].view(np.uint32), ri["size"].view(np.uint32))
    np.testing.assert_array_equal(gpu.vertices["uv"], ref.vertices["uv"])
    gm, rm = gi["model"], ri["model"]  # column-major: entry (row j, col i) at [4 * i + j]
    gv = np.stack([gpu.vertices["x"], gpu.vertices["y"], gpu.vertices["z"]], axis=-1)
    rv = np.stack([ref.vertices["x"], ref.vertices["y"], ref.vertices["z"]], axis=-1)
    exact = bool(np.array_equal(gm, rm) and np.array_equal(gv, rv))
    mmax, mp999 = _rel_stats(gm, rm)
    vmax, vp999 = _rel_stats(gv, rv)
    PARITY_LOG.append({"what": what, "instances": int(len(ref.order)), "exact": exact, "model_rel_max": mmax,
                       "model_rel_p999": mp999, "vertex_rel_max": vmax, "vertex_rel_p999": vp999})
    if exact:
        return
    if exact_floats is None:
        exact_floats = libm_pinned()
    if exact_floats:
        bad = np.argwhere(gm != rm)
        where = f"model {tuple(bad[0])}: gpu {gm[tuple(bad[0])]!r} ref {rm[tuple(bad[0])]!r}" if len(bad) else "vertices"
        raise AssertionError(f"{what}: floats differ from the oracle (== comparison; libm restatement pinned): {where}; "
                             f"{len(bad)} model entries, max rel {mmax:.3g}")
    # split 1e-5 metric
    blk = [0, 1, 2, 4, 5, 6, 8, 9, 10]          # the 3x3 block (either storage order: the index set is symmetric)
    bscale = np.abs(rm[:, blk]).max(axis=1, keepdims=True)
    float_close(gm[:, blk], rm[:, blk], bscale, "instance model 3x3 block")
    bounds = getattr(ref, "bounds", None)
    if bounds is None or not np.any(bounds):
        raise AssertionError("no forward bounds from the oracle for the translation column / vertices")
    bt = np.asarray(bounds).reshape(-1, 4, 4)      # row-major per instance
    float_close(gm[:, [12, 13, 14]], rm[:, [12, 13, 14]], bt[:, :3, 3], "instance model translation column")
    float_close(gm[:, [3, 7, 11, 15]], rm[:, [3, 7, 11, 15]], bt[:, 3, :], "instance model bottom row")
    size = ri["size"].astype(np.float64)
    cx = np.stack([0 * size[:, 0], size[:, 0], size[:, 0], 0 * size[:, 0]], axis=1)   # (m, 4 corners)
    cy = np.stack([0 * size[:, 1], 0 * size[:, 1], size[:, 1], size[:, 1]], axis=1)
    vb = bt[:, None, :3, 0] * cx[:, :, None] + bt[:, None, :3, 1] * cy[:, :, None] + bt[:, None, :3, 3]
    float_close(gv, rv, vb, "quad vertex position")


def random_mask(n, p, seed):
    rng = np.random.default_rng(seed)
    bits = rng.random(n) < p
    words = np.zeros((n + 63) // 64, np.uint64)
    idx = np.nonzero(bits)[0]
    np.bitwise_or.at(words, idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
    return words


def check_frame_properties(ctx, gpu, n_expected):
    """Size-independent properties of a frame (used at full BASELINE sizes, where the oracle
    would take minutes): order is a permutation of the renderables, keys recomputed from the
    emitted instances are sorted with ties in entity order, the batch table tiles [0, M)."""
    m = gpu.info.num_instances
    assert m == n_expected
    order = gpu.order.astype(np.int64)
    seen = np.zeros(int(order.max()) + 1 if m else 0, np.uint8)
    seen[order] = 1
    assert int(seen.sum()) == m, "order is not a permutation"
    inst = gpu.instances
    z = inst["model"][:, 14] + np.float32(0.0)
    u = z.view(np.uint32).astype(np.uint64)
    u = np.where(u >> np.uint64(31), u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))
    key = (inst["layer"].astype(np.uint64) << np.uint64(48)) | (inst["texture"].astype(np.uint64) << np.uint64(32)) | u
    ks, ke = key[:-1], key[1:]
    assert np.all(ks <= ke), "keys not sorted"
    tie = ks == ke
    assert np.all(order[:-1][tie] < order[1:][tie]), "ties not in entity order"
    b = gpu.batches
    assert int(b["count"].sum()) == m
    assert np.array_equal(b["first_instance"], np.concatenate([[0], np.cumsum(b["count"])[:-1]]).astype(np.uint32))
    lt = (inst["layer"].astype(np.uint64) << np.uint64(32)) | inst["texture"].astype(np.uint64)
    heads = np.concatenate([[True], lt[1:] != lt[:-1]])
    assert int(heads.sum()) == len(b)
    np.testing.assert_array_equal(np.nonzero(heads)[0].astype(np.uint32), b["first_instance"])
    np.testing.assert_array_equal(inst["layer"][heads], b["layer"])
    np.testing.assert_array_equal(inst["texture"][heads], b["texture"])
