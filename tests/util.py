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


def assert_frame_parity(gpu, ref, exact_floats=False):
    """Integer outputs bit-exact; float outputs within REL_TOL norm-wise per matrix / quad.
    Hierarchy frames: a world translation is a sum of ancestors' terms that can cancel to ~0, so the
    scale is at least the scene extent (largest entry of any reference world matrix)."""
    assert gpu.info.num_instances == len(ref.order), (gpu.info.num_instances, len(ref.order))
    np.testing.assert_array_equal(gpu.order, ref.order, err_msg="sorted entity order")
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
    gm, rm = gi["model"], ri["model"]
    gv = np.stack([gpu.vertices["x"], gpu.vertices["y"], gpu.vertices["z"]], axis=-1)
    rv = np.stack([ref.vertices["x"], ref.vertices["y"], ref.vertices["z"]], axis=-1)
    if exact_floats:
        assert np.array_equal(gm, rm), "model matrices differ (== comparison)"
        assert np.array_equal(gv, rv), "vertex positions differ (== comparison)"
        return
    scale = np.abs(rm).max(axis=1, keepdims=True)
    if getattr(gpu.info, "hierarchy_levels", 0):
        scale = np.maximum(scale, np.abs(rm).max())
    float_close(gm, rm, scale, "instance model matrix")
    vscale = np.abs(rv).max(axis=(1, 2), keepdims=True)
    float_close(gv, rv, np.maximum(vscale, scale[:, :, None]), "quad vertex position")


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
