"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded scenes (sizes the oracle finishes in seconds). Integer outputs (sorted order, batch
table, texture/layer fields) are compared bit-exact. Floats: the device evaluates the host libm's
sinf / cosf bit for bit and rounds every other operation like the reference, so matrices and vertex
positions must EQUAL the oracle's (`==`) wherever that restatement is pinned on the box (util.libm_pinned);
elsewhere the split 1e-5 metric of util.assert_frame_parity applies.
"""
import numpy as np
import pytest

from oracle import oracle
from tuber_legacy_b200 import capi, scenes

import util

pytestmark = pytest.mark.gpu


def run_both(scene, masks=None, ticks=0, dt=0.01, **options):
    ctx = util.make_ctx(scene, masks, **options)
    ref = util.make_oracle(scene, masks)
    try:
        for _ in range(ticks):
            ctx.system_integrate(dt)
            ref.step(dt)
        g = util.GpuFrame(ctx)
        r = ref.frame()
        return ctx, ref, g, r
    except Exception:
        ctx.close()
        raise


# ---- BASELINE configs at oracle-sized N ---------------------------------------------------------

def test_config1_10k_flat():
    """configs[0]: 10k flat sprites, 1 texture, no hierarchy — the reference's own CPU-runnable case
    (fits the reference's 65 536-entity presence bitset)."""
    sc = scenes.config1()
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r)
    assert g.info.path == capi.TB_PATH_BUCKET and g.info.num_batches == 1
    ctx.close()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 2047, 2048, 2049, 3072, 6145, 70_001])
def test_flat_sizes(n):
    """ragged sizes around the warp / tile boundaries of the rank and emit kernels"""
    sc = scenes.config2(n=n, seed=20 + n % 7)
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r)
    ctx.close()


@pytest.mark.parametrize("z_mode", ["steps", "random"])
def test_config2_flat_64_textures_16_layers(z_mode):
    sc = scenes.config2(n=150_000, z_mode=z_mode)
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r)
    assert g.info.path == capi.TB_PATH_SORT
    if z_mode == "random":
        assert g.info.key_bits > 32  # 64-bit short keys
    ctx.close()


@pytest.mark.parametrize("random_parents", [False, True])
def test_config3_hierarchy_depth8(random_parents):
    sc = scenes.config3(n=120_000, random_parents=random_parents)
    ctx, ref, g, r = run_both(sc)
    assert g.info.hierarchy_levels == 8
    util.assert_frame_parity(g, r)
    w = ctx.download_world_matrices(0, sc.n)
    if util.libm_pinned():
        assert np.array_equal(w, r.world), "world matrices differ from the oracle (== comparison)"
    else:
        util.float_close(w[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]], r.world[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]],
                         np.abs(r.world[:, [0, 1, 2, 4, 5, 6, 8, 9, 10]]).max(axis=1, keepdims=True), "world 3x3 block")
    ctx.close()


def test_hierarchy_shuffled_ids_uses_level_lists():
    """ids not in BFS order: levels are not id ranges, the level kernels walk sorted lists"""
    sc = scenes.config3(n=30_000, seed=31)
    rng = np.random.default_rng(5)
    perm = rng.permutation(sc.n).astype(np.uint32)  # new id of old entity i
    inv = np.empty_like(perm)
    inv[perm] = np.arange(sc.n, dtype=np.uint32)
    sc.transforms = sc.transforms[inv]
    sc.sprites = sc.sprites[inv]
    old_parents = sc.parents[inv]
    sc.parents = np.where(old_parents == capi.TB_NO_PARENT, capi.TB_NO_PARENT, perm[np.minimum(old_parents, sc.n - 1)]).astype(np.uint32)
    ctx, ref, g, r = run_both(sc)
    assert g.info.hierarchy_levels == 8
    util.assert_frame_parity(g, r)
    ctx.close()


def test_config4_particles_velocity_system():
    sc = scenes.config4(n=200_000)
    ctx, ref, g, r = run_both(sc, ticks=1, dt=sc.dt)
    util.assert_frame_parity(g, r)
    assert g.info.path == capi.TB_PATH_BUCKET and g.info.num_batches == 4
    # the velocity system's write-back is bit-exact
    t = ctx.download_transforms(0, sc.n)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), ref.transforms().view(np.uint32))
    # several more ticks and frames (steady state reuses the cached plan)
    for _ in range(3):
        ctx.system_integrate(sc.dt)
        ref.step(sc.dt)
        g = util.GpuFrame(ctx)
        r = ref.frame()
        util.assert_frame_parity(g, r)
        assert g.info.replanned == 0
    # two ticks before one frame: both are applied, in order
    for _ in range(2):
        ctx.system_integrate(sc.dt)
        ref.step(sc.dt)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, ref.frame())
    ctx.close()


def test_velocity_moves_z_and_replans():
    """z velocities change the key bits between frames without any upload: the cached plan goes
    stale, the frame re-plans itself and still matches"""
    sc = scenes.config4(n=50_000, seed=44)
    idx = np.arange(sc.n)
    sc.velocities["v"][:, 2] = ((idx % 7) - 3).astype(np.float32) * 10
    ctx = util.make_ctx(sc)
    ref = util.make_oracle(sc)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, ref.frame())
    replans = 0
    for _ in range(4):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        g = util.GpuFrame(ctx)
        util.assert_frame_parity(g, ref.frame())
        replans += g.info.replanned
    assert replans >= 1
    t = ctx.download_transforms(0, sc.n)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), ref.transforms().view(np.uint32))
    ctx.close()


def test_hierarchy_with_velocity():
    sc = scenes.config3(n=40_000, seed=33)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = np.linspace(-50, 50, sc.n, dtype=np.float32)
    v["v"][:, 1] = 3.0
    sc.velocities = v
    ctx, ref, g, r = run_both(sc, ticks=2, dt=0.01)
    util.assert_frame_parity(g, r)
    ctx.close()


# ---- presence masks / query ------------------------------------------------------------------------

def test_sparse_presence_masks():
    sc = scenes.config4(n=90_001, seed=7)
    sc.sprites["texture"] = np.arange(sc.n) % 5
    sc.sprites["layer"] = np.arange(sc.n) % 3
    masks = {capi.TB_TRANSFORM: util.random_mask(sc.n, 0.8, 1), capi.TB_SPRITE: util.random_mask(sc.n, 0.6, 2),
             capi.TB_VELOCITY: util.random_mask(sc.n, 0.5, 3)}
    ctx, ref, g, r = run_both(sc, masks, ticks=1)
    assert 0 < g.info.num_instances < sc.n
    util.assert_frame_parity(g, r)
    t = ctx.download_transforms(0, sc.n)
    tm = masks[capi.TB_TRANSFORM]
    present = ((tm[np.arange(sc.n) // 64] >> (np.arange(sc.n) % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12)[present], ref.transforms().view(np.uint32)[present])
    ctx.close()


def test_sparse_masks_multipass():
    sc = scenes.config2(n=80_000, seed=9, z_mode="random")
    masks = {capi.TB_TRANSFORM: util.random_mask(sc.n, 0.7, 4), capi.TB_SPRITE: util.random_mask(sc.n, 0.7, 5)}
    ctx, ref, g, r = run_both(sc, masks)
    assert g.info.sort_passes > 1
    util.assert_frame_parity(g, r)
    ctx.close()


def test_query_matches_reference_iteration():
    """Ecs::query: same matches as the reference's QueryIterator (which yields them descending)"""
    sc = scenes.config4(n=70_003, seed=8)
    masks = {capi.TB_TRANSFORM: util.random_mask(sc.n, 0.9, 11), capi.TB_SPRITE: util.random_mask(sc.n, 0.5, 12),
             capi.TB_VELOCITY: util.random_mask(sc.n, 0.3, 13)}
    ctx = util.make_ctx(sc, masks)
    ref = util.make_oracle(sc, masks)
    L = oracle.lib()
    import ctypes as C
    from oracle.oracle import _p
    TY = {capi.TB_TRANSFORM: 0x7472616e73666f72, capi.TB_SPRITE: 0x7370726974650000, capi.TB_VELOCITY: 0x76656c6f63697479}
    for comps in ([capi.TB_TRANSFORM], [capi.TB_TRANSFORM, capi.TB_SPRITE], [capi.TB_TRANSFORM, capi.TB_SPRITE, capi.TB_VELOCITY]):
        got = ctx.query(comps)
        types = np.array([TY[c] for c in comps], np.uint64)
        req = np.ones(len(comps), np.uint8)
        exc = np.zeros(len(comps), np.uint8)
        out = np.zeros(sc.n, np.uint64)
        cnt = L.orc_ecs_query(ref._h, len(comps), _p(types), _p(req), _p(exc), _p(out), C.c_void_p(0), sc.n)
        want = out[:cnt][::-1].astype(np.uint32)  # reference pops from the back: descending
        np.testing.assert_array_equal(got, want)
        assert ctx.query(comps, want_indices=False) == cnt
    # a required component with no store matches nothing (query.rs:110)
    assert ctx.query([capi.TB_PARENT], want_indices=False) == 0
    ctx.close()


# ---- compose -----------------------------------------------------------------------------------------

def test_as_matrix4_matches_reference_compose():
    rng = np.random.default_rng(0)
    n = 50_000
    t = np.zeros((n, 12), np.float32)
    t[:, 0:3] = rng.uniform(-1000, 1000, (n, 3))
    t[:, 3:6] = rng.uniform(-7, 7, (n, 3))
    t[:, 6:9] = rng.uniform(-64, 64, (n, 3))
    t[:, 9:12] = rng.uniform(-3, 3, (n, 3))
    t[: n // 2, 3:5] = 0  # planar half: rotation about z only
    # the survey's golden vectors + Transform::default()
    t[0] = [10, 20, 0, 0, 0, 0.5, 0, 0, 0, 2, 3, 1]
    t[1] = [1.5, -2.25, 0.75, 0.1, 0.2, 0.3, 4, 5, 6, 1, 2, 3]
    t[2] = [100, -50, 3, 0, 0, 2.5, 16, 16, 0, 1.5, 1.5, 1]
    t[3] = [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1]
    with capi.Context(16) as ctx:
        got = ctx.as_matrix4(t)
    want = oracle.as_matrix4(t)
    util.assert_floats_match(got, want, "as_matrix4")
    assert np.array_equal(got[3], np.eye(4, dtype=np.float32).reshape(-1))
    # everything that does not pass through sin/cos is exact: translation column of planar rows with c == 0
    t2 = t[: n // 2].copy()
    t2[:, 6:9] = 0
    with capi.Context(16) as ctx:
        got2 = ctx.as_matrix4(t2)
    want2 = oracle.as_matrix4(t2)
    assert np.array_equal(got2[:, [3, 7, 11]], want2[:, [3, 7, 11]])


def test_3d_rotation():
    """General Euler rotations about a NON-ZERO rotation centre: world z passes through all six sin / cos, so the
    sort key itself depends on them. The device evaluates the host libm's sinf / cosf bit for bit
    (csrc/libm_sincosf.h), so order, batch table and every float are the oracle's exactly."""
    sc = scenes.config2(n=60_000, seed=12, z_mode="random")
    rng = np.random.default_rng(3)
    sc.transforms["angle"][:, 0] = rng.uniform(-3, 3, sc.n)
    sc.transforms["angle"][:, 1] = rng.uniform(-3, 3, sc.n)
    sc.transforms["rotation_center"][:, 2] = rng.uniform(-20, 20, sc.n)
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r, what="3-D rotation, non-zero centre")
    if util.libm_pinned():
        assert np.array_equal(g.instances["model"], r.instances["model"])
    ctx.close()


def test_3d_rotation_near_ties_in_z():
    """Many entities whose world z differ in the last ulps only (same depth before rotation, tiny rotation
    differences): any sin / cos discrepancy would swap neighbours in the draw order."""
    sc = scenes.config2(n=40_000, seed=13)
    rng = np.random.default_rng(4)
    sc.sprites["texture"][:] = 0
    sc.sprites["layer"][:] = 0
    sc.transforms["translation"][:, 2] = 5.0
    sc.transforms["angle"][:, 0] = (0.3 + rng.integers(0, 64, sc.n) * 1e-7).astype(np.float32)
    sc.transforms["angle"][:, 1] = (-0.2 + rng.integers(0, 64, sc.n) * 1e-7).astype(np.float32)
    sc.transforms["rotation_center"][:, 0] = 7.0
    sc.transforms["rotation_center"][:, 1] = -3.0
    sc.transforms["rotation_center"][:, 2] = 2.0
    ctx, ref, g, r = run_both(sc)
    assert len(np.unique(r.keys)) > 100  # the keys really are spread over neighbouring floats
    util.assert_frame_parity(g, r, what="3-D rotation, near-ties")
    ctx.close()


@pytest.mark.parametrize("random_parents", [False, True])
def test_hierarchy_with_3d_rotated_parents(random_parents):
    """Rotated parents: a child's world z is a sum over the parent's rotated rows, i.e. sin / cos of every ancestor
    reach the sort key."""
    sc = scenes.config3(n=60_000, seed=21, random_parents=random_parents, z_mode="random")
    rng = np.random.default_rng(6)
    sc.transforms["angle"][:, 0] = rng.uniform(-1, 1, sc.n)
    sc.transforms["angle"][:, 1] = rng.uniform(-1, 1, sc.n)
    sc.transforms["rotation_center"][:, 2] = rng.uniform(-5, 5, sc.n)
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r, what="hierarchy, 3-D rotated parents")
    w = ctx.download_world_matrices(0, sc.n)
    if util.libm_pinned():
        assert np.array_equal(w, r.world)
    ctx.close()


# ---- sin/cos-free scenes: every float must match with == whatever the libm ------------------------------------

def _no_rotation(sc):
    sc.transforms["angle"][:] = 0  # sinf(0) = 0, cosf(0) = 1 on any libm: every product is exactly rounded
    return sc


@pytest.mark.parametrize("kind", ["particles", "sorted", "sorted64", "hierarchy", "hierarchy_random"])
def test_unrotated_scenes_match_exactly_on_every_emit_path(kind):
    sc = {"particles": lambda: scenes.config4(n=70_000, seed=31),
          "sorted": lambda: scenes.config2(n=70_000, seed=32),
          "sorted64": lambda: scenes.config2(n=70_000, seed=33, z_mode="random"),
          "hierarchy": lambda: scenes.config3(n=70_000, seed=34),
          "hierarchy_random": lambda: scenes.config3(n=70_000, seed=35, random_parents=True)}[kind]()
    _no_rotation(sc)
    ticks = 1 if sc.velocities is not None else 0
    ctx, ref, g, r = run_both(sc, ticks=ticks)
    util.assert_frame_parity(g, r, exact_floats=True, what=f"unrotated {kind}, first frame")
    vp = capi.orthographic(-1200, 1200, -900, 900, -100, 100)
    # steady frames: single-kernel / kept-order / draw-ordered-copy paths, then the same under a camera
    for k in range(6):
        for _ in range(ticks):
            ctx.system_integrate(0.01)
            ref.step(0.01)
        cam = vp if k >= 3 else None
        ctx.set_view_projection(cam)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(view_proj=cam), exact_floats=True,
                                 what=f"unrotated {kind}, steady frame {k}")
    ctx.close()


def test_tick_then_world_download_then_frame_revalidates_the_draw_order():
    """A velocity tick flushed by tb_download_world_matrices (outside a frame) must not let the next hierarchy frame
    replay the kept draw order: the tick moved z."""
    sc = scenes.config3(n=30_000, seed=41)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    rng = np.random.default_rng(8)
    v["v"][:, 2] = rng.uniform(-300, 300, sc.n)
    sc.velocities = v
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r)
    for _ in range(3):  # reach the kept-order replay path
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ctx.system_integrate(0.01)
    ref.step(0.01)
    w = ctx.download_world_matrices(0, sc.n)  # flushes the tick and recomputes the world rows
    r2 = ref.frame()
    assert not np.array_equal(r2.order, r.order)  # the tick really changed the order
    if util.libm_pinned():
        assert np.array_equal(w, r2.world)
    util.assert_frame_parity(util.GpuFrame(ctx), r2, what="frame after tick + world download")
    ctx.close()


def test_resharding_and_option_toggles_drop_kept_draw_orders():
    sc = scenes.config2(n=40_000, seed=43)
    ctx, ref, g, r = run_both(sc, sorted_cache=0)
    for _ in range(3):
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ctx.set_option("sorted_cache", 1)  # a graph captured without the kept order must not be replayed with it
    for _ in range(4):
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ctx.shard_set(0, 1, 1000)  # ids shift: kept orders hold the old global ids
    g = util.GpuFrame(ctx)
    assert np.array_equal(g.order, r.order + 1000)
    ctx.shard_set(0, 1, 0)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ctx.close()


# ---- every internal path gives the same bytes --------------------------------------------------------

@pytest.mark.parametrize("options", [
    {"interleaved_rows": 1},
    {"staged_emit": 0},
    {"digit_bits": 3},
    {"lt_limit": 0},
    {"lt_limit": 0, "digit_bits": 5, "staged_emit": 0, "interleaved_rows": 1},
    {"sorted_cache": 0},
    {"static_buckets": 0},
    {"cuda_graphs": 0},
    {"z_dictionary": 0},
    {"sorted_cache": 0, "static_buckets": 0, "cuda_graphs": 0, "z_dictionary": 0, "small_emit": 0},
    {"hier_fused": 0},
    {"chunked_sort": 1},
    {"chunked_sort": 1, "chunk_shift": 11},
    {"chunked_sort": 1, "chunk_shift": 13, "sorted_cache": 0},
    {"chunked_sort": 1, "chunk_shift": 12, "cuda_graphs": 0, "z_dictionary": 0},
])
def test_options_do_not_change_results(options):
    for sc in (scenes.config2(n=50_000, seed=14), scenes.config4(n=50_000, seed=15), scenes.config3(n=30_000, seed=16)):
        ctx, ref, g, r = run_both(sc, ticks=1 if sc.velocities is not None else 0, **options)
        util.assert_frame_parity(g, r)
        for _ in range(3):  # steady-state frames under the same options
            if sc.velocities is not None:
                ctx.system_integrate(0.01)
                ref.step(0.01)
            util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
        ctx.close()


def test_force_sort_path_on_bucket_scene():
    sc = scenes.config4(n=40_000, seed=17)
    ctx = util.make_ctx(sc)
    ctx.force_path(capi.TB_PATH_SORT)
    g = util.GpuFrame(ctx)
    assert g.info.path == capi.TB_PATH_SORT and g.info.sort_passes >= 2
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    ctx.close()


def test_many_layer_texture_pairs_generic_batches():
    """more than 2^16 distinct (layer, texture) bins: batch boundaries by run-head detection"""
    n = 200_000
    sc = scenes.config2(n=n, seed=18)
    rng = np.random.default_rng(1)
    sc.sprites["texture"] = rng.integers(0, 65536, n)
    sc.sprites["layer"] = rng.integers(0, 65535, n)
    ctx = util.make_ctx(sc)
    ctx.set_max_batches(n)
    g = util.GpuFrame(ctx)
    r = util.make_oracle(sc).frame()
    util.assert_frame_parity(g, r)
    assert g.info.num_batches > 65536
    ctx.close()


# ---- edge cases --------------------------------------------------------------------------------------

def test_empty_and_missing_stores():
    with capi.Context(1024) as ctx:
        out = ctx.frame()
        assert out.num_instances == 0 and out.num_batches == 0
        ctx.set_entity_count(100)
        assert ctx.frame().num_instances == 0  # no stores at all
        sc = scenes.config1(n=100)
        ctx.upload_transforms(0, sc.transforms)
        assert ctx.frame().num_instances == 0  # Transform only: (&Transform, &Sprite) matches nothing
        ctx.upload_sprites(0, sc.sprites)
        assert ctx.frame().num_instances == 100


def test_identical_keys_keep_entity_order():
    sc = scenes.config1(n=5000)
    sc.transforms["translation"][:, 2] = 0
    ctx, ref, g, r = run_both(sc)
    np.testing.assert_array_equal(g.order, np.arange(sc.n, dtype=np.uint32))
    util.assert_frame_parity(g, r)
    assert g.info.key_bits == 0
    ctx.close()


def test_negative_zero_and_negative_z():
    sc = scenes.config2(n=20_000, seed=19)
    z = np.linspace(-5, 5, sc.n).astype(np.float32)
    z[::97] = -0.0
    z[::89] = 0.0
    sc.transforms["translation"][:, 2] = z
    ctx, ref, g, r = run_both(sc)
    util.assert_frame_parity(g, r)
    ctx.close()


def test_errors_are_statuses_not_crashes():
    with capi.Context(1000) as ctx:
        with pytest.raises(capi.TuberError) as e:
            ctx.set_entity_count(1001)
        assert e.value.status == capi.TB_ERR_CAPACITY
        ctx.set_entity_count(10)
        sc = scenes.config1(n=11)
        with pytest.raises(capi.TuberError) as e:
            ctx.upload_transforms(0, sc.transforms)
        assert e.value.status == capi.TB_ERR_INVALID
        bad = scenes.config1(n=10)
        bad.sprites["layer"][3] = 70000
        with pytest.raises(capi.TuberError) as e:
            ctx.upload_sprites(0, bad.sprites)
        assert e.value.status == capi.TB_ERR_RANGE
        # a cycle in the Parent links
        ok = scenes.config1(n=10)
        ctx.upload_transforms(0, ok.transforms)
        ctx.upload_sprites(0, ok.sprites)
        parents = np.full(10, capi.TB_NO_PARENT, np.uint32)
        parents[2], parents[5] = 5, 2
        ctx.upload_parents(0, parents)
        with pytest.raises(capi.TuberError) as e:
            ctx.frame()
        assert e.value.status == capi.TB_ERR_CYCLE
        parents[5] = capi.TB_NO_PARENT
        ctx.upload_parents(0, parents)
        assert ctx.frame().num_instances == 10


def test_parent_without_transform_is_a_root():
    sc = scenes.config3(n=5000, seed=21)
    masks = {capi.TB_TRANSFORM: util.random_mask(sc.n, 0.9, 6)}
    ctx, ref, g, r = run_both(sc, masks)
    util.assert_frame_parity(g, r)
    ctx.close()


def test_reupload_and_growth():
    sc = scenes.config2(n=30_000, seed=22)
    ctx = util.make_ctx(sc, max_entities=60_000)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    # grow the entity count and change some sprites
    sc2 = scenes.config2(n=60_000, seed=22)
    sc2.sprites["texture"][:100] = 63
    ctx.set_entity_count(sc2.n)
    ctx.upload_transforms(30_000, sc2.transforms[30_000:])
    ctx.upload_sprites(30_000, sc2.sprites[30_000:])
    ctx.upload_sprites(0, sc2.sprites[:100])
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, util.make_oracle(sc2).frame())
    ctx.close()


def test_frames_are_idempotent():
    sc = scenes.config2(n=40_000, seed=23)
    ctx = util.make_ctx(sc)
    a = util.GpuFrame(ctx)
    b = util.GpuFrame(ctx)
    assert a.instances.tobytes() == b.instances.tobytes()
    assert a.vertices.tobytes() == b.vertices.tobytes()
    assert a.order.tobytes() == b.order.tobytes() and a.batches.tobytes() == b.batches.tobytes()
    ctx.close()


# ---- one scene over several contexts: exact global draw order (multi-GPU placement) ---------------

def _sharded_frame(sc, nranks, instances_only=False, **options):
    """Emulates `nranks` ranks as contexts on this GPU (run one after another, no waiting between
    them), all emitting into one set of output buffers; returns the frame read back from them."""
    n = sc.n
    bufs = [capi.device_alloc(max(n, 1) * 80), capi.device_alloc(max(n, 1) * 64), capi.device_alloc(max(n, 1) * 4)]
    ctxs = []
    for r in range(nranks):
        first, cnt = scenes.shard_range(n, r, nranks)
        part = scenes.Scene(sc.name, n, first, cnt)
        part.transforms, part.sprites = sc.transforms[first:first + cnt], sc.sprites[first:first + cnt]
        part.velocities = None if sc.velocities is None else sc.velocities[first:first + cnt]
        ctx = util.make_ctx(part, **options)
        ctx.shard_set(r, nranks, first)
        ctx.set_output(bufs[0], 0 if instances_only else bufs[1], bufs[2], n)
        ctxs.append(ctx)
    return ctxs, bufs


def _run_sharded(ctxs, bufs, ticks=0, dt=0.01, regenerate_vertices=False):
    for _ in range(ticks):
        for c in ctxs:
            c.system_integrate(dt)
    stats = [c.shard_key_stats() for c in ctxs]
    from tuber_legacy_b200.shard import reduce_stats
    g = reduce_stats(stats)
    nbins = [c.shard_plan(g) for c in ctxs]
    assert len(set(nbins)) == 1
    nb = nbins[0]
    d_all = capi.device_alloc(len(ctxs) * nb * 4)
    for r, c in enumerate(ctxs):
        c.shard_key_histogram(d_all + r * nb * 4)
    totals = [c.shard_place(d_all) for c in ctxs]
    assert len(set(totals)) == 1 and totals[0] == g[2]
    outs = [c.frame() for c in ctxs]
    m, b = totals[0], outs[0].num_batches
    if regenerate_vertices:  # the presenter rebuilds every quad from the instances it received
        ctxs[0].vertices_from_instances(bufs[0], 0, m, bufs[1])

    class G:
        pass
    gpu = G()
    gpu.info = outs[0]
    c0 = ctxs[0]
    gpu.order = c0.device_read(bufs[2], m * 4, np.uint32)
    gpu.instances = c0.device_read(bufs[0], m * 80, np.uint8).view(capi.INSTANCE_DTYPE)
    gpu.vertices = c0.device_read(bufs[1], m * 64, np.uint8).view(capi.VERTEX_DTYPE).reshape(-1, 4)
    gpu.batches = c0.download_batches(b)
    for c in ctxs[1:]:
        np.testing.assert_array_equal(c.download_batches(b), gpu.batches)  # every rank has the global table
    capi.device_free(d_all)
    return gpu


@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_sharded_scene_matches_single_oracle_frame(nranks):
    """ranks own contiguous id ranges; their emits interleave into the exact global order"""
    sc = scenes.config2(n=90_001, seed=40)  # 21 key bits: layers x textures x z steps, 3 radix passes
    ctxs, bufs = _sharded_frame(sc, nranks)
    g = _run_sharded(ctxs, bufs)
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    for c in ctxs:
        c.close()
    for b in bufs:
        capi.device_free(b)


def test_sharded_particles_with_velocity_system():
    sc = scenes.config4(n=120_000, seed=41)
    ref = util.make_oracle(sc)
    ctxs, bufs = _sharded_frame(sc, 4)
    for _ in range(2):
        ref.step(sc.dt)
        g = _run_sharded(ctxs, bufs, ticks=1, dt=sc.dt)
        util.assert_frame_parity(g, ref.frame())
    for c in ctxs:
        c.close()
    for b in bufs:
        capi.device_free(b)


def test_sharded_sparse_and_empty_shards():
    sc = scenes.config2(n=40_000, seed=42)
    masks = {capi.TB_SPRITE: util.random_mask(sc.n, 0.5, 9)}
    words = masks[capi.TB_SPRITE]
    words[: 10_000 // 64 + 1] = 0  # the first shard ends up with no renderables at all
    ref = util.make_oracle(sc, masks)
    n, nranks = sc.n, 4
    bufs = [capi.device_alloc(n * 80), capi.device_alloc(n * 64), capi.device_alloc(n * 4)]
    ctxs = []
    bits = ((words[np.arange(n) // 64] >> (np.arange(n) % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
    for r in range(nranks):
        first, cnt = scenes.shard_range(n, r, nranks)
        part = scenes.Scene(sc.name, n, first, cnt)
        part.transforms, part.sprites = sc.transforms[first:first + cnt], sc.sprites[first:first + cnt]
        local = np.zeros((cnt + 63) // 64, np.uint64)
        idx = np.nonzero(bits[first:first + cnt])[0]
        np.bitwise_or.at(local, idx // 64, np.uint64(1) << (idx % 64).astype(np.uint64))
        ctx = util.make_ctx(part, {capi.TB_SPRITE: local})
        ctx.shard_set(r, nranks, first)
        ctx.set_output(bufs[0], bufs[1], bufs[2], n)
        ctxs.append(ctx)
    g = _run_sharded(ctxs, bufs)
    util.assert_frame_parity(g, ref.frame())
    for c in ctxs:
        c.close()
    for b in bufs:
        capi.device_free(b)


def test_sharded_wide_keys_are_refused_not_wrong():
    sc = scenes.config2(n=20_000, seed=43, z_mode="random")
    ctxs, bufs = _sharded_frame(sc, 2)
    from tuber_legacy_b200.shard import reduce_stats
    g = reduce_stats([c.shard_key_stats() for c in ctxs])
    with pytest.raises(capi.TuberError) as e:
        ctxs[0].shard_plan(g)
    assert e.value.status == capi.TB_ERR_UNSUPPORTED
    for c in ctxs:
        c.close()
    for b in bufs:
        capi.device_free(b)


# ---- few-bucket frames: the warp-autonomous emit -------------------------------------------------------

@pytest.mark.parametrize("textures,layers,zvals", [(1, 1, 1), (2, 1, 1), (4, 2, 1), (2, 1, 2), (8, 1, 1), (3, 1, 2)])
def test_few_buckets_all_key_shapes(textures, layers, zvals):
    """<= 8 buckets made of any mix of layer / texture / z bits, ragged size, sparse sprites"""
    sc = scenes.config4(n=77_777, seed=50 + textures)
    i = np.arange(sc.n)
    sc.sprites["texture"] = i % textures
    sc.sprites["layer"] = (i // 3) % layers
    sc.transforms["translation"][:, 2] = (2 + (i // 5) % zvals).astype(np.float32)  # 2.0 / 3.0: one bit apart
    masks = {capi.TB_SPRITE: util.random_mask(sc.n, 0.85, 21)}
    for opts in ({}, {"small_emit": 0}, {"interleaved_rows": 1}):
        ctx, ref, g, r = run_both(sc, masks, ticks=1, **opts)
        assert g.info.path == capi.TB_PATH_BUCKET
        util.assert_frame_parity(g, r)
        ctx.close()


def test_few_buckets_hierarchy():
    sc = scenes.config3(n=50_000, seed=52)
    sc.sprites["texture"] = np.arange(sc.n) % 3
    sc.sprites["layer"] = 0
    sc.transforms["translation"][:, 2] = 0
    ctx, ref, g, r = run_both(sc)
    assert g.info.path == capi.TB_PATH_BUCKET and g.info.hierarchy_levels == 8
    util.assert_frame_parity(g, r)
    ctx.close()


# ---- single-kernel steady-state frames (cached per-tile bucket offsets) -------------------------------

def test_static_bucket_frames_are_one_kernel_and_stay_exact():
    sc = scenes.config4(n=150_001, seed=60)
    masks = {capi.TB_VELOCITY: util.random_mask(sc.n, 0.7, 31), capi.TB_SPRITE: util.random_mask(sc.n, 0.9, 32)}
    ctx = util.make_ctx(sc, masks)
    ref = util.make_oracle(sc, masks)
    launches = []
    for _ in range(4):
        ctx.system_integrate(sc.dt)
        ref.step(sc.dt)
        g = util.GpuFrame(ctx)
        util.assert_frame_parity(g, ref.frame())
        launches.append(g.info.kernel_launches)
    assert launches[0] > 1 and launches[1:] == [1, 1, 1], launches   # full frame once, then one kernel per frame
    t = ctx.download_transforms(0, sc.n)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), ref.transforms().view(np.uint32))
    # any component upload drops the cached offsets: textures change -> buckets change -> still exact
    sc.sprites["texture"][::3] = 3 - sc.sprites["texture"][::3]
    ctx.upload_sprites(0, sc.sprites)
    ctx.upload_presence(capi.TB_SPRITE, 0, masks[capi.TB_SPRITE])
    ref2 = oracle.Scene(sc.n, t, sc.sprites, sc.velocities, mask_sprite=masks[capi.TB_SPRITE], mask_velocity=masks[capi.TB_VELOCITY])
    g = util.GpuFrame(ctx)
    assert g.info.kernel_launches > 1
    util.assert_frame_parity(g, ref2.frame())
    ctx.system_integrate(sc.dt)
    ref2.step(sc.dt)
    g = util.GpuFrame(ctx)
    assert g.info.kernel_launches == 1
    util.assert_frame_parity(g, ref2.frame())
    # the option switches it off without changing results
    ctx.set_option("static_buckets", 0)
    ctx.system_integrate(sc.dt)
    ref2.step(sc.dt)
    g = util.GpuFrame(ctx)
    assert g.info.kernel_launches > 1
    util.assert_frame_parity(g, ref2.frame())
    ctx.close()


def test_static_bucket_frame_detects_a_z_change():
    """the cached plan says z is constant; a tick moves z of some entities: the single-kernel frame's own
    statistics expose it, the frame is redone the long way, and the tick is still applied exactly once"""
    sc = scenes.config4(n=60_000, seed=61)
    ctx = util.make_ctx(sc)
    ref = util.make_oracle(sc)
    for _ in range(2):
        ctx.system_integrate(sc.dt)
        ref.step(sc.dt)
        g = util.GpuFrame(ctx)
    assert g.info.kernel_launches == 1
    v = sc.velocities.copy()
    v["v"][::11, 2] = 50.0
    ctx.upload_velocities(0, v)   # velocities do not touch placement: the cached offsets survive this upload
    ref3 = oracle.Scene(sc.n, ctx.download_transforms(0, sc.n), sc.sprites, v)
    ctx.system_integrate(sc.dt)
    ref3.step(sc.dt)
    g = util.GpuFrame(ctx)
    assert g.info.replanned == 1 and g.info.kernel_launches > 1
    util.assert_frame_parity(g, ref3.frame())
    t = ctx.download_transforms(0, sc.n)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), ref3.transforms().view(np.uint32))
    ctx.close()


def test_sharded_scene_graph_with_ghost_ancestors():
    """Parent links across shards: every shard also holds its entities' out-of-shard ancestors (no Sprite)
    and recomputes their chain; the merged result equals the single-process oracle frame."""
    from tuber_legacy_b200.shard import ancestor_closure, reduce_stats
    for random_parents in (False, True):
        sc = scenes.config3(n=60_000, seed=70, random_parents=random_parents)
        v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
        v["v"][:, 0] = np.linspace(-30, 30, sc.n, dtype=np.float32)
        sc.velocities = v
        ref = util.make_oracle(sc)
        n, nranks = sc.n, 3
        bufs = [capi.device_alloc(n * 80), capi.device_alloc(n * 64), capi.device_alloc(n * 4)]
        ctxs = []
        for r in range(nranks):
            first, cnt = scenes.shard_range(n, r, nranks)
            ghosts, local_parents = ancestor_closure(sc.parents, first, cnt)
            ids = np.concatenate([np.arange(first, first + cnt), ghosts.astype(np.int64)])
            m = len(ids)
            ctx = capi.Context(m)
            ctx.shard_set(r, nranks, first)
            ctx.set_entity_count(m)
            ctx.upload_transforms(0, sc.transforms[ids])
            ctx.upload_velocities(0, sc.velocities[ids])
            ctx.upload_sprites(0, sc.sprites[first:first + cnt])     # own entities only: ghosts are never drawn
            ctx.upload_parents(0, local_parents)
            ctx.set_output(bufs[0], bufs[1], bufs[2], n)
            ctxs.append(ctx)
            if r > 0 and not random_parents:
                assert 0 < len(ghosts) < cnt // 2   # BFS-ordered tree: a slice of ancestors, not the whole scene
        ref.step(0.01)
        g = _run_sharded(ctxs, bufs, ticks=1, dt=0.01)
        assert g.info.hierarchy_levels == 8
        util.assert_frame_parity(g, ref.frame())
        for c in ctxs:
            c.close()
        for b in bufs:
            capi.device_free(b)


# ---- misuse of the C ABI is an error status, never a crash -------------------------------------------

def test_c_abi_misuse_returns_statuses():
    import ctypes as C
    lib = capi.load()
    out = capi.FrameOut()
    assert lib.tb_frame_transform_batch(None, C.byref(out)) == capi.TB_ERR_INVALID
    assert lib.tb_set_entity_count(None, 3) == capi.TB_ERR_INVALID
    h = C.c_void_p()
    assert lib.tb_ctx_create(0, 0, C.byref(h)) == capi.TB_ERR_INVALID            # zero capacity
    assert lib.tb_ctx_create(0, 1 << 31, C.byref(h)) == capi.TB_ERR_INVALID      # beyond 2^30 entity indices
    assert lib.tb_ctx_create(999, 16, C.byref(h)) == capi.TB_ERR_NO_DEVICE       # no such GPU
    with capi.Context(64) as ctx:
        sc = scenes.config1(n=64)
        ctx.set_entity_count(64)
        ctx.upload_transforms(0, sc.transforms)
        ctx.upload_sprites(0, sc.sprites)
        # downloads before any frame, and beyond the last frame's output
        with pytest.raises(capi.TuberError) as e:
            ctx.download_instances(1)
        assert e.value.status == capi.TB_ERR_INVALID
        assert ctx.frame().num_instances == 64
        with pytest.raises(capi.TuberError):
            ctx.download_instances(65)
        # the multi-GPU calls must follow their protocol order
        with pytest.raises(capi.TuberError) as e:
            ctx.shard_plan([0, 0, 0])
        assert e.value.status == capi.TB_ERR_INVALID
        with pytest.raises(capi.TuberError):
            ctx.shard_key_histogram(0)
        with pytest.raises(capi.TuberError):
            ctx.set_option("no_such_option", 1)
        with pytest.raises(capi.TuberError):
            ctx.upload_presence(7, 0, np.zeros(1, np.uint64))
        # output buffers smaller than the frame
        buf = capi.device_alloc(1024)
        ctx.set_output(buf, buf, buf, 8)
        with pytest.raises(capi.TuberError) as e:
            ctx.frame()
        assert e.value.status == capi.TB_ERR_CAPACITY
        ctx.set_output(0, 0, 0, 0)
        assert ctx.frame().num_instances == 64      # and the context is still usable
        capi.device_free(buf)
        # more batches than the table holds
        sc.sprites["texture"] = np.arange(64)
        ctx.upload_sprites(0, sc.sprites)
        ctx.set_max_batches(10)
        with pytest.raises(capi.TuberError) as e:
            ctx.frame()
        assert e.value.status == capi.TB_ERR_CAPACITY


def test_sharded_instances_only_then_vertices_regenerated():
    """only instances cross to the presenter; its vertices_from_instances gives the very same quads"""
    sc = scenes.config2(n=70_003, seed=45)
    ctxs, bufs = _sharded_frame(sc, 4, instances_only=True)
    g = _run_sharded(ctxs, bufs, regenerate_vertices=True)
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    # bit-identical to the vertices a single context emits itself
    one = util.make_ctx(sc)
    ref = util.GpuFrame(one)
    assert g.vertices.tobytes() == ref.vertices.tobytes() and g.instances.tobytes() == ref.instances.tobytes()
    one.close()
    for c in ctxs:
        c.close()
    for b in bufs:
        capi.device_free(b)


# ---- z dictionary (few distinct depths sort by rank, not by their varying float bits) ------------------

def test_z_dictionary_saves_a_sort_pass():
    """64 textures x 16 layers x depths 0.0..15.0: 10 + 11 varying key bits (3 passes) become 10 + 4 (2)"""
    sc = scenes.config2(n=150_000, seed=61)
    r = util.make_oracle(sc).frame()
    ctx = util.make_ctx(sc)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, r)
    assert g.info.key_bits == 14 and g.info.sort_passes == 2
    g2 = util.GpuFrame(ctx)  # cached plan, dictionary still valid
    assert g2.info.replanned == 0
    util.assert_frame_parity(g2, r)
    ctx.close()
    ctx = util.make_ctx(sc, z_dictionary=0)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, r)
    assert g.info.key_bits > 14 and g.info.sort_passes == 3
    ctx.close()


def test_z_dictionary_follows_uploads():
    """uploads drop the cached plan: the next frame collects the depths again"""
    sc = scenes.config2(n=60_000, seed=62)
    ctx = util.make_ctx(sc)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    # 2.5 lies inside the bit envelope of {0..15}: a 17th depth, 5 rank bits
    sc.transforms["translation"][1000:1200, 2] = 2.5
    ctx.upload_transforms(1000, sc.transforms[1000:1200])
    g = util.GpuFrame(ctx)
    assert g.info.key_bits == 10 + 5
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    # a new layer bit: the (layer, texture) half is validated by the key statistics
    sc.sprites["layer"][5:50] = 40
    ctx.upload_sprites(5, sc.sprites[5:50])
    g = util.GpuFrame(ctx)
    assert g.info.key_bits == 11 + 5
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    # more than 64 depths: no dictionary any more
    sc.transforms["translation"][:, 2] = (np.arange(sc.n) % 100).astype(np.float32) * np.float32(0.25)
    ctx.upload_transforms(0, sc.transforms)
    g = util.GpuFrame(ctx)
    assert g.info.key_bits > 16
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    ctx.close()


def test_z_dictionary_goes_stale_and_replans():
    """the velocity system moves every depth without an upload: the cached dictionary misses (the
    varying-bit envelope alone would not notice 1.0 -> 3.0) and the frame re-plans itself"""
    sc = scenes.config4(n=40_000, seed=63)
    sc.sprites["texture"] = np.arange(sc.n) % 64
    sc.sprites["layer"] = np.arange(sc.n) % 5
    sc.transforms["translation"][:, 2] = (np.arange(sc.n) % 3).astype(np.float32)
    sc.velocities["v"][:, 2] = 16.0  # every depth moves each tick: the dictionary is rebuilt each frame
    ctx = util.make_ctx(sc)
    ref = util.make_oracle(sc)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    for _ in range(3):
        ctx.system_integrate(0.0625)
        ref.step(0.0625)
        g = util.GpuFrame(ctx)
        assert g.info.replanned == 1 and g.info.key_bits == 6 + 3 + 2
        util.assert_frame_parity(g, ref.frame())
    t = ctx.download_transforms(0, sc.n)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), ref.transforms().view(np.uint32))
    ctx.close()


@pytest.mark.parametrize("opts", [{}, {"small_emit": 0}, {"interleaved_rows": 1, "staged_emit": 0}])
def test_z_dictionary_single_pass_and_sparse(opts):
    """2 textures x depths {0, 1, 5}: 1 + 2 rank bits, one pass, the few-bucket emit with a dictionary"""
    sc = scenes.config4(n=70_003, seed=64)
    i = np.arange(sc.n)
    sc.sprites["texture"] = i % 2
    sc.sprites["layer"] = 0
    sc.transforms["translation"][:, 2] = np.array([0.0, 1.0, 5.0], np.float32)[(i // 7) % 3]
    masks = {capi.TB_SPRITE: util.random_mask(sc.n, 0.8, 31), capi.TB_TRANSFORM: util.random_mask(sc.n, 0.9, 32)}
    ctx, ref, g, r = run_both(sc, masks, ticks=1, **opts)
    assert g.info.key_bits == 3 and g.info.path == capi.TB_PATH_BUCKET
    util.assert_frame_parity(g, r)
    g = util.GpuFrame(ctx)
    assert g.info.replanned == 0
    util.assert_frame_parity(g, r)
    ctx.close()


def test_z_dictionary_hierarchy():
    """every node 1.0 deep below its parent: world depths 1..8, three rank bits"""
    sc = scenes.config3(n=60_000, seed=65)
    sc.transforms["translation"][:, 2] = 1.0
    ctx, ref, g, r = run_both(sc)
    assert g.info.hierarchy_levels == 8 and g.info.key_bits == 10 + 3
    util.assert_frame_parity(g, r)
    ctx.close()


# ---- row layout ------------------------------------------------------------------------------------------

def test_row_layout_switches_keep_the_rows():
    """sorted frames move the storage to 64-byte rows on their own; an explicit switch converts in place"""
    sc = scenes.config2(n=70_001, seed=71, z_mode="random")
    ctx = util.make_ctx(sc)
    ctx.profile_enable(True)
    r = util.make_oracle(sc).frame()
    util.assert_frame_parity(util.GpuFrame(ctx), r)
    util.assert_frame_parity(util.GpuFrame(ctx), r)
    assert ctx.profile()["relayout_rows"][2] == 1  # converted once, at the first sorted frame
    for layout in (0, 1, 0, -1):
        ctx.set_option("interleaved_rows", layout)
        util.assert_frame_parity(util.GpuFrame(ctx), r)
        t = ctx.download_transforms(0, sc.n)
        np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), sc.transforms.view(np.uint32).reshape(-1, 12))
    # in-order single-pass frames stay on the sector arrays
    sc = scenes.config4(n=30_000, seed=72)
    ctx2 = util.make_ctx(sc)
    ctx2.profile_enable(True)
    util.assert_frame_parity(util.GpuFrame(ctx2), util.make_oracle(sc).frame())
    assert "relayout_rows" not in ctx2.profile()
    ctx2.close()
    ctx.close()


# ---- steady-state frames replay a CUDA graph ---------------------------------------------------------------

@pytest.mark.parametrize("kind", ["flat", "hierarchy"])
def test_graph_replayed_frames_stay_exact(kind):
    """frame 1 plans, frame 2 runs kernel by kernel, frame 3 captures, frames 4.. replay: all equal the
    oracle, through velocity ticks, and an upload in between is picked up"""
    if kind == "flat":
        sc = scenes.config4(n=60_000, seed=81)
        sc.sprites["texture"] = np.arange(sc.n) % 64
        sc.sprites["layer"] = np.arange(sc.n) % 7
    else:
        sc = scenes.config3(n=40_000, seed=82)
        sc.velocities = scenes.config4(n=sc.n, seed=83).velocities  # every node moves in x / y: propagation every frame
    has_vel = sc.velocities is not None
    ctx = util.make_ctx(sc)
    ref = util.make_oracle(sc)
    launches = []
    for k in range(7):
        if has_vel:
            ctx.system_integrate(0.01)
            ref.step(0.01)
        g = util.GpuFrame(ctx)
        util.assert_frame_parity(g, ref.frame())
        launches.append(g.info.kernel_launches)
        assert g.info.sort_passes >= 2
    # after two frames that kept their sort the rows (scene graph: the drawn entities' local rows) move into draw order
    # (one extra copy launch) and every later frame streams: flat one kernel, scene graph inner level passes + one emit
    assert len(set(launches[3:])) == 1, launches
    # same frames without graphs
    ctx2 = util.make_ctx(sc, cuda_graphs=0)
    for k in range(7):
        if has_vel:
            ctx2.system_integrate(0.01)
    g2 = util.GpuFrame(ctx2)
    np.testing.assert_array_equal(g2.order, g.order)
    np.testing.assert_array_equal(g2.instances.view(np.uint32), g.instances.view(np.uint32))
    np.testing.assert_array_equal(g2.vertices.view(np.uint32), g.vertices.view(np.uint32))
    ctx2.close()
    # an upload changes data (and here the plan): the graph must not serve a stale frame
    sc.sprites["texture"][: sc.n // 2] = 100
    sc.sprites["size"][:, 0] += 1.0
    ctx.upload_sprites(0, sc.sprites)
    ref.close()
    if has_vel:
        sc.transforms = ctx.download_transforms(0, sc.n)
    ref = util.make_oracle(sc)
    for k in range(4):
        if has_vel:
            ctx.system_integrate(0.02)  # a different dt is a different signature
            ref.step(0.02)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ctx.close()


def test_several_ticks_before_a_frame_run_as_one_pass():
    """the update loop may tick more than once per render: the ticks fuse into the frame's first pass and
    give the bits of running them one after the other (also across a dt change and a download)"""
    for sc, opts in ((scenes.config4(n=50_000, seed=91), {}), (scenes.config4(n=50_000, seed=92), {"static_buckets": 0}),
                     (scenes.config3(n=30_000, seed=93), {})):
        if sc.velocities is None:
            sc.velocities = scenes.config4(n=sc.n, seed=94).velocities
        ctx = util.make_ctx(sc, **opts)
        ref = util.make_oracle(sc)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
        for ticks, dt in ((3, 0.01), (1, 0.01), (2, 0.02), (5, 0.01)):
            for _ in range(ticks):
                ctx.system_integrate(dt)
                ref.step(dt)
            g = util.GpuFrame(ctx)
            util.assert_frame_parity(g, ref.frame())
        # mixed dt without a frame in between, then a download (flushes the queue)
        for dt in (0.01, 0.01, 0.03, 0.01):
            ctx.system_integrate(dt)
            ref.step(dt)
        t = ctx.download_transforms(0, sc.n)
        np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12), ref.transforms().view(np.uint32))
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
        ctx.close()


# ---- camera: view-projection folded into the emitted geometry ------------------------------------------------

def test_camera_view_projection_folded_into_every_emit_path():
    """instances carry view_proj * world, corners are transformed by it; order and batches are unchanged"""
    vp = capi.orthographic(-1200.0, 1200.0, -900.0, 900.0, -100.0, 100.0)
    np.testing.assert_array_equal(vp.view(np.uint32), oracle.orthographic(-1200.0, 1200.0, -900.0, 900.0, -100.0, 100.0).view(np.uint32))
    general = vp.copy()
    general[3] = [0.001, -0.002, 0.003, 1.5]  # a non-affine last row is carried through as well
    cases = [(scenes.config4(n=40_000, seed=101), {}),                    # single-kernel frames
             (scenes.config4(n=40_000, seed=102), {"small_emit": 0}),     # CTA-tile emit, in order
             (scenes.config2(n=60_000, seed=103), {}),                    # sorted stream, gathered rows, graph replay
             (scenes.config3(n=30_000, seed=104), {})]                    # hierarchy: world rows
    for sc, opts in cases:
        ctx = util.make_ctx(sc, **opts)
        ref = util.make_oracle(sc)
        plain = ref.frame()
        for m in (vp, None, general, general, general, vp):
            ctx.set_view_projection(m)
            if sc.velocities is not None:
                ctx.system_integrate(0.01)
                ref.step(0.01)
            g = util.GpuFrame(ctx)
            r = ref.frame(m)
            util.assert_frame_parity(g, r)
            np.testing.assert_array_equal(r.order, ref.frame().order)
        assert not np.array_equal(plain.instances["model"], r.instances["model"])
        ctx.close()


# ---- sorted frames keep the previous frame's sort when no key changed ------------------------------------------

@pytest.mark.parametrize("kind", ["flat", "flat64", "hierarchy"])
def test_sorted_cache_skips_the_sort_only_when_no_key_changed(kind):
    """x/y velocities move sprites without touching their (layer, texture, z) keys: the sort passes are skipped
    on the device and the frame still matches; a z change, a sprite upload or a presence change re-sorts"""
    if kind == "hierarchy":
        sc = scenes.config3(n=40_000, seed=111)
        sc.velocities = scenes.config4(n=sc.n, seed=112).velocities
    else:
        sc = scenes.config2(n=60_000, seed=113, z_mode="random" if kind == "flat64" else "steps")
        sc.velocities = scenes.config4(n=sc.n, seed=114).velocities
    sc.velocities["v"][:, 2] = 0.0
    masks = {capi.TB_SPRITE: util.random_mask(sc.n, 0.9, 41)}
    ctx = util.make_ctx(sc, masks)
    ref = util.make_oracle(sc, masks)
    ctx.profile_enable(True)  # per-kernel times: skipped sort passes return at once
    for k in range(5):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        g = util.GpuFrame(ctx)
        assert g.info.sort_passes >= 2 and g.info.replanned == 0
        util.assert_frame_parity(g, ref.frame())
    ctx.profile_enable(False)
    # a z velocity on a few entities changes their keys: the same frame re-sorts
    v = sc.velocities.copy()
    v["v"][100:140, 2] = 50.0
    ctx.upload_velocities(0, v)
    ref.close()
    sc.transforms = ctx.download_transforms(0, sc.n)
    sc.velocities = v
    ref = util.make_oracle(sc, masks)
    for k in range(4):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    # presence change and sprite change
    sc.sprites["texture"][::7] = 3
    ctx.upload_sprites(0, sc.sprites)  # marks every uploaded sprite present ...
    masks[capi.TB_SPRITE] = util.random_mask(sc.n, 0.8, 42)
    ctx.upload_presence(capi.TB_SPRITE, 0, masks[capi.TB_SPRITE])  # ... so the mask goes second
    ref.close()
    sc.transforms = ctx.download_transforms(0, sc.n)
    ref = util.make_oracle(sc, masks)
    for k in range(3):
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    # the cache off gives the same bytes
    c2 = util.make_ctx(sc, masks, sorted_cache=0)
    g1, g2 = util.GpuFrame(ctx), util.GpuFrame(c2)
    np.testing.assert_array_equal(g1.order, g2.order)
    np.testing.assert_array_equal(g1.instances.view(np.uint32), g2.instances.view(np.uint32))
    c2.close()
    ctx.close()


def test_kept_frames_survive_output_and_batch_table_changes():
    """caller-side changes between steady frames: a larger batch table (re-allocated), other output buffers, a
    smaller batch cap (error, then recovery)"""
    for sc in (scenes.config4(n=30_000, seed=121), scenes.config2(n=30_000, seed=122)):
        ctx = util.make_ctx(sc)
        ref = util.make_oracle(sc)
        r = ref.frame()
        for _ in range(3):
            util.assert_frame_parity(util.GpuFrame(ctx), r)
        ctx.set_max_batches(200_000)  # the table moves to a bigger buffer
        util.assert_frame_parity(util.GpuFrame(ctx), r)
        util.assert_frame_parity(util.GpuFrame(ctx), r)
        bufs = [capi.device_alloc(sc.n * b) for b in (80, 64, 4)]
        ctx.set_output(bufs[0], bufs[1], bufs[2], sc.n)
        info = ctx.frame()
        assert info.num_instances == len(r.order) and info.num_batches == len(r.batches)
        order = ctx.device_read(bufs[2], len(r.order) * 4, np.uint32)
        np.testing.assert_array_equal(order, r.order)
        ctx.set_output(None, None, None, 0)
        if len(r.batches) > 1:
            ctx.set_max_batches(1)
            with pytest.raises(capi.TuberError) as e:
                ctx.frame()
            assert e.value.status == capi.TB_ERR_CAPACITY
            ctx.set_max_batches(65536)
        util.assert_frame_parity(util.GpuFrame(ctx), r)
        for b in bufs:
            capi.device_free(b)
        ctx.close()


def test_per_frame_uploads_keep_plan_and_sort_when_keys_do_not_change():
    """a host system rewrites some transforms every frame (dirty-range upload): the plan is not rediscovered and
    the sort is kept as long as no key moves; a changed depth re-sorts; a texture outside the plan re-plans"""
    sc = scenes.config2(n=50_000, seed=131)
    ctx = util.make_ctx(sc)
    util.assert_frame_parity(util.GpuFrame(ctx), util.make_oracle(sc).frame())
    rng = np.random.default_rng(7)
    for k in range(4):
        lo = int(rng.integers(0, sc.n - 5000))
        sc.transforms["translation"][lo:lo + 5000, 0] += np.float32(1.5)  # x only: keys unchanged
        ctx.upload_transforms(lo, sc.transforms[lo:lo + 5000])
        g = util.GpuFrame(ctx)
        assert g.info.replanned == 0
        util.assert_frame_parity(g, util.make_oracle(sc).frame())
    sc.transforms["translation"][100:200, 2] = 7.0  # a depth of the dictionary: same plan, new order
    ctx.upload_transforms(100, sc.transforms[100:200])
    g = util.GpuFrame(ctx)
    assert g.info.replanned == 0
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    sc.sprites["texture"][300:310] = 500  # bits the plan does not cover
    ctx.upload_sprites(300, sc.sprites[300:310])
    g = util.GpuFrame(ctx)
    assert g.info.replanned == 1
    util.assert_frame_parity(g, util.make_oracle(sc).frame())
    for _ in range(3):  # and back to steady single-kernel frames
        g = util.GpuFrame(ctx)
        util.assert_frame_parity(g, util.make_oracle(sc).frame())
    assert g.info.kernel_launches == 1
    ctx.close()


# ---- chunked sort (csrc/chunked.cuh): many small chunks at test sizes ----------------------------------------------

@pytest.mark.parametrize("shift", [11, 12, 14])
@pytest.mark.parametrize("kind", ["flat", "flat_sparse", "hierarchy", "hierarchy_random", "textures_only"])
def test_chunked_sort_matches_the_oracle(kind, shift):
    masks = None
    if kind == "flat":
        sc = scenes.config2(n=70_001, seed=71)
    elif kind == "flat_sparse":
        sc = scenes.config2(n=50_000, seed=72)
        masks = {capi.TB_TRANSFORM: util.random_mask(sc.n, 0.7, 1), capi.TB_SPRITE: util.random_mask(sc.n, 0.5, 2)}
    elif kind == "hierarchy":
        sc = scenes.config3(n=60_000, seed=73)
    elif kind == "hierarchy_random":
        sc = scenes.config3(n=60_000, seed=74, random_parents=True)
    else:  # one pass of 6 bits: 64 textures, nothing else varies
        sc = scenes.config2(n=66_000, seed=75)
        sc.sprites["layer"][:] = 3
        sc.transforms["translation"][:, 2] = 2.0
    ctx, ref, g, r = run_both(sc, masks=masks, chunked_sort=1, chunk_shift=shift)
    assert g.info.path in (capi.TB_PATH_SORT, capi.TB_PATH_BUCKET)
    util.assert_frame_parity(g, r, what=f"chunked sort {kind} shift {shift}")
    for k in range(4):  # kept list: validated single-kernel frames (flat) / key-compare frames (scene graph)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"chunked sort {kind} steady {k}")
    ctx.close()


@pytest.mark.parametrize("shift", [11, 13])
def test_chunked_sort_moving_entities_and_key_changes(shift):
    """x / y motion keeps the list (the velocity system runs inside the list emit); a z velocity changes keys: the
    validation catches it and the frame re-sorts."""
    sc = scenes.config2(n=60_000, seed=76)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    rng = np.random.default_rng(9)
    v["v"][:, 0] = rng.uniform(-100, 100, sc.n)
    v["v"][:, 1] = rng.uniform(-100, 100, sc.n)
    sc.velocities = v
    ctx, ref, g, r = run_both(sc, ticks=1, chunked_sort=1, chunk_shift=shift)
    util.assert_frame_parity(g, r)
    for k in range(4):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        gf = util.GpuFrame(ctx)
        util.assert_frame_parity(gf, ref.frame(), what=f"chunked, moving, frame {k}")
    assert gf.info.kernel_launches == 2  # the steady frame: the list emit + the tick of entities that move undrawn
    # now some entities move in z by whole depth steps: keys change
    v["v"][: sc.n // 3, 2] = 100.0
    ctx.upload_velocities(0, v)
    ref.close()
    ref = util.make_oracle(sc)
    ctx.upload_transforms(0, sc.transforms)
    for k in range(3):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"chunked, z velocity, frame {k}")
    ctx.close()


def test_chunked_and_global_sort_emit_the_same_bytes():
    sc = scenes.config2(n=200_000, seed=77)
    a = util.make_ctx(sc, chunked_sort=1, chunk_shift=14)
    b = util.make_ctx(sc, chunked_sort=0)
    ga, gb = util.GpuFrame(a), util.GpuFrame(b)
    for f in ("order", "instances", "vertices", "batches"):
        assert getattr(ga, f).tobytes() == getattr(gb, f).tobytes(), f
    a.close()
    b.close()


# ---- draw-ordered working storage: moving sorted scenes stream their rows --------------------------------------------

@pytest.mark.parametrize("z_mode", ["steps", "random"])
def test_moving_sorted_scene_lives_in_draw_order(z_mode):
    """Every entity moves in x / y: after two validated frames the rows (and velocities) are copied into draw order and
    the velocity system runs on the copy, one streaming kernel per frame. Entity-order rows are synchronised lazily: a
    transform download in between must see every tick; a z velocity (keys change) falls back to a re-sort."""
    sc = scenes.config2(n=70_000, seed=101, z_mode=z_mode)
    rng = np.random.default_rng(11)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-100, 100, sc.n)
    v["v"][:, 1] = rng.uniform(-100, 100, sc.n)
    sc.velocities = v
    # some entities move without being drawn (Velocity, no Sprite): they are not in the draw-ordered copy and are ticked
    # by a side pass on the entity-order rows
    masks = {capi.TB_VELOCITY: util.random_mask(sc.n, 0.7, 7), capi.TB_SPRITE: util.random_mask(sc.n, 0.85, 8)}
    ctx, ref, g, r = run_both(sc, masks=masks, ticks=1)
    util.assert_frame_parity(g, r)
    launches = []
    for k in range(8):
        ticks = 2 if k == 5 else 1  # two engine ticks before one frame
        for _ in range(ticks):
            ctx.system_integrate(0.01)
            ref.step(0.01)
        gf = util.GpuFrame(ctx)
        launches.append(gf.info.kernel_launches)
        util.assert_frame_parity(gf, ref.frame(), what=f"moving sorted scene, frame {k}")
        if k == 6:  # forces the lazy write-back of the translations
            got = ctx.download_transforms(0, sc.n)
            want = ref.transforms().view(capi.TRANSFORM_DTYPE).reshape(-1)
            assert np.array_equal(got["translation"], want["translation"])
    assert launches[-1] == 2 and launches[4] == 2, launches  # the steady frame: the streaming emit + the undrawn movers' tick
    # a camera-only frame (nothing pending) and a tick after it
    vp = capi.orthographic(-1200, 1200, -900, 900, -100, 100)
    ctx.set_view_projection(vp)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(view_proj=vp), what="camera frame on the working storage")
    ctx.set_view_projection(None)
    # some entities start moving in z by whole depth steps: keys change, the validation catches it
    v["v"][: sc.n // 4, 2] = 100.0
    ctx.upload_velocities(0, v)
    ctx.upload_presence(capi.TB_VELOCITY, 0, masks[capi.TB_VELOCITY])
    moved = ctx.download_transforms(0, sc.n)
    want = ref.transforms().view(capi.TRANSFORM_DTYPE).reshape(-1)
    assert np.array_equal(moved["translation"], want["translation"])  # undrawn movers included
    ref.close()
    sc2 = scenes.config2(n=70_000, seed=101, z_mode=z_mode)
    sc2.transforms = moved
    sc2.velocities = v
    ref = util.make_oracle(sc2, masks)
    for k in range(4):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"z velocity, frame {k}")
    ctx.close()


def test_working_storage_survives_downloads_and_is_dropped_by_uploads():
    sc = scenes.config2(n=40_000, seed=102)
    rng = np.random.default_rng(12)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-10, 10, sc.n)
    sc.velocities = v
    ctx, ref, g, r = run_both(sc, ticks=1)
    for k in range(5):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    # world matrices of the moved entities (flushes nothing pending, but needs current rows)
    w = ctx.download_world_matrices(0, sc.n)
    rf = ref.frame()
    if util.libm_pinned():
        assert np.array_equal(w, rf.world)
    # a partial transform upload lands in the entity-order rows: nothing of the other entities' motion may be lost
    t2 = ctx.download_transforms(0, 100)
    t2["translation"][:, 0] += np.float32(5.0)
    ctx.upload_transforms(0, t2)
    sc3 = scenes.config2(n=40_000, seed=102)
    sc3.transforms = ctx.download_transforms(0, sc.n)
    sc3.velocities = v
    ref2 = util.make_oracle(sc3)
    for k in range(4):
        ctx.system_integrate(0.01)
        ref2.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref2.frame(), what=f"after partial upload, frame {k}")
    ctx.close()


# ---- scene graphs: the last level composed inside the kept-order emit -------------------------------------------------

@pytest.mark.parametrize("random_parents", [False, True])
def test_moving_scene_graph_composes_its_leaves_inside_the_emit(random_parents):
    """Every node moves in x / y. Once the draw order is kept, the level pass runs for the inner levels only and the emit
    computes world = parent * local for the leaves; leaves without a Sprite are ticked by a side pass. Entity-order
    transforms, world matrices and frames must stay the oracle's through ticks, a camera-only frame and a z velocity."""
    sc = scenes.config3(n=50_000, seed=111, random_parents=random_parents)
    rng = np.random.default_rng(13)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-30, 30, sc.n)
    v["v"][:, 1] = rng.uniform(-30, 30, sc.n)
    sc.velocities = v
    masks = {capi.TB_SPRITE: util.random_mask(sc.n, 0.8, 3), capi.TB_VELOCITY: util.random_mask(sc.n, 0.9, 4)}
    ctx, ref, g, r = run_both(sc, masks=masks, ticks=1)
    util.assert_frame_parity(g, r)
    launches = []
    for k in range(6):
        ticks = 2 if k == 3 else 1
        for _ in range(ticks):
            ctx.system_integrate(0.01)
            ref.step(0.01)
        gf = util.GpuFrame(ctx)
        launches.append(gf.info.kernel_launches)
        util.assert_frame_parity(gf, ref.frame(), what=f"moving scene graph, frame {k}")
    # inner levels + (side tick) + emit: fewer launches than the full frame, which also runs tile_hist / scans / scatters
    assert launches[-1] <= 10, launches
    got = ctx.download_transforms(0, sc.n)
    want = ref.transforms().view(capi.TRANSFORM_DTYPE).reshape(-1)
    assert np.array_equal(got["translation"], want["translation"])  # non-renderable leaves were ticked too
    rf = ref.frame()
    w = ctx.download_world_matrices(0, sc.n)
    if util.libm_pinned():
        assert np.array_equal(w[rf.has_world.astype(bool)], rf.world[rf.has_world.astype(bool)])
    # camera-only frames after the download (world rows valid again), then ticks again
    vp = capi.orthographic(-500, 500, -400, 400, -100, 100)
    ctx.set_view_projection(vp)
    for k in range(3):
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(view_proj=vp), what=f"scene graph, camera frame {k}")
    ctx.set_view_projection(None)
    # a z velocity on some leaves and inner nodes: keys change, the fused frame detects it and the frame re-sorts
    v["v"][::5, 2] = 100.0
    ctx.upload_velocities(0, v)
    ctx.upload_presence(capi.TB_VELOCITY, 0, masks[capi.TB_VELOCITY])
    sc2 = scenes.config3(n=50_000, seed=111, random_parents=random_parents)
    sc2.transforms = ctx.download_transforms(0, sc.n)
    sc2.velocities = v
    ref.close()
    ref = util.make_oracle(sc2, masks)
    for k in range(4):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"scene graph, z velocity, frame {k}")
    ctx.close()


def test_keys_changing_every_frame_back_off_from_single_kernel_frames():
    """z velocities change sort keys on every tick: after one failed single-kernel attempt the frames go straight to the
    full path (no failed attempt on top of every frame); when the keys settle the single-kernel frames come back."""
    sc = scenes.config2(n=50_000, seed=121)
    rng = np.random.default_rng(21)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-50, 50, sc.n)
    sc.velocities = v
    ctx, ref, g, r = run_both(sc, ticks=1)
    for k in range(4):  # settle into the single-kernel frames
        ctx.system_integrate(0.01)
        ref.step(0.01)
        g = util.GpuFrame(ctx)
    assert g.info.kernel_launches == 1
    v2 = v.copy()
    v2["v"][::3, 2] = 100.0  # whole depth steps per tick: keys change every frame
    ctx.upload_velocities(0, v2)
    sc2 = scenes.config2(n=50_000, seed=121)
    sc2.transforms = ctx.download_transforms(0, sc.n)
    sc2.velocities = v2
    ref.close()
    ref = util.make_oracle(sc2)
    ctx.profile_enable(True)
    for k in range(8):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"keys change every frame, frame {k}")
    prof = ctx.profile()
    ctx.profile_enable(False)
    # a failed single-kernel attempt shows as a standalone velocity pass (+ validated emit) or a validated streaming emit;
    # full frames fuse the tick into their key pass
    wasted = prof.get("emit_sequential", (0, 0, 0))[2] + prof.get("integrate", (0, 0, 0))[2]
    assert wasted <= 3, prof  # not one failed attempt per frame
    # keys settle: z velocities off again
    ctx.upload_velocities(0, v)
    sc3 = scenes.config2(n=50_000, seed=121)
    sc3.transforms = ctx.download_transforms(0, sc.n)
    sc3.velocities = v
    ref.close()
    ref = util.make_oracle(sc3)
    launches = []
    for k in range(8):
        ctx.system_integrate(0.01)
        ref.step(0.01)
        gf = util.GpuFrame(ctx)
        launches.append(gf.info.kernel_launches)
        util.assert_frame_parity(gf, ref.frame(), what=f"keys settled, frame {k}")
    assert launches[-1] == 1, launches
    ctx.close()
