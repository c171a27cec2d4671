"""Device-side structural changes (Ecs::insert / delete_by_ids / delete_by_query / add_component /
remove_component, crates/tuber-ecs/src/ecs.rs:93-124) and the second device system with the access-set scheduler
(§8f.2: SystemBundle::step, system.rs:17-23), against the oracle's restated Ecs driven through the same operations."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from tuber_legacy_b200 import capi, scenes

import util

pytestmark = pytest.mark.gpu

TY = {capi.TB_TRANSFORM: 0x7472616e73666f72, capi.TB_SPRITE: 0x7370726974650000, capi.TB_VELOCITY: 0x76656c6f63697479,
      capi.TB_PARENT: 0x706172656e740000}


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def ref_delete_by_ids(ref, ids):
    a = np.ascontiguousarray(ids, dtype=np.uint64)
    oracle.lib().orc_ecs_delete_by_ids(ref._h, _ptr(a), len(a))


def ref_delete_by_query(ref, comps):
    types = np.array([TY[c] for c in comps], np.uint64)
    req = np.ones(len(comps), np.uint8)
    exc = np.zeros(len(comps), np.uint8)
    oracle.lib().orc_ecs_delete_by_query(ref._h, len(comps), _ptr(types), _ptr(req), _ptr(exc))


def ref_add(ref, comp, index, value):
    v = np.ascontiguousarray(value)
    oracle.lib().orc_ecs_add_component(ref._h, C.c_uint64(TY[comp]), _ptr(v), C.c_uint64(v.nbytes), C.c_uint64(index))


def ref_remove(ref, comp, index):
    oracle.lib().orc_ecs_remove_component(ref._h, C.c_uint64(TY[comp]), C.c_uint64(index))


def test_delete_by_ids_and_generations():
    sc = scenes.config2(n=30_000, seed=91)
    ctx, ref = util.make_ctx(sc), util.make_oracle(sc)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    gens0 = ctx.entity_generations(0, sc.n)
    assert not gens0.any()
    rng = np.random.default_rng(1)
    ids = rng.choice(sc.n, 5000, replace=False).astype(np.uint32)
    ctx.delete_by_ids(ids)
    ref_delete_by_ids(ref, ids)
    g = util.GpuFrame(ctx)
    assert g.info.num_instances == sc.n - 5000
    util.assert_frame_parity(g, ref.frame(), what="after delete_by_ids")
    gens = ctx.entity_generations(0, sc.n)
    assert gens[ids].min() == 1 and gens.sum() == 5000  # exactly the deleted slots were bumped
    # handles issued before the delete no longer resolve; the others do
    probe = np.arange(sc.n, dtype=np.uint32)
    alive = ctx.entities_alive(probe, gens0)
    assert alive.sum() == sc.n - 5000 and not alive[ids].any()
    assert ctx.entities_alive(ids, gens[ids]).all()  # re-issued handles (new generation) resolve again
    # an index beyond entity_count: the reference panics, the library reports it
    with pytest.raises(capi.TuberError):
        ctx.delete_by_ids([sc.n + 3])
    assert ctx.query([capi.TB_TRANSFORM, capi.TB_SPRITE], want_indices=False) == sc.n - 5000
    ctx.close()


def test_delete_by_query_add_and_remove_component():
    sc = scenes.config4(n=20_000, seed=92)
    masks = {capi.TB_VELOCITY: util.random_mask(sc.n, 0.3, 5)}
    ctx, ref = util.make_ctx(sc, masks=masks), util.make_oracle(sc, masks=masks)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    # delete everything that has (Transform, Velocity)
    deleted = ctx.delete_by_query((1 << capi.TB_TRANSFORM) | (1 << capi.TB_VELOCITY))
    ref_delete_by_query(ref, [capi.TB_TRANSFORM, capi.TB_VELOCITY])
    have_v = np.unpackbits(masks[capi.TB_VELOCITY].view(np.uint8), bitorder="little")[:sc.n].astype(bool)
    assert deleted == have_v.sum()
    g = util.GpuFrame(ctx)
    assert g.info.num_instances == sc.n - deleted
    util.assert_frame_parity(g, ref.frame(), what="after delete_by_query")
    # remove the Sprite of a few survivors, give one deleted entity its Transform + Sprite back
    survivors = np.nonzero(~have_v)[0][:50]
    for e in survivors[:10]:
        ctx.remove_component(capi.TB_SPRITE, int(e))
        ref_remove(ref, capi.TB_SPRITE, int(e))
    back = int(np.nonzero(have_v)[0][0])
    ctx.add_component(capi.TB_TRANSFORM, back, sc.transforms[back])
    ref_add(ref, capi.TB_TRANSFORM, back, sc.transforms[back:back + 1])
    ctx.add_component(capi.TB_SPRITE, back, sc.sprites[back])
    ref_add(ref, capi.TB_SPRITE, back, sc.sprites[back:back + 1])
    g = util.GpuFrame(ctx)
    assert g.info.num_instances == sc.n - deleted - 10 + 1
    util.assert_frame_parity(g, ref.frame(), what="after remove / add component")
    # a component type without a store: add_component is a no-op, as in the reference (ecs.rs:114-118)
    ctx.add_component(capi.TB_PARENT, 3, np.uint32(1))
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ctx.close()


def test_insert_appends_entities_one_by_one():
    sc = scenes.config2(n=300, seed=93)
    ctx = capi.Context(1000)
    for e in range(sc.n):
        idx = ctx.insert(transform=sc.transforms[e], sprite=sc.sprites[e] if e % 3 else None)
        assert idx == e  # ecs.rs:94-97: monotonically increasing, never reused
    masks = {capi.TB_SPRITE: np.zeros((sc.n + 63) // 64, np.uint64)}
    for e in range(sc.n):
        if e % 3:
            masks[capi.TB_SPRITE][e // 64] |= np.uint64(1) << np.uint64(e % 64)
    ref = util.make_oracle(sc, masks=masks)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what="after 300 inserts")
    ctx.close()


def test_deletes_in_a_scene_graph_reroot_the_children():
    sc = scenes.config3(n=20_000, seed=94)
    ctx, ref = util.make_ctx(sc), util.make_oracle(sc)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    ids = np.arange(5, 400, 7, dtype=np.uint32)  # inner nodes of the upper levels
    ctx.delete_by_ids(ids)
    ref_delete_by_ids(ref, ids)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what="scene graph after deleting inner nodes")
    ctx.close()


# ---- systems ---------------------------------------------------------------------------------------------------------

def _spin_scene(n, seed, three_d=False):
    sc = scenes.config2(n=n, seed=seed)
    rng = np.random.default_rng(seed)
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-50, 50, sc.n)
    v["v"][:, 1] = rng.uniform(-50, 50, sc.n)
    sc.velocities = v
    spins = rng.uniform(-3, 3, sc.n).astype(np.float32)
    if three_d:
        sc.transforms["angle"][:, 0] = rng.uniform(-1, 1, sc.n)
        sc.transforms["rotation_center"][:, 2] = rng.uniform(-5, 5, sc.n)
    return sc, spins


def test_access_sets_and_stage_assignment():
    rv, wv = capi.system_access(capi.TB_SYSTEM_VELOCITY)
    rs, ws = capi.system_access(capi.TB_SYSTEM_SPIN)
    assert wv & (rs | ws) == 0 and ws & (rv | wv) == 0  # disjoint: the two systems may overlap
    sc, spins = _spin_scene(1000, 95)
    ctx = util.make_ctx(sc)
    ctx.upload_spins(0, spins)
    V, S = capi.TB_SYSTEM_VELOCITY, capi.TB_SYSTEM_SPIN
    assert list(ctx.bundle_step([V, S], 0.01)) == [0, 0]
    assert list(ctx.bundle_step([V, S, V, S], 0.01)) == [0, 0, 1, 1]
    assert list(ctx.bundle_step([S, S, V], 0.01)) == [0, 1, 0]
    ctx.close()


@pytest.mark.parametrize("three_d", [False, True])
@pytest.mark.parametrize("order", [(0, 1), (1, 0), (0, 1, 0), (1,)])
def test_bundle_of_two_systems_matches_sequential_execution(order, three_d):
    sc, spins = _spin_scene(30_000, 96, three_d)
    masks = {capi.TB_VELOCITY: util.random_mask(sc.n, 0.8, 2)}
    mask_spin = util.random_mask(sc.n, 0.6, 3)
    ctx = util.make_ctx(sc, masks=masks)
    ctx.upload_spins(0, spins)
    ctx.upload_presence(capi.TB_SPIN, 0, mask_spin)
    ref = oracle.Scene(sc.n, sc.transforms, sc.sprites, sc.velocities, mask_velocity=masks[capi.TB_VELOCITY], spins=spins,
                       mask_spin=mask_spin)
    util.assert_frame_parity(util.GpuFrame(ctx), ref.frame())
    for k in range(4):
        ctx.bundle_step(list(order), 0.01)
        ref.step_systems(0.01, list(order))
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"bundle {order}, tick {k}")
    got = ctx.download_transforms(0, sc.n)
    want = ref.transforms().view(capi.TRANSFORM_DTYPE).reshape(-1)
    assert np.array_equal(got["angle"], want["angle"]) and np.array_equal(got["translation"], want["translation"])
    ctx.close()


def test_spin_system_alone_and_in_a_scene_graph():
    sc = scenes.config3(n=20_000, seed=97)
    rng = np.random.default_rng(4)
    spins = rng.uniform(-2, 2, sc.n).astype(np.float32)
    ctx = util.make_ctx(sc)
    ctx.upload_spins(0, spins)
    ref = oracle.Scene(sc.n, sc.transforms, sc.sprites, None, sc.parents, spins=spins)
    for k in range(3):
        ctx.system_spin(0.02)
        ref.step_systems(0.02, [1])
        util.assert_frame_parity(util.GpuFrame(ctx), ref.frame(), what=f"scene graph, spin tick {k}")  # parents rotate: world z moves
    ctx.close()
