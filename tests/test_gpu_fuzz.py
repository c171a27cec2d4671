"""Randomised operation sequences against the oracle: the frame keeps plans, sorts, draw orders, draw-ordered working
storage and world rows between frames and brings entity-order rows up to date lazily — so every order of ticks, frames,
partial uploads, downloads, spins, structural changes, camera and option changes must still give the oracle's frame."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from tuber_legacy_b200 import capi, scenes

import util

pytestmark = pytest.mark.gpu

TY_T, TY_S = 0x7472616e73666f72, 0x7370726974650000


def build(kind, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(3_000, 9_000))
    if kind == "flat":
        sc = scenes.config2(n=n, seed=seed)
    elif kind == "particles":
        sc = scenes.config4(n=n, seed=seed)
    else:
        sc = scenes.config3(n=n, seed=seed, random_parents=bool(seed % 2))
    v = np.zeros(sc.n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-50, 50, sc.n)
    v["v"][:, 1] = rng.uniform(-50, 50, sc.n)
    sc.velocities = v
    spins = rng.uniform(-2, 2, sc.n).astype(np.float32)
    masks = {capi.TB_VELOCITY: util.random_mask(sc.n, 0.8, seed + 1), capi.TB_SPRITE: util.random_mask(sc.n, 0.9, seed + 2)}
    mask_spin = util.random_mask(sc.n, 0.5, seed + 3)
    return sc, spins, masks, mask_spin


@pytest.mark.parametrize("kind", ["flat", "particles", "hierarchy"])
@pytest.mark.parametrize("seed", list(range(1, 17)))
def test_random_operation_sequences(kind, seed):
    sc, spins, masks, mask_spin = build(kind, 1000 * seed + {"flat": 0, "particles": 1, "hierarchy": 2}[kind])
    options = [{}, {"chunked_sort": 1, "chunk_shift": 11}, {"cuda_graphs": 0}, {"hier_fused": 0, "sorted_cache": 1}][seed % 4]
    ctx = util.make_ctx(sc, masks=masks, **options)
    ctx.upload_spins(0, spins)
    ctx.upload_presence(capi.TB_SPIN, 0, mask_spin)
    ref = oracle.Scene(sc.n, sc.transforms, sc.sprites, sc.velocities, sc.parents, mask_sprite=masks[capi.TB_SPRITE],
                       mask_velocity=masks[capi.TB_VELOCITY], spins=spins, mask_spin=mask_spin)
    rng = np.random.default_rng(seed * 77 + 5)
    vp = capi.orthographic(-800, 800, -600, 600, -100, 100)
    cam = None
    log = []
    try:
        ops = ["frame", "tick", "tick2", "spin", "bundle", "download", "upload", "camera", "delete", "world", "option", "zvel"]
        # mostly frames and ticks, so that the kept-order / working-storage states are reached before something disturbs them
        weights = np.array([8, 6, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0.5])
        for step in range(120):
            op = rng.choice(ops, p=weights / weights.sum())
            log.append(str(op))
            if op == "frame":
                g = util.GpuFrame(ctx)
                util.assert_frame_parity(g, ref.frame(view_proj=cam), what=f"{kind} seed {seed} step {step} after {log[-12:]}")
            elif op == "tick":
                ctx.system_integrate(0.01)
                ref.step(0.01)
            elif op == "tick2":  # another dt: flushes the queued tick
                ctx.system_integrate(0.02)
                ref.step(0.02)
            elif op == "spin":
                ctx.system_spin(0.01)
                ref.step_systems(0.01, [1])
            elif op == "bundle":
                order = [int(x) for x in rng.permutation(2)]
                ctx.bundle_step(order, 0.01)
                ref.step_systems(0.01, order)
            elif op == "download":
                got = ctx.download_transforms(0, sc.n)
                want = ref.transforms().view(capi.TRANSFORM_DTYPE).reshape(-1)
                alive = ctx.query([capi.TB_TRANSFORM])  # deleted entities: the oracle reports zeros, the device keeps the old row
                assert np.array_equal(got["translation"][alive], want["translation"][alive]), (step, log[-6:])
                assert np.array_equal(got["angle"][alive], want["angle"][alive]), (step, log[-6:])
            elif op == "upload":  # a host system rewrites a range of Transforms
                first = int(rng.integers(0, sc.n - 50))
                cnt = int(rng.integers(1, 50))
                t = ctx.download_transforms(first, cnt)
                t["translation"][:, 0] += np.float32(1.5)
                t["scale"][:, 1] *= np.float32(1.25)
                ctx.upload_transforms(first, t)
                tt = np.ascontiguousarray(t)
                for k in range(cnt):  # ref: add_component overwrites (ecs.rs:114-118)
                    one = tt[k:k + 1]
                    oracle.lib().orc_ecs_add_component(ref._h, C.c_uint64(TY_T), C.c_void_p(one.ctypes.data), C.c_uint64(48),
                                                       C.c_uint64(first + k))
            elif op == "camera":
                cam = None if cam is not None else vp
                ctx.set_view_projection(cam)
            elif op == "delete":
                ids = rng.choice(sc.n, 20, replace=False).astype(np.uint32)
                for e in ids:  # the host-side copy of the Velocity mask (re-uploaded by "zvel") forgets them too
                    masks[capi.TB_VELOCITY][int(e) // 64] &= ~(np.uint64(1) << np.uint64(int(e) % 64))
                ctx.delete_by_ids(ids)
                a = ids.astype(np.uint64)
                oracle.lib().orc_ecs_delete_by_ids(ref._h, C.c_void_p(a.ctypes.data), len(a))
            elif op == "world":
                w = ctx.download_world_matrices(0, sc.n)
                rf = ref.frame(bounds=False)
                has = rf.has_world.astype(bool)
                if util.libm_pinned():
                    assert np.array_equal(w[has], rf.world[has]), (step, log[-6:])
            elif op == "option":
                name = str(rng.choice(["sorted_cache", "static_buckets", "cuda_graphs", "hier_fused"]))
                ctx.set_option(name, int(rng.integers(0, 2)))
            elif op == "zvel":  # some entities get a z velocity by whole depth steps: later ticks change keys
                v = sc.velocities
                sel = rng.choice(sc.n, 200, replace=False)
                v["v"][sel, 2] = np.float32(100.0)
                ctx.upload_velocities(0, v)
                ctx.upload_presence(capi.TB_VELOCITY, 0, masks[capi.TB_VELOCITY])
                # ref: rewrite the Velocity of those that have one
                have = np.unpackbits(masks[capi.TB_VELOCITY].view(np.uint8), bitorder="little")[:sc.n].astype(bool)
                for e in sel[have[sel]]:
                    one = np.ascontiguousarray(v[e:e + 1])
                    oracle.lib().orc_ecs_add_component(ref._h, C.c_uint64(0x76656c6f63697479), C.c_void_p(one.ctypes.data),
                                                       C.c_uint64(12), C.c_uint64(int(e)))
        g = util.GpuFrame(ctx)
        util.assert_frame_parity(g, ref.frame(view_proj=cam), what=f"{kind} seed {seed} final after {log[-8:]}")
    finally:
        ctx.close()
