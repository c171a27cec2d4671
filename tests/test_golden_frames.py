"""Committed golden frames (tests/golden/frames.npz, written by tests/golden/make_frames.py from the CPU
oracle): the oracle must keep reproducing them bit for bit, and the CUDA path is checked against the
committed bytes (integers exact, floats within the parity tolerance), independent of the oracle build."""
import os
import types

import numpy as np
import pytest

from tuber_legacy_b200 import capi

import util

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["particles", "flat_sorted", "flat_random_z", "hierarchy", "camera"]


def load_case(name):
    z = np.load(os.path.join(HERE, "golden", "frames.npz"))
    get = lambda k: z[f"{name}/{k}"] if f"{name}/{k}" in z.files else None
    sc = types.SimpleNamespace()
    sc.transforms = np.ascontiguousarray(get("transforms")).view(capi.TRANSFORM_DTYPE).reshape(-1)
    sc.sprites = np.ascontiguousarray(get("sprites")).view(capi.SPRITE_DTYPE).reshape(-1)
    v = get("velocities")
    sc.velocities = None if v is None else np.ascontiguousarray(v).view(capi.VELOCITY_DTYPE).reshape(-1)
    sc.parents = get("parents")
    sc.n = len(sc.transforms)
    masks = {c: get(f"mask{c}") for c in range(4) if get(f"mask{c}") is not None}
    ticks, dt = get("ticks_dt")
    gold = types.SimpleNamespace()
    gold.order = get("order")
    gold.batches = np.ascontiguousarray(get("batches")).view(capi.BATCH_DTYPE).reshape(-1)
    gold.instances = np.ascontiguousarray(get("instances")).view(capi.INSTANCE_DTYPE).reshape(-1)
    gold.vertices = np.ascontiguousarray(get("vertices")).view(capi.VERTEX_DTYPE).reshape(-1, 4)
    gold.transforms_after = get("transforms_after")
    return sc, masks or None, int(ticks), float(dt), get("view_proj"), gold


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_golden_frames(name):
    sc, masks, ticks, dt, vp, gold = load_case(name)
    ref = util.make_oracle(sc, masks)
    for _ in range(ticks):
        ref.step(dt)
    fr = ref.frame(vp)
    np.testing.assert_array_equal(fr.order, gold.order)
    np.testing.assert_array_equal(fr.batches.view(np.uint32), gold.batches.view(np.uint32))
    np.testing.assert_array_equal(fr.instances.view(np.uint32), gold.instances.view(np.uint32))
    np.testing.assert_array_equal(fr.vertices.view(np.uint32), gold.vertices.view(np.uint32))
    np.testing.assert_array_equal(ref.transforms().view(np.uint32).reshape(-1, 12), gold.transforms_after)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_path_matches_the_golden_frames(name):
    sc, masks, ticks, dt, vp, gold = load_case(name)
    ctx = util.make_ctx(sc, masks)
    ctx.set_view_projection(vp)
    for _ in range(ticks):
        ctx.system_integrate(dt)
    g = util.GpuFrame(ctx)
    util.assert_frame_parity(g, gold)
    t = ctx.download_transforms(0, sc.n)
    present = np.ones(sc.n, bool)
    if masks and capi.TB_TRANSFORM in masks:
        i = np.arange(sc.n)
        present = ((masks[capi.TB_TRANSFORM][i // 64] >> (i % 64).astype(np.uint64)) & np.uint64(1)).astype(bool)
    np.testing.assert_array_equal(t.view(np.uint32).reshape(-1, 12)[present], gold.transforms_after[present])
    ctx.close()
