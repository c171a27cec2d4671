"""The library-owned multi-GPU frame (tb_shard_init / tb_frame_sharded): exchanges inside the library.

* local transport: N contexts of this process, one thread each, all on cuda:0 — the whole protocol (statistics and
  histogram all-gathers, list mapping, barrier, direct stores into the presenting rank's list, vertex regeneration)
  against the single-process oracle frame. Runs on a single-GPU box.
* NCCL transport: one process per GPU (tests/helpers/shard_worker.py), world = min(4, device_count); the presenting
  rank's global order / instances / vertices / batch table are compared byte for byte with a single-context frame of
  the whole scene. Needs >= 2 GPUs (NCCL refuses two ranks on one device); with one GPU the same code path still runs
  with world = 1."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from tuber_legacy_b200 import capi, scenes

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_local(full, nranks, ticks=0, frames=1, root=0, devices=None, hook=None):
    """`full`: the whole scene. Returns the presenting rank's frame of the last of `frames` frames (`g.infos`: every
    frame's info on that rank). `hook(rank, ctx, frame_index, first, count)` runs before every frame."""
    sid = capi.shard_local_id()
    results, errors = [None] * nranks, []

    def worker(rank):
        try:
            first, count = scenes.shard_range(full.n, rank, nranks)
            ctx = capi.Context(max(count, 1), device=(devices[rank] if devices else 0))
            ctx.shard_init(rank, nranks, sid, first, root)
            ctx.set_entity_count(count)
            if count:
                ctx.upload_transforms(0, full.transforms[first:first + count])
                ctx.upload_sprites(0, full.sprites[first:first + count])
                if full.velocities is not None:
                    ctx.upload_velocities(0, full.velocities[first:first + count])
            infos = []
            for f in range(frames):
                if hook is not None:
                    hook(rank, ctx, f, first, count)
                for _ in range(ticks):
                    ctx.system_integrate(0.01)
                info = ctx.frame_sharded()
                infos.append(info)
            if rank == root:
                class G:
                    pass
                g = G()
                g.info = info
                g.infos = infos
                m, b = info.num_instances, info.num_batches
                g.order = ctx.download_order(m)
                g.instances = ctx.download_instances(m)
                g.vertices = ctx.download_vertices(m)
                g.batches = ctx.download_batches(b)
                results[rank] = g
            ctx.shard_finalize()
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    return results[root]


@pytest.mark.parametrize("nranks,root", [(1, 0), (2, 0), (3, 2), (8, 0)])
def test_sharded_frame_over_the_local_transport_matches_the_oracle(nranks, root):
    sc = scenes.config2(n=90_001, seed=51)
    g = run_local(sc, nranks, root=root)
    r = util.make_oracle(sc).frame()
    util.assert_frame_parity(g, r, what=f"tb_frame_sharded, local transport, {nranks} ranks")


def test_sharded_particles_tick_and_steady_frames():
    sc = scenes.config4(n=70_000, seed=52)
    ref = util.make_oracle(sc)
    for _ in range(3):
        ref.step(0.01)
    g = run_local(sc, 4, ticks=1, frames=3)
    # a plan of one pass: placed once, then the single-kernel frame with the kept placement on every rank
    assert [i.sort_passes for i in g.infos] == [1, 0, 0]
    util.assert_frame_parity(g, ref.frame(), what="tb_frame_sharded, particles, third frame")


def moving_sorted_scene(n, seed):
    """Sorted sprites (64 textures x 16 layers x 16 depths) that all move in x / y: the keys stay, the matrices change."""
    sc = scenes.config2(n=n, seed=seed)
    sc.velocities = scenes.config4(n=n, seed=seed + 1000).velocities
    sc.velocities["v"][:, 2] = 0.0
    return sc


@pytest.mark.parametrize("nranks,root", [(1, 0), (2, 1), (4, 0), (8, 0)])
def test_sharded_steady_frames_stream_the_kept_list(nranks, root):
    """From the second frame on nobody re-sorts: every rank streams its draw-ordered working storage (velocity system
    fused) to its kept positions of the global list; one word of agreement before, one after."""
    sc = moving_sorted_scene(70_003, 56)
    ref = util.make_oracle(sc)
    for _ in range(5):
        ref.step(0.01)
    g = run_local(sc, nranks, ticks=1, frames=5, root=root)
    assert g.infos[0].sort_passes > 1
    assert [i.sort_passes for i in g.infos[1:]] == [0, 0, 0, 0], "steady frames must not sort"
    util.assert_frame_parity(g, ref.frame(), what=f"kept list frames, {nranks} ranks")


def test_sharded_kept_list_gives_way_when_keys_move():
    """A few entities drift in z: their keys change every frame, every steady attempt is voided by some rank's key check
    and the frame is redone in full by all of them; the ticks run exactly once either way."""
    sc = moving_sorted_scene(50_000, 57)
    sc.velocities["v"][::5000, 2] = 0.37
    ref = util.make_oracle(sc)
    for _ in range(6):
        ref.step(0.01)
    g = run_local(sc, 3, ticks=1, frames=6)
    assert all(i.sort_passes >= 1 for i in g.infos), "no frame of this scene can keep the list"
    util.assert_frame_parity(g, ref.frame(), what="kept list voided by moving keys")


def test_sharded_kept_list_after_an_upload_on_one_rank():
    """One rank replaces its sprites between steady frames: it votes against the kept list, everybody re-sorts once,
    and the steady frames resume."""
    sc = scenes.config2(n=60_000, seed=58)
    changed = sc.sprites.copy()
    first1, count1 = scenes.shard_range(sc.n, 1, 4)
    changed["layer"][first1:first1 + count1] = (changed["layer"][first1:first1 + count1] + 3) % 16

    def hook(rank, ctx, f, first, count):
        if rank == 1 and f == 3:
            ctx.upload_sprites(0, changed[first:first + count])

    g = run_local(sc, 4, frames=6, hook=hook)
    assert [i.sort_passes > 0 for i in g.infos] == [True, False, False, True, False, False]
    sc.sprites = changed
    util.assert_frame_parity(g, util.make_oracle(sc).frame(), what="kept list after one rank's upload")


def test_sharded_empty_ranks():
    sc = scenes.config2(n=5, seed=53)          # fewer entities than ranks: some shards are empty
    g = run_local(sc, 8)
    util.assert_frame_parity(g, util.make_oracle(sc).frame())


@pytest.mark.parametrize("nranks", [1, 3, 8])
def test_sharded_wide_keys_are_served_by_co_ranking(nranks):
    """continuous z: 40+ key bits, far beyond histogram placement. tb_frame_sharded all-gathers the ranks' sorted full keys
    and co-ranks; the list must be the oracle's, ties between ranks in entity order included."""
    sc = scenes.config2(n=60_000, seed=54, z_mode="random")
    sc.transforms["translation"][::7, 2] = 0.25  # many equal keys across ranks: the tie-break matters
    g = run_local(sc, nranks, frames=2)
    assert g.info.key_bits > 22
    util.assert_frame_parity(g, util.make_oracle(sc).frame(), what=f"sharded wide keys, {nranks} ranks")


def test_sharded_wide_keys_with_ticks():
    sc = scenes.config4(n=40_000, seed=55)
    rng = np.random.default_rng(5)
    sc.transforms["translation"][:, 2] = rng.uniform(-1, 1, sc.n).astype(np.float32)
    sc.velocities["v"][:, 2] = rng.uniform(-1, 1, sc.n).astype(np.float32)
    ref = util.make_oracle(sc)
    for _ in range(2):
        ref.step(0.01)
    g = run_local(sc, 4, ticks=1, frames=2)
    util.assert_frame_parity(g, ref.frame(), what="sharded wide keys, z velocities")


@pytest.mark.parametrize("kind", ["config2", "config4", "config2_random_z", "config2_moving"])
def test_nccl_multiprocess_list_equals_single_context_frame(kind):
    import torch
    world = min(4, torch.cuda.device_count())
    n = 300_000
    out = os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "/tmp"
    id_file = os.path.join("/tmp", f"shard_id_{os.getpid()}_{kind}.bin")
    if os.path.exists(id_file):
        os.remove(id_file)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "helpers", "shard_worker.py"), str(r), str(world),
                               str(n), id_file, kind], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["world"] == world and res["instances"] == n
    assert res["order_equal"] and res["instances_equal"] and res["vertices_equal"] and res["batches_equal"], res
    if kind in ("config2", "config2_moving"):  # sorted in two passes: every frame after the first keeps the list
        assert res["sort_passes_per_frame"][0] > 1 and not any(res["sort_passes_per_frame"][1:]), res
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", f"nccl_parity_world{world}_{kind}.json"), "w") as f:
            json.dump(res, f)


def _shard_fuzz_ops(rng, n, nranks, steps):
    """A deterministic operation list every rank executes in lockstep (frames are collective)."""
    names = ["frame", "tick", "upload", "delete", "sprites", "zvel", "option", "camera"]
    weights = np.array([10, 7, 1, 1, 1, 0.5, 1, 0.5])
    ops = []
    for _ in range(steps):
        op = str(rng.choice(names, p=weights / weights.sum()))
        rank = int(rng.integers(0, nranks))
        first, count = scenes.shard_range(n, rank, nranks)
        if op == "upload":
            f = first + int(rng.integers(0, max(count - 40, 1)))
            ops.append((op, rank, f, int(rng.integers(1, 40))))
        elif op == "delete":
            ops.append((op, rank, (first + rng.choice(count, 10, replace=False)).astype(np.uint32)))
        elif op == "sprites":
            f = first + int(rng.integers(0, max(count - 60, 1)))
            ops.append((op, rank, f, int(rng.integers(1, 60)), int(rng.integers(1, 15))))
        elif op == "zvel":
            ops.append((op, rank, (first + rng.choice(count, 30, replace=False)).astype(np.int64)))
        elif op == "option":
            ops.append((op, rank, str(rng.choice(["sorted_cache", "static_buckets"])), int(rng.integers(0, 2))))
        else:
            ops.append((op,))
    ops.append(("frame",))
    return ops


@pytest.mark.parametrize("kind", ["flat", "particles"])
@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_sharded_random_operation_sequences(kind, seed):
    """The kept list, its working storage and the completion flags survive any interleaving of ticks, partial uploads,
    deletions, sprite changes, key-changing velocities and per-rank option changes: every frame is the oracle's."""
    import ctypes as C

    from oracle import oracle

    nranks, root = 3, seed % 3
    rng = np.random.default_rng(900 + seed)
    n = int(rng.integers(5_000, 9_000))
    sc = scenes.config2(n=n, seed=70 + seed) if kind == "flat" else scenes.config4(n=n, seed=70 + seed)
    v = np.zeros(n, capi.VELOCITY_DTYPE)
    v["v"][:, 0] = rng.uniform(-50, 50, n)
    v["v"][:, 1] = rng.uniform(-50, 50, n)
    sc.velocities = v
    ops = _shard_fuzz_ops(rng, n, nranks, 70)
    vp = capi.orthographic(-800, 800, -600, 600, -100, 100)

    # the oracle's run: the expected frame at every "frame" op
    ref = util.make_oracle(sc)
    want, cam = [], None
    tr, sp, ve = sc.transforms.copy(), sc.sprites.copy(), sc.velocities.copy()
    alive = np.ones(n, bool)
    TY = {"t": 0x7472616e73666f72, "s": 0x7370726974650000, "v": 0x76656c6f63697479}

    def ref_add(ty, rec, size, e):
        one = np.ascontiguousarray(rec)
        oracle.lib().orc_ecs_add_component(ref._h, C.c_uint64(TY[ty]), C.c_void_p(one.ctypes.data), C.c_uint64(size), C.c_uint64(int(e)))

    for op in ops:
        if op[0] == "frame":
            want.append(ref.frame(view_proj=cam))
        elif op[0] == "tick":
            ref.step(0.01)
        elif op[0] == "camera":
            cam = None if cam is not None else vp
        elif op[0] == "upload":
            _, _, f, cnt = op
            cur = ref.transforms().view(capi.TRANSFORM_DTYPE).reshape(-1)[f:f + cnt].copy()
            cur["translation"][:, 0] += np.float32(2.5)
            for k in range(cnt):
                if alive[f + k]:
                    ref_add("t", cur[k:k + 1], 48, f + k)
        elif op[0] == "delete":
            a = op[2].astype(np.uint64)
            alive[op[2]] = False
            oracle.lib().orc_ecs_delete_by_ids(ref._h, C.c_void_p(a.ctypes.data), len(a))
        elif op[0] == "sprites":
            _, _, f, cnt, add = op
            sp["layer"][f:f + cnt] = (sp["layer"][f:f + cnt] + add) % 16
            for k in range(cnt):
                if alive[f + k]:
                    ref_add("s", sp[f + k:f + k + 1], 16, f + k)
        elif op[0] == "zvel":
            ve["v"][op[2], 2] = np.float32(100.0)
            for e in op[2]:
                if alive[e]:
                    ref_add("v", ve[e:e + 1], 12, e)

    sid = capi.shard_local_id()
    got, errors = [], []

    def worker(rank):
        try:
            first, count = scenes.shard_range(n, rank, nranks)
            ctx = capi.Context(count, device=0)
            ctx.shard_init(rank, nranks, sid, first, root)
            ctx.set_entity_count(count)
            ctx.upload_transforms(0, sc.transforms[first:first + count])
            ctx.upload_sprites(0, sc.sprites[first:first + count])
            ctx.upload_velocities(0, sc.velocities[first:first + count])
            my_sp, my_ve = sc.sprites[first:first + count].copy(), sc.velocities[first:first + count].copy()
            my_alive = np.ones(count, bool)
            on = None
            for op in ops:
                mine = len(op) > 1 and op[1] == rank
                if op[0] == "frame":
                    info = ctx.frame_sharded()
                    if rank == root:
                        class G:
                            pass
                        g = G()
                        g.info = info
                        g.order = ctx.download_order(info.num_instances)
                        g.instances = ctx.download_instances(info.num_instances)
                        g.vertices = ctx.download_vertices(info.num_instances)
                        g.batches = ctx.download_batches(info.num_batches)
                        got.append(g)
                elif op[0] == "tick":
                    ctx.system_integrate(0.01)
                elif op[0] == "camera":
                    on = None if on is not None else vp
                    ctx.set_view_projection(on)
                elif op[0] == "upload" and mine:
                    _, _, f, cnt = op
                    for k in range(cnt):  # entity by entity: deleted ones stay deleted (add_component on a dead index is a no-op in the oracle)
                        if my_alive[f - first + k]:
                            t = ctx.download_transforms(f - first + k, 1)
                            t["translation"][:, 0] += np.float32(2.5)
                            ctx.upload_transforms(f - first + k, t)
                elif op[0] == "delete" and mine:
                    my_alive[op[2] - first] = False
                    ctx.delete_by_ids((op[2] - first).astype(np.uint32))
                elif op[0] == "sprites" and mine:
                    _, _, f, cnt, add = op
                    lo = f - first
                    my_sp["layer"][lo:lo + cnt] = (my_sp["layer"][lo:lo + cnt] + add) % 16
                    for k in range(cnt):
                        if my_alive[lo + k]:
                            ctx.upload_sprites(lo + k, my_sp[lo + k:lo + k + 1])
                elif op[0] == "zvel" and mine:
                    for e in op[2] - first:
                        my_ve["v"][e, 2] = np.float32(100.0)
                        if my_alive[e]:
                            ctx.upload_velocities(int(e), my_ve[e:e + 1])
                elif op[0] == "option" and mine:
                    ctx.set_option(op[2], op[3])
            ctx.shard_finalize()
            ctx.close()
        except Exception as e:  # noqa: BLE001
            import traceback
            errors.append((rank, repr(e), traceback.format_exc()[-800:]))

    threads = [threading.Thread(target=worker, args=(r,), daemon=True) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert len(got) == len(want)
    frames = [i for i, op in enumerate(ops) if op[0] == "frame"]
    kept = 0
    for k, (g, r) in enumerate(zip(got, want)):
        kept += int(g.info.sort_passes == 0 and g.info.num_instances > 0)
        util.assert_frame_parity(g, r, what=f"sharded {kind} seed {seed} frame {k} (op {frames[k]}): {[o[0] for o in ops[max(0, frames[k] - 8):frames[k]]]}")
    if not any(op[0] in ("option", "zvel") for op in ops):  # those may legitimately keep every later frame off the kept list
        assert kept > 0, "no frame of the sequence kept the list"
