"""The multi-GPU frame protocol (tuber_legacy_b200/shard.py) with world_size 2 and 3 over the gloo
backend on CPU. Each rank's device context is replaced by a numpy test double that computes the
same quantities the C ABI's tb_shard_* calls return (key statistics, short-key histogram, placement),
from the CPU oracle's frame of that rank's id range. What is tested here is the host-side exchange:
the statistics reduction, the histogram all-gather, and that histogram placement interleaves the
ranks into exactly the single-process global order."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, result_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from tuber_legacy_b200 import scenes
    from tuber_legacy_b200.shard import ShardBackend, ShardedFrame
    from test_plan import FramePlan
    import subprocess
    so = os.path.join(HERE, "helpers", "_plan_capi.so")
    if rank == 0 and not os.path.exists(so):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "helpers", "plan_capi.cpp")], check=True)
    dist.barrier()
    L = C.CDLL(so)
    L.plan_build.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(FramePlan)]
    L.plan_pext.restype = C.c_uint64
    L.plan_pext.argtypes = [C.POINTER(FramePlan), C.c_uint64]

    first, cnt = scenes.shard_range(n, rank, world)
    sc = scenes.config2(n=n, first=first, count=cnt, seed=seed)
    local = oracle.Scene(cnt, sc.transforms, sc.sprites).frame()  # this rank's own sorted frame

    class Double(ShardBackend):
        def key_stats(self):
            k = local.keys
            k_or = int(np.bitwise_or.reduce(k)) if len(k) else 0
            k_nand = int(np.bitwise_or.reduce(~k)) if len(k) else 0
            return [k_or, k_nand, len(k)]

        def plan(self, g):
            self.f = FramePlan()
            L.plan_build(int(g[0]), (~int(g[1])) & 0xFFFFFFFFFFFFFFFF, C.byref(self.f))
            self.short = np.array([L.plan_pext(C.byref(self.f), int(k)) for k in local.keys], np.int64)
            return 1 << self.f.pext.nbits

        def key_histogram(self, hist):
            hist.zero_()
            hist += torch.from_numpy(np.bincount(self.short, minlength=hist.numel()).astype(np.int32))

        def place(self, all_hists):
            H = all_hists.numpy().astype(np.int64)
            total = H.sum(axis=0)
            g_excl = np.cumsum(total) - total
            lower = H[:rank].sum(axis=0)
            l_excl = np.cumsum(H[rank]) - H[rank]
            self.remap = g_excl + lower - l_excl
            return int(total.sum())

        def frame(self):
            # local sorted position + remap[short key] = global position
            return np.arange(len(self.short)) + self.remap[self.short]

    drv = ShardedFrame(Double(), rank, world)
    gpos, total = drv.step()
    np.savez(os.path.join(result_dir, f"rank{rank}.npz"), gpos=gpos, ent=local.order.astype(np.int64) + first,
             total=total)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_histogram_placement_over_gloo(world, tmp_path):
    n, seed = 30_001, 77
    mp.spawn(_worker, args=(world, _free_port(), n, seed, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import oracle
    from tuber_legacy_b200 import scenes
    sc = scenes.config2(n=n, seed=seed)
    want = oracle.Scene(n, sc.transforms, sc.sprites).frame().order.astype(np.int64)
    got = np.full(n, -1, np.int64)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert int(d["total"]) == n
        assert np.all(got[d["gpos"]] == -1), "two elements placed at the same global position"
        got[d["gpos"]] = d["ent"]
    np.testing.assert_array_equal(got, want)


def test_shard_ranges_tile_the_scene():
    sys.path.insert(0, ROOT)
    from tuber_legacy_b200 import scenes
    for n, w in ((10, 3), (1 << 20, 8), (7, 8), (64, 1)):
        spans = [scenes.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    # a shard generated on its own equals the same id range of the whole scene
    whole = scenes.config2(n=5000, seed=3)
    part = scenes.config2(n=5000, first=1234, count=1000, seed=3)
    assert whole.transforms[1234:2234].tobytes() == part.transforms.tobytes()
    assert whole.sprites[1234:2234].tobytes() == part.sprites.tobytes()
