"""One rank of the NCCL multi-process parity test (tests/test_gpu_sharded_lib.py): python shard_worker.py RANK WORLD N ID_FILE.
Rank 0 creates the exchange id (tb_shard_unique_id) and publishes it through ID_FILE — no torch.distributed anywhere:
the collectives live inside the library. Rank 0 then compares the global draw list it presents with a single-context
frame of the whole scene, byte for byte, and prints one JSON line."""
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from tuber_legacy_b200 import capi, scenes  # noqa: E402


def main():
    rank, world, n, id_file = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    kind = sys.argv[5] if len(sys.argv) > 5 else "config2"
    if rank == 0:
        sid = capi.shard_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(sid)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise SystemExit("no exchange id from rank 0")
            time.sleep(0.05)
        sid = open(id_file, "rb").read()
    def moving(**kw):  # sorted sprites, all moving in x / y: the steady list frames (kept order, chunked completion flags)
        sc = scenes.config2(n=n, seed=64, **kw)
        sc.velocities = scenes.config4(n=n, seed=65, **kw).velocities
        sc.velocities["v"][:, 2] = 0.0
        return sc

    make = {"config2": lambda **kw: scenes.config2(n=n, seed=61, **kw), "config4": lambda **kw: scenes.config4(n=n, seed=62, **kw),
            "config2_random_z": lambda **kw: scenes.config2(n=n, seed=63, z_mode="random", **kw), "config2_moving": moving}[kind]
    frames = 5
    first, count = scenes.shard_range(n, rank, world)
    sc = make(first=first, count=count)
    ctx = capi.Context(max(count, 1), device=rank)
    ctx.shard_init(rank, world, sid, first, 0)
    ctx.set_entity_count(count)
    ctx.upload_transforms(0, sc.transforms)
    ctx.upload_sprites(0, sc.sprites)
    ticks = 0
    if sc.velocities is not None:
        ctx.upload_velocities(0, sc.velocities)
        ticks = 1
    passes = []
    for _ in range(frames):
        for _ in range(ticks):
            ctx.system_integrate(0.01)
        info = ctx.frame_sharded()
        passes.append(int(info.sort_passes))
    res = None
    if rank == 0:
        m, b = info.num_instances, info.num_batches
        got = (ctx.download_order(m), ctx.download_instances(m), ctx.download_vertices(m), ctx.download_batches(b))
        full = make()
        one = capi.Context(full.n, device=0)
        one.set_entity_count(full.n)
        one.upload_transforms(0, full.transforms)
        one.upload_sprites(0, full.sprites)
        if full.velocities is not None:
            one.upload_velocities(0, full.velocities)
        for _ in range(frames):
            for _ in range(ticks):
                one.system_integrate(0.01)
            oi = one.frame()
        want = (one.download_order(oi.num_instances), one.download_instances(oi.num_instances),
                one.download_vertices(oi.num_instances), one.download_batches(oi.num_batches))
        eq = [g.shape == w.shape and g.tobytes() == w.tobytes() for g, w in zip(got, want)]
        res = {"world": world, "instances": int(m), "batches": int(b), "order_equal": eq[0], "instances_equal": eq[1],
               "vertices_equal": eq[2], "batches_equal": eq[3], "sort_passes_per_frame": passes, "crc_order": zlib.crc32(got[0].tobytes()),
               "crc_instances": zlib.crc32(got[1].tobytes()), "transport": "NCCL (libnccl.so.2 opened by the library)"}
        one.close()
    ctx.shard_finalize()
    ctx.close()
    if res is not None:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
