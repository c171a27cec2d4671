"""Timeline of the e2e leg (host buffers through the C ABI): per-call wall times of every lane."""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tuber_legacy_b200 import capi, scenes

n = int(os.environ.get("TB_TUNE_N", 1 << 24))
nlanes = int(os.environ.get("TB_LANES", 3))
steps = int(os.environ.get("TB_STEPS", 4))
sc = scenes.config4(n=n)
torch.cuda.init()


def pinned(a):
    buf, ptr = capi.host_alloc(a.nbytes)
    out = buf.view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


h_t, h_s, h_v = pinned(sc.transforms), pinned(sc.sprites), pinned(sc.velocities)
t_origin = [0.0]
use_locks = int(os.environ.get("TB_LOCKS", 0))
up_lock, dn_lock = threading.Lock(), threading.Lock()


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def lock(l):
    return l if use_locks else _NoLock()


class Lane:
    def __init__(self):
        self.ctx = capi.Context(n, device=0)
        self.ctx.set_entity_count(n)
        self.inst = capi.host_alloc(n * 80)[0].view(capi.INSTANCE_DTYPE)
        self.vtx = capi.host_alloc(n * 64)[0].view(capi.VERTEX_DTYPE).reshape(n, 4)
        self.order = capi.host_alloc(n * 4)[0].view(np.uint32)
        self.log = []

    def timed(self, name, fn, *a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        self.log.append((name, round((t0 - t_origin[0]) * 1e3, 1), round((time.perf_counter() - t0) * 1e3, 1)))
        return r

    def step(self):
        c = self.ctx
        with lock(up_lock):
            self.timed("up_t", c.upload_transforms, 0, h_t)
            self.timed("up_s", c.upload_sprites, 0, h_s)
            self.timed("up_v", c.upload_velocities, 0, h_v)
            if use_locks:
                self.timed("sync", c.synchronize)
        self.timed("tick", c.system_integrate, sc.dt)
        out = self.timed("frame", c.frame)
        with lock(dn_lock):
            self.timed("dn_i", c.download_instances, out.num_instances, out=self.inst)
            self.timed("dn_v", c.download_vertices, out.num_instances, out=self.vtx)
            self.timed("dn_o", c.download_order, out.num_instances, out=self.order)
        self.timed("dn_b", c.download_batches, out.num_batches)


lanes = [Lane() for _ in range(nlanes)]
for ln in lanes:
    ln.step()
    ln.log.clear()
torch.cuda.synchronize()
t_origin[0] = time.perf_counter()


def loop(ln, delay):
    torch.cuda.set_device(0)
    time.sleep(delay)
    for _ in range(steps):
        ln.step()


threads = [threading.Thread(target=loop, args=(ln, i * 0.070 / nlanes)) for i, ln in enumerate(lanes)]
for t in threads:
    t.start()
for t in threads:
    t.join()
torch.cuda.synchronize()
total = (time.perf_counter() - t_origin[0]) * 1e3
print(json.dumps({"lanes": nlanes, "steps": steps * nlanes, "ms_per_step": round(total / (steps * nlanes), 2), "locks": use_locks}))
for i, ln in enumerate(lanes):
    print(i, " ".join(f"{nm}@{t0}+{d}" for nm, t0, d in ln.log[: 9 * 3]))
