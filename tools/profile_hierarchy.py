"""4M-entity scene graph, every entity moving: per-kernel CUDA-event times of the steady frame, with the last level
fused into the emit (hier_fused=1) and without."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from tuber_legacy_b200 import scenes
import util

sc = scenes.config3()
sc.velocities = scenes.config4(n=sc.n, seed=77).velocities
for fused in (1, 0):
    ctx = util.make_ctx(sc, hier_fused=fused)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    for k in range(8):
        ctx.system_integrate(0.01)
        info = ctx.frame()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for k in range(20):
        ctx.system_integrate(0.01)
        ctx.frame()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ctx.profile_enable(True)
    for k in range(10):
        ctx.system_integrate(0.01)
        info = ctx.frame()
    print({"hier_fused": fused, "ms_per_frame": round(ms, 4), "launches": info.kernel_launches,
           "kernels": {k: (round(v[0] / 10, 4), v[2] // 10) for k, v in ctx.profile().items()}}, flush=True)
    ctx.close()
