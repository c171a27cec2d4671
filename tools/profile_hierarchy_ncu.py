"""4M-entity scene graph, every entity moving, steady frames on the working storage (for ncu: no timing, few frames)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from tuber_legacy_b200 import scenes
import util

sc = scenes.config3()
sc.velocities = scenes.config4(n=sc.n, seed=77).velocities
ctx = util.make_ctx(sc, cuda_graphs=0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for k in range(12):
    ctx.system_integrate(0.01)
    info = ctx.frame()
print(info.kernel_launches)
