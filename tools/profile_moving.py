"""2^24 sorted sprites, every entity moving in x / y: steady frames on the draw-ordered working storage (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from tuber_legacy_b200 import scenes
import util

n = int(os.environ.get("TB_TUNE_N", 1 << 24))
sc = scenes.config2(n=n)
sc.velocities = scenes.config4(n=n, seed=77).velocities
ctx = util.make_ctx(sc)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
for k in range(10):
    ctx.system_integrate(0.01)
    info = ctx.frame()
ctx.profile_enable(True)
for k in range(5):
    ctx.system_integrate(0.01)
    info = ctx.frame()
print({k: round(v[0] / max(v[2], 1), 4) for k, v in ctx.profile().items()}, info.kernel_launches)
