"""The headline kernel alone: 2^24 particles, steady single-kernel frames, CUDA-event time per launch."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.time()
from tuber_legacy_b200 import capi, scenes

n = 1 << 24
sc = scenes.config4(n=n)
ctx = capi.Context(n)
ctx.set_entity_count(n)
ctx.upload_transforms(0, sc.transforms)
ctx.upload_sprites(0, sc.sprites)
ctx.upload_velocities(0, sc.velocities)
for _ in range(5):
    ctx.system_integrate(0.01)
    info = ctx.frame()
ctx.profile_enable(True)
for _ in range(30):
    ctx.system_integrate(0.01)
    info = ctx.frame()
print({k: (round(v[0] / max(v[2], 1), 5), v[2]) for k, v in ctx.profile().items()}, info.kernel_launches, round(time.time() - t0, 1))
