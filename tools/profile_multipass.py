"""A multi-pass frame (BASELINE configs[1] shape at 2^24 entities: 64 textures x 16 layers x z steps,
21 key bits, 3 radix passes) for ncu captures of radix_scatter / tile_hist / emit_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tuber_legacy_b200 import scenes
import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
sc = scenes.config2(n=n)
ctx = util.make_ctx(sc)
for _ in range(3):
    info = ctx.frame()
ctx.profile_enable(True)
for _ in range(5):
    info = ctx.frame()
print(info.num_instances, info.sort_passes, {k: round(v[0] / v[2], 4) for k, v in ctx.profile().items()})
ctx.close()
