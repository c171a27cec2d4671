"""Sorted sprites, static: the kept-order frame BEFORE the draw-ordered working storage exists (emit_ordered_kernel:
rows gathered at random, outputs written in order), next to the full sort frame's kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from tuber_legacy_b200 import scenes
import util

n = int(os.environ.get("TB_TUNE_N", 1 << 24))
sc = scenes.config2(n=n)
out = {}
for rep in range(3):
    ctx = util.make_ctx(sc, cuda_graphs=0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.profile_enable(True)
    ctx.frame()
    ctx.frame()
    for k, v in ctx.profile().items():
        out.setdefault(k, []).append(round(v[0] / max(v[2], 1), 4))
    ctx.close()
print({k: min(v) for k, v in out.items()})
