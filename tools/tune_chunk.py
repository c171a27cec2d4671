"""Tuning run for the chunked sort: 2^24 sorted sprites (64 textures x 16 layers x 16 depths), static frames and
full-recompute frames, over chunk sizes. One JSON line per setting with per-kernel CUDA-event times."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from tuber_legacy_b200 import scenes
import util

n = int(os.environ.get("TB_TUNE_N", 1 << 24))
sc = scenes.config2(n=n)
shifts = [int(x) for x in os.environ.get("TB_TUNE_SHIFTS", "16,17,18,19,20").split(",") if x]
settings = [dict(chunked_sort=0)]
for extra in ({}, {"chunk_l2_hint": 0}, {"interleaved_rows": 1}):
    settings += [dict(chunked_sort=1, chunk_shift=s, **extra) for s in shifts]
modes = os.environ.get("TB_TUNE_MODES", "static").split(",")
steps = 10
for opts in settings:
    for mode in modes:
        o = dict(opts)
        if mode == "keys_change":
            o.update(sorted_cache=0, static_buckets=0)
        ctx = util.make_ctx(sc, **o)
        stream = torch.cuda.current_stream()
        ctx.set_stream(stream.cuda_stream)
        for _ in range(5):
            info = ctx.frame()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            ctx.frame()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        ctx.profile_enable(True)
        for _ in range(steps):
            ctx.frame()
        prof = ctx.profile()
        print(json.dumps({"n": n, "opts": opts, "mode": mode, "ms": round(ms, 4), "passes": info.sort_passes, "bits": info.key_bits,
                          "kernels": {k: round(v[0] / steps, 4) for k, v in prof.items()}}), flush=True)
        ctx.close()
