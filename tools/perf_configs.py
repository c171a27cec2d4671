"""Per-config device timings of the frame (not the bench line: that is configs[3] in bench.py).
Prints one JSON line per BASELINE config with the per-kernel CUDA-event times."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from tuber_legacy_b200 import capi, scenes
import util

PEAK = 6546.2


def run(name, sc, alg_bytes, steps=int(os.environ.get("TB_PERF_STEPS", "20")), integrate=False, **opts):
    ctx = util.make_ctx(sc, **opts)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    def frames(k):
        for _ in range(k):
            if integrate:
                ctx.system_integrate(sc.dt)
            info = ctx.frame()
        return info

    frames(4)
    # frame time as a user sees it (steady-state frames replay a CUDA graph) ...
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    info = frames(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # ... then the per-kernel CUDA-event times of the same frames, launched one by one
    ctx.profile_enable(True)
    frames(steps)
    prof = ctx.profile()
    out = {"config": name, "n": sc.n, "ms_per_frame": round(ms, 4), "entities_per_s": round(sc.n / ms * 1e3),
           "frame_roofline_frac": round(sc.n * alg_bytes / (ms * 1e-3) / 1e9 / PEAK, 3), "alg_bytes": alg_bytes,
           "key_bits": info.key_bits, "passes": info.sort_passes, "levels": info.hierarchy_levels,
           "batches": info.num_batches, "options": opts,
           "kernels_ms": {k: round(v[0] / steps, 4) for k, v in prof.items()}}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "2r", "3", "3r", "4"]
    if "1" in which:
        run("configs[0] 10k flat", scenes.config1(), 208)
    if "2" in which:
        run("configs[1] 1M flat, z in {0..15}", scenes.config2(), 208)
    if "2r" in which:
        run("configs[1] 1M flat, z ~ U(-1,1)", scenes.config2(z_mode="random"), 208)
    if "3" in which:
        run("configs[2] 4M hierarchy depth 8", scenes.config3(), 212)
    if "3r" in which:
        run("configs[2] 4M hierarchy, random parents", scenes.config3(random_parents=True), 212)
    if "4" in which:
        run("configs[3] 16M particles", scenes.config4(), 232, integrate=True)
    if "5" in which:
        run("configs[4] 64M flat on one GPU", scenes.config5(), 208)
    if "5i" in which:
        run("configs[4] 64M flat on one GPU", scenes.config5(), 208, interleaved_rows=1)
    if "4s" in which:
        run("configs[3] 16M particles, nothing cached between frames", scenes.config4(), 232, integrate=True, static_buckets=0)
    if "2n" in which:
        run("configs[1] 1M flat, z in {0..15}, no z dictionary", scenes.config2(), 208, z_dictionary=0)
    if "5n" in which:
        run("configs[4] 64M flat on one GPU, no z dictionary", scenes.config5(), 208, z_dictionary=0)
    if "16" in which:
        run("16M flat, 64 textures x 16 layers, z in {0..15}", scenes.config2(n=1 << 24), 208)
    if "16n" in which:
        run("16M flat, 64 textures x 16 layers, z in {0..15}, no z dictionary", scenes.config2(n=1 << 24), 208, z_dictionary=0)
    if "16i" in which:
        run("16M flat, 64 textures x 16 layers, z in {0..15}, interleaved rows", scenes.config2(n=1 << 24), 208, interleaved_rows=1)
    if "2i" in which:
        run("configs[1] 1M flat, z in {0..15}, interleaved rows", scenes.config2(), 208, interleaved_rows=1)
    if "4i" in which:
        run("configs[3] 16M particles, interleaved rows", scenes.config4(), 232, integrate=True, interleaved_rows=1)
    if "3i" in which:
        run("configs[2] 4M hierarchy depth 8, interleaved rows", scenes.config3(), 212, interleaved_rows=1)
    if "2c" in which:
        run("configs[1] 1M flat, z in {0..15}, keys treated as changed every frame", scenes.config2(), 208, sorted_cache=0)
    if "3c" in which:
        run("configs[2] 4M hierarchy depth 8, keys treated as changed every frame", scenes.config3(), 212, sorted_cache=0)
    if "16c" in which:
        run("16M flat, 64 textures x 16 layers, z in {0..15}, keys treated as changed every frame", scenes.config2(n=1 << 24), 208,
            sorted_cache=0)
    if "5c" in which:
        run("configs[4] 64M flat on one GPU, keys treated as changed every frame", scenes.config5(), 208, sorted_cache=0)
    if "3v" in which:
        sc = scenes.config3()
        sc.velocities = scenes.config4(n=sc.n, seed=33).velocities  # x / y motion: world rows change every frame, keys do not
        sc.dt = 0.01
        run("configs[2] 4M hierarchy depth 8, every entity moving (propagation every frame, sort kept)", sc, 224, integrate=True)
    if "16v" in which:
        sc = scenes.config2(n=1 << 24)
        sc.velocities = scenes.config4(n=sc.n, seed=34).velocities
        sc.dt = 0.01
        run("16M flat, 64 textures x 16 layers, z in {0..15}, every entity moving (sort kept)", sc, 232, integrate=True)
