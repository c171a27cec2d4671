"""Writes tests/golden/frames.npz: small [SPEC] frames (inputs + the CPU oracle's outputs).

Provenance: the reference has no batching / sorting / hierarchy / velocity code, so
these are NOT reference outputs — "parity unpinned" by the reference. They freeze what the oracle
(oracle/ref_frame.hpp, built on the restated reference math and ECS) produced in this container with
g++ -O2 -ffp-contract=off and this image's glibc sinf/cosf, so that
  * tests/test_golden_frames.py::test_oracle_reproduces_the_golden_frames catches any drift of the
    oracle itself (compiler, libm, edits), and
  * the -m gpu test checks the CUDA path against committed bytes, not only against a freshly built oracle.
Run from the repository root:  python tests/golden/make_frames.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle  # noqa: E402
from tuber_legacy_b200 import capi, scenes  # noqa: E402
import util  # noqa: E402


def cases():
    """name -> (scene, masks, ticks, dt, view_proj)"""
    out = {}
    sc = scenes.config4(n=257, seed=201)  # particles: few buckets, ragged size, sparse components, 2 ticks
    sc.velocities["v"][:, 2] = 0.0
    masks = {capi.TB_TRANSFORM: util.random_mask(sc.n, 0.9, 1), capi.TB_SPRITE: util.random_mask(sc.n, 0.8, 2),
             capi.TB_VELOCITY: util.random_mask(sc.n, 0.7, 3)}
    out["particles"] = (sc, masks, 2, 0.01, None)
    out["flat_sorted"] = (scenes.config2(n=600, seed=202), None, 0, 0.0, None)  # 64 textures x 16 layers x 16 depths
    out["flat_random_z"] = (scenes.config2(n=300, seed=203, z_mode="random"), None, 0, 0.0, None)
    out["hierarchy"] = (scenes.config3(n=400, seed=204, depth=5), None, 0, 0.0, None)
    vp = oracle.orthographic(-1000.0, 1000.0, -1000.0, 1000.0, -50.0, 50.0)
    out["camera"] = (scenes.config2(n=200, seed=205), None, 0, 0.0, vp)
    return out


def main():
    data = {}
    for name, (sc, masks, ticks, dt, vp) in cases().items():
        ref = util.make_oracle(sc, masks)
        for _ in range(ticks):
            ref.step(dt)
        fr = ref.frame(vp)
        data[name + "/transforms"] = sc.transforms.view(np.float32).reshape(-1, 12)
        data[name + "/sprites"] = sc.sprites.view(np.uint32).reshape(-1, 4)
        if sc.velocities is not None:
            data[name + "/velocities"] = sc.velocities.view(np.float32).reshape(-1, 3)
        if sc.parents is not None:
            data[name + "/parents"] = np.asarray(sc.parents, np.uint32)
        for comp, words in (masks or {}).items():
            data[f"{name}/mask{comp}"] = words
        data[name + "/ticks_dt"] = np.array([ticks, dt], np.float64)
        if vp is not None:
            data[name + "/view_proj"] = np.asarray(vp, np.float32)
        data[name + "/order"] = fr.order
        data[name + "/batches"] = fr.batches.view(np.uint32).reshape(-1, 4)
        data[name + "/instances"] = fr.instances.view(np.uint32).reshape(-1, 20)
        data[name + "/vertices"] = fr.vertices.view(np.uint32).reshape(-1, 4)
        data[name + "/transforms_after"] = ref.transforms().view(np.uint32).reshape(-1, 12)
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **data)
    print({k: v.shape for k, v in data.items() if k.endswith("/order")})


if __name__ == "__main__":
    main()
