"""profiles/emit_traffic.json (the ncu capture bench.py quotes as roofline.traffic) must say which build it came from, and
the tie bench.py checks at run time must be computable from the built library (cuobjdump is part of the CUDA toolkit)."""
import json
import os
import shutil
import warnings

import pytest

from tuber_legacy_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_evidence_names_its_build():
    tr = json.load(open(os.path.join(ROOT, "profiles", "emit_traffic.json")))
    assert len(tr["source_sha256"]) == 64 and len(tr["kernel_sass_sha256"]) == 64
    assert tr["entities"] == 1 << 24 and tr["algorithmic_bytes_per_launch"] == (1 << 24) * 232
    assert 1.0 <= tr["traffic_over_algorithmic"] < 1.2


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_headline_kernel_hash_is_computable_and_stable():
    build.build()
    h = build.kernel_sass_hash()
    assert len(h) == 64 and h == build.kernel_sass_hash()
    sass = build.kernel_sass(build.LIB_PATH, build.HEADLINE_KERNEL)
    assert any(line.startswith("UBLKCP") for line in sass), "the headline kernel streams its rows with TMA bulk copies"
    assert not any(line.startswith("BAR.SYNC") for line in sass), "warp-autonomous: no block barrier in the tile loop"
    tr = json.load(open(os.path.join(ROOT, "profiles", "emit_traffic.json")))
    if tr["kernel_sass_sha256"] != h and tr["source_sha256"] != build.source_hash():
        warnings.warn("profiles/emit_traffic.json was captured from another headline kernel: bench.py will report "
                      "roofline.traffic = null until tools/ncu_traffic.py is re-run on a B200")
