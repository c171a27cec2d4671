import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    return oracle.lib()


def pytest_sessionfinish(session, exitstatus):
    """Writes the element-wise error statistics every parity assertion recorded (tests/util.py PARITY_LOG) to
    gpurun_out/parity_report.json, when that directory exists (GPU runs)."""
    import json
    try:
        import util
    except Exception:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not util.PARITY_LOG or not os.path.isdir(out_dir):
        return
    log = util.PARITY_LOG
    summary = {
        "frames_compared": len(log),
        "frames_exact": sum(1 for e in log if e["exact"]),
        "instances_compared": sum(e["instances"] for e in log),
        "model_rel_max": max(e["model_rel_max"] for e in log),
        "model_rel_p999_max": max(e["model_rel_p999"] for e in log),
        "vertex_rel_max": max(e["vertex_rel_max"] for e in log),
        "vertex_rel_p999_max": max(e["vertex_rel_p999"] for e in log),
        "libm_pinned": util.libm_pinned(),
    }
    with open(os.path.join(out_dir, "parity_report.json"), "w") as f:
        json.dump({"summary": summary, "worst": sorted(log, key=lambda e: -e["model_rel_max"])[:20]}, f, indent=1)
