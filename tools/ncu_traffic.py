"""Writes profiles/emit_traffic.json from an `ncu --set full` capture of the headline kernel, tagged with the hash of the
sources the library was built from and with the hash of the kernel's own SASS (bench.py reports roofline.traffic only when
the running build matches one of them: the same sources, or instruction for instruction the same kernel).
    python tools/ncu_traffic.py gpurun_out/<report>.ncu-rep 16777216"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tuber_legacy_b200 import build  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, entities):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "emit_small_kernel" not in d.get("Kernel Name", ""):
            continue
        rd = float(d["dram__bytes_read.sum"]) * UNIT[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"]) * UNIT[units[hdr.index("dram__bytes_write.sum")]]
        best = {"kernel": d["Kernel Name"], "entities": entities, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd,
                "dram_bytes_write": wr, "algorithmic_bytes_per_launch": entities * 232,
                "traffic_over_algorithmic": (rd + wr) / (entities * 232), "gpu_time_us_under_ncu": d.get("gpu__time_duration.sum"),
                "source": "ncu --set full --clock-control none (" + os.path.basename(path) + "), dram__bytes_read.sum + dram__bytes_write.sum",
                "source_sha256": build.source_hash(), "kernel_sass_sha256": build.kernel_sass_hash()}
    if best is None:
        raise SystemExit("no emit_small_kernel launch in the report")
    with open(os.path.join(ROOT, "profiles", "emit_traffic.json"), "w") as f:
        json.dump(best, f, indent=1)
    print(json.dumps(best))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24)
