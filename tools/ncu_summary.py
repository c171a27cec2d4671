"""Summarises an .ncu-rep (`ncu --set full`) into JSON: one object per profiled launch with the
counters DESIGN.md / profiles/README.md quote. Usage: python tools/ncu_summary.py report.ncu-rep > out.json"""
import csv
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        o = {"kernel": d.get("Kernel Name", "")}
        for k in WANT:
            if k in d:
                o[k] = (d[k] + " " + u.get(k, "")).strip()
        stalls = {}
        for k, v in d.items():
            if k.startswith(STALL) and k.endswith("_per_issue_active.ratio"):
                try:
                    stalls[k[len(STALL):-len("_per_issue_active.ratio")]] = round(float(v), 2)
                except ValueError:
                    pass
        o["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        res.append(o)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
