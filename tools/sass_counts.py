"""Regenerates profiles/sass_counts.json from the built library: per kernel, how often the SASS mnemonics that prove
the design occur (TMA bulk copies, mbarrier, cp.async, evict-first stores, block barriers, FP64, tensor-core ops).
    python tools/sass_counts.py            # writes profiles/sass_counts.json
Runs on the CPU (cuobjdump only needs the .so)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tuber_legacy_b200 import build  # noqa: E402

MNEMONICS = {
    "UBLKCP": r"\bUBLKCP",                 # cp.async.bulk (TMA bulk copy)
    "SYNCS": r"\bSYNCS",                   # mbarrier
    "LDGSTS": r"\bLDGSTS",                 # cp.async
    "STG.E.EF.128": r"\bSTG\.E\.EF\.128",  # 128-bit evict-first stores
    "BAR.SYNC": r"\bBAR\.SYNC",
    "MATCH": r"\bMATCH",
    "DFMA": r"\bDFMA", "DMUL": r"\bDMUL", "DADD": r"\bDADD",   # the bit-exact libm sinf / cosf (FP64)
    "FFMA": r"\bFFMA",                     # must be absent from parity arithmetic (only address / libm-free helpers)
    "MUFU": r"\bMUFU",
    "HMMA|UTC*MMA (tensor core)": r"\b(HMMA|UTCHMMA|UTCQMMA|UTCIMMA|UTCMMA)",
}


def main():
    lib = build.build()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = {}
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = {k: 0 for k in MNEMONICS}
            continue
        if name is None:
            continue
        for k, pat in MNEMONICS.items():
            if re.search(pat, line):
                kernels[name][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = {"library": os.path.relpath(lib, ROOT), "arch": arch, "kernels": {}}
    for mangled, pretty in zip(kernels, demangled):
        short = re.sub(r"\(.*", "", pretty).replace("tb::", "").replace("void ", "")
        out["kernels"][short] = {k: v for k, v in kernels[mangled].items() if v}
    path = os.path.join(ROOT, "profiles", "sass_counts.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(path, len(out["kernels"]), "kernels", arch)


if __name__ == "__main__":
    main()
