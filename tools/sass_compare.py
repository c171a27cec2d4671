"""Offline comparison of two builds, kernel by kernel (cuobjdump -sass; addresses and encodings dropped, constant-bank
offsets normalised): which kernels are instruction for instruction the same, which changed, which are new.
    python tools/sass_compare.py old.cubin tuber_legacy_b200/lib/libtuber_b200.so [old_mangled=new_mangled ...]
Used this round to show, without GPU time, that templating the world-row output left every non-list kernel exactly as in
the build whose full GPU suite and timings are committed (profiles/README.md)."""
import re, subprocess, sys, collections
def funcs(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []; continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur: d[cur].append(re.sub(r"c\[0x0\]\[(R\d+\+)?0x[0-9a-f]+\]", "c[0][X]", m.group(2).strip()))
    return d
A, B = funcs(sys.argv[1]), funcs(sys.argv[2])
same = diff = 0
for k in B:
    if k in A:
        if A[k] == B[k]: same += 1
        else:
            diff += 1; print("DIFFERS", k[:90], len(A[k]), len(B[k]))
    else:
        print("NEW    ", k[:110], len(B[k]))
for k in A:
    if k not in B: print("GONE   ", k[:110], len(A[k]))
print("identical:", same, "different:", diff)
if len(sys.argv) > 3:
    for pair in sys.argv[3:]:
        a, b = pair.split("=")
        print("PAIR", a[:60], "vs", b[:60], "->", "IDENTICAL" if A.get(a) == B.get(b) and A.get(a) else "different", len(A.get(a, [])), len(B.get(b, [])))
