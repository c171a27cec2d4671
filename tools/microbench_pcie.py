"""PCIe rates of the box: pinned H2D alone, D2H alone, both at once on two streams (what the e2e leg can hope for)."""
import json
import torch

h2d_bytes, d2h_bytes = 1275068416, 2483028032
hs = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
hd = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
ds = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
dd = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        s1.synchronize()
        s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def h2d(chunks=1):
    with torch.cuda.stream(s1):
        s1.wait_stream(torch.cuda.current_stream())
        n = h2d_bytes // chunks
        for i in range(chunks):
            ds[i * n:(i + 1) * n].copy_(hs[i * n:(i + 1) * n], non_blocking=True)


def d2h(chunks=1):
    with torch.cuda.stream(s2):
        s2.wait_stream(torch.cuda.current_stream())
        n = d2h_bytes // chunks
        for i in range(chunks):
            hd[i * n:(i + 1) * n].copy_(dd[i * n:(i + 1) * n], non_blocking=True)


a = timed(h2d)
b = timed(d2h)
c = timed(lambda: (h2d(), d2h()))
d = timed(lambda: (h2d(8), d2h(8)))
print(json.dumps({"h2d_alone_ms": round(a, 2), "h2d_GBps": round(h2d_bytes / a / 1e6, 1), "d2h_alone_ms": round(b, 2),
                  "d2h_GBps": round(d2h_bytes / b / 1e6, 1), "both_ms": round(c, 2), "both_8chunks_ms": round(d, 2),
                  "both_total_GBps": round((h2d_bytes + d2h_bytes) / c / 1e6, 1)}))
