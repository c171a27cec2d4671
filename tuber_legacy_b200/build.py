"""Builds libtuber_b200.so (sm_100a only) in-tree with nvcc.

The library is the product: a C-ABI shared object (include/tuber_b200.h) with no Python or
PyTorch types in its interface. It is built next to this file so that it travels to the GPU
box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtuber_b200.so")

SOURCES = ["tuber_b200.cu"]
HEADERS = ["kernels.cuh", "common.cuh", "storage.cuh", "prepare.cuh", "placement.cuh", "emit.cuh", "chunked.cuh", "shard.cuh", "mathops.cuh", "present.cuh", "comm.h",
           "compose.cuh", "libm_sincosf.h", "plan.h", os.path.join("..", "..", "include", "tuber_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",          # parity: rustc never contracts a*b+c; the kernels also use __fmul_rn/__fadd_rn
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall",
    "-shared", "-cudart", "static",
    "-ldl",                  # NCCL is opened at run time (comm.h), not linked
]


def source_hash():
    """sha256 over the CUDA / C++ sources the library is built from: ties committed ncu evidence (profiles/emit_traffic.json)
    to the build that produced it."""
    import hashlib
    h = hashlib.sha256()
    for name in sorted(SOURCES + [x for x in HEADERS if not x.startswith("..")]):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()


def kernel_sass(path, mangled_regex):
    """Instruction text of the kernel whose mangled name matches `mangled_regex` in a cubin / shared library (cuobjdump
    -sass), without addresses and encodings and with constant-bank OFFSETS normalised: the offset of a kernel parameter
    moves when an unrelated field is added to the parameter struct, the code does not."""
    import re
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, text = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1) if re.search(mangled_regex, m.group(1)) else None
            if cur:
                text[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            t = re.sub(r"c\[0x0\]\[((?:U?R\d+)\+)?0x[0-9a-f]+\]", "c[0][X]", m.group(1).strip())
            if re.match(r"(@!?U?P\d+ )?(MOV|UMOV|IMAD\.MOV\.U32|HFMA2) ", t):  # immediates that carry such offsets
                t = re.sub(r"(0x[0-9a-f]+|-?[0-9.]+(e[-+]?\d+)?)$", "IMM", t)
            text[cur].append(t)
    if len(text) != 1:
        raise RuntimeError(f"{len(text)} kernels match {mangled_regex!r}")
    return next(iter(text.values()))


HEADLINE_KERNEL = r"emit_small_kernelIjLb1E(Lb0E)?EEv"  # emit_small_kernel<u32, true> (not the list-frame instantiation)


def kernel_sass_hash(path=None, mangled_regex=HEADLINE_KERNEL):
    """sha256 over kernel_sass(): ties the ncu capture of the headline kernel to the kernel BINARY that runs."""
    import hashlib
    return hashlib.sha256("\n".join(kernel_sass(path or LIB_PATH, mangled_regex)).encode()).hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compiles the library if it is missing or older than its sources. Returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtuber_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
