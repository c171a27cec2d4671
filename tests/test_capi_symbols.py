"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports
exactly the symbols include/tuber_b200.h declares. No compute calls (there is no GPU here)."""
import os
import re
import subprocess

from tuber_legacy_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tuber_b200.h")).read()
    return sorted(set(re.findall(r"TB_API\s+[\w\s\*]+?\b(tb_\w+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(tb_\w+)", out)))
    assert exported == header_symbols()


def test_ctypes_binding_covers_the_header():
    capi.load()
    assert sorted(capi.declared_symbols()) == header_symbols()


def test_library_contains_only_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_means_error_not_fallback():
    """Without a B200 tb_ctx_create must fail with TB_ERR_NO_DEVICE (there is no CPU path)."""
    import ctypes as C
    lib = capi.load()
    h = C.c_void_p()
    st = lib.tb_ctx_create(0, 1024, C.byref(h))
    import torch
    if not torch.cuda.is_available():
        assert st == capi.TB_ERR_NO_DEVICE and not h.value
    else:
        assert st == capi.TB_OK
        lib.tb_ctx_destroy(h)


def test_product_does_not_reference_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "tuber_legacy_b200")
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|oracle/|oracle\.|ref_math|ref_ecs|ref_frame")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), os.path.join(dirpath, f)


def test_header_is_plain_c99():
    """include/tuber_b200.h compiles as strict C99 and the struct layouts are the documented ones"""
    src = os.path.join(ROOT, "tests", "helpers", "c_abi_check.c")
    out = os.path.join(ROOT, "tests", "helpers", "_c_abi_check.o")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-c", src, "-o", out], check=True)


def test_cpp_host_header_compiles_standalone():
    """include/tuber_host.hpp is self-contained C++17 (syntax check; no GPU needed)"""
    hdr = os.path.join(ROOT, "include", "tuber_host.hpp")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
