"""The C++ host side (include/tuber_host.hpp: Ecs / SystemBundle / GraphicsAPI mirroring the
reference's Rust API) — compiled here, run on the GPU box. tests/cpp/host_test.cpp replays the
reference's own ecs.rs / system.rs unit tests against it and checks render_scene against the oracle."""
import os
import subprocess

import pytest

from tuber_legacy_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_host_test")


def compile_host_test():
    lib = build.build()
    libdir = os.path.dirname(lib)
    deps = [SRC, os.path.join(ROOT, "include", "tuber_host.hpp"), os.path.join(ROOT, "include", "tuber_b200.h"), lib]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-Wall", "-Werror", "-o", EXE, SRC,
                        "-L" + libdir, "-ltuber_b200", "-Wl,-rpath," + libdir], check=True)
    return EXE


def test_cpp_host_compiles_and_fails_loudly_without_a_gpu():
    exe = compile_host_test()
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "no sm_100 device" in r.stdout  # no CPU fallback


@pytest.mark.gpu
def test_cpp_host_replays_reference_tests_and_matches_oracle():
    exe = compile_host_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("ecs_new_and_insert", "ecs_query", "ecs_query_one_and_optional", "systems", "source_only_semantics", "query_items_and_borrows", "device_bundle",
                 "frame_parity(flat)", "frame_parity(hierarchy)"):
        assert "ok " + name in r.stdout
    assert "ALL OK" in r.stdout
