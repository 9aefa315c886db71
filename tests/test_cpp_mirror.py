"""Builds tests/cpp/mirror_test.cpp against include/rulinalg_b200.hpp + librla_b200.so.
CPU box: the binary must report 'no device' (exit 77) -- the C++ mirror has no CPU fallback either.
GPU box (-m gpu): it must reproduce the reference KATs (exit 0)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_binary(tmp_path):
    import __graft_entry__ as ge
    ge.build_library()
    exe = str(tmp_path / "mirror_test")
    libdir = os.path.join(ROOT, "rulinalg_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp"), "-o", exe,
                           "-L" + libdir, "-lrla_b200", "-Wl,-rpath," + libdir])
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = build_binary(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 77, (r.returncode, r.stderr)
        assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_kats_on_gpu(tmp_path):
    exe = build_binary(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "mirror_test ok" in r.stdout
