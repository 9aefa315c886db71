"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/rla_b200.h declares, and FAILS LOUDLY (no CPU fallback) when no B200 is present."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rla():
    import __graft_entry__ as ge
    ge.build_library()
    import rulinalg_b200 as r
    return r


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rla_b200.h")).read()
    return sorted(set(re.findall(r"RLA_API\s+[\w\s\*]+?\b(rla_\w+)\s*\(", text)))


def test_header_declares_what_python_binds(rla):
    assert header_symbols() == sorted(rla.SYMBOLS)


def test_library_exports_every_declared_symbol(rla):
    out = subprocess.check_output(["nm", "-D", "--defined-only", rla.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, missing
    # nothing but the declared ABI leaks out of the library
    leaked = [s for s in exported if not s.startswith("rla_")]
    assert not leaked, leaked[:10]
    lib = rla.lib()
    for s in header_symbols():
        assert hasattr(lib, s)


def test_library_is_sm100a_only(rla):
    out = subprocess.run(["cuobjdump", "-lelf", rla.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_kernels_use_the_intended_pipes(rla):
    """SASS evidence: FP64 GEMM on the tensor pipe (DMMA), FP32 GEMM on the FMA pipe (Blackwell packed
    FFMA2 = two IEEE fp32 FMAs per lane) with no HMMA/TF32, operands staged with cp.async (LDGSTS)."""
    sass = subprocess.run(["cuobjdump", "-sass", rla.LIB_PATH], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    by_name = {b.split("\n", 1)[0].strip(): b for b in blocks[1:]}
    dg = [v for k, v in by_name.items() if "dgemm_dmma_kernel" in k]
    sg = [v for k, v in by_name.items() if "sgemm_ffma_kernel" in k]
    assert dg and sg
    for v in dg:
        assert v.count("DMMA.8x8x4") >= 64 and "LDGSTS" in v
    for v in sg:
        assert v.count("FFMA2") >= 512 and "LDGSTS" in v
        assert "HMMA" not in v and "DMMA" not in v


def test_no_cpu_fallback_without_device(rla):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for CPU boxes")
    lib = rla.lib()
    assert lib.rla_device_count() == 0
    assert lib.rla_init(0) == -3
    a = rla.Matrix.ones(2, 2)
    with pytest.raises(rla.RlaError, match="no CPU fallback"):
        a * a
    with pytest.raises(rla.RlaError):
        rla.PartialPivLu.decompose(rla.Matrix.identity(3))
    c = np.zeros((2, 2))
    assert lib.rla_dgemm(2, 2, 2, 1.0, c.ctypes.data, 2, 1, c.ctypes.data, 2, 1, 0.0, c.ctypes.data, 2, 1) == -3
    assert b"no CPU fallback" in lib.rla_strerror(-3)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under rulinalg_b200/ (or include/) may reference it."""
    bad = []
    for base in ("rulinalg_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle|liboracle|orc_", txt, re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_mirror_logic_cpu(rla):
    """Host-only pieces of the mirror (no device needed): permutation algebra, panics, unpack."""
    P = rla.PermutationMatrix
    p = P([2, 0, 1])
    assert p.inverse().perm().tolist() == [1, 2, 0]
    v = p * rla.Vector([10.0, 20.0, 30.0])          # out[perm[i]] = b[i]
    assert v.data().tolist() == [20.0, 30.0, 10.0]
    assert P.identity(4).det() == 1.0 and P([1, 0, 2]).det() == -1.0 and P([1, 2, 0]).det() == 1.0
    assert np.array_equal(P([1, 0]).as_matrix().to_numpy(), [[0, 1], [1, 0]])
    m = p * rla.Matrix(3, 2, [1, 2, 3, 4, 5, 6])
    assert m.into_vec() == [3, 4, 5, 6, 1, 2]
    with pytest.raises(rla.Panic, match="Matrix dimensions do not agree."):
        rla.Matrix.ones(2, 3) * rla.Matrix.ones(2, 3)
    with pytest.raises(rla.Panic, match="square"):
        rla.PartialPivLu.decompose(rla.Matrix.ones(2, 3))
    lu = rla.PartialPivLu(rla.Matrix(2, 2, [4.0, 3.0, 0.5, 2.0]), P.identity(2))
    f = lu.unpack()
    assert f.l.into_vec() == [1.0, 0.0, 0.5, 1.0] and f.u.into_vec() == [4.0, 3.0, 0.0, 2.0]
    assert lu.det() == 8.0
    s = rla.MatrixSlice.from_matrix(rla.Matrix(3, 3, range(9)), [1, 1], 2, 2)
    assert s.row_stride() == 3 and s.into_vec() == [4.0, 5.0, 7.0, 8.0]
    # Matrix::det special cases that never reach the device (impl_mat.rs:407-420)
    assert rla.Matrix(2, 2, [2., 3., 1., 2.]).det() == 1.0
    assert rla.Matrix(3, 3, [1., 2., 3., 4., 5., 6., 7., 8., 9.]).det() == 0.0
    assert rla.Matrix(3, 3, [2., 0., 0., 0., 3., 0., 0., 0., 4.]).det() == 24.0


def test_rust_shim_lists_every_source_and_symbol():
    """rust/build.rs must compile exactly the sources the Makefile builds (round 1 shipped a build.rs without
    cholesky.cu: undefined symbols at link), and rust/src/ffi.rs may only name symbols the header declares."""
    mk = open(os.path.join(ROOT, "rulinalg_b200", "csrc", "Makefile")).read()
    mk_srcs = sorted(re.search(r"^SRCS\s*=\s*(.+)$", mk, re.M).group(1).split())
    on_disk = sorted(f for f in os.listdir(os.path.join(ROOT, "rulinalg_b200", "csrc")) if f.endswith(".cu"))
    assert mk_srcs == on_disk
    rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    rs_srcs = sorted(re.findall(r'"(\w+\.cu)"', rs))
    assert rs_srcs == mk_srcs, (rs_srcs, mk_srcs)
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    bound = set(re.findall(r"pub fn (rla_\w+)\s*\(", ffi))
    assert bound and bound <= set(header_symbols()), bound - set(header_symbols())
    for must in ("rla_dgemm", "rla_sgemm", "rla_dgetrf", "rla_dgetrs", "rla_set_devices", "rla_shutdown"):
        assert must in bound


def test_multi_device_entry_points_fail_loudly_without_gpu(rla):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-box check")
    l = rla.lib()
    assert l.rla_device_count() == 0
    assert l.rla_set_devices(2) == rla._lib.RLA_ERR_NO_DEVICE if hasattr(rla, "_lib") else l.rla_set_devices(2) == -3
    assert l.rla_get_devices() == 1
    assert l.rla_shutdown() == 0


def test_operand_hold_bookkeeping_needs_no_device(rla):
    # rla_operand_hold / rla_operand_release are host-side declarations: nesting, mismatched ranges and
    # unknown pointers are answered without touching a GPU; nothing is resident until a product actually runs
    lib = rla.lib()
    a = np.zeros((64, 64))
    p, nb = a.ctypes.data, a.nbytes
    assert lib.rla_operand_resident_bytes() == 0
    assert lib.rla_operand_hold(0, nb) != 0 and lib.rla_operand_hold(p, 0) != 0
    assert lib.rla_operand_hold(p, nb) == 0
    assert lib.rla_operand_hold(p, nb) == 0            # nests
    assert lib.rla_operand_hold(p, nb - 8) != 0        # same base, different range
    assert lib.rla_operand_release(p + 8) != 0         # not a held base
    assert lib.rla_operand_release(p) == 0
    assert lib.rla_operand_release(p) == 0
    assert lib.rla_operand_release(p) != 0             # nothing left
    assert lib.rla_operand_resident_bytes() == 0
    m = rla.Matrix.from_numpy(a, copy=False)
    with m.held() as same:
        assert same is m and not a.flags.writeable     # numpy's stand-in for the shared borrow
        with pytest.raises(ValueError):
            a[0, 0] = 1.0
    assert a.flags.writeable


def test_every_tuning_key_is_documented_in_the_header():
    # rla_set_tuning keys are part of what a maintainer sees: api.cu and the header comment must list the same set
    src = open(os.path.join(ROOT, "rulinalg_b200", "csrc", "api.cu")).read()
    keys = set(re.findall(r'strcmp\(key, "(\w+)"\)', src))
    hdr = open(os.path.join(ROOT, "include", "rla_b200.h")).read()
    doc = hdr[hdr.index("Tuning knobs"):hdr.index("RLA_API int rla_set_tuning")]
    documented = set(re.findall(r'"(\w+)"', doc))
    assert keys and keys == documented, (sorted(keys - documented), sorted(documented - keys))
