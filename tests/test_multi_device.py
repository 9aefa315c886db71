"""The multi-GPU path behind the C ABI (rla_set_devices): sharded results must be bit-identical to one GPU.
Needs >= 2 B200s in ONE process (gpurun --gpus 2); skipped on a single-GPU box."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rla():
    import rulinalg_b200 as r
    st = r.lib().rla_init(0)
    assert st == 0, r.lib().rla_strerror(st)
    yield r
    r.lib().rla_set_devices(1)


def ngpu(rla):
    return int(rla.lib().rla_device_count())


def host_gemm(rla, a, b):
    m, k = a.shape
    n = b.shape[1]
    c = np.full((m, n), np.nan, dtype=a.dtype)
    fn = rla.lib().rla_dgemm if a.dtype == np.float64 else rla.lib().rla_sgemm
    assert rla.check(fn(m, k, n, 1.0, a.ctypes.data, k, 1, b.ctypes.data, n, 1, 0.0, c.ctypes.data, n, 1)) == 0
    return c


def test_set_devices_validation(rla):
    l = rla.lib()
    assert l.rla_set_devices(0) == 2
    assert l.rla_set_devices(ngpu(rla) + 1) == 2
    assert l.rla_set_devices(1) == 0 and l.rla_get_devices() == 1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_sharded_bit_identical(rla, oracle, dtype):
    if ngpu(rla) < 2:
        pytest.skip("needs 2 GPUs in one process")
    l = rla.lib()
    m, k, n = 9000, 4096, 6016                       # ragged row split; 4.4e11 flop -> sharded path
    a = oracle.fill_uniform((m, k), 12, dtype)
    b = oracle.fill_uniform((k, n), 2049, dtype)
    assert l.rla_set_devices(1) == 0
    c1 = host_gemm(rla, a, b)
    for g in sorted({2, ngpu(rla)}):
        assert l.rla_set_devices(g) == 0
        cg = host_gemm(rla, a, b)
        assert np.array_equal(c1, cg), f"{g} GPUs: {np.count_nonzero(c1 != cg)} elements differ"
    l.rla_set_devices(1)
    # the oracle on a row sample (tolerance as in test_gpu_parity: ulp <= ceil(4 sqrt k))
    rows = np.array([0, 1, 4499, 4500, 4607, 4608, 8999])
    ref = oracle.gemm(np.ascontiguousarray(a[rows]), b)
    oracle.assert_matrix_eq(c1[rows], ref, comp="ulp", tol=int(np.ceil(4 * np.sqrt(k))))


def test_gemm_pinned_equals_pageable(rla, oracle):
    """pinned operands are DMA'd in place, pageable ones travel through the staging ring: same bits"""
    l = rla.lib()
    m = k = n = 2048
    a = oracle.fill_uniform((m, k), 12)
    b = oracle.fill_uniform((k, n), 2049)
    c_page = host_gemm(rla, a, b)
    ptrs = []
    arrs = []
    for src in (a, b, np.empty((m, n))):
        p = C.c_void_p()
        assert l.rla_host_alloc_pinned(C.byref(p), src.nbytes) == 0
        ptrs.append(p)
        arr = np.frombuffer((C.c_char * src.nbytes).from_address(p.value), dtype=np.float64).reshape(src.shape)
        arr[...] = src
        arrs.append(arr)
    assert rla.check(l.rla_dgemm(m, k, n, 1.0, arrs[0].ctypes.data, k, 1, arrs[1].ctypes.data, n, 1, 0.0, arrs[2].ctypes.data, n, 1)) == 0
    assert np.array_equal(c_page, arrs[2])
    l.rla_set_tuning(b"host_stage", 0)
    c_plain = host_gemm(rla, a, b)
    l.rla_set_tuning(b"host_stage", 1)
    assert np.array_equal(c_page, c_plain)
    del arrs
    for p in ptrs:
        l.rla_host_free_pinned(p)


def test_lu_sharded_bit_identical(rla, oracle):
    if ngpu(rla) < 2:
        pytest.skip("needs 2 GPUs in one process")
    l = rla.lib()
    n = 8192 + 192                                   # ragged last block
    a = oracle.fill_uniform((n, n), 12)
    out = {}
    for g in sorted({1, 2, ngpu(rla)}):
        assert l.rla_set_devices(g) == 0
        lu = a.copy()
        perm = np.empty(n, dtype=np.uint64)
        assert rla.check(l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data)) == 0
        out[g] = (lu, perm)
    l.rla_set_devices(1)
    for g in out:
        assert np.array_equal(out[1][1], out[g][1]), f"perm differs at {g} GPUs"
        assert np.array_equal(out[1][0], out[g][0]), f"factors differ at {g} GPUs"
    # residual of a solve with the sharded factors (HPL scaled residual <= 16)
    lu, perm = out[max(out)]
    b = np.ones(n)
    x = b.copy()
    assert rla.check(l.rla_dgetrs(n, lu.ctypes.data, perm.ctypes.data, x.ctypes.data)) == 0
    r = a @ x - b
    eps = np.finfo(np.float64).eps
    assert np.max(np.abs(r)) / (np.max(np.sum(np.abs(a), axis=1)) * np.max(np.abs(x)) * n * eps) <= 16


def test_lu_sharded_pinned_streams_rows_back(rla, oracle):
    """pinned host matrix + regular block-cyclic layout (n a multiple of 256 * N): finished block rows travel back during
    the factorisation as strided 3-D copies.  Same bits as one GPU."""
    if ngpu(rla) < 2:
        pytest.skip("needs 2 GPUs in one process")
    l = rla.lib()
    n = 8192
    a = oracle.fill_uniform((n, n), 12)
    p = C.c_void_p()
    assert l.rla_host_alloc_pinned(C.byref(p), a.nbytes) == 0
    lu = np.frombuffer((C.c_char * a.nbytes).from_address(p.value), dtype=np.float64).reshape(n, n)
    out = {}
    for g in sorted({1, 2, ngpu(rla)}):
        if n % (256 * g):
            continue
        assert l.rla_set_devices(g) == 0
        lu[...] = a
        perm = np.empty(n, dtype=np.uint64)
        assert rla.check(l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data)) == 0
        out[g] = (lu.copy(), perm)
    l.rla_set_devices(1)
    for g in out:
        assert np.array_equal(out[1][1], out[g][1]), f"perm differs at {g} GPUs"
        assert np.array_equal(out[1][0], out[g][0]), f"factors differ at {g} GPUs"
    del lu
    l.rla_host_free_pinned(p)


def test_lu_sharded_singular(rla, oracle):
    if ngpu(rla) < 2:
        pytest.skip("needs 2 GPUs in one process")
    l = rla.lib()
    n = 8192
    a = oracle.fill_uniform((n, n), 12)
    a[:, 5000] = 0.0                                 # a zero column -> |pivot| < eps at column 5000 (lu.rs:179-183)
    perm = np.empty(n, dtype=np.uint64)
    assert l.rla_set_devices(2) == 0
    assert l.rla_dgetrf(n, a.ctypes.data, perm.ctypes.data) == 1
    l.rla_set_devices(1)


def test_shutdown_and_thread_exit(rla, oracle):
    """rla_shutdown returns every resource and the next call re-initialises; a host thread's context dies with it"""
    l = rla.lib()
    a = oracle.fill_uniform((300, 200), 12)
    b = oracle.fill_uniform((200, 100), 2049)
    c0 = host_gemm(rla, a, b)
    assert l.rla_shutdown() == 0
    assert np.array_equal(c0, host_gemm(rla, a, b))
    res = {}

    def worker():
        assert l.rla_init(0) == 0
        res["c"] = host_gemm(rla, a, b)
        big = oracle.fill_uniform((1024, 1024), 5)          # large enough to start the staging ring + drainer thread
        res["big"] = host_gemm(rla, big, big)
    for _ in range(3):
        t = threading.Thread(target=worker)
        t.start()
        t.join()
        assert np.array_equal(c0, res["c"])


def test_cpp_multi_gpu_through_c_abi(rla, tmp_path):
    if ngpu(rla) < 2:
        pytest.skip("needs 2 GPUs in one process")
    exe = str(tmp_path / "multi_test")
    libdir = os.path.join(ROOT, "rulinalg_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "multi_test.cpp"), "-o", exe,
                           "-L" + libdir, "-lrla_b200", "-Wl,-rpath," + libdir])
    r = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "multi_test ok" in r.stdout
