"""Held operands: device-resident GEMM operands across host-API calls, declared immutable by the
caller (rla_operand_hold / rla_operand_release, `Matrix.held()`), for the repeated products of lu.rs:789,907 and
eigen.rs:114-148.  Bit-identical to the un-held path; the upload is skipped on every call after the first."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rla():
    import rulinalg_b200 as r
    st = r.lib().rla_init(0)
    assert st == 0, r.lib().rla_strerror(st)
    return r


def product(rla, a, b):
    return (rla.Matrix.from_numpy(a, copy=False) * rla.Matrix.from_numpy(b, copy=False)).to_numpy()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_held_operand_bit_identical_and_resident(rla, oracle, dtype):
    l = rla.lib()
    m, k, n = 1536, 2048, 1792                      # the row-panel pipeline; (3000, 3100, 2500) below takes the 2-D one
    a1 = oracle.fill_uniform((m, k), 5, dtype)
    a2 = oracle.fill_uniform((m, k), 6, dtype)
    b = oracle.fill_uniform((k, n), 7, dtype)
    ref1 = product(rla, a1, b)
    ref2 = product(rla, a2, b)
    assert l.rla_operand_resident_bytes() == 0
    mb = rla.Matrix.from_numpy(b, copy=False)
    with mb.held():
        got1 = (rla.Matrix.from_numpy(a1, copy=False) * mb).to_numpy()
        kept = l.rla_operand_resident_bytes()
        assert kept >= b.nbytes                     # b's device copy stays (padded row stride)
        got2 = (rla.Matrix.from_numpy(a2, copy=False) * mb).to_numpy()
        assert l.rla_operand_resident_bytes() == kept          # re-used, not duplicated
        # held as the LEFT operand too: b^T-shaped product b[:m2] * c
        c = oracle.fill_uniform((n, 640), 8, dtype)
        sl = mb.sub_slice([0, 0], 1024, n)
        got3 = (sl * rla.Matrix.from_numpy(c, copy=False)).to_numpy()
        got3b = (sl * rla.Matrix.from_numpy(c, copy=False)).to_numpy()
        assert l.rla_operand_resident_bytes() > kept           # a second view (different shape) of the held range
    assert l.rla_operand_resident_bytes() == 0                  # release frees the copies
    assert np.array_equal(got1, ref1) and np.array_equal(got2, ref2)
    assert np.array_equal(got3, got3b)
    assert np.array_equal(got3, product(rla, np.ascontiguousarray(b[:1024]), c))
    # after the release the host data may change, and the next product sees the change
    b[0, 0] += 1.0
    assert not np.array_equal(product(rla, a1, b), ref1)


def test_held_operand_2d_pipeline_and_nesting(rla, oracle):
    l = rla.lib()
    m, k, n = 3000, 3100, 2500
    a = oracle.fill_uniform((m, k), 12, np.float64)
    b = oracle.fill_uniform((k, n), 2049, np.float64)
    ref = product(rla, a, b)
    assert l.rla_operand_hold(a.ctypes.data, a.nbytes) == 0
    assert l.rla_operand_hold(a.ctypes.data, a.nbytes) == 0      # holds nest
    assert l.rla_operand_hold(a.ctypes.data, a.nbytes - 8) != 0  # same base, different range: refused
    got = [product(rla, a, b) for _ in range(3)]
    assert all(np.array_equal(g, ref) for g in got)
    assert l.rla_operand_release(a.ctypes.data) == 0
    assert l.rla_operand_resident_bytes() > 0                    # one hold still outstanding
    assert l.rla_operand_release(a.ctypes.data) == 0
    assert l.rla_operand_resident_bytes() == 0
    assert l.rla_operand_release(a.ctypes.data) != 0             # nothing left to release


def test_held_operand_skips_the_upload(rla, oracle):
    # 8192 x 8192 f64 B (512 MiB): its H2D is ~10 ms of PCIe time per call; with B held a thin product is several times faster
    l = rla.lib()
    k = n = 8192
    m = 512
    a = oracle.fill_uniform((m, k), 3, np.float64)
    b = oracle.fill_uniform((k, n), 4, np.float64)

    def best(reps):
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            c = product(rla, a, b)
            t.append(time.perf_counter() - t0)
        return min(t), c

    t_plain, c_plain = best(3)
    with rla.Matrix.from_numpy(b, copy=False).held():
        product(rla, a, b)                                       # first call uploads and keeps
        t_held, c_held = best(3)
    assert np.array_equal(c_plain, c_held)
    print(f"thin product with B pageable: {t_plain * 1e3:.2f} ms, with B held: {t_held * 1e3:.2f} ms")
    assert t_held < 0.6 * t_plain


def test_held_matrix_in_solves_and_matvec(rla, oracle):
    """The matrix operand of rla_dgetrs / rla_dtrsv / rla_dgemv is served from HBM while its host range is held: for these
    O(n^2)-flop calls the upload is the whole cost (lu.rs:203-206: "multiple such linear systems involving the same A")."""
    import ctypes as C
    l = rla.lib()
    n = 3000
    a = oracle.fill_uniform((n, n), 21, np.float64) + n * np.eye(n)
    lu = a.copy()
    perm = np.zeros(n, dtype=np.uint64)
    assert l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data) == 0
    rhs = [oracle.fill_uniform((n,), 40 + i, np.float64) for i in range(3)]

    def solve(b):
        x = b.copy()
        assert l.rla_dgetrs(n, lu.ctypes.data, perm.ctypes.data, x.ctypes.data) == 0
        return x

    def timed(fn, *args):
        t0 = time.perf_counter(); out = fn(*args); return out, time.perf_counter() - t0

    plain = [solve(b) for b in rhs]
    _, t_plain = timed(solve, rhs[0])
    assert l.rla_operand_resident_bytes() == 0
    assert l.rla_operand_hold(lu.ctypes.data, lu.nbytes) == 0
    try:
        first = solve(rhs[0])                                   # uploads and keeps the factors
        assert l.rla_operand_resident_bytes() >= lu.nbytes
        held = [solve(b) for b in rhs]
        _, t_held = timed(solve, rhs[0])
        assert np.array_equal(first, plain[0])
        for p_, h_ in zip(plain, held):
            assert np.array_equal(p_, h_)
        # the same held matrix as a triangle and in a matrix-vector product
        xt = rhs[1].copy(); xt2 = rhs[1].copy()
        assert l.rla_dtrsv(0, n, lu.ctypes.data, n, xt.ctypes.data) == 0
        y = np.empty(n); y2 = np.empty(n)
        assert l.rla_dgemv(n, n, lu.ctypes.data, n, rhs[2].ctypes.data, y.ctypes.data) == 0
        kept = l.rla_operand_resident_bytes()
        assert l.rla_dtrsv(0, n, lu.ctypes.data, n, xt2.ctypes.data) == 0
        assert l.rla_dgemv(n, n, lu.ctypes.data, n, rhs[2].ctypes.data, y2.ctypes.data) == 0
        assert l.rla_operand_resident_bytes() == kept == first.nbytes * 0 + kept
    finally:
        assert l.rla_operand_release(lu.ctypes.data) == 0
    assert l.rla_operand_resident_bytes() == 0
    xt_plain = rhs[1].copy()
    assert l.rla_dtrsv(0, n, lu.ctypes.data, n, xt_plain.ctypes.data) == 0
    y_plain = np.empty(n)
    assert l.rla_dgemv(n, n, lu.ctypes.data, n, rhs[2].ctypes.data, y_plain.ctypes.data) == 0
    assert np.array_equal(xt, xt_plain) and np.array_equal(xt2, xt_plain)
    assert np.array_equal(y, y_plain) and np.array_equal(y2, y_plain)
    print(f"solve n={n}: factors re-uploaded {t_plain * 1e3:.2f} ms, held {t_held * 1e3:.2f} ms")
    assert t_held < 0.6 * t_plain
