"""Parity tests proper: the CUDA path (through the C ABI / the Matrix mirror) against the CPU oracle
and the reference's known-answer vectors.  Run on the B200 box:  pytest -m gpu.

Stated tolerances (DESIGN.md "Parity"):
  GEMM   hard gate  |c_gpu - c_ref| <= 2*gamma_k*(|A||B|)_ij, gamma_k = k*u/(1-k*u)   (any order)
         asserted   assert_matrix_eq!(gpu, ref, comp = ulp, tol = ceil(4*sqrt(k)))  on U[0,1) data
         mixed sign comp = float, eps = 2*gamma_k*max(|A||B|), ulp = ceil(4*sqrt(k))
         KATs       comp = exact
  LU     n <= 64 (one panel): factors and perm BIT-EXACT vs the oracle
         n  > 64: identical pivot sequence, P^-1 L U vs A  comp = abs, tol = 8*n*u*rho*max|A|
  solve  n <= 64: BIT-EXACT; larger: scaled residual <= 16 and relative error vs oracle
"""
import ctypes as C
import math

import numpy as np
import pytest

from tests.golden import reference_kats as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rla():
    import rulinalg_b200 as r
    st = r.lib().rla_init(0)
    assert st == 0, r.lib().rla_strerror(st)
    return r


def U(dtype):
    return 2.0 ** -53 if np.dtype(dtype) == np.float64 else 2.0 ** -24


def gamma(k, dtype):
    u = U(dtype)
    return k * u / (1 - k * u)


def M(rla, x, dtype=np.float64):
    return rla.Matrix.from_numpy(np.array(x, dtype=dtype))


def gpu_gemm_host(rla, a, b, alpha=1.0, beta=0.0, c=None):
    """straight through the host-pointer C ABI with arbitrary numpy strides"""
    m, k = a.shape
    n = b.shape[1]
    it = a.dtype.itemsize
    if c is None:
        c = np.full((m, n), np.nan, dtype=a.dtype)     # poison: beta == 0 must never read C
    fn = rla.lib().rla_dgemm if a.dtype == np.float64 else rla.lib().rla_sgemm
    st = fn(m, k, n, alpha, a.ctypes.data, a.strides[0] // it, a.strides[1] // it,
            b.ctypes.data, b.strides[0] // it, b.strides[1] // it, beta,
            c.ctypes.data, c.strides[0] // it, c.strides[1] // it)
    assert rla.check(st) == 0
    return c


def gpu_gemm_dev(rla, a, b, alpha=1.0, beta=0.0, c=None, lda=None, ldb=None, ldc=None):
    """device-resident twin with explicit leading dimensions (odd ld exercises the 8-byte path)"""
    m, k = a.shape
    n = b.shape[1]
    dt = a.dtype
    lda = lda or max(k, 1)
    ldb = ldb or n
    ldc = ldc or n
    ap = np.zeros((m, lda), dt); ap[:, :k] = a
    bp = np.zeros((max(k, 1), ldb), dt); bp[:k, :n] = b
    cp = np.full((m, ldc), np.nan, dt)
    if c is not None:
        cp[:, :n] = c
    da, db, dc = (rla.DeviceBuffer(max(x.nbytes, 16)) for x in (ap, bp, cp))
    da.upload(ap); db.upload(bp); dc.upload(cp)
    fn = rla.lib().rla_dgemm_dev if dt == np.float64 else rla.lib().rla_sgemm_dev
    assert rla.check(fn(m, k, n, alpha, da.ptr, lda, db.ptr, ldb, beta, dc.ptr, ldc, None)) == 0
    out = dc.download((m, ldc), dt)
    for d in (da, db, dc):
        d.free()
    return out[:, :n]


def check_gemm(oracle, a, b, got, positive):
    k = a.shape[1]
    ref = oracle.gemm(a, b)
    bound = 2 * gamma(max(k, 1), a.dtype) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert np.all(err <= bound + 0.0), f"Higham gate violated: max err {err.max()} bound {bound.max()}"
    ulp_tol = math.ceil(4 * math.sqrt(max(k, 1)))
    if positive:
        info = oracle.assert_matrix_eq(got, ref, comp="ulp", tol=ulp_tol)
    else:
        oracle.assert_matrix_eq(got, ref, comp="float", eps=float(bound.max()), ulp=ulp_tol)
        info = {}
    return info


# ============================================================================ GEMM
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_matrix_mul_kat(rla, oracle, dtype):
    # src/matrix/mat_mul.rs:293-341 through the Matrix mirror (exact)
    a = M(rla, K.GEMM_3x2_2x3["a"], dtype)
    b = M(rla, K.GEMM_3x2_2x3["b"], dtype)
    c = a * b
    assert (c.rows(), c.cols()) == (3, 3)
    oracle.assert_matrix_eq(c.to_numpy(), np.array(K.GEMM_3x2_2x3["c"], dtype), comp="exact")


def test_mul_slice_kats(rla, oracle):
    # src/matrix/mat_mul.rs:368-412: operands with row_stride != cols
    g = K.GEMM_SLICE_BASIC
    c = M(rla, g["parent"])
    d = rla.MatrixSlice.from_matrix(c, list(g["start"]), g["rows"], g["cols"])
    assert d.row_stride() == 3
    assert (d * rla.Matrix.ones(2, 2)).into_vec() == [4.0] * 4
    assert (d * d).into_vec() == [8.0] * 4
    g = K.GEMM_SLICE_UNEVEN
    d = rla.MatrixSlice.from_matrix(M(rla, g["parent"]), [0, 0], 2, 2)
    e = d * M(rla, g["rhs"])
    oracle.assert_matrix_eq(e.to_numpy(), np.array(g["c"]), comp="exact")


def test_mul_dimension_mismatch_panics(rla):
    # mat_mul.rs:21
    with pytest.raises(rla.Panic, match="Matrix dimensions do not agree."):
        rla.Matrix.ones(2, 3) * rla.Matrix.ones(2, 3)


SHAPES = [(1, 1, 1), (3, 2, 3), (17, 5, 9), (64, 64, 64), (128, 128, 128), (129, 130, 131), (200, 333, 77),
          (255, 1000, 257), (512, 512, 512), (1024, 1024, 1024), (4096, 256, 256), (65536, 256, 256)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_vs_oracle_positive(rla, oracle, dtype, shape):
    m, k, n = shape
    a = oracle.fill_uniform((m, k), 12, dtype)
    b = oracle.fill_uniform((k, n), 2049, dtype)
    got = (rla.Matrix.from_numpy(a) * rla.Matrix.from_numpy(b)).to_numpy()
    check_gemm(oracle, a, b, got, positive=True)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(129, 130, 131), (512, 512, 512), (300, 2000, 100)])
def test_gemm_vs_oracle_mixed_sign(rla, oracle, dtype, shape):
    m, k, n = shape
    a = oracle.fill_uniform((m, k), 12, dtype, lo=-1.0, scale=2.0)
    b = oracle.fill_uniform((k, n), 2049, dtype, lo=-1.0, scale=2.0)
    got = gpu_gemm_host(rla, a, b)
    check_gemm(oracle, a, b, got, positive=False)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_strides_alpha_beta(rla, oracle, dtype):
    m, k, n = 150, 70, 90
    big_a = oracle.fill_uniform((m, k + 7), 1, dtype)
    big_b = oracle.fill_uniform((k, n + 3), 2, dtype)
    a, b = big_a[:, 2:2 + k], big_b[:, 1:1 + n]                 # row_stride > cols, offset start
    c0 = oracle.fill_uniform((m, n), 3, dtype)
    ref = oracle.gemm(a, b, alpha=0.5, beta=2.0, c=c0.copy())
    got = gpu_gemm_host(rla, a, b, alpha=0.5, beta=2.0, c=c0.copy())
    tol = 4 * gamma(k + 2, dtype) * (0.5 * np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64) + 2 * np.abs(c0))
    assert np.all(np.abs(got.astype(np.float64) - ref) <= tol)
    # general strides: transposed (column-major) operands and a transposed output
    at = np.ascontiguousarray(a.T).T
    bt = np.ascontiguousarray(b.T).T
    ct = np.full((n, m), np.nan, dtype).T
    got2 = gpu_gemm_host(rla, at, bt, c=ct)
    check_gemm(oracle, np.ascontiguousarray(a), np.ascontiguousarray(b), np.ascontiguousarray(got2), positive=True)
    # negative row stride on an input
    ar = np.ascontiguousarray(a)[::-1]
    got3 = gpu_gemm_host(rla, ar, np.ascontiguousarray(b))
    check_gemm(oracle, np.ascontiguousarray(ar), np.ascontiguousarray(b), got3, positive=True)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_device_twin_odd_ld(rla, oracle, dtype):
    # odd leading dimensions => unaligned (8-byte / 4-byte cp.async) kernel variant
    m, k, n = 131, 67, 129
    a = oracle.fill_uniform((m, k), 5, dtype)
    b = oracle.fill_uniform((k, n), 6, dtype)
    got = gpu_gemm_dev(rla, a, b, lda=k + 2 if k % 2 else k + 1, ldb=n + 2, ldc=n + 4)
    check_gemm(oracle, a, b, got, positive=True)
    got = gpu_gemm_dev(rla, a, b, lda=k + 4 - k % 4 + 4, ldb=n + 4 - n % 4, ldc=n + 4 - n % 4)   # aligned variant
    check_gemm(oracle, a, b, got, positive=True)
    # beta = 1 accumulation (the LU trailing-update form)
    c0 = oracle.fill_uniform((m, n), 7, dtype)
    got = gpu_gemm_dev(rla, a, b, alpha=-1.0, beta=1.0, c=c0)
    ref = c0.astype(np.float64) - a.astype(np.float64) @ b.astype(np.float64)
    assert np.max(np.abs(got - ref)) <= 4 * gamma(k + 1, dtype) * (k + 1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_host_2d_wavefront_pipeline(rla, oracle, dtype):
    # large host-pointer products take the 2-D wavefront H2D/kernel/D2H pipeline (ragged panels and chunks here);
    # it must agree bit-for-bit with the row-panel pipeline (same per-element arithmetic) and with the oracle
    m, k, n = (3000, 3100, 2500) if dtype == np.float64 else (4000, 4100, 3500)
    a = oracle.fill_uniform((m, k), 12, dtype)
    b = oracle.fill_uniform((k, n), 2049, dtype)
    l = rla.lib()
    assert l.rla_set_tuning(b"host_gemm_2d", 1) == 0
    got2d = gpu_gemm_host(rla, a, b)
    assert l.rla_set_tuning(b"host_gemm_2d", 0) == 0
    got1d = gpu_gemm_host(rla, a, b)
    assert l.rla_set_tuning(b"host_gemm_2d", 1) == 0
    assert not np.isnan(got2d).any()
    assert np.array_equal(got2d, got1d)
    check_gemm(oracle, a, b, got2d, positive=True)
    if True:
        # k-prefix: the first part of k uploaded and multiplied as rank-512 / rank-256 updates of all of C that continue the
        # accumulators (dgemm / sgemm ACC_C), the wavefront on the rest: every element is still one accumulation chain => same bits
        try:
            for pre, kc in ((4, 512), (7, 256)):
                assert l.rla_set_tuning(b"host_gemm_kprefix", pre) == 0 and l.rla_set_tuning(b"host_gemm_kchunk", kc) == 0
                gotk = gpu_gemm_host(rla, a, b, alpha=-0.75)
                assert l.rla_set_tuning(b"host_gemm_kprefix", 0) == 0
                plain = gpu_gemm_host(rla, a, b, alpha=-0.75)
                assert not np.isnan(gotk).any()
                assert np.array_equal(gotk, plain), (pre, kc)
        finally:
            l.rla_set_tuning(b"host_gemm_kprefix", -1)
            l.rla_set_tuning(b"host_gemm_kchunk", 256)


@pytest.mark.parametrize("mode", [2, 3])
def test_gemm_streamk_variant_vs_oracle(rla, oracle, mode):
    # the opt-in stream-K DGEMM (rla_set_tuning("dgemm_streamk", 2 | 3); off by default): uneven k ranges per CTA,
    # partial tiles reduced in ascending k order by a fix-up kernel.  Same error bound as the tiled kernels (ragged shapes,
    # alpha / beta, one-CTA-per-tile and many-CTAs-per-tile cases), and run-to-run deterministic.
    l = rla.lib()
    try:
        for (m, k, n), alpha, beta in (((300, 1000, 200), 1.0, 0.0), ((129, 77, 513), -1.5, 0.75), ((1024, 1024, 1024), 1.0, 0.0),
                                       ((64, 4096, 64), 1.0, 1.0), ((2048, 96, 2048), 1.0, 0.0)):
            a = oracle.fill_uniform((m, k), 31, np.float64)
            b = oracle.fill_uniform((k, n), 32, np.float64)
            c0 = oracle.fill_uniform((m, n), 33, np.float64)
            assert l.rla_set_tuning(b"dgemm_streamk", mode) == 0
            got = gpu_gemm_dev(rla, a, b, alpha=alpha, beta=beta, c=c0.copy())
            again = gpu_gemm_dev(rla, a, b, alpha=alpha, beta=beta, c=c0.copy())
            assert l.rla_set_tuning(b"dgemm_streamk", 0) == 0
            tiled = gpu_gemm_dev(rla, a, b, alpha=alpha, beta=beta, c=c0.copy())
            assert np.array_equal(got, again)
            bound = 2 * gamma(k + 2, np.float64) * (abs(alpha) * (np.abs(a) @ np.abs(b)) + abs(beta) * np.abs(c0))
            assert np.all(np.abs(got - tiled) <= 2 * bound)
            ref = alpha * oracle.gemm(a, b) + beta * c0
            assert np.all(np.abs(got - ref) <= 2 * bound)
    finally:
        l.rla_set_tuning(b"dgemm_streamk", 0)


def test_gemm_degenerate(rla):
    # m*n == 0 -> no-op; k == 0 with beta == 0 -> zero fill
    assert (rla.Matrix.zeros(0, 3) * rla.Matrix.zeros(3, 4)).rows() == 0
    c = rla.Matrix.zeros(2, 0) * rla.Matrix.zeros(0, 3)
    assert (c.rows(), c.cols()) == (2, 3) and np.all(c.to_numpy() == 0.0)
    a = np.zeros((2, 0)); b = np.zeros((0, 3))
    c0 = np.full((2, 3), 3.0)
    got = gpu_gemm_host(rla, a, b, alpha=1.0, beta=0.5, c=c0.copy())
    assert np.all(got == 1.5)


def test_gemm_linearity_and_freivalds_full_size(rla, oracle):
    # BASELINE size (n = 8192 f64) through size-independent properties: Freivalds C x == A (B x) within
    # the Higham bound, plus 2048 sampled entries against extended-precision dots.
    n = 8192
    l = rla.lib()
    bufs = [rla.DeviceBuffer(n * n * 8) for _ in range(3)]
    assert l.rla_fill_uniform_f64_dev(bufs[0].ptr, n, n, n, 12, 0, 0.0, 1.0, None) == 0
    assert l.rla_fill_uniform_f64_dev(bufs[1].ptr, n, n, n, 2049, 0, 0.0, 1.0, None) == 0
    assert l.rla_dgemm_dev(n, n, n, 1.0, bufs[0].ptr, n, bufs[1].ptr, n, 0.0, bufs[2].ptr, n, None) == 0
    c = bufs[2].download((n, n), np.float64)
    a = oracle.fill_uniform((n, n), 12)
    b = oracle.fill_uniform((n, n), 2049)
    # the device generator is bit-identical to the oracle's
    a_dev = bufs[0].download((n, n), np.float64)
    assert np.array_equal(a_dev, a)
    for d in bufs:
        d.free()
    x = oracle.fill_uniform((n,), 4000)
    lhs = c @ x
    rhs = a @ (b @ x)
    bound = 3 * gamma(n, np.float64) * (np.abs(a) @ (np.abs(b) @ np.abs(x)))
    assert np.all(np.abs(lhs - rhs) <= bound)
    rng = np.random.default_rng(7)
    ii = rng.integers(0, n, 2048); jj = rng.integers(0, n, 2048)
    truth, absd = oracle.gemm_truth_samples(a, b, ii, jj)
    assert np.all(np.abs(c[ii, jj] - truth) <= gamma(n, np.float64) * absd)
    rel = np.max(np.abs(c[ii, jj] - truth) / np.abs(truth))
    assert rel < 8 * math.sqrt(n) * U(np.float64), rel


# ============================================================================ LU
def decompose(rla, a):
    return rla.PartialPivLu.decompose(rla.Matrix.from_numpy(a))


def test_lu_kats_through_api(rla, oracle):
    # tests/mat/mod.rs:100-170, lu.rs:775-793: exact factors and reconstructions, comp = float
    g = K.LU_EXACT_3x3
    f = decompose(rla, np.array(g["a"])).unpack()
    oracle.assert_matrix_eq(f.l.to_numpy(), np.array(g["l"]), comp="float")
    oracle.assert_matrix_eq(f.u.to_numpy(), np.array(g["u"]), comp="float")
    oracle.assert_matrix_eq(f.p.as_matrix().to_numpy(), np.array(g["p"]), comp="float")
    for a in K.LU_RECONSTRUCT:
        a = np.array(a)
        f = decompose(rla, a).unpack()
        k = f.p.inverse() * (f.l * f.u)
        oracle.assert_matrix_eq(k.to_numpy(), a, comp="float")
        assert oracle.is_lower_triangular(f.l.to_numpy()) and oracle.is_upper_triangular(f.u.to_numpy())


def test_lu_inverse_det_solve_kats(rla, oracle):
    # lu.rs:796-862, impl_mat.rs:648-693, tests/mat/mod.rs:4-26
    g = K.LU_INVERSE_4x4
    oracle.assert_matrix_eq(decompose(rla, np.array(g["a"])).inverse().to_numpy(), np.array(g["inv"]), comp="float")
    g = K.LU_DET_4x4
    oracle.assert_matrix_eq(np.array([decompose(rla, np.array(g["a"])).det()]), np.array([g["det"]]), comp="float")
    g = K.LU_SOLVE_4x4
    y = decompose(rla, np.array(g["a"])).solve(rla.Vector(g["b"]))
    oracle.assert_matrix_eq(y.data(), np.array(g["x"]), comp="ulp", tol=g["ulp_tol"])
    g = K.SOLVE_LAPLACIAN
    c = M(rla, g["a"]).solve(rla.Vector(g["b"]))
    oracle.assert_matrix_eq(c.data(), np.array(g["x"]), comp="abs", tol=g["abs_tol"])
    g = K.SOLVE_2x2
    x = M(rla, g["a"]).solve(rla.Vector(g["b"]))
    assert x.size() == 2 and x[0] == 1.0 and x[1] == 2.0
    g = K.SOLVE_IDENTITY
    y = rla.PartialPivLu.decompose(rla.Matrix.identity(g["n"])).solve(rla.Vector(g["b"]))
    oracle.assert_matrix_eq(y.data(), np.array(g["b"]), comp="float")
    # Matrix::det incl. the special cases and singular -> 0 (impl_mat.rs:648-678)
    assert M(rla, [[2., 3.], [1., 2.]]).det() == 1.0
    assert M(rla, [[1., 2., 3.], [4., 5., 6.], [7., 8., 9.]]).det() == 0.0
    oracle.assert_matrix_eq(np.array([M(rla, K.DET_5x5["a"]).det()]), np.array([K.DET_5x5["det"]]), comp="float")
    assert M(rla, K.LU_SINGULAR).det() == 0.0


def test_lu_errors_and_panics(rla):
    # lu.rs:749-772: non-square panics, singular -> ErrorKind::DivByZero
    with pytest.raises(rla.Panic):
        rla.PartialPivLu.decompose(rla.Matrix.ones(2, 3))
    with pytest.raises(rla.Error) as ei:
        decompose(rla, np.array(K.LU_SINGULAR))
    assert ei.value.kind() == rla.ErrorKind.DivByZero
    lu = decompose(rla, np.eye(3))
    with pytest.raises(rla.Panic):
        lu.solve(rla.Vector([1.0, 2.0]))
    # back_substitution's |u_ii| < eps check (mod.rs:333-336) through a hand-made factorisation
    bad = rla.PartialPivLu(M(rla, [[1.0, 2.0], [0.5, 1e-17]]), rla.PermutationMatrix.identity(2))
    with pytest.raises(rla.Error) as ei:
        bad.solve(rla.Vector([1.0, 1.0]))
    assert ei.value.kind() == rla.ErrorKind.DivByZero
    # empty matrix
    e = rla.PartialPivLu.decompose(rla.Matrix.zeros(0, 0))
    assert e.solve(rla.Vector([])).size() == 0 and e.det() == 1.0


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 31, 33, 63, 64])
def test_lu_bit_exact_within_one_panel(rla, oracle, dtype, n):
    # F13: n <= panel width => factors, perm and solve bit-identical to the reference restatement
    for seed, lo, scale in ((12, 0.0, 1.0), (99, -1.0, 2.0)):
        a = oracle.fill_uniform((n, n), seed, dtype, lo=lo, scale=scale)
        ref_lu, ref_perm = oracle.lu_decompose(a)
        f = decompose(rla, a)
        assert f.p.perm().tolist() == ref_perm.tolist()
        oracle.assert_matrix_eq(f.lu.to_numpy(), ref_lu, comp="exact")
        b = oracle.fill_uniform((n,), 4000, dtype)
        x = f.solve(rla.Vector(b)).data()
        oracle.assert_matrix_eq(x, oracle.lu_solve(ref_lu, ref_perm, b), comp="exact")


def test_lu_pivot_rule_ties_and_nan(rla, oracle):
    # first row attaining the max wins (lu.rs:173-178); ties across CTAs of the panel kernel too
    n = 200
    a = oracle.fill_uniform((n, n), 5, lo=-0.5, scale=1.0)
    a[[3, 77, 150], 0] = 0.75          # three-way tie in column 0; rows live in different CTAs
    a[150, 0] = -0.75
    f = decompose(rla, a)
    ref_lu, ref_perm = oracle.lu_decompose(a)
    assert f.p.perm()[3] == 0 and ref_perm[3] == 0
    assert f.p.perm().tolist() == ref_perm.tolist()
    # a NaN below the diagonal never wins a comparison
    a = oracle.fill_uniform((8, 8), 6)
    a[5, 0] = np.nan
    ref_lu, ref_perm = oracle.lu_decompose(a)
    f = decompose(rla, a)
    assert f.p.perm().tolist() == ref_perm.tolist()
    assert np.array_equal(f.lu.to_numpy(), ref_lu, equal_nan=True)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_lu_cluster_panel_identical_to_grid_panel(rla, oracle, dtype):
    """The panel kernels (grid-wide exchange through L2; one thread-block cluster with the panel in registers and
    the exchange over DSMEM, pull and push variants; column slabs with a CTA-local pivot search and one-way hand-overs
    through L2) must produce bit-identical factors, permutations and status: same pivot rule
    (lu.rs:170-178), same unfused arithmetic.  Sizes cover 1/2/4/8/16-CTA clusters, ragged last panels, the hand-over
    from the grid kernel (n > 4096), ties, a NaN diagonal and a singular matrix."""
    import torch
    l = rla.lib()
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    fn = l.rla_dgetrf_dev if dtype == np.float64 else l.rla_sgetrf_dev
    s = torch.cuda.current_stream().cuda_stream

    def factor(a_np, mode):
        assert l.rla_set_tuning(b"lu_cluster", mode) == 0
        n = a_np.shape[0]
        a = torch.from_numpy(a_np.copy()).cuda()
        perm = torch.empty(n, dtype=torch.int64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        rla.check(fn(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s))
        torch.cuda.synchronize()
        return a.cpu().numpy(), perm.cpu().numpy(), int(info.item())

    try:
        cases = []
        for n in (5, 64, 65, 200, 257, 300, 777, 1030, 2100, 4096, 4500):
            cases.append(oracle.fill_uniform((n, n), 700 + n, dtype, lo=-0.5, scale=1.0))
        ties = oracle.fill_uniform((600, 600), 11, dtype, lo=-0.5, scale=1.0)
        ties[:, 0] = 0.25                       # every row ties in column 0: the first one must win in both kernels
        ties[300:, 3] = ties[:300, 3]           # ties between rows that live in different CTAs
        cases.append(ties)
        nan_diag = oracle.fill_uniform((300, 300), 12, dtype, lo=-0.5, scale=1.0)
        nan_diag[0, 0] = np.nan                 # a NaN diagonal stays the pivot (lu.rs:170-171)
        cases.append(nan_diag)
        sing = oracle.fill_uniform((400, 400), 13, dtype, lo=-0.5, scale=1.0)
        sing[:, 130] = 0.0                      # exact zero column -> DivByZero at column 130 in both kernels
        cases.append(sing)
        for a in cases:
            g = factor(a, 0)
            for mode in (1, 2, 3, 4, 5):        # automatic; pushed-row; column-slab wherever it fits; pull only; grid kernel in cluster mode
                c = factor(a, mode)
                assert g[2] == c[2], (a.shape, mode)
                if g[2] == 0:
                    assert np.array_equal(g[1], c[1]), (a.shape, mode)
                    assert np.array_equal(g[0], c[0], equal_nan=True), (a.shape, mode)
    finally:
        l.rla_set_tuning(b"lu_cluster", LU_CLUSTER_DEFAULT)


LU_CLUSTER_DEFAULT = 1


@pytest.mark.parametrize("f32", [0, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_multiplier_division_is_ieee(rla, f32, mode):
    """The pushed-row panel kernel forms m = a / pivot from a correctly rounded reciprocal + five FMAs (lu.cu
    div_via_rcp); the reference divides (lu.rs:606).  2^32 operand pairs per case (arbitrary bit patterns incl. NaN, Inf,
    subnormals; and LU-like |a| <= |b|): every result must equal the IEEE division bit for bit."""
    bad = C.c_uint64(1)
    assert rla.lib().rla_debug_divcheck(f32, mode, 20240 + mode, 1 << 32, C.byref(bad)) == 0
    assert bad.value == 0


def test_lu_panel_kernels_randomized_stress(rla, oracle):
    """a long randomized A/B run of the panel kernels (grid exchange through L2 with hashed
    self-validating messages; cluster pull; cluster push; column slabs).  60 seeded matrices with sizes that hit every
    cluster size, every slab shape and ragged panels, each factored by all of them: bit-identical factors, permutations,
    status."""
    import torch
    l = rla.lib()
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(2024)

    def factor(a_dev0, mode):
        assert l.rla_set_tuning(b"lu_cluster", mode) == 0
        n = a_dev0.shape[0]
        a = a_dev0.clone()
        perm = torch.empty(n, dtype=torch.int64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s))
        return a, perm, info

    try:
        for trial in range(60):
            n = int(rng.choice([130, 256, 300, 511, 640, 1000, 1500, 2048, 2500, 3333, 4096]))
            a0 = torch.from_numpy(oracle.fill_uniform((n, n), 9000 + trial, np.float64, lo=-0.5, scale=1.0)).cuda()
            g = factor(a0, 0)
            for mode in (1, 2, 3, 4, 5):
                c = factor(a0, mode)
                torch.cuda.synchronize()
                assert int(g[2].item()) == int(c[2].item()) == 0
                assert torch.equal(g[1], c[1]), (trial, n, mode)
                assert torch.equal(g[0], c[0]), (trial, n, mode)
    finally:
        l.rla_set_tuning(b"lu_cluster", LU_CLUSTER_DEFAULT)


def lu_checks(rla, oracle, a, f, ref=None):
    n = a.shape[0]
    dt = a.dtype
    lu = f.lu.to_numpy().astype(np.float64)
    perm = f.p.perm()
    assert sorted(perm.tolist()) == list(range(n))
    l = np.tril(lu, -1) + np.eye(n)
    u = np.triu(lu)
    assert np.max(np.abs(np.tril(lu, -1))) <= 1.0                      # partial pivoting bound
    rho = max(1.0, np.max(np.abs(u)) / np.max(np.abs(a)))              # observed growth
    recon = np.empty_like(lu)
    recon[:] = (l @ u)[perm.astype(np.int64)]                          # P^-1 (L U): row i of A sits at perm[i]
    tol = 8 * n * U(dt) * rho * float(np.max(np.abs(a)))
    oracle.assert_matrix_eq(recon, a.astype(np.float64), comp="abs", tol=tol)
    if ref is not None:
        ref_lu, ref_perm = ref
        assert perm.tolist() == ref_perm.tolist(), "pivot sequence differs from the reference restatement"
        return oracle.max_ulp(f.lu.to_numpy(), ref_lu)
    return None


@pytest.mark.parametrize("dtype,n", [(np.float64, 65), (np.float64, 100), (np.float64, 128), (np.float64, 129),
                                     (np.float64, 256), (np.float64, 257), (np.float64, 500), (np.float64, 1024),
                                     (np.float64, 2048), (np.float32, 100), (np.float32, 300), (np.float32, 1024)])
def test_lu_vs_oracle_blocked(rla, oracle, dtype, n):
    a = oracle.fill_uniform((n, n), 12, dtype)
    ref = oracle.lu_decompose(a, fast=True)
    f = decompose(rla, a)
    lu_checks(rla, oracle, a, f, ref)
    # solve parity: HPL-style scaled residual and distance to the oracle solution
    b = oracle.fill_uniform((n,), 4000, dtype)
    x = f.solve(rla.Vector(b)).data().astype(np.float64)
    a64 = a.astype(np.float64)
    dt = a.dtype
    eps = 2 * U(dt)
    res = np.max(np.abs(a64 @ x - b)) / (np.max(np.sum(np.abs(a64), axis=1)) * np.max(np.abs(x)) * n * eps)
    assert res <= 16, res
    x_ref = oracle.lu_solve(ref[0], ref[1], b, fast=True).astype(np.float64)
    kappa = np.linalg.cond(a64, 1)
    assert np.max(np.abs(x - x_ref)) / np.max(np.abs(x_ref)) <= 8 * n * U(dt) * kappa


def test_lu_diag_dominant_no_pivoting(rla, oracle):
    n = 300
    a = oracle.fill_uniform((n, n), 12) + n * np.eye(n)
    f = decompose(rla, a)
    assert f.p.perm().tolist() == list(range(n))
    lu_checks(rla, oracle, a, f, oracle.lu_decompose(a))


def test_lu_full_size_properties(rla, oracle):
    # BASELINE C3: n = 4096 f64 decompose + solve; oracle too slow => residual / backward error only
    n = 4096
    a = oracle.fill_uniform((n, n), 12)
    f = decompose(rla, a)
    lu_checks(rla, oracle, a, f)
    b = np.ones(n)                                           # benches/linalg/lu.rs:125,137
    x = f.solve(rla.Vector(b)).data()
    eps = 2.0 ** -52
    res = np.max(np.abs(a @ x - b)) / (np.max(np.sum(np.abs(a), axis=1)) * np.max(np.abs(x)) * n * eps)
    assert res <= 16, res
    # repeated solves reuse the device-resident factors (handle path) and agree with rla_dgetrs
    x2 = np.array(b, copy=True)
    st = rla.lib().rla_dgetrs(n, f.lu.as_ptr(), f.p.perm().ctypes.data, x2.ctypes.data)
    assert st == 0 and np.array_equal(x2, x)


@pytest.mark.parametrize("n", [1, 3, 17, 64])
def test_inverse_bit_exact_small(rla, oracle, n):
    # PartialPivLu::inverse = n solves of unit vectors (lu.rs:251-285); n <= 64 keeps the exact operation order
    a = oracle.fill_uniform((n, n), 21, lo=-1.0, scale=2.0) + np.eye(n)
    ref_lu, ref_perm = oracle.lu_decompose(a)
    inv = decompose(rla, a).inverse().to_numpy()
    oracle.assert_matrix_eq(inv, oracle.lu_inverse(ref_lu, ref_perm), comp="exact")


@pytest.mark.parametrize("dtype,n", [(np.float64, 65), (np.float64, 300), (np.float64, 700), (np.float64, 1024), (np.float32, 300)])
def test_inverse_blocked_vs_oracle(rla, oracle, dtype, n):
    # blocked multi-RHS inverse (rla_?getri): residual and distance to the reference's column-by-column result
    a = oracle.fill_uniform((n, n), 12, dtype)
    f = decompose(rla, a)
    inv = f.inverse().to_numpy().astype(np.float64)
    a64 = a.astype(np.float64)
    u = U(dtype)
    kappa = np.linalg.cond(a64, 1)
    resid = np.max(np.abs(a64 @ inv - np.eye(n)))
    assert resid <= 8 * n * u * kappa, (resid, kappa)
    if dtype == np.float64:
        ref_lu, ref_perm = oracle.lu_decompose(a, fast=True)
        ref_inv = oracle.lu_inverse(ref_lu, ref_perm)
        assert np.max(np.abs(inv - ref_inv)) / np.max(np.abs(ref_inv)) <= 8 * n * u * kappa
    # Matrix::inverse wrapper (impl_mat.rs:386-388) takes the same path
    inv2 = rla.Matrix.from_numpy(a).inverse().to_numpy()
    assert np.array_equal(inv2.astype(np.float64), inv)


def test_inverse_singular_is_div_by_zero(rla):
    n = 100
    lu = np.triu(np.ones((n, n))) + np.tril(np.full((n, n), 0.5), -1)
    lu[70, 70] = 1e-17
    bad = rla.PartialPivLu(rla.Matrix.from_numpy(lu), rla.PermutationMatrix.identity(n))
    with pytest.raises(rla.Error) as ei:
        bad.inverse()
    assert ei.value.kind() == rla.ErrorKind.DivByZero
    with pytest.raises(rla.Panic):
        rla.Matrix.ones(2, 3).inverse()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_triangular_solves(rla, oracle, dtype):
    # solve_u_triangular / solve_l_triangular (base/mod.rs:1015-1067; tests/mat/mod.rs:28-39; benches/linalg/triangular.rs)
    for fn in ("solve_l_triangular", "solve_u_triangular"):
        with pytest.raises(rla.Error) as ei:
            getattr(M(rla, [[0.0]], dtype), fn)(rla.Vector(np.array([1.0], dtype)))
        assert ei.value.kind() == rla.ErrorKind.DivByZero
    x = M(rla, [[1.0, 0.0], [2.0, 1.0]], dtype).solve_l_triangular(rla.Vector(np.array([1.0, 3.0], dtype)))
    assert x.data().tolist() == [1.0, 1.0]
    with pytest.raises(rla.Panic):
        M(rla, [[1.0, 0.0], [2.0, 1.0]], dtype).solve_u_triangular(rla.Vector(np.array([1.0], dtype)))
    u = U(dtype)
    for n in (5, 64, 65, 300, 1000):
        a = oracle.fill_uniform((n, n), 31, dtype, lo=-1.0, scale=2.0)
        a[np.arange(n), np.arange(n)] = 2.0 + np.arange(n) % 3          # well-conditioned triangles
        a = (a / np.sqrt(n)).astype(dtype)
        a[np.arange(n), np.arange(n)] *= np.sqrt(n).astype(dtype)
        y = oracle.fill_uniform((n,), 32, dtype)
        xl = rla.Matrix.from_numpy(a).solve_l_triangular(rla.Vector(y)).data().astype(np.float64)
        xu = rla.Matrix.from_numpy(a).solve_u_triangular(rla.Vector(y)).data().astype(np.float64)
        rl = oracle.forward_substitution(np.tril(a), y).astype(np.float64)
        ru = oracle.back_substitution(np.triu(a), y).astype(np.float64)
        tol = 64 * n * u
        assert np.max(np.abs(xl - rl)) <= tol * np.max(np.abs(rl)), (fn, n)
        assert np.max(np.abs(xu - ru)) <= tol * np.max(np.abs(ru)), (fn, n)
    # identity of order 10000 (benches/linalg/triangular.rs:6-58) and a strided slice operand
    n = 2000
    eye = rla.Matrix.identity(n, dtype)
    y = oracle.fill_uniform((n,), 33, dtype)
    assert np.array_equal(eye.solve_u_triangular(rla.Vector(y)).data(), y)
    assert np.array_equal(eye.solve_l_triangular(rla.Vector(y)).data(), y)
    big = rla.Matrix.from_numpy(np.pad(np.triu(np.ones((100, 100), dtype)) + 99 * np.eye(100, dtype=dtype), ((3, 0), (5, 2))))
    sl = rla.MatrixSlice.from_matrix(big, [3, 5], 100, 100)
    xs = sl.solve_u_triangular(rla.Vector(np.ones(100, dtype))).data()
    ref = oracle.back_substitution(np.ascontiguousarray(sl._arr), np.ones(100, dtype))
    assert np.max(np.abs(xs.astype(np.float64) - ref)) <= 64 * 100 * u


def test_lu_solve_device_resident_large_property(rla):
    # BASELINE C5-sized path on one GPU through size-independent properties (no CPU oracle at this size):
    # device-resident decompose + solve of A x = A*1 must return ones; |l_ij| <= 1; perm is a permutation.
    import torch
    n = 16384
    l = rla.lib()
    torch.cuda.set_device(0)
    s = torch.cuda.current_stream().cuda_stream
    a0 = torch.empty(n, n, dtype=torch.float64, device="cuda")
    assert l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, s) == 0
    b = a0 @ torch.ones(n, dtype=torch.float64, device="cuda")          # checker-side product (torch), not the product path
    lu = a0.clone()
    perm = torch.empty(n, dtype=torch.int64, device="cuda")
    info = torch.ones(1, dtype=torch.int32, device="cuda")
    assert l.rla_dgetrf_dev(n, lu.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s) == 0
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device="cuda"))
    assert float(torch.tril(lu, -1).abs().max().item()) <= 1.0
    x = b.clone()
    assert l.rla_dgetrs_dev(n, lu.data_ptr(), n, perm.data_ptr(), x.data_ptr(), info.data_ptr(), s) == 0
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    err = float((x - 1.0).abs().max().item())
    resid = float((a0 @ x - b).abs().max().item()) / (float(a0.abs().sum(dim=1).max().item()) * float(x.abs().max().item()) * n * 2.0 ** -52)
    assert resid <= 16, resid
    assert err <= 1e-5, err


# ------------------------------------------------------------------ Cholesky
def spd(oracle, n, seed, dtype=np.float64):
    """well-conditioned symmetric positive definite test matrix: M M^T / n + I from the seeded generator"""
    m = oracle.fill_uniform((n, n), seed, np.float64, lo=-1.0, scale=2.0)
    a = m @ m.T / n + np.eye(n)
    a = (a + a.T) / 2
    return np.ascontiguousarray(a.astype(dtype))


def test_cholesky_kats_through_api(rla, oracle):
    # cholesky.rs doc-tests :39-79 and tests :384-539 through the mirror API, with the reference's comparators
    g = K.CHOL_DOC_3x3
    oracle.assert_matrix_eq(rla.Cholesky.decompose(M(rla, g["a"])).unpack().to_numpy(), np.array(g["l"]), comp="float")
    for g in K.CHOL_UNPACK:
        oracle.assert_matrix_eq(rla.Cholesky.decompose(M(rla, g["a"])).unpack().to_numpy(), np.array(g["l"]), comp="float")
    e = rla.Cholesky.decompose(rla.Matrix.zeros(0, 0))
    assert e.unpack().to_numpy().shape == (0, 0) and e.det() == 1.0 and e.solve(rla.Vector([])).size() == 0
    assert e.inverse().to_numpy().shape == (0, 0)
    for n in (1, 2, 7, 30):                                  # quickcheck property :541-557
        oracle.assert_matrix_eq(rla.Cholesky.decompose(M(rla, np.eye(n))).unpack().to_numpy(), np.eye(n), comp="float")
    for g in K.CHOL_DET:
        oracle.assert_matrix_eq(np.array([rla.Cholesky.decompose(M(rla, g["a"])).det()]), np.array([g["det"]]), comp="float")
    for g in K.CHOL_SOLVE:
        x = rla.Cholesky.decompose(M(rla, g["a"])).solve(rla.Vector(g["b"])).data()
        oracle.assert_matrix_eq(x, np.array(g["x"]), comp="float")
    g = K.CHOL_DOC_SOLVE
    c = rla.Cholesky.decompose(M(rla, g["a"]))
    oracle.assert_matrix_eq(c.solve(rla.Vector(g["b1"])).data(), np.array(g["y1"]), comp="exact")
    oracle.assert_matrix_eq(c.solve(rla.Vector(g["b2"])).data(), np.array(g["y2"]), comp="exact")
    for g in K.CHOL_INVERSE:
        oracle.assert_matrix_eq(rla.Cholesky.decompose(M(rla, g["a"])).inverse().to_numpy(), np.array(g["inv"]), comp="float")


def test_cholesky_errors_and_panics(rla):
    for bad in K.CHOL_SINGULAR:                              # cholesky.rs:418-442
        with pytest.raises(rla.Error) as ei:
            rla.Cholesky.decompose(M(rla, bad))
        assert ei.value.kind() == rla.ErrorKind.DecompFailure and "singular" in str(ei.value)
    with pytest.raises(rla.Error) as ei:                     # negative diagonal (:155-158)
        rla.Cholesky.decompose(M(rla, [[1.0, 0.0], [0.0, -4.0]]))
    assert ei.value.kind() == rla.ErrorKind.DecompFailure and "not all positive" in str(ei.value)
    with pytest.raises(rla.Panic):                           # :117-118, :377-381
        rla.Cholesky.decompose(rla.Matrix.ones(2, 3))
    with pytest.raises(rla.Panic):
        rla.Cholesky.decompose(M(rla, np.eye(3))).solve(rla.Vector([1.0, 2.0]))
    # a failure deep inside a blocked factorisation (column 300 of 400) is reported, not ignored
    a = np.eye(400)
    a[300, 300] = -1.0
    with pytest.raises(rla.Error) as ei:
        rla.Cholesky.decompose(M(rla, a))
    assert "not all positive" in str(ei.value)
    a[300, 300] = 0.0
    with pytest.raises(rla.Error) as ei:
        rla.Cholesky.decompose(M(rla, a))
    assert "singular" in str(ei.value)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 3, 8, 9, 17, 33, 63, 64])
def test_cholesky_bit_exact_within_one_block(rla, oracle, dtype, n):
    # n <= 64: the diagonal-block kernel follows cholesky.rs:134-165 operation by operation (utils::dot's 8 partial sums)
    a = spd(oracle, n, 100 + n, dtype)
    a_upper_garbage = a.copy()
    a_upper_garbage[np.triu_indices(n, 1)] = 12345.0          # only the lower triangle may be read
    ref = oracle.cholesky_decompose(a)
    c = rla.Cholesky.decompose(M(rla, a_upper_garbage, dtype))
    oracle.assert_matrix_eq(c.unpack().to_numpy(), np.tril(ref), comp="exact")
    b = oracle.fill_uniform((n,), 4000, dtype)
    oracle.assert_matrix_eq(c.solve(rla.Vector(b)).data(), oracle.cholesky_solve(ref, b), comp="exact")
    if dtype == np.float64:
        oracle.assert_matrix_eq(c.inverse().to_numpy(), oracle.cholesky_inverse(ref), comp="exact")
        assert c.det() == oracle.cholesky_det(ref)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [65, 200, 257, 300, 777, 1030])
def test_cholesky_blocked_vs_oracle(rla, oracle, dtype, n):
    # beyond one block the trailing updates run on the GEMM kernels (different summation order): gates are the
    # reconstruction error (backward stability: |L L^T - A| <= c n u max|A|) and closeness to the oracle's factor
    a = spd(oracle, n, 300 + n, dtype)
    u = U(dtype)
    ref = np.tril(oracle.cholesky_decompose(a)).astype(np.float64)
    c = rla.Cholesky.decompose(M(rla, a, dtype))
    l = c.unpack().to_numpy().astype(np.float64)
    assert np.all(np.triu(l, 1) == 0)
    recon = np.abs(l @ l.T - a.astype(np.float64)).max()
    assert recon <= 8 * n * u * np.abs(a).max(), recon
    assert np.abs(l - ref).max() <= 64 * n * u * np.abs(ref).max()
    b = oracle.fill_uniform((n,), 4000, dtype)
    x = c.solve(rla.Vector(b)).data().astype(np.float64)
    xr = np.linalg.solve(a.astype(np.float64), b.astype(np.float64))
    kappa = np.linalg.cond(a.astype(np.float64))
    assert np.abs(x - xr).max() <= 64 * n * u * kappa * np.abs(xr).max()
    if n <= 300:
        inv = c.inverse().to_numpy().astype(np.float64)
        assert np.abs(inv @ a.astype(np.float64) - np.eye(n)).max() <= 64 * n * u * kappa


def test_cholesky_device_resident_large_property(rla):
    # n = 8192 on the device: L L^T reproduces A on sampled rows, L is lower triangular with a positive diagonal
    import torch
    n = 8192
    l = rla.lib()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(7)
    m = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    a0 = m @ m.T / n + torch.eye(n, dtype=torch.float64, device="cuda")     # checker-side product (torch)
    a0 = (a0 + a0.T) / 2
    a = a0.clone()
    ws = torch.empty(int(l.rla_potrf_workspace_bytes(n, 8)), dtype=torch.uint8, device="cuda")
    info = torch.ones(1, dtype=torch.int32, device="cuda")
    assert l.rla_dpotrf_dev(n, a.data_ptr(), n, ws.data_ptr(), info.data_ptr(), s) == 0
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    lo = torch.tril(a)
    assert float(torch.diagonal(lo).min().item()) > 0
    rows = torch.tensor([0, 1, 63, 64, 255, 256, 1000, 4095, 4096, 8191], device="cuda")
    recon = lo[rows] @ lo.T
    err = float((recon - a0[rows]).abs().max().item())
    assert err <= 8 * n * 2.0 ** -53 * float(a0.abs().max().item()), err


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_matrix_vector_product(rla, oracle, dtype):
    # &Matrix * &Vector (impl_ops.rs:298-314): row-wise utils::dot in the reference
    a = M(rla, [[1, 2, 3], [4, 5, 6]], dtype)
    assert (a * rla.Vector(np.array([1, 2, 3], dtype))).data().tolist() == [14.0, 32.0]
    with pytest.raises(rla.Panic, match="Matrix and Vector dimensions do not agree."):
        a * rla.Vector(np.array([1, 2], dtype))
    u = U(dtype)
    for (m, n) in ((1, 1), (7, 5), (300, 1000), (1000, 301), (2048, 4096)):
        an = oracle.fill_uniform((m, n), 41, dtype, lo=-1.0, scale=2.0)
        xn = oracle.fill_uniform((n,), 42, dtype, lo=-1.0, scale=2.0)
        got = (rla.Matrix.from_numpy(an) * rla.Vector(xn)).data().astype(np.float64)
        ref = np.array([oracle.dot(an[i], xn) for i in range(min(m, 64))], dtype=np.float64)
        bound = 2 * gamma(n, dtype) * (np.abs(an.astype(np.float64)) @ np.abs(xn.astype(np.float64)))
        assert np.all(np.abs(got[:len(ref)] - ref) <= bound[:len(ref)] + 0.0)
        assert np.all(np.abs(got - an.astype(np.float64) @ xn.astype(np.float64)) <= bound)
    assert (rla.Matrix.zeros(0, 3, dtype) * rla.Vector(np.zeros(3, dtype))).size() == 0
    assert np.all((rla.Matrix.zeros(2, 0, dtype) * rla.Vector(np.zeros(0, dtype))).data() == 0)


def test_launch_counter_and_version(rla):
    l = rla.lib()
    l.rla_launch_count_reset()
    _ = rla.Matrix.ones(4, 4) * rla.Matrix.ones(4, 4)
    assert l.rla_launch_count() >= 1
    assert b"sm_100a" in l.rla_version()


# ============================================================================ parity holes named by the round-1 review
def _freivalds_and_samples(rla, oracle, dtype, m, k, n, samples=1024):
    """C = A B on device-generated U[0,1) operands at a size the CPU oracle cannot sweep: Freivalds C x vs A (B x) through
    the library's gemv (a different, HBM-bound kernel) within 6 gamma_k |A||B||x|, plus sampled entries against
    extended-precision dots within the Higham bound gamma_k sum|a||b| and 8 sqrt(k) u relative."""
    import torch
    l = rla.lib()
    s = torch.cuda.current_stream().cuda_stream
    f64 = np.dtype(dtype) == np.float64
    tdt = torch.float64 if f64 else torch.float32
    fill = l.rla_fill_uniform_f64_dev if f64 else l.rla_fill_uniform_f32_dev
    gemm = l.rla_dgemm_dev if f64 else l.rla_sgemm_dev
    gemv = l.rla_dgemv_dev if f64 else l.rla_sgemv_dev
    a = torch.empty(m, k, dtype=tdt, device="cuda"); b = torch.empty(k, n, dtype=tdt, device="cuda")
    c = torch.full((m, n), float("nan"), dtype=tdt, device="cuda")
    assert fill(a.data_ptr(), m, k, k, 12, 0, 0.0, 1.0, s) == 0
    assert fill(b.data_ptr(), k, n, n, 2049, 0, 0.0, 1.0, s) == 0
    assert rla.check(gemm(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, s)) == 0
    x = torch.empty(n, dtype=tdt, device="cuda")
    assert fill(x.data_ptr(), 1, n, n, 4000, 0, 0.0, 1.0, s) == 0
    y = torch.empty(k, dtype=tdt, device="cuda"); z = torch.empty(m, dtype=tdt, device="cuda"); w = torch.empty(m, dtype=tdt, device="cuda")
    assert rla.check(gemv(k, n, b.data_ptr(), n, x.data_ptr(), y.data_ptr(), s)) == 0
    assert rla.check(gemv(m, k, a.data_ptr(), k, y.data_ptr(), z.data_ptr(), s)) == 0
    assert rla.check(gemv(m, n, c.data_ptr(), n, x.data_ptr(), w.data_ptr(), s)) == 0
    torch.cuda.synchronize()
    g = gamma(max(k, n), dtype)
    assert bool(torch.all((w.double() - z.double()).abs() <= 6 * g * z.double().abs()))
    rng = np.random.default_rng(7)
    ii = rng.integers(0, m, samples); jj = rng.integers(0, n, samples)
    it, jt = torch.as_tensor(ii, device="cuda"), torch.as_tensor(jj, device="cuda")
    a_rows = a.index_select(0, it).cpu().numpy()
    b_cols = b.index_select(1, jt).t().contiguous().cpu().numpy()
    got = c[it, jt].double().cpu().numpy()
    idx = np.arange(samples)
    truth, absd = oracle.gemm_truth_samples(a_rows, b_cols.T, idx, idx)
    err = np.abs(got - truth)
    assert np.all(err <= gamma(k, dtype) * absd)
    assert np.max(err / np.abs(truth)) < 8 * math.sqrt(k) * U(dtype)


def test_gemm_f32_8192_freivalds_and_samples(rla, oracle):
    _freivalds_and_samples(rla, oracle, np.float32, 8192, 8192, 8192)


def test_gemm_f64_16384_freivalds_and_samples(rla, oracle):
    _freivalds_and_samples(rla, oracle, np.float64, 16384, 16384, 16384)


def test_gemm_f32_small_tile_shape_vs_oracle(rla, oracle):
    """the 64 x 128 SGEMM shape (products with fewer 128 x 128 tiles than SMs) against the oracle, incl. ragged edges"""
    for (m, k, n) in ((1024, 1024, 1024), (1000, 515, 777), (70, 33, 130)):
        a = oracle.fill_uniform((m, k), 12, np.float32)
        b = oracle.fill_uniform((k, n), 2049, np.float32)
        assert rla.lib().rla_set_tuning(b"sgemm_cfg", 1) == 0
        try:
            c = gpu_gemm_dev(rla, a, b)
        finally:
            rla.lib().rla_set_tuning(b"sgemm_cfg", -1)
        oracle.assert_matrix_eq(c, oracle.gemm(a, b), comp="ulp", tol=int(math.ceil(4 * math.sqrt(k))))


def test_lu_4096_pivot_sequence_matches_oracle(rla, oracle):
    """BASELINE config C3 (n = 4096): SURVEY 8d promises the reference's pivot sequence at this size (the oracle's
    unblocked elimination takes ~15 s on one host core), the reconstruction gate of the smaller sizes and solve parity."""
    n = 4096
    a = oracle.fill_uniform((n, n), 12)
    ref = oracle.lu_decompose(a, fast=True)
    f = decompose(rla, a)
    lu_checks(rla, oracle, a, f, ref)            # asserts the identical pivot sequence + |P^-1 L U - A| <= 8 n u rho max|A|
    b = np.ones(n)
    x = f.solve(rla.Vector(b)).data()
    x_ref = oracle.lu_solve(ref[0], ref[1], b, fast=True)
    eps = np.finfo(np.float64).eps
    assert np.max(np.abs(a @ x - b)) / (np.max(np.sum(np.abs(a), axis=1)) * np.max(np.abs(x)) * n * eps) <= 16
    kappa = np.linalg.cond(a, 1)
    assert np.max(np.abs(x - x_ref)) / np.max(np.abs(x_ref)) <= 8 * n * U(np.float64) * kappa


def test_lu_size_limit_is_reported_up_front(rla):
    """panels taller than the row CTAs' shared memory are refused BEFORE anything is enqueued, with the limit
    queryable (rla_lu_max_n)."""
    l = rla.lib()
    nmax = int(l.rla_lu_max_n(8))
    assert 50000 < nmax < 70000 and 2 * nmax <= int(l.rla_lu_max_n(4)) < 2 * nmax + 300
    buf = rla.DeviceBuffer(1024)
    assert l.rla_dgetrf_dev(nmax + 1, buf.ptr, nmax + 1, buf.ptr, buf.ptr, None) == 2      # RLA_ERR_INVALID, nothing touched
    assert l.rla_stream_sync(None) == 0
    buf.free()
