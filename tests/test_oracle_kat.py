"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md 8c).  CPU only.  Each test names the reference test it reproduces."""
import numpy as np
import pytest

from tests.golden import reference_kats as K


def A(x, dtype=np.float64):
    return np.array(x, dtype=dtype)


# ------------------------------------------------------------------ GEMM KATs (exact)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_matrix_mul_3x2_2x3(oracle, dtype):
    # src/matrix/mat_mul.rs:293-341
    c = oracle.gemm(A(K.GEMM_3x2_2x3["a"], dtype), A(K.GEMM_3x2_2x3["b"], dtype))
    oracle.assert_matrix_eq(c, A(K.GEMM_3x2_2x3["c"], dtype), comp="exact")


def test_mul_slice_basic(oracle):
    # src/matrix/mat_mul.rs:368-396 (row_stride 3, cols 2)
    g = K.GEMM_SLICE_BASIC
    parent = A(g["parent"])
    r0, c0 = g["start"]
    d = parent[r0:r0 + g["rows"], c0:c0 + g["cols"]]
    assert d.strides[0] == 3 * 8
    oracle.assert_matrix_eq(oracle.gemm(d, np.ones((2, 2))), g["times_ones"], comp="exact")
    oracle.assert_matrix_eq(oracle.gemm(d, d), g["times_self"], comp="exact")


def test_mul_slice_uneven_data(oracle):
    # src/matrix/mat_mul.rs:398-412
    g = K.GEMM_SLICE_UNEVEN
    d = A(g["parent"])[0:2, 0:2]
    oracle.assert_matrix_eq(oracle.gemm(d, A(g["rhs"])), A(g["c"]), comp="exact")


def test_gemm_blocked_order_vs_ikj_and_truth(oracle):
    # the kc=256 blocked order must agree with the naive i-k-j branch (mat_mul.rs:76-99)
    # exactly while k <= 256 (single block: identical per-element order) ...
    a = oracle.fill_uniform((37, 200), 12)
    b = oracle.fill_uniform((200, 53), 2049)
    oracle.assert_matrix_eq(oracle.gemm(a, b), oracle.gemm_ikj(a, b), comp="exact")
    # ... and within the Higham bound of the extended-precision truth for k > 256
    a = oracle.fill_uniform((19, 1000), 12)
    b = oracle.fill_uniform((1000, 23), 2049)
    c = oracle.gemm(a, b)
    ii, jj = np.meshgrid(np.arange(19), np.arange(23), indexing="ij")
    truth, absd = oracle.gemm_truth_samples(a, b, ii.ravel(), jj.ravel())
    u = 2.0 ** -53
    gamma = 1000 * u / (1 - 1000 * u)
    assert np.all(np.abs(c.ravel() - truth) <= gamma * absd)
    # strided/transposed operands and general alpha/beta
    c0 = oracle.fill_uniform((19, 23), 7)
    got = oracle.gemm(a, b, alpha=0.5, beta=2.0, c=c0.copy())
    np.testing.assert_allclose(got, 0.5 * (a @ b) + 2.0 * c0, rtol=1e-13)
    at = np.ascontiguousarray(a.T).T          # column-major view of a
    oracle.assert_matrix_eq(oracle.gemm(at, b), c, comp="exact")


def test_gemm_degenerate(oracle):
    assert oracle.gemm(np.zeros((0, 3)), np.zeros((3, 4))).shape == (0, 4)
    c = oracle.gemm(np.zeros((2, 0)), np.zeros((0, 3)))
    assert c.shape == (2, 3) and np.all(c == 0.0)


# ------------------------------------------------------------------ utils::dot order
def test_dot_order(oracle):
    # src/utils.rs:20-51: 8 partial sums then tail; compare to a literal Python restatement
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 16, 23, 64, 100):
        u = rng.random(n)
        v = rng.random(n)
        p = [0.0] * 8
        m = n - n % 8
        for i in range(0, m, 8):
            for q in range(8):
                p[q] = p[q] + u[i + q] * v[i + q]
        s = 0.0
        s = s + p[0] + p[4]
        s = s + p[1] + p[5]
        s = s + p[2] + p[6]
        s = s + p[3] + p[7]
        for i in range(m, n):
            s = s + u[i] * v[i]
        assert oracle.dot(u, v) == s


# ------------------------------------------------------------------ LU KATs
def reconstruct(oracle, lu, perm):
    l, u = oracle.lu_unpack(lu)
    # p.inverse() * l * u  (tests/mat/mod.rs:141-143) using the reference's own product order
    lu_prod = oracle.gemm(l, u)
    return oracle.perm_mul_matrix(oracle.perm_inverse(perm), lu_prod), l, u


def test_matrix_partial_piv_lu_exact_factors(oracle):
    # tests/mat/mod.rs:100-124
    g = K.LU_EXACT_3x3
    lu, perm = oracle.lu_decompose(A(g["a"]))
    l, u = oracle.lu_unpack(lu)
    oracle.assert_matrix_eq(l, A(g["l"]), comp="float")
    oracle.assert_matrix_eq(u, A(g["u"]), comp="float")
    oracle.assert_matrix_eq(oracle.perm_as_matrix(perm), A(g["p"]), comp="float")


@pytest.mark.parametrize("idx", range(len(K.LU_RECONSTRUCT)))
def test_partial_piv_lu_reconstruct(oracle, idx):
    # tests/mat/mod.rs:127-170, lu.rs:775-793
    a = A(K.LU_RECONSTRUCT[idx])
    lu, perm = oracle.lu_decompose(a)
    k, l, u = reconstruct(oracle, lu, perm)
    oracle.assert_matrix_eq(k, a, comp="float")
    assert oracle.is_lower_triangular(l)
    assert oracle.is_upper_triangular(u)


def test_partial_piv_lu_inverse(oracle):
    # lu.rs:796-822
    inv = oracle.lu_inverse(np.eye(3), np.arange(3, dtype=np.uintp))
    oracle.assert_matrix_eq(inv, np.eye(3), comp="float")
    g = K.LU_INVERSE_4x4
    lu, perm = oracle.lu_decompose(A(g["a"]))
    oracle.assert_matrix_eq(oracle.lu_inverse(lu, perm), A(g["inv"]), comp="float")


def test_partial_piv_lu_det(oracle):
    # lu.rs:825-845 (F13: passes with ~4 ULP, no slack) and impl_mat.rs:648-678
    assert oracle.lu_det(np.eye(3), np.arange(3, dtype=np.uintp)) == 1.0
    g = K.LU_DET_4x4
    lu, perm = oracle.lu_decompose(A(g["a"]))
    oracle.assert_matrix_eq(A([oracle.lu_det(lu, perm)]), A([g["det"]]), comp="float")
    lu, perm = oracle.lu_decompose(A(K.DET_5x5["a"]))
    oracle.assert_matrix_eq(A([oracle.lu_det(lu, perm)]), A([K.DET_5x5["det"]]), comp="float")


def test_partial_piv_lu_solve(oracle):
    # lu.rs:848-862
    g = K.LU_SOLVE_4x4
    lu, perm = oracle.lu_decompose(A(g["a"]))
    x = oracle.lu_solve(lu, perm, A(g["b"]))
    oracle.assert_matrix_eq(x, A(g["x"]), comp="ulp", tol=g["ulp_tol"])


def test_solve_laplacian(oracle):
    # tests/mat/mod.rs:4-26
    g = K.SOLVE_LAPLACIAN
    lu, perm = oracle.lu_decompose(A(g["a"]))
    x = oracle.lu_solve(lu, perm, A(g["b"]))
    oracle.assert_matrix_eq(x, A(g["x"]), comp="abs", tol=g["abs_tol"])


def test_matrix_solve_2x2_exact(oracle):
    # impl_mat.rs:681-693
    g = K.SOLVE_2x2
    lu, perm = oracle.lu_decompose(A(g["a"]))
    oracle.assert_matrix_eq(oracle.lu_solve(lu, perm, A(g["b"])), A(g["x"]), comp="exact")


def test_solve_identity_doc(oracle):
    # doc-test lu.rs:218-230
    g = K.SOLVE_IDENTITY
    lu, perm = oracle.lu_decompose(np.eye(g["n"]))
    oracle.assert_matrix_eq(oracle.lu_solve(lu, perm, A(g["b"])), A(g["b"]), comp="float")


def test_singular_is_div_by_zero(oracle):
    # lu.rs:759-772
    with pytest.raises(oracle.DivByZero):
        oracle.lu_decompose(A(K.LU_SINGULAR))
    # tests/mat/mod.rs:28-39 (1x1 zero triangular solves)
    with pytest.raises(oracle.DivByZero):
        oracle.back_substitution(A([[0.0]]), A([1.0]))
    with pytest.raises(oracle.DivByZero):
        oracle.forward_substitution(A([[0.0]]), A([1.0]))


def test_non_square_panics(oracle):
    # lu.rs:749-755
    with pytest.raises(AssertionError):
        oracle.lu_decompose(np.ones((2, 3)))


def test_lu_forward_substitution(oracle):
    # lu.rs:865-889
    for g in K.FORWARD_SUBST:
        x = oracle.lu_forward_substitution(A(g["lu"]).reshape(len(g["b"]), len(g["b"])), A(g["b"]))
        assert x.tolist() == g["x"]


def test_pivot_rule_first_max_wins(oracle):
    # lu.rs:173-178: strict '>' in ascending i => first row attaining the max is the pivot
    a = A([[1., 2., 3.], [-4., 1., 0.], [4., 0., 1.]])
    lu, perm = oracle.lu_decompose(a)
    # original row 1 (|-4|, first max) must land at position 0
    assert perm[1] == 0
    # f32 twin works and uses FLT_EPSILON
    lu32, perm32 = oracle.lu_decompose(a.astype(np.float32))
    assert perm32.tolist() == perm.tolist()
    with pytest.raises(oracle.DivByZero):
        oracle.lu_decompose(A([[1e-8, 0.], [0., 1.]], np.float32))
    oracle.lu_decompose(A([[1e-8, 0.], [0., 1.]], np.float64))


# ------------------------------------------------------------------ comparators
def test_ulp_diff_semantics(oracle):
    # src/ulp.rs:41-65 and its tests :67-198
    assert oracle.ulp_diff(0.0, -0.0) == ("exact", 0)
    assert oracle.ulp_diff(1.0, np.nextafter(1.0, 2.0)) == ("diff", 1)
    assert oracle.ulp_diff(np.nextafter(1.0, 2.0), 1.0) == ("diff", 1)
    assert oracle.ulp_diff(1.0, -1.0)[0] == "signs"
    assert oracle.ulp_diff(float("nan"), 1.0)[0] == "nan"
    assert oracle.ulp_diff(float("inf"), float("inf")) == ("exact", 0)
    assert oracle.ulp_diff(np.finfo(np.float64).max, float("inf")) == ("diff", 1)
    assert oracle.ulp_diff(1.0, float(np.nextafter(np.float32(1.0), np.float32(2.0))), np.float32) == ("diff", 1)


def test_comparators(oracle):
    # src/macros/comparison.rs:46-197
    one = A([1.0])
    eps = np.finfo(np.float64).eps
    oracle.assert_matrix_eq(one, one + eps, comp="float")                 # abs eps passes
    oracle.assert_matrix_eq(A([1e10]), A([np.nextafter(1e10, 2e10)]), comp="float")   # ulp 1 passes
    with pytest.raises(AssertionError):
        oracle.assert_matrix_eq(A([1e10]), A([1e10 * (1 + 10 * eps)]), comp="float")
    oracle.assert_matrix_eq(A([1.0]), A([1.5]), comp="abs", tol=0.5)      # inclusive
    with pytest.raises(AssertionError):
        oracle.assert_matrix_eq(A([1.0]), A([1.5]), comp="abs", tol=0.49)
    with pytest.raises(AssertionError):
        oracle.assert_matrix_eq(A([float("nan")]), A([float("nan")]), comp="exact")
    with pytest.raises(AssertionError):
        oracle.assert_matrix_eq(A([float("nan")]), A([1.0]), comp="ulp", tol=2 ** 62)
    with pytest.raises(AssertionError):
        oracle.assert_matrix_eq(np.zeros((2, 3)), np.zeros((3, 2)), comp="exact")   # dimension mismatch
    info = oracle.assert_matrix_eq(A([1.0, 2.0]), A([np.nextafter(1.0, 2), 2.0]), comp="ulp", tol=1)
    assert info["max_ulp"] == 1


# ------------------------------------------------------------------ Cholesky KATs
def test_cholesky_unpack_kats(oracle):
    # cholesky.rs:39-59 (doc), :384-415
    g = K.CHOL_DOC_3x3
    oracle.assert_matrix_eq(oracle.cholesky_unpack(oracle.cholesky_decompose(A(g["a"]))), A(g["l"]), comp="float")
    for g in K.CHOL_UNPACK:
        oracle.assert_matrix_eq(oracle.cholesky_unpack(oracle.cholesky_decompose(A(g["a"]))), A(g["l"]), comp="float")
    e = oracle.cholesky_unpack(oracle.cholesky_decompose(np.zeros((0, 0))))
    assert e.shape == (0, 0)
    # quickcheck property :541-557: cholesky(I) = I
    for n in (1, 2, 7, 30):
        oracle.assert_matrix_eq(oracle.cholesky_unpack(oracle.cholesky_decompose(np.eye(n))), np.eye(n), comp="float")


def test_cholesky_only_lower_triangle_is_touched(oracle):
    # decompose ignores the strict upper triangle (cholesky.rs:126-131): garbage there changes nothing below
    a = A(K.CHOL_DOC_3x3["a"])
    b = a.copy()
    b[np.triu_indices(3, 1)] = 1e300
    la, lb = oracle.cholesky_decompose(a), oracle.cholesky_decompose(b)
    assert np.array_equal(np.tril(la), np.tril(lb))
    assert np.array_equal(lb[np.triu_indices(3, 1)], b[np.triu_indices(3, 1)])


def test_cholesky_failures(oracle):
    # cholesky.rs:418-442 and the two messages :151-158
    for bad in K.CHOL_SINGULAR:
        with pytest.raises(oracle.DecompFailure, match="singular to working precision"):
            oracle.cholesky_decompose(A(bad))
    with pytest.raises(oracle.DecompFailure, match="not all positive"):
        oracle.cholesky_decompose(A([[1.0, 0.0], [0.0, -4.0]]))
    with pytest.raises(AssertionError):
        oracle.cholesky_decompose(np.ones((2, 3)))


def test_cholesky_det_solve_inverse_kats(oracle):
    # cholesky.rs:445-466 (det), :469-499 (solve), :502-539 (inverse), doc :65-79
    assert oracle.cholesky_det(oracle.cholesky_decompose(np.zeros((0, 0)))) == 1.0
    for g in K.CHOL_DET:
        oracle.assert_matrix_eq(np.array([oracle.cholesky_det(oracle.cholesky_decompose(A(g["a"])))]), np.array([g["det"]]), comp="float")
    for g in K.CHOL_SOLVE:
        oracle.assert_matrix_eq(oracle.cholesky_solve(oracle.cholesky_decompose(A(g["a"])), A(g["b"])), A(g["x"]), comp="float")
    g = K.CHOL_DOC_SOLVE
    l = oracle.cholesky_decompose(A(g["a"]))
    oracle.assert_matrix_eq(oracle.cholesky_solve(l, A(g["b1"])), A(g["y1"]), comp="exact")
    oracle.assert_matrix_eq(oracle.cholesky_solve(l, A(g["b2"])), A(g["y2"]), comp="exact")
    for g in K.CHOL_INVERSE:
        oracle.assert_matrix_eq(oracle.cholesky_inverse(oracle.cholesky_decompose(A(g["a"]))), A(g["inv"]), comp="float")
    assert oracle.cholesky_solve(oracle.cholesky_decompose(np.zeros((0, 0))), np.zeros(0)).size == 0


def test_transpose_back_substitution(oracle):
    # cholesky.rs:329-365: L^T x = b
    l = A([[2.0, 0.0, 0.0], [1.0, 3.0, 0.0], [4.0, 5.0, 6.0]])
    x = A([1.0, -2.0, 0.5])
    b = l.T @ x
    oracle.assert_matrix_eq(oracle.transpose_back_substitution(l, b), x, comp="float")
    with pytest.raises(oracle.DivByZero):
        oracle.transpose_back_substitution(A([[1.0, 0.0], [1.0, 0.0]]), A([1.0, 1.0]))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 5, 33, 64, 100])
def test_cholesky_oracle_properties(oracle, dtype, n):
    # not a reference test: the restatement must also BE a Cholesky factorisation (L L^T = A, agreement with LAPACK's
    # factor within rounding, solve residual), so that the GPU parity tests compare against something meaningful
    m = oracle.fill_uniform((n, n), 900 + n, np.float64, lo=-1.0, scale=2.0)
    a = (m @ m.T / n + np.eye(n)).astype(dtype)
    a = np.ascontiguousarray((a + a.T) / 2)
    u = 2.0 ** (-52 if dtype == np.float64 else -23)
    l = oracle.cholesky_unpack(oracle.cholesky_decompose(a)).astype(np.float64)
    assert np.abs(l @ l.T - a.astype(np.float64)).max() <= 8 * n * u * np.abs(a).max()
    assert np.abs(l - np.linalg.cholesky(a.astype(np.float64))).max() <= 64 * n * u * np.abs(l).max()
    b = oracle.fill_uniform((n,), 4000, dtype)
    x = oracle.cholesky_solve(oracle.cholesky_decompose(a), b).astype(np.float64)
    r = np.abs(a.astype(np.float64) @ x - b.astype(np.float64)).max()
    assert r <= 64 * n * u * np.abs(a).max() * max(np.abs(x).max(), 1.0)
