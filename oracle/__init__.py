"""ctypes/numpy front-end for the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY.  Import sites allowed: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / ``--impl reference`` legs.  Nothing under rulinalg_b200/ may
import this module; the product path fails loudly without its CUDA library instead.

Pinning status (see oracle.c header and DESIGN.md): LU/solve/det/inverse/comparators are
pinned by the reference's own known-answer tests (tests/test_oracle_kat.py); GEMM rounding
order is "parity unpinned" (third-party matrixmultiply 0.1.x, not vendored, no rustc).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ORC_OK = 0
ORC_ERR_SINGULAR = 1


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_fast.so with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "oracle.c")
    need = force
    for name in ("liboracle.so", "liboracle_fast.so"):
        so = os.path.join(_HERE, name)
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            need = True
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])


_libs: dict = {}


def _lib(fast: bool = False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    sz, pd, dbl, flt, u64 = C.c_size_t, C.c_ssize_t, C.c_double, C.c_float, C.c_uint64
    P = C.c_void_p
    lib.orc_fill_f64.argtypes = [P, sz, u64, u64, dbl, dbl]
    lib.orc_fill_f32.argtypes = [P, sz, u64, u64, flt, flt]
    lib.orc_dgemm.argtypes = [sz, sz, sz, dbl, P, pd, pd, P, pd, pd, dbl, P, pd, pd]
    lib.orc_sgemm.argtypes = [sz, sz, sz, flt, P, pd, pd, P, pd, pd, flt, P, pd, pd]
    lib.orc_dgemm_ikj.argtypes = [sz, sz, sz, P, sz, P, sz, P]
    lib.orc_ddot_ld.argtypes = [sz, P, pd, P, pd]
    lib.orc_ddot_ld.restype = dbl
    lib.orc_sdot_d.argtypes = [sz, P, pd, P, pd]
    lib.orc_sdot_d.restype = dbl
    lib.orc_dabsdot.argtypes = [sz, P, pd, P, pd]
    lib.orc_dabsdot.restype = dbl
    lib.orc_ddot.argtypes = [P, P, sz]
    lib.orc_ddot.restype = dbl
    lib.orc_sdot.argtypes = [P, P, sz]
    lib.orc_sdot.restype = flt
    for pre in ("d", "s"):
        getattr(lib, f"orc_{pre}getrf").argtypes = [sz, P, P]
        getattr(lib, f"orc_{pre}getrs").argtypes = [sz, P, P, P]
        getattr(lib, f"orc_{pre}lu_forward_substitution").argtypes = [sz, P, P]
        getattr(lib, f"orc_{pre}lu_forward_substitution").restype = None
        getattr(lib, f"orc_{pre}back_substitution").argtypes = [sz, P, sz, P]
        getattr(lib, f"orc_{pre}forward_substitution").argtypes = [sz, P, sz, P]
    for pre in ("d", "s"):
        getattr(lib, f"orc_{pre}potrf").argtypes = [sz, P]
        getattr(lib, f"orc_{pre}potrf").restype = C.c_long
        getattr(lib, f"orc_{pre}potrs").argtypes = [sz, P, P]
        getattr(lib, f"orc_{pre}transpose_back_substitution").argtypes = [sz, P, sz, P]
    lib.orc_dpotri.argtypes = [sz, P, P]
    lib.orc_dpotrf_det.argtypes = [sz, P]
    lib.orc_dpotrf_det.restype = dbl
    lib.orc_dgetri.argtypes = [sz, P, P, P]
    lib.orc_perm_sign.argtypes = [sz, P]
    lib.orc_ddet.argtypes = [sz, P, P]
    lib.orc_ddet.restype = dbl
    lib.orc_dunpack.argtypes = [sz, P, P, P]
    lib.orc_dunpack.restype = None
    lib.orc_ulp_diff_f64.argtypes = [dbl, dbl, P]
    lib.orc_ulp_diff_f32.argtypes = [flt, flt, P]
    for suf, t in (("f64", dbl), ("f32", flt)):
        getattr(lib, f"orc_cmp_exact_{suf}").argtypes = [sz, P, P, P]
        getattr(lib, f"orc_cmp_exact_{suf}").restype = sz
        getattr(lib, f"orc_cmp_abs_{suf}").argtypes = [sz, P, P, t, P, P]
        getattr(lib, f"orc_cmp_abs_{suf}").restype = sz
        getattr(lib, f"orc_cmp_ulp_{suf}").argtypes = [sz, P, P, u64, P, P]
        getattr(lib, f"orc_cmp_ulp_{suf}").restype = sz
        getattr(lib, f"orc_cmp_float_{suf}").argtypes = [sz, P, P, t, u64, P]
        getattr(lib, f"orc_cmp_float_{suf}").restype = sz
    lib.orc_is_lower_triangular.argtypes = [sz, sz, P]
    lib.orc_is_upper_triangular.argtypes = [sz, sz, P]
    _libs[name] = lib
    return lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _pre(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "d"
    if dtype == np.float32:
        return "s"
    raise TypeError(f"oracle handles f32/f64 only, got {dtype}")


# ----------------------------------------------------------------------------- generators
def fill_uniform(shape, seed: int, dtype=np.float64, lo: float = 0.0, scale: float = 1.0,
                 offset: int = 0) -> np.ndarray:
    """Seeded U[lo, lo+scale) matrix; bit-identical to rulinalg_b200's device generator."""
    out = np.empty(shape, dtype=dtype)
    lib = _lib()
    if out.dtype == np.float64:
        lib.orc_fill_f64(_p(out), out.size, seed, offset, lo, scale)
    else:
        lib.orc_fill_f32(_p(out), out.size, seed, offset, lo, scale)
    return out


# ----------------------------------------------------------------------------- GEMM
def gemm(a: np.ndarray, b: np.ndarray, alpha=1.0, beta=0.0, c: np.ndarray | None = None,
         fast: bool = False) -> np.ndarray:
    """C = alpha*A*B + beta*C with matrixmultiply-0.1.x summation order.  Any 2-D strides."""
    assert a.ndim == 2 and b.ndim == 2 and a.shape[1] == b.shape[0], "Matrix dimensions do not agree."
    assert a.dtype == b.dtype
    m, k = a.shape
    n = b.shape[1]
    if c is None:
        c = np.empty((m, n), dtype=a.dtype)
    it = a.dtype.itemsize
    fn = getattr(_lib(fast), f"orc_{_pre(a.dtype)}gemm")
    fn(m, k, n, alpha, _p(a), a.strides[0] // it, a.strides[1] // it,
       _p(b), b.strides[0] // it, b.strides[1] // it,
       beta, _p(c), c.strides[0] // it, c.strides[1] // it)
    return c


def gemm_ikj(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    c = np.empty((a.shape[0], b.shape[1]))
    _lib().orc_dgemm_ikj(a.shape[0], a.shape[1], b.shape[1], _p(a), a.shape[1], _p(b), b.shape[1], _p(c))
    return c


def gemm_truth_samples(a: np.ndarray, b: np.ndarray, idx_i, idx_j):
    """(truth, absdot) at sampled (i,j): extended-precision dot and sum|a||b| (Higham bound)."""
    lib = _lib()
    it = a.dtype.itemsize
    k = a.shape[1]
    truth = np.empty(len(idx_i))
    absd = np.empty(len(idx_i))
    for t, (i, j) in enumerate(zip(idx_i, idx_j)):
        pa = C.c_void_p(a.ctypes.data + int(i) * a.strides[0])
        pb = C.c_void_p(b.ctypes.data + int(j) * b.strides[1])
        if a.dtype == np.float64:
            truth[t] = lib.orc_ddot_ld(k, pa, a.strides[1] // it, pb, b.strides[0] // it)
            absd[t] = lib.orc_dabsdot(k, pa, a.strides[1] // it, pb, b.strides[0] // it)
        else:
            truth[t] = lib.orc_sdot_d(k, pa, a.strides[1] // it, pb, b.strides[0] // it)
            absd[t] = float(np.abs(a[int(i), :].astype(np.float64)) @ np.abs(b[:, int(j)].astype(np.float64)))
    return truth, absd


def dot(u: np.ndarray, v: np.ndarray):
    n = min(u.size, v.size)
    u = np.ascontiguousarray(u[:n])
    v = np.ascontiguousarray(v[:n])
    return getattr(_lib(), f"orc_{_pre(u.dtype)}dot")(_p(u), _p(v), n)


# ----------------------------------------------------------------------------- LU / solve
class DivByZero(Exception):
    """Mirror of ErrorKind::DivByZero (src/error.rs:10-34)."""


def lu_decompose(a: np.ndarray, fast: bool = False):
    """PartialPivLu::decompose -> (lu, perm) with perm = p.inverse().perm; raises DivByZero."""
    assert a.ndim == 2 and a.shape[0] == a.shape[1], "Matrix must be square for LU decomposition."
    lu = np.array(a, order="C", copy=True)
    n = lu.shape[0]
    perm = np.zeros(n, dtype=np.uintp)
    rc = getattr(_lib(fast), f"orc_{_pre(lu.dtype)}getrf")(n, _p(lu), _p(perm))
    if rc != ORC_OK:
        raise DivByZero("The matrix is too ill-conditioned for LU decomposition with partial pivoting.")
    return lu, perm


def lu_solve(lu: np.ndarray, perm: np.ndarray, b: np.ndarray, fast: bool = False) -> np.ndarray:
    n = lu.shape[0]
    assert b.size == n, "Right-hand side vector must have compatible size."
    lu = np.ascontiguousarray(lu)
    perm = np.ascontiguousarray(perm, dtype=np.uintp)
    x = np.array(b, dtype=lu.dtype, copy=True).reshape(n)
    rc = getattr(_lib(fast), f"orc_{_pre(lu.dtype)}getrs")(n, _p(lu), _p(perm), _p(x))
    if rc != ORC_OK:
        raise DivByZero("Lower triangular matrix is singular to working precision.")
    return x


def lu_forward_substitution(lu: np.ndarray, b: np.ndarray) -> np.ndarray:
    lu = np.ascontiguousarray(lu)
    x = np.array(b, dtype=lu.dtype, copy=True)
    getattr(_lib(), f"orc_{_pre(lu.dtype)}lu_forward_substitution")(lu.shape[0], _p(lu), _p(x))
    return x


def back_substitution(u: np.ndarray, y: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(u)
    x = np.array(y, dtype=u.dtype, copy=True)
    rc = getattr(_lib(), f"orc_{_pre(u.dtype)}back_substitution")(u.shape[0], _p(u), u.shape[1], _p(x))
    if rc != ORC_OK:
        raise DivByZero("Lower triangular matrix is singular to working precision.")
    return x


def forward_substitution(l: np.ndarray, y: np.ndarray) -> np.ndarray:
    l = np.ascontiguousarray(l)
    x = np.array(y, dtype=l.dtype, copy=True)
    rc = getattr(_lib(), f"orc_{_pre(l.dtype)}forward_substitution")(l.shape[0], _p(l), l.shape[1], _p(x))
    if rc != ORC_OK:
        raise DivByZero("Lower triangular matrix is singular to working precision.")
    return x


def lu_inverse(lu: np.ndarray, perm: np.ndarray) -> np.ndarray:
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    perm = np.ascontiguousarray(perm, dtype=np.uintp)
    n = lu.shape[0]
    inv = np.zeros((n, n))
    rc = _lib().orc_dgetri(n, _p(lu), _p(perm), _p(inv))
    if rc != ORC_OK:
        raise DivByZero("Lower triangular matrix is singular to working precision.")
    return inv


def lu_det(lu: np.ndarray, perm: np.ndarray) -> float:
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    perm = np.ascontiguousarray(perm, dtype=np.uintp)
    return _lib().orc_ddet(lu.shape[0], _p(lu), _p(perm))


def lu_unpack(lu: np.ndarray):
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    n = lu.shape[0]
    l = np.empty((n, n))
    u = np.empty((n, n))
    _lib().orc_dunpack(n, _p(lu), _p(l), _p(u))
    return l, u


# ----------------------------------------------------------------------------- Cholesky
class DecompFailure(Exception):
    """Mirror of ErrorKind::DecompFailure (src/error.rs)."""


def cholesky_decompose(a: np.ndarray, fast: bool = False) -> np.ndarray:
    """Cholesky::decompose (cholesky.rs:116-170): returns the packed factor (lower triangle = L, strict upper triangle =
    the input's, untouched); raises DecompFailure with the reference's two messages."""
    assert a.ndim == 2 and a.shape[0] == a.shape[1], "Matrix must be square for Cholesky decomposition."
    l = np.array(a, order="C", copy=True)
    rc = getattr(_lib(fast), f"orc_{_pre(l.dtype)}potrf")(l.shape[0], _p(l))
    if rc > 0:
        raise DecompFailure("Matrix is singular to working precision.")
    if rc < 0:
        raise DecompFailure("Diagonal entries of matrix are not all positive.")
    return l


def cholesky_unpack(l: np.ndarray) -> np.ndarray:
    """Decomposition::unpack (cholesky.rs:237-245): zero the strict upper triangle."""
    return np.tril(l)


def cholesky_solve(l: np.ndarray, b: np.ndarray, fast: bool = False) -> np.ndarray:
    l = np.ascontiguousarray(l)
    n = l.shape[0]
    assert b.size == n, "RHS vector and coefficient matrix must be dimensionally compatible."
    x = np.array(b, dtype=l.dtype, copy=True).reshape(n)
    rc = getattr(_lib(fast), f"orc_{_pre(l.dtype)}potrs")(n, _p(l), _p(x))
    if rc != ORC_OK:
        raise DivByZero("Matrix L is singular to working precision.")
    return x


def transpose_back_substitution(l: np.ndarray, y: np.ndarray) -> np.ndarray:
    l = np.ascontiguousarray(l)
    x = np.array(y, dtype=l.dtype, copy=True)
    rc = getattr(_lib(), f"orc_{_pre(l.dtype)}transpose_back_substitution")(l.shape[0], _p(l), l.shape[1], _p(x))
    if rc != ORC_OK:
        raise DivByZero("Matrix L is singular to working precision.")
    return x


def cholesky_inverse(l: np.ndarray) -> np.ndarray:
    l = np.ascontiguousarray(l, dtype=np.float64)
    n = l.shape[0]
    inv = np.zeros((n, n))
    rc = _lib().orc_dpotri(n, _p(l), _p(inv))
    if rc != ORC_OK:
        raise DivByZero("Matrix L is singular to working precision.")
    return inv


def cholesky_det(l: np.ndarray) -> float:
    l = np.ascontiguousarray(l, dtype=np.float64)
    return _lib().orc_dpotrf_det(l.shape[0], _p(l))


def perm_as_matrix(perm: np.ndarray) -> np.ndarray:
    """PermutationMatrix::as_matrix (permutation_matrix.rs:241-249): M[i, perm[i]] = 1."""
    n = len(perm)
    m = np.zeros((n, n))
    m[np.arange(n), np.asarray(perm, dtype=np.int64)] = 1.0
    return m


def perm_inverse(perm: np.ndarray) -> np.ndarray:
    inv = np.zeros(len(perm), dtype=np.uintp)
    inv[np.asarray(perm, dtype=np.int64)] = np.arange(len(perm), dtype=np.uintp)
    return inv


def perm_mul_matrix(perm: np.ndarray, a: np.ndarray) -> np.ndarray:
    """P * A (permute_rows_into_buffer, permutation_matrix.rs:318-331): out[perm[i]] = a[i]."""
    out = np.empty_like(a)
    out[np.asarray(perm, dtype=np.int64)] = a
    return out


# ----------------------------------------------------------------------------- comparators
def ulp_diff(a, b, dtype=np.float64):
    d = C.c_uint64(0)
    if np.dtype(dtype) == np.float64:
        code = _lib().orc_ulp_diff_f64(float(a), float(b), C.byref(d))
    else:
        code = _lib().orc_ulp_diff_f32(float(a), float(b), C.byref(d))
    return ("exact", "diff", "signs", "nan")[code], d.value


def _cmp_prep(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b, dtype=a.dtype)
    if a.shape != b.shape:
        raise AssertionError(f"Dimension mismatch: {a.shape} vs {b.shape}")
    suf = "f64" if a.dtype == np.float64 else "f32"
    return a, b, suf


class MatrixEqFailure(AssertionError):
    pass


def assert_matrix_eq(x, y, comp: str = "float", tol=None, eps=None, ulp: int | None = None):
    """assert_matrix_eq!(x, y, comp = exact|abs|ulp|float, ...) (assert_matrix_eq.rs:329-406).

    Returns a dict of measured extremes (max_ulp / max_abs where the comparator computes them).
    """
    a, b, suf = _cmp_prep(x, y)
    lib = _lib()
    first = C.c_size_t(0)
    info: dict = {}
    n = a.size
    if comp == "exact":
        bad = getattr(lib, f"orc_cmp_exact_{suf}")(n, _p(a), _p(b), C.byref(first))
    elif comp == "abs":
        mx = C.c_double(0)
        bad = getattr(lib, f"orc_cmp_abs_{suf}")(n, _p(a), _p(b), tol, C.byref(first), C.byref(mx))
        info["max_abs"] = mx.value
    elif comp == "ulp":
        mu = C.c_uint64(0)
        bad = getattr(lib, f"orc_cmp_ulp_{suf}")(n, _p(a), _p(b), int(tol), C.byref(first), C.byref(mu))
        info["max_ulp"] = mu.value
    elif comp == "float":
        e = float(np.finfo(a.dtype).eps) if eps is None else eps
        u = 4 if ulp is None else ulp
        bad = getattr(lib, f"orc_cmp_float_{suf}")(n, _p(a), _p(b), e, u, C.byref(first))
    else:
        raise ValueError(comp)
    if bad:
        i = first.value
        raise MatrixEqFailure(
            f"Matrices X and Y have {bad} mismatched element pairs (comp={comp}); first at flat index {i}: "
            f"x={a.reshape(-1)[i]!r} y={b.reshape(-1)[i]!r} {info}")
    return info


def max_ulp(x, y) -> int:
    a, b, suf = _cmp_prep(x, y)
    first = C.c_size_t(0)
    mu = C.c_uint64(0)
    getattr(_lib(), f"orc_cmp_ulp_{suf}")(a.size, _p(a), _p(b), 2 ** 63, C.byref(first), C.byref(mu))
    return mu.value


def is_lower_triangular(m: np.ndarray) -> bool:
    m = np.ascontiguousarray(m, dtype=np.float64)
    return bool(_lib().orc_is_lower_triangular(m.shape[0], m.shape[1], _p(m)))


def is_upper_triangular(m: np.ndarray) -> bool:
    m = np.ascontiguousarray(m, dtype=np.float64)
    return bool(_lib().orc_is_upper_triangular(m.shape[0], m.shape[1], _p(m)))
