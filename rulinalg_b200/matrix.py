"""Host-side mirror of the rulinalg types on the hot path, over the C ABI.

Mirrors (file:line relative to the rulinalg tree):
  Matrix / MatrixSlice          src/matrix/mod.rs:45-64   (row-major Vec<T>; slices carry row_stride)
  &Matrix * &Matrix             src/matrix/mat_mul.rs:17-143,145-287  -> rla_dgemm / rla_sgemm
  Vector                        src/vector/mod.rs:13-16
  PermutationMatrix             src/matrix/permutation_matrix.rs:103-148,216-260,369-382
  PartialPivLu / LUP            src/matrix/decomposition/lu.rs:17-24,130-300 -> rla_dgetrf / rla_dgetrs
  Matrix::{solve,inverse,det}   src/matrix/impl_mat.rs:354-356,386-388,407-427

Same names, argument meaning and error behaviour: shape violations panic (`Panic`), numerical
failure is `Error(ErrorKind.DivByZero, <reference message>)`.  Storage is host-resident numpy, like
the reference's Vec<T>; all arithmetic on the path happens in librla_b200.so on the B200 -- there is
no CPU fallback in this module (CUDA-side failures raise `RlaError`).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .error import Error, ErrorKind, Panic

_LU_ILL_MSG = "The matrix is too ill-conditioned for\n                     LU decomposition with partial pivoting."
_TRI_SINGULAR_MSG = "Lower triangular matrix is singular to working precision."


def _dtype_pre(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "d"
    if dtype == np.float32:
        return "s"
    raise TypeError(f"the B200 path covers f32/f64 only (mat_mul.rs:27-75), got {dtype}")


class Vector:
    """rulinalg::vector::Vector<T>."""

    def __init__(self, data, dtype=None):
        self._data = np.array(data, dtype=dtype if dtype is not None else None, copy=True).reshape(-1)
        if self._data.dtype not in (np.float32, np.float64):
            self._data = self._data.astype(np.float64)

    @staticmethod
    def new(data):
        return Vector(data)

    @staticmethod
    def zeros(n, dtype=np.float64):
        return Vector(np.zeros(n, dtype=dtype))

    @staticmethod
    def ones(n, dtype=np.float64):
        return Vector(np.ones(n, dtype=dtype))

    def size(self) -> int:
        return self._data.size

    def data(self) -> np.ndarray:
        return self._data

    def __getitem__(self, i):
        return self._data[i]

    def __eq__(self, other):
        return isinstance(other, Vector) and np.array_equal(self._data, other._data)

    def __repr__(self):
        return f"Vector({self._data.tolist()})"


class _BaseMatrix:
    """The five accessors the hot path reads (BaseMatrix: rows, cols, row_stride, as_ptr; base/mod.rs:44-81)."""

    _arr: np.ndarray  # 2-D view, unit column stride

    def rows(self) -> int:
        return self._arr.shape[0]

    def cols(self) -> int:
        return self._arr.shape[1]

    def row_stride(self) -> int:
        return self._arr.shape[1]

    def as_ptr(self) -> int:
        return self._arr.ctypes.data

    def __getitem__(self, ij):
        i, j = ij
        if not (0 <= i < self.rows() and 0 <= j < self.cols()):
            raise Panic("index out of bounds")
        return self._arr[i, j]

    def into_vec(self):
        return self._arr.reshape(-1).tolist()

    def to_numpy(self) -> np.ndarray:
        return np.array(self._arr, copy=True)

    # ---- mat_mul_general! (mat_mul.rs:17-75) -------------------------------------------------
    def __mul__(self, m):
        if isinstance(m, (int, float)):
            return Matrix._from_array(self._arr * self._arr.dtype.type(m))
        if isinstance(m, Vector):
            # &Matrix * &Vector (impl_ops.rs:298-314) -> rla_?gemv
            if m.size() != self.cols():
                raise Panic("Matrix and Vector dimensions do not agree.")
            xin = np.ascontiguousarray(m.data(), dtype=self._arr.dtype)
            out = np.empty(self.rows(), dtype=self._arr.dtype)
            fnv = getattr(_lib.lib(), f"rla_{_dtype_pre(self._arr.dtype)}gemv")
            _lib.check(fnv(self.rows(), self.cols(), self.as_ptr(), self.row_stride(), xin.ctypes.data, out.ctypes.data))
            return Vector(out)
        if not isinstance(m, _BaseMatrix):
            return NotImplemented
        if self.cols() != m.rows():
            raise Panic("Matrix dimensions do not agree.")
        if self._arr.dtype != m._arr.dtype:
            raise TypeError("operand element types differ")
        p, q, r = self.rows(), self.cols(), m.cols()
        pre = _dtype_pre(self._arr.dtype)
        new_data = np.empty((p, r), dtype=self._arr.dtype)          # uninitialised, like set_len (mat_mul.rs:52-55)
        fn = getattr(_lib.lib(), f"rla_{pre}gemm")
        st = fn(p, q, r, 1.0, self.as_ptr(), self.row_stride(), 1,
                m.as_ptr(), m.row_stride(), 1, 0.0, new_data.ctypes.data, r, 1)
        _lib.check(st)
        return Matrix._from_array(new_data)

    # ---- solve_u_triangular / solve_l_triangular (base/mod.rs:1015-1067) ---------------------------
    def _tri_solve(self, y: "Vector", lower: bool) -> "Vector":
        if self.cols() != y.size():
            raise Panic(f"Vector size {y.size()} != {self.cols()} Matrix column count.")
        if self.rows() != self.cols():
            raise Panic("Matrix U must be square." if not lower else "Matrix L must be square.")
        n = self.rows()
        x = np.array(y.data(), dtype=self._arr.dtype, copy=True)
        pre = _dtype_pre(self._arr.dtype)
        st = getattr(_lib.lib(), f"rla_{pre}trsv")(1 if lower else 0, n, self.as_ptr(), self.row_stride(), x.ctypes.data)
        if _lib.check(st) == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DivByZero, _TRI_SINGULAR_MSG)
        return Vector(x)

    def solve_u_triangular(self, y: "Vector") -> "Vector":
        return self._tri_solve(y, lower=False)

    def solve_l_triangular(self, y: "Vector") -> "Vector":
        return self._tri_solve(y, lower=True)

    def sub_slice(self, start, rows, cols):
        return MatrixSlice.from_matrix(self, start, rows, cols)

    def held(self):
        """Guard (context manager) during which this matrix is promised immutable, so products that read it keep
        its device copy and skip the upload on later calls (rla_operand_hold / rla_operand_release; SURVEY 8f rank 3,
        the repeated products of lu.rs:789,907 and eigen.rs:114-148).  The Rust counterpart borrows `&Matrix<T>`."""
        return _Held(self)


class _Held:
    def __init__(self, m: "_BaseMatrix"):
        self._m = m
        self._ptr = None

    def __enter__(self):
        a = self._m._arr
        if a.size == 0:
            return self._m
        nbytes = ((a.shape[0] - 1) * (a.strides[0] // a.itemsize) + a.shape[1]) * a.itemsize
        a.flags.writeable = False                      # numpy's stand-in for the shared borrow
        self._ptr = a.ctypes.data
        _lib.check(_lib.lib().rla_operand_hold(self._ptr, nbytes))
        return self._m

    def __exit__(self, *exc):
        if self._ptr is not None:
            _lib.check(_lib.lib().rla_operand_release(self._ptr))
            try:
                self._m._arr.flags.writeable = True
            except ValueError:
                pass
            self._ptr = None
        return False


class Matrix(_BaseMatrix):
    """rulinalg::matrix::Matrix<T>: rows, cols, contiguous row-major data."""

    def __init__(self, rows, cols, data, dtype=None):
        arr = np.array(data, dtype=dtype, copy=True)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        if arr.size != rows * cols:
            raise Panic("Data does not match given dimensions.")
        self._arr = np.ascontiguousarray(arr.reshape(rows, cols))

    @staticmethod
    def new(rows, cols, data):
        return Matrix(rows, cols, data)

    @staticmethod
    def _from_array(arr: np.ndarray) -> "Matrix":
        m = Matrix.__new__(Matrix)
        m._arr = np.ascontiguousarray(arr)
        return m

    @staticmethod
    def from_numpy(arr, copy=True) -> "Matrix":
        """copy=False adopts a C-contiguous f32/f64 array as the matrix's storage (Matrix::new takes the Vec by value)."""
        arr = np.asarray(arr)
        if arr.ndim != 2:
            raise Panic("expected a 2-D array")
        if not copy and arr.flags.c_contiguous and arr.dtype in (np.float32, np.float64):
            return Matrix._from_array(arr)
        return Matrix._from_array(np.array(arr, copy=True))

    @staticmethod
    def zeros(rows, cols, dtype=np.float64):
        return Matrix._from_array(np.zeros((rows, cols), dtype=dtype))

    @staticmethod
    def ones(rows, cols, dtype=np.float64):
        return Matrix._from_array(np.ones((rows, cols), dtype=dtype))

    @staticmethod
    def identity(n, dtype=np.float64):
        return Matrix._from_array(np.eye(n, dtype=dtype))

    def data(self) -> np.ndarray:
        return self._arr.reshape(-1)

    def clone(self) -> "Matrix":
        return Matrix._from_array(self._arr.copy())

    def __eq__(self, other):
        return isinstance(other, _BaseMatrix) and self._arr.shape == other._arr.shape and np.array_equal(self._arr, other._arr)

    def __repr__(self):
        return f"Matrix({self.rows()}x{self.cols()}, {self._arr.dtype})"

    # ---- impl_mat.rs:354-356,386-388,407-427 -----------------------------------------------------
    def solve(self, y: Vector) -> Vector:
        return PartialPivLu.decompose(self.clone()).solve(y)

    def inverse(self) -> "Matrix":
        if self.rows() != self.cols():
            raise Panic("Matrix is not square.")
        return PartialPivLu.decompose(self.clone()).inverse()

    def det(self):
        if self.rows() != self.cols():
            raise Panic("Matrix is not square.")
        n = self.cols()
        a = self._arr
        if n == 0:
            return a.dtype.type(1)
        if np.count_nonzero(a - np.diag(np.diagonal(a))) == 0:      # is_diag
            out = a.dtype.type(1)
            for v in np.diagonal(a):
                out = out * v
            return out
        if n == 2:
            return a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]
        if n == 3:
            return ((a[0, 0] * a[1, 1] * a[2, 2]) + (a[0, 1] * a[1, 2] * a[2, 0]) + (a[0, 2] * a[1, 0] * a[2, 1])
                    - (a[0, 0] * a[1, 2] * a[2, 1]) - (a[0, 1] * a[1, 0] * a[2, 2]) - (a[0, 2] * a[1, 1] * a[2, 0]))
        try:
            return PartialPivLu.decompose(self.clone()).det()
        except Error:
            return a.dtype.type(0)          # "singular => det 0" (impl_mat.rs:421-426)


class MatrixSlice(_BaseMatrix):
    """rulinalg::matrix::MatrixSlice<T>: a strided window (ptr, rows, cols, row_stride)."""

    @staticmethod
    def from_matrix(mat: _BaseMatrix, start, rows, cols) -> "MatrixSlice":
        r0, c0 = start
        if r0 + rows > mat.rows():
            raise Panic("View dimensions exceed matrix dimensions.")
        if c0 + cols > mat.cols():
            raise Panic("View dimensions exceed matrix dimensions.")
        s = MatrixSlice.__new__(MatrixSlice)
        s._arr = mat._arr[r0:r0 + rows, c0:c0 + cols]
        s._parent_stride = mat._arr.strides[0] // mat._arr.itemsize
        return s

    def row_stride(self) -> int:
        return self._parent_stride

    def __repr__(self):
        return f"MatrixSlice({self.rows()}x{self.cols()}, row_stride={self.row_stride()})"


MatrixSliceMut = MatrixSlice


class PermutationMatrix:
    """rulinalg::matrix::PermutationMatrix<T> (perm: Vec<usize>)."""

    def __init__(self, perm):
        self._perm = np.array(perm, dtype=np.uintp, copy=True)

    @staticmethod
    def identity(n):
        return PermutationMatrix(np.arange(n, dtype=np.uintp))

    def size(self) -> int:
        return self._perm.size

    def perm(self) -> np.ndarray:
        return self._perm

    def swap_rows(self, i, j):
        self._perm[[i, j]] = self._perm[[j, i]]

    def inverse(self) -> "PermutationMatrix":
        inv = np.zeros_like(self._perm)
        inv[self._perm.astype(np.int64)] = np.arange(self._perm.size, dtype=np.uintp)
        return PermutationMatrix(inv)

    def as_matrix(self, dtype=np.float64) -> Matrix:
        n = self.size()
        m = np.zeros((n, n), dtype=dtype)
        m[np.arange(n), self._perm.astype(np.int64)] = 1
        return Matrix._from_array(m)

    def parity_sign(self) -> int:
        p = self._perm.astype(np.int64).copy()
        sign = 1
        for i in range(p.size):
            while p[i] != i:
                t = p[i]
                p[i], p[t] = p[t], t
                sign = -sign
        return sign

    def det(self, dtype=np.float64):
        return np.dtype(dtype).type(1) if self.parity_sign() > 0 else np.dtype(dtype).type(0) - np.dtype(dtype).type(1)

    def __mul__(self, rhs):
        if isinstance(rhs, Vector):                      # impl_permutation_mul.rs:21-41
            if rhs.size() != self.size():
                raise Panic("Permutation matrix and Vector dimensions are not compatible.")
            out = np.empty_like(rhs.data())
            out[self._perm.astype(np.int64)] = rhs.data()
            return Vector(out)
        if isinstance(rhs, _BaseMatrix):                 # permute_rows_into_buffer (:318-331)
            if rhs.rows() != self.size():
                raise Panic("Permutation matrix and right-hand side matrix dimensions are not compatible.")
            out = np.empty_like(rhs._arr)
            out[self._perm.astype(np.int64)] = rhs._arr
            return Matrix._from_array(out)
        return NotImplemented


class LUP:
    """Result of PartialPivLu::unpack (lu.rs:17-24)."""

    def __init__(self, l: Matrix, u: Matrix, p: PermutationMatrix):
        self.l, self.u, self.p = l, u, p


class PartialPivLu:
    """rulinalg::matrix::decomposition::PartialPivLu<T> over librla_b200 (lu.rs:130-300)."""

    # factors of at most this many bytes also stay resident in HBM for repeated solves (rla_dgetrf_keep); larger ones are
    # re-uploaded by solve() -- an object that pins gigabytes of HBM until Python's GC runs is a bad default.  release() /
    # the context-manager protocol return the device copy early.
    KEEP_RESIDENT_BYTES = 512 << 20

    def __init__(self, lu: Matrix, p: PermutationMatrix, handle=None):
        self.lu = lu
        self.p = p
        self._handle = handle        # device-resident copy of the factors for repeated solves (f64)

    @staticmethod
    def decompose(matrix: Matrix, keep_resident: bool | None = None) -> "PartialPivLu":
        n = matrix.cols()
        if matrix.rows() != n:
            raise Panic("Matrix must be square for LU decomposition.")
        lu = matrix                                          # moved in, factorised in place (lu.rs:166)
        pre = _dtype_pre(lu._arr.dtype)
        perm = np.zeros(n, dtype=np.uintp)
        handle = None
        if keep_resident is None:
            keep_resident = n * n * 8 <= PartialPivLu.KEEP_RESIDENT_BYTES
        if pre == "d" and keep_resident:
            h = C.c_void_p()
            st = _lib.lib().rla_dgetrf_keep(n, lu.as_ptr(), perm.ctypes.data, C.byref(h))
            handle = h
        else:
            st = getattr(_lib.lib(), f"rla_{pre}getrf")(n, lu.as_ptr(), perm.ctypes.data)
        if _lib.check(st) == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DivByZero, _LU_ILL_MSG)
        return PartialPivLu(lu, PermutationMatrix(perm), handle)

    def release(self) -> None:
        """Return the device-resident copy of the factors (solve() then re-uploads `lu` / `p`)."""
        self.__del__()

    def held(self):
        """`with lu.held(): ...` -- the factors are promised immutable inside the block (rla_operand_hold), so solve() and
        inverse() find them resident in HBM after the first call: f32 as well, and without the f64 handle.  The Rust and C++
        mirrors hold for the whole lifetime of the struct (its `lu` is private there); here the fields are public, so it is
        explicit."""
        return self.lu.held()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()
        return False

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                _lib.lib().rla_lu_free(h)
            except Exception:
                pass
            self._handle = None

    def solve(self, b: Vector) -> Vector:
        n = self.lu.rows()
        if b.size() != n:
            raise Panic("Right-hand side vector must have compatible size.")
        x = np.array(b.data(), dtype=self.lu._arr.dtype, copy=True)
        pre = _dtype_pre(self.lu._arr.dtype)
        l = _lib.lib()
        if self._handle is not None:
            st = l.rla_dlu_solve(self._handle, x.ctypes.data)
        else:
            st = getattr(l, f"rla_{pre}getrs")(n, self.lu.as_ptr(), self.p.perm().ctypes.data, x.ctypes.data)
        if _lib.check(st) == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DivByZero, _TRI_SINGULAR_MSG)
        return Vector(x)

    def inverse(self) -> Matrix:
        # lu.rs:251-285 (n solves of unit vectors) as ONE blocked multi-RHS solve on the device (rla_?getri);
        # n <= 64 keeps the reference's exact per-column order
        n = self.lu.rows()
        dt = self.lu._arr.dtype
        inv = np.empty((n, n), dtype=dt)
        pre = _dtype_pre(dt)
        st = getattr(_lib.lib(), f"rla_{pre}getri")(n, self.lu.as_ptr(), self.p.perm().ctypes.data, inv.ctypes.data)
        if _lib.check(st) == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DivByZero, _TRI_SINGULAR_MSG)
        return Matrix._from_array(inv)

    def det(self):
        # lu.rs:291-300: fold over the diagonal, times p.det()
        dt = self.lu._arr.dtype.type
        u_det = dt(1)
        for v in np.diagonal(self.lu._arr):
            u_det = u_det * v
        return self.p.det(self.lu._arr.dtype) * u_det

    def unpack(self) -> LUP:
        # lu.rs:138-149, :644-663; internal_utils.rs:5-13
        a = self.lu._arr
        l = np.tril(a, -1) + np.eye(a.shape[0], dtype=a.dtype)
        u = np.triu(a)
        return LUP(Matrix._from_array(l), Matrix._from_array(u), self.p)


class Cholesky:
    """rulinalg::matrix::decomposition::Cholesky<T> over librla_b200 (cholesky.rs:94-245)."""

    _SINGULAR_MSG = "Matrix is singular to working precision."
    _NEGATIVE_MSG = "Diagonal entries of matrix are not all positive."
    _L_SINGULAR_MSG = "Matrix L is singular to working precision."

    def __init__(self, l: Matrix):
        self.l = l                      # packed: lower triangle = L (strict upper triangle unspecified)

    @staticmethod
    def decompose(matrix: Matrix) -> "Cholesky":
        n = matrix.cols()
        if matrix.rows() != n:
            raise Panic("Matrix must be square for Cholesky decomposition.")
        a = matrix                                           # moved in, factorised in place (cholesky.rs:132)
        pre = _dtype_pre(a._arr.dtype)
        st = _lib.check(getattr(_lib.lib(), f"rla_{pre}potrf")(n, a.as_ptr()))
        if st == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DecompFailure, Cholesky._SINGULAR_MSG)
        if st == _lib.RLA_ERR_NOT_POSITIVE:
            raise Error(ErrorKind.DecompFailure, Cholesky._NEGATIVE_MSG)
        return Cholesky(a)

    def det(self):
        # cholesky.rs:175-180: fold over the diagonal of L, squared
        dt = self.l._arr.dtype.type
        l_det = dt(1)
        for v in np.diagonal(self.l._arr):
            l_det = l_det * v
        return l_det * l_det

    def solve(self, b: Vector) -> Vector:
        n = self.l.rows()
        if b.size() != n:
            raise Panic("RHS vector and coefficient matrix must be dimensionally compatible.")
        x = np.array(b.data(), dtype=self.l._arr.dtype, copy=True)
        pre = _dtype_pre(self.l._arr.dtype)
        st = _lib.check(getattr(_lib.lib(), f"rla_{pre}potrs")(n, self.l.as_ptr(), x.ctypes.data))
        if st == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DivByZero, Cholesky._L_SINGULAR_MSG)
        return Vector(x)

    def inverse(self) -> Matrix:
        # cholesky.rs:209-233 (n solves of unit vectors) as one blocked multi-RHS solve; n <= 64 keeps the exact order
        n = self.l.rows()
        dt = self.l._arr.dtype
        inv = np.empty((n, n), dtype=dt)
        pre = _dtype_pre(dt)
        st = _lib.check(getattr(_lib.lib(), f"rla_{pre}potri")(n, self.l.as_ptr(), inv.ctypes.data))
        if st == _lib.RLA_ERR_SINGULAR:
            raise Error(ErrorKind.DivByZero, Cholesky._L_SINGULAR_MSG)
        return Matrix._from_array(inv)

    def unpack(self) -> Matrix:
        # cholesky.rs:237-245: nullify_upper_triangular_part
        return Matrix._from_array(np.tril(self.l._arr))
