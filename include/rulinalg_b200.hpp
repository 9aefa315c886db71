// rulinalg_b200.hpp -- header-only C++ mirror of the rulinalg types on the hot path, over the C ABI
// (rla_b200.h).  Same names, argument meaning and error behaviour as the reference:
//   Matrix<T> / MatrixSlice<T>   src/matrix/mod.rs:45-64            (row-major, slices carry row_stride)
//   operator*                    src/matrix/mat_mul.rs:17-143       -> rla_dgemm / rla_sgemm
//   Vector<T>                    src/vector/mod.rs:13-16
//   PermutationMatrix            src/matrix/permutation_matrix.rs:103-148,369-382
//   PartialPivLu<T>              src/matrix/decomposition/lu.rs:130-300 -> rla_?getrf / rla_?getrs
//   Error / ErrorKind            src/error.rs:10-63
// Rust `assert!`/`panic!` -> rla::Panic (std::logic_error); `Result::Err` -> rla::Error (thrown);
// CUDA/environment failures -> rla::RlaFailure.  There is no CPU fallback.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "rla_b200.h"

namespace rla {

enum class ErrorKind { InvalidArg, DecompFailure, AlgebraFailure, DivByZero, ScalarConversionFailure, InvalidPermutation };

class Error : public std::runtime_error {
  public:
    Error(ErrorKind k, const std::string &m) : std::runtime_error(m), kind_(k) {}
    ErrorKind kind() const { return kind_; }
  private:
    ErrorKind kind_;
};
struct Panic : std::logic_error { using std::logic_error::logic_error; };
struct RlaFailure : std::runtime_error {
    int status;
    explicit RlaFailure(int s) : std::runtime_error(std::string("librla_b200: ") + rla_strerror(s)), status(s) {}
};

namespace detail {
inline int check(int st) {
    if (st == RLA_OK || st == RLA_ERR_SINGULAR || st == RLA_ERR_NOT_POSITIVE) return st;
    throw RlaFailure(st);
}
}  // namespace detail

// Multi-GPU opt-in (rla_set_devices): from here on `a * b` and `PartialPivLu<T>::decompose` shard large problems over
// GPUs 0..n-1 of the box from this one process (row panels / 1D block-cyclic, peer-memory exchange); results are
// bit-identical to one GPU.  use_gpus(1) restores single-GPU behaviour; shutdown() returns every library resource.
inline void use_gpus(int n) { detail::check(rla_set_devices(n)); }
inline int gpus_in_use() { return rla_get_devices(); }
inline void shutdown() { detail::check(rla_shutdown()); }

namespace detail {
template <typename T> struct Abi;
template <> struct Abi<double> {
    static int gemm(size_t m, size_t k, size_t n, const double *a, ptrdiff_t rsa, const double *b, ptrdiff_t rsb, double *c) {
        return rla_dgemm(m, k, n, 1.0, a, rsa, 1, b, rsb, 1, 0.0, c, ptrdiff_t(n), 1);
    }
    static int getrf(size_t n, double *lu, size_t *p) { return rla_dgetrf(n, lu, p); }
    static int getrs(size_t n, const double *lu, const size_t *p, double *b) { return rla_dgetrs(n, lu, p, b); }
    static int potrf(size_t n, double *a) { return rla_dpotrf(n, a); }
    static int potrs(size_t n, const double *l, double *b) { return rla_dpotrs(n, l, b); }
    static int potri(size_t n, const double *l, double *inv) { return rla_dpotri(n, l, inv); }
};
template <> struct Abi<float> {
    static int gemm(size_t m, size_t k, size_t n, const float *a, ptrdiff_t rsa, const float *b, ptrdiff_t rsb, float *c) {
        return rla_sgemm(m, k, n, 1.0f, a, rsa, 1, b, rsb, 1, 0.0f, c, ptrdiff_t(n), 1);
    }
    static int getrf(size_t n, float *lu, size_t *p) { return rla_sgetrf(n, lu, p); }
    static int getrs(size_t n, const float *lu, const size_t *p, float *b) { return rla_sgetrs(n, lu, p, b); }
    static int potrf(size_t n, float *a) { return rla_spotrf(n, a); }
    static int potrs(size_t n, const float *l, float *b) { return rla_spotrs(n, l, b); }
    static int potri(size_t n, const float *l, float *inv) { return rla_spotri(n, l, inv); }
};
}  // namespace detail

template <typename T>
class Vector {
  public:
    Vector() = default;
    explicit Vector(std::vector<T> d) : data_(std::move(d)) {}
    static Vector zeros(size_t n) { return Vector(std::vector<T>(n, T(0))); }
    static Vector ones(size_t n) { return Vector(std::vector<T>(n, T(1))); }
    size_t size() const { return data_.size(); }
    const std::vector<T> &data() const { return data_; }
    std::vector<T> &mut_data() { return data_; }
    T operator[](size_t i) const { return data_.at(i); }
  private:
    std::vector<T> data_;
};

template <typename T> class Matrix;

// BaseMatrix: the accessors the hot path reads (base/mod.rs:44-81)
template <typename T>
class MatrixSlice {
  public:
    MatrixSlice(const T *ptr, size_t rows, size_t cols, size_t row_stride) : ptr_(ptr), rows_(rows), cols_(cols), rs_(row_stride) {}
    static MatrixSlice from_matrix(const Matrix<T> &m, size_t r0, size_t c0, size_t rows, size_t cols);
    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    size_t row_stride() const { return rs_; }
    const T *as_ptr() const { return ptr_; }
  private:
    const T *ptr_;
    size_t rows_, cols_, rs_;
};

template <typename T>
class Matrix {
  public:
    Matrix() : rows_(0), cols_(0) {}
    Matrix(size_t rows, size_t cols, std::vector<T> data) : rows_(rows), cols_(cols), data_(std::move(data)) {
        if (data_.size() != rows * cols) throw Panic("Data does not match given dimensions.");
    }
    static Matrix zeros(size_t r, size_t c) { return Matrix(r, c, std::vector<T>(r * c, T(0))); }
    static Matrix ones(size_t r, size_t c) { return Matrix(r, c, std::vector<T>(r * c, T(1))); }
    static Matrix identity(size_t n) {
        Matrix m = zeros(n, n);
        for (size_t i = 0; i < n; ++i) m.data_[i * n + i] = T(1);
        return m;
    }
    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    size_t row_stride() const { return cols_; }
    const T *as_ptr() const { return data_.data(); }
    T *as_mut_ptr() { return data_.data(); }
    const std::vector<T> &data() const { return data_; }
    T operator()(size_t i, size_t j) const {
        if (i >= rows_ || j >= cols_) throw Panic("index out of bounds");
        return data_[i * cols_ + j];
    }
    Vector<T> solve(const Vector<T> &y) const;   // impl_mat.rs:354-356
  private:
    size_t rows_, cols_;
    std::vector<T> data_;
};

template <typename T>
MatrixSlice<T> MatrixSlice<T>::from_matrix(const Matrix<T> &m, size_t r0, size_t c0, size_t rows, size_t cols) {
    if (r0 + rows > m.rows() || c0 + cols > m.cols()) throw Panic("View dimensions exceed matrix dimensions.");
    return MatrixSlice(m.as_ptr() + r0 * m.row_stride() + c0, rows, cols, m.row_stride());
}

// mat_mul_general! (mat_mul.rs:17-75) for every Matrix / MatrixSlice pairing
template <typename T, typename L, typename R>
Matrix<T> mat_mul_general(const L &a, const R &m) {
    if (a.cols() != m.rows()) throw Panic("Matrix dimensions do not agree.");
    const size_t p = a.rows(), q = a.cols(), r = m.cols();
    std::vector<T> new_data(p * r);     // the reference hands over uninitialised memory (mat_mul.rs:52-55)
    detail::check(detail::Abi<T>::gemm(p, q, r, a.as_ptr(), ptrdiff_t(a.row_stride()), m.as_ptr(), ptrdiff_t(m.row_stride()),
                                       new_data.data()));
    return Matrix<T>(p, r, std::move(new_data));
}
template <typename T> Matrix<T> operator*(const Matrix<T> &a, const Matrix<T> &b) { return mat_mul_general<T>(a, b); }
template <typename T> Matrix<T> operator*(const MatrixSlice<T> &a, const Matrix<T> &b) { return mat_mul_general<T>(a, b); }
template <typename T> Matrix<T> operator*(const Matrix<T> &a, const MatrixSlice<T> &b) { return mat_mul_general<T>(a, b); }
template <typename T> Matrix<T> operator*(const MatrixSlice<T> &a, const MatrixSlice<T> &b) { return mat_mul_general<T>(a, b); }

// Held operands (rla_operand_hold / rla_operand_release): while the guard lives the matrix is promised immutable -- it is
// only reachable through `const Matrix<T> &` here -- so products, solves and matrix-vector calls that read it find its copy
// resident in HBM instead of uploading it again (the repeated products of lu.rs:789,907, eigen.rs:114-148).
template <typename T>
class Held {
  public:
    explicit Held(const Matrix<T> &m) : p_(m.rows() * m.cols() ? m.as_ptr() : nullptr) {
        if (p_) detail::check(rla_operand_hold(p_, m.rows() * m.cols() * sizeof(T)));
    }
    ~Held() { if (p_) rla_operand_release(p_); }
    Held(const Held &) = delete;
    Held &operator=(const Held &) = delete;
  private:
    const void *p_;
};

class PermutationMatrix {
  public:
    PermutationMatrix() = default;
    explicit PermutationMatrix(std::vector<size_t> perm) : perm_(std::move(perm)) {}
    static PermutationMatrix identity(size_t n) {
        std::vector<size_t> p(n);
        for (size_t i = 0; i < n; ++i) p[i] = i;
        return PermutationMatrix(std::move(p));
    }
    size_t size() const { return perm_.size(); }
    const std::vector<size_t> &perm() const { return perm_; }
    PermutationMatrix inverse() const {
        std::vector<size_t> inv(perm_.size());
        for (size_t s = 0; s < perm_.size(); ++s) inv[perm_[s]] = s;
        return PermutationMatrix(std::move(inv));
    }
    int parity_sign() const {
        std::vector<size_t> p = perm_;
        int sign = 1;
        for (size_t i = 0; i < p.size(); ++i)
            while (p[i] != i) { size_t t = p[i]; p[i] = p[t]; p[t] = t; sign = -sign; }
        return sign;
    }
    template <typename T> Vector<T> operator*(const Vector<T> &v) const {   // impl_permutation_mul.rs:21-41
        if (v.size() != size()) throw Panic("Permutation matrix and Vector dimensions are not compatible.");
        std::vector<T> out(v.size());
        for (size_t i = 0; i < v.size(); ++i) out[perm_[i]] = v.data()[i];
        return Vector<T>(std::move(out));
    }
  private:
    std::vector<size_t> perm_;
};

template <typename T>
class PartialPivLu {
  public:
    // lu.rs:163-195: consumes the matrix, factorises in place
    static PartialPivLu decompose(Matrix<T> matrix) {
        const size_t n = matrix.cols();
        if (matrix.rows() != n) throw Panic("Matrix must be square for LU decomposition.");
        std::vector<size_t> perm(n);
        if (detail::check(detail::Abi<T>::getrf(n, matrix.as_mut_ptr(), perm.data())) == RLA_ERR_SINGULAR)
            throw Error(ErrorKind::DivByZero, "The matrix is too ill-conditioned for\n                     LU decomposition with partial pivoting.");
        PartialPivLu out;
        out.lu_ = std::move(matrix);
        out.p_ = PermutationMatrix(std::move(perm));
        out.hold_();         // the factors never change again: every solve() after the first finds them resident in HBM
        return out;
    }
    PartialPivLu() = default;
    PartialPivLu(const PartialPivLu &o) : lu_(o.lu_), p_(o.p_) { hold_(); }
    PartialPivLu(PartialPivLu &&o) noexcept : lu_(std::move(o.lu_)), p_(std::move(o.p_)), held_(o.held_) { o.held_ = false; }
    PartialPivLu &operator=(PartialPivLu o) {
        release_();
        lu_ = std::move(o.lu_);          // the vector's heap block (the held range) moves with it
        p_ = std::move(o.p_);
        held_ = o.held_;
        o.held_ = false;
        return *this;
    }
    ~PartialPivLu() { release_(); }
    // lu.rs:231-244
    Vector<T> solve(Vector<T> b) const {
        if (b.size() != lu_.rows()) throw Panic("Right-hand side vector must have compatible size.");
        if (detail::check(detail::Abi<T>::getrs(lu_.rows(), lu_.as_ptr(), p_.perm().data(), b.mut_data().data())) == RLA_ERR_SINGULAR)
            throw Error(ErrorKind::DivByZero, "Lower triangular matrix is singular to working precision.");
        return b;
    }
    // lu.rs:291-300
    T det() const {
        T u_det = T(1);
        for (size_t i = 0; i < lu_.rows(); ++i) u_det = u_det * lu_(i, i);
        return (p_.parity_sign() > 0 ? T(1) : T(0) - T(1)) * u_det;
    }
    const Matrix<T> &lu() const { return lu_; }
    const PermutationMatrix &p() const { return p_; }
  private:
    void hold_() {
        const size_t bytes = lu_.rows() * lu_.cols() * sizeof(T);
        held_ = bytes != 0 && rla_operand_hold(lu_.as_ptr(), bytes) == RLA_OK;
    }
    void release_() {
        if (held_) rla_operand_release(lu_.as_ptr());
        held_ = false;
    }
    Matrix<T> lu_;
    PermutationMatrix p_;
    bool held_ = false;
};

// Cholesky<T> (src/matrix/decomposition/cholesky.rs:94-245)
template <typename T>
class Cholesky {
  public:
    // :116-170: consumes the matrix, factorises its lower triangle in place
    static Cholesky decompose(Matrix<T> matrix) {
        const size_t n = matrix.cols();
        if (matrix.rows() != n) throw Panic("Matrix must be square for Cholesky decomposition.");
        const int st = detail::check(detail::Abi<T>::potrf(n, matrix.as_mut_ptr()));
        if (st == RLA_ERR_SINGULAR) throw Error(ErrorKind::DecompFailure, "Matrix is singular to working precision.");
        if (st == RLA_ERR_NOT_POSITIVE) throw Error(ErrorKind::DecompFailure, "Diagonal entries of matrix are not all positive.");
        Cholesky out;
        out.l_ = std::move(matrix);
        return out;
    }
    // :175-180
    T det() const {
        T l_det = T(1);
        for (size_t i = 0; i < l_.rows(); ++i) l_det = l_det * l_(i, i);
        return l_det * l_det;
    }
    // :194-203
    Vector<T> solve(Vector<T> b) const {
        if (b.size() != l_.rows()) throw Panic("RHS vector and coefficient matrix must be dimensionally compatible.");
        if (detail::check(detail::Abi<T>::potrs(l_.rows(), l_.as_ptr(), b.mut_data().data())) == RLA_ERR_SINGULAR)
            throw Error(ErrorKind::DivByZero, "Matrix L is singular to working precision.");
        return b;
    }
    // :209-233
    Matrix<T> inverse() const {
        const size_t n = l_.rows();
        Matrix<T> inv = Matrix<T>::zeros(n, n);
        if (detail::check(detail::Abi<T>::potri(n, l_.as_ptr(), inv.as_mut_ptr())) == RLA_ERR_SINGULAR)
            throw Error(ErrorKind::DivByZero, "Matrix L is singular to working precision.");
        return inv;
    }
    // Decomposition::unpack (:237-245): L with the strict upper triangle zeroed
    Matrix<T> unpack() const {
        Matrix<T> l = l_;
        T *p = l.as_mut_ptr();
        for (size_t i = 0; i < l.rows(); ++i)
            for (size_t j = i + 1; j < l.cols(); ++j) p[i * l.cols() + j] = T(0);
        return l;
    }
  private:
    Matrix<T> l_;
};

template <typename T>
Vector<T> Matrix<T>::solve(const Vector<T> &y) const {
    return PartialPivLu<T>::decompose(*this).solve(y);
}

}  // namespace rla
