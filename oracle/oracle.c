/*
 * oracle.c -- CPU restatement of rulinalg's dense hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This file is the parity oracle for rulinalg_b200.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path (rulinalg_b200/csrc) never links or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 *
 * PINNING STATUS
 *   LU / solve / det / inverse / unpack / permutation / comparators: fully in-tree in the
 *     reference and PINNED here against the reference's own known-answer tests
 *     (tests/test_oracle_kat.py reproduces tests/mat/mod.rs:4-26,100-170 and
 *     src/matrix/decomposition/lu.rs:759-889).
 *   GEMM: the arithmetic lives in the third-party crate `matrixmultiply` ("0.1.13" caret
 *     requirement, Cargo.toml:19; Cargo.lock git-ignored => version unpinned; source not
 *     vendored under /root/reference; no Rust toolchain in this image).  What is restated
 *     is its published algorithm (gemm_loop: k split into kc=256 blocks; micro-kernel
 *     accumulates ab += a*b sequentially in k with separate multiply and add; first
 *     k-block stores C = alpha*ab (+ beta*C if beta != 0), later blocks C = C + alpha*ab).
 *     The reference's own GEMM tests (src/matrix/mat_mul.rs:293-427) are exact-integer
 *     KATs that pin shape/stride handling but not rounding: for real-valued data GEMM is
 *     "PARITY UNPINNED" w.r.t. rounding order; the gate used instead is stated in DESIGN.md.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  -ffp-contract=off
 * matters: rustc never contracts a*b+c into an FMA, so neither may this file.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define ORC_OK 0
#define ORC_ERR_SINGULAR 1   /* maps to ErrorKind::DivByZero (src/error.rs:10-34) */

/* ------------------------------------------------------------------------------------------
 * Seeded counter-based generator (NOT from the reference: benches/linalg/util.rs:5-10 uses
 * rand-0.3 StdRng whose stream no reference output depends on).  splitmix64 keyed by
 * (seed, linear index) so the oracle and every GPU shard generate identical data.
 * The same function is restated in rulinalg_b200/csrc/fill.cu.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
/* U[lo, lo+scale) with 53 random mantissa bits */
void orc_fill_f64(double *dst, size_t count, uint64_t seed, uint64_t offset, double lo, double scale) {
    for (size_t i = 0; i < count; ++i) {
        uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + (offset + i));
        dst[i] = lo + scale * ((double)(r >> 11) * (1.0 / 9007199254740992.0));
    }
}
/* U[lo, lo+scale) with 24 random mantissa bits */
void orc_fill_f32(float *dst, size_t count, uint64_t seed, uint64_t offset, float lo, float scale) {
    for (size_t i = 0; i < count; ++i) {
        uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + (offset + i));
        dst[i] = lo + scale * ((float)(r >> 40) * (1.0f / 16777216.0f));
    }
}

/* ------------------------------------------------------------------------------------------
 * GEMM  (src/matrix/mat_mul.rs:17-75 -> matrixmultiply::{sgemm,dgemm}, call sites :33-43,:57-67)
 *
 * Per-element order (matrixmultiply 0.1.x gemm_loop / kernel, restated from its published
 * algorithm):  c_ij = ((S_0 + S_1) + S_2) ...,  S_b = sum over k in block b (ascending, kc=256)
 * of fl(a_ik*b_kj), one rounding per multiply and per add.  MR/NR/mc/nc blocking does not
 * change the per-element order, so this routine is free to block i and j for cache reuse.
 * ---------------------------------------------------------------------------------------- */
#define ORC_KC 256
#define ORC_MC 64
#define ORC_NC 512

#define DEFINE_GEMM(NAME, T)                                                                   \
void NAME(size_t m, size_t k, size_t n, T alpha,                                               \
          const T *a, ptrdiff_t rsa, ptrdiff_t csa,                                            \
          const T *b, ptrdiff_t rsb, ptrdiff_t csb,                                            \
          T beta, T *c, ptrdiff_t rsc, ptrdiff_t csc)                                          \
{                                                                                              \
    if (m == 0 || n == 0) return;                                                              \
    if (k == 0) {                                                                              \
        /* beta*C; with beta==0 C is zero-filled and never read (mat_mul.rs:28-31: uninit) */  \
        for (size_t i = 0; i < m; ++i)                                                         \
            for (size_t j = 0; j < n; ++j) {                                                   \
                T *cp = c + (ptrdiff_t)i * rsc + (ptrdiff_t)j * csc;                           \
                *cp = (beta == (T)0) ? (T)0 : (*cp) * beta;                                    \
            }                                                                                  \
        return;                                                                                \
    }                                                                                          \
    T *ab = (T *)malloc(sizeof(T) * ORC_MC * ORC_NC);                                          \
    T *bp = (T *)malloc(sizeof(T) * ORC_KC * ORC_NC);                                          \
    for (size_t j0 = 0; j0 < n; j0 += ORC_NC) {                                                \
        size_t nc = (n - j0 < ORC_NC) ? n - j0 : ORC_NC;                                       \
        for (size_t k0 = 0; k0 < k; k0 += ORC_KC) {                                            \
            size_t kc = (k - k0 < ORC_KC) ? k - k0 : ORC_KC;                                   \
            /* pack B block (kc x nc) contiguous, any strides */                               \
            for (size_t kk = 0; kk < kc; ++kk)                                                 \
                for (size_t j = 0; j < nc; ++j)                                                \
                    bp[kk * ORC_NC + j] =                                                      \
                        b[(ptrdiff_t)(k0 + kk) * rsb + (ptrdiff_t)(j0 + j) * csb];             \
            for (size_t i0 = 0; i0 < m; i0 += ORC_MC) {                                        \
                size_t mc = (m - i0 < ORC_MC) ? m - i0 : ORC_MC;                               \
                for (size_t i = 0; i < mc; ++i) {                                              \
                    T *abr = ab + i * ORC_NC;                                                  \
                    for (size_t j = 0; j < nc; ++j) abr[j] = (T)0;                             \
                    for (size_t kk = 0; kk < kc; ++kk) {                                       \
                        T av = a[(ptrdiff_t)(i0 + i) * rsa + (ptrdiff_t)(k0 + kk) * csa];      \
                        const T *bpr = bp + kk * ORC_NC;                                       \
                        for (size_t j = 0; j < nc; ++j) {                                      \
                            T prod = av * bpr[j];      /* separate multiply ...           */  \
                            abr[j] = abr[j] + prod;    /* ... then add (no FMA)           */  \
                        }                                                                      \
                    }                                                                          \
                    for (size_t j = 0; j < nc; ++j) {                                          \
                        T *cp = c + (ptrdiff_t)(i0 + i) * rsc + (ptrdiff_t)(j0 + j) * csc;     \
                        if (k0 == 0) {                                                         \
                            if (beta == (T)0) *cp = alpha * abr[j];                            \
                            else              *cp = (*cp) * beta + alpha * abr[j];             \
                        } else {                                                               \
                            *cp = (*cp) + alpha * abr[j];                                      \
                        }                                                                      \
                    }                                                                          \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }                                                                                          \
    free(ab); free(bp);                                                                        \
}
DEFINE_GEMM(orc_dgemm, double)
DEFINE_GEMM(orc_sgemm, float)

/* Generic-T branch of the same macro (mat_mul.rs:76-99): naive i-k-j accumulation straight into
 * C, no k-blocking.  Kept as an independent cross-check of orc_dgemm's indexing. */
void orc_dgemm_ikj(size_t p, size_t q, size_t r, const double *a, size_t rsa,
                   const double *b, size_t rsb, double *c) {
    for (size_t i = 0; i < p * r; ++i) c[i] = 0.0;
    for (size_t i = 0; i < p; ++i)
        for (size_t k = 0; k < q; ++k)
            for (size_t j = 0; j < r; ++j)
                c[i * r + j] = c[i * r + j] + a[i * rsa + k] * b[k * rsb + j];
}

/* "Truth" for error-bound gates: one C entry with long-double accumulation (not the reference;
 * used so that GEMM parity conclusions do not hinge on matrixmultiply's unverifiable order). */
double orc_ddot_ld(size_t k, const double *a, ptrdiff_t csa, const double *b, ptrdiff_t rsb) {
    long double s = 0.0L;
    for (size_t i = 0; i < k; ++i) s += (long double)a[(ptrdiff_t)i * csa] * (long double)b[(ptrdiff_t)i * rsb];
    return (double)s;
}
double orc_sdot_d(size_t k, const float *a, ptrdiff_t csa, const float *b, ptrdiff_t rsb) {
    double s = 0.0;
    for (size_t i = 0; i < k; ++i) s += (double)a[(ptrdiff_t)i * csa] * (double)b[(ptrdiff_t)i * rsb];
    return s;
}
/* sum_k |a_ik| |b_kj| for the Higham bound */
double orc_dabsdot(size_t k, const double *a, ptrdiff_t csa, const double *b, ptrdiff_t rsb) {
    long double s = 0.0L;
    for (size_t i = 0; i < k; ++i) s += fabsl((long double)a[(ptrdiff_t)i * csa] * (long double)b[(ptrdiff_t)i * rsb]);
    return (double)s;
}

/* ------------------------------------------------------------------------------------------
 * utils::dot  (src/utils.rs:20-51): 8 partial sums, combined (s+p0+p4),(+p1+p5),(+p2+p6),(+p3+p7),
 * then a scalar tail.  The order matters for bit parity of back_substitution.
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_DOT(NAME, T)                                                                    \
T NAME(const T *xs, const T *ys, size_t len)                                                   \
{                                                                                              \
    T s = (T)0, p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;                \
    while (len >= 8) {                                                                         \
        p0 = p0 + xs[0] * ys[0]; p1 = p1 + xs[1] * ys[1];                                      \
        p2 = p2 + xs[2] * ys[2]; p3 = p3 + xs[3] * ys[3];                                      \
        p4 = p4 + xs[4] * ys[4]; p5 = p5 + xs[5] * ys[5];                                      \
        p6 = p6 + xs[6] * ys[6]; p7 = p7 + xs[7] * ys[7];                                      \
        xs += 8; ys += 8; len -= 8;                                                            \
    }                                                                                          \
    s = s + p0 + p4;                                                                           \
    s = s + p1 + p5;                                                                           \
    s = s + p2 + p6;                                                                           \
    s = s + p3 + p7;                                                                           \
    for (size_t i = 0; i < len; ++i) s = s + xs[i] * ys[i];                                    \
    return s;                                                                                  \
}
DEFINE_DOT(orc_ddot, double)
DEFINE_DOT(orc_sdot, float)

/* ------------------------------------------------------------------------------------------
 * PartialPivLu::decompose  (src/matrix/decomposition/lu.rs:163-195) with
 * gaussian_elimination (lu.rs:603-616), PermutationMatrix::{identity,swap_rows,inverse}
 * (src/matrix/permutation_matrix.rs:124-148) and BaseMatrixMut::swap_rows
 * (src/matrix/base/mod.rs:1394-1417).
 *   - pivot = FIRST row attaining max |a_ik| (strict '>', ascending i); NaN never wins
 *   - |pivot| < epsilon (ABSOLUTE) -> DivByZero
 *   - elimination: mult = a_ik / piv; a_ij = a_ij - mult * a_kj   (separate mul and sub)
 *   - returned perm is p.inverse():  perm[original_row] = final_position,  P*A = L*U
 * `lu` is n x n row-major contiguous, factorised in place.
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_GETRF(NAME, T, EPS, FABS)                                                       \
int NAME(size_t n, T *lu, size_t *perm)                                                        \
{                                                                                              \
    size_t *p = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));                                \
    for (size_t i = 0; i < n; ++i) p[i] = i;               /* identity, :124-129 */            \
    for (size_t index = 0; index < n; ++index) {                                               \
        size_t curr_max_idx = index;                                                           \
        T curr_max = lu[index * n + index];                                                    \
        for (size_t i = index + 1; i < n; ++i) {                                               \
            if (FABS(lu[i * n + index]) > FABS(curr_max)) {                                    \
                curr_max = lu[i * n + index];                                                  \
                curr_max_idx = i;                                                              \
            }                                                                                  \
        }                                                                                      \
        if (FABS(curr_max) < EPS) { free(p); return ORC_ERR_SINGULAR; }                        \
        if (curr_max_idx != index) {                       /* swap whole rows */               \
            for (size_t j = 0; j < n; ++j) {                                                   \
                T t = lu[index * n + j];                                                       \
                lu[index * n + j] = lu[curr_max_idx * n + j];                                  \
                lu[curr_max_idx * n + j] = t;                                                  \
            }                                                                                  \
        }                                                                                      \
        { size_t t = p[index]; p[index] = p[curr_max_idx]; p[curr_max_idx] = t; }              \
        {   /* gaussian_elimination, lu.rs:603-616 */                                          \
            T piv_val = lu[index * n + index];                                                 \
            for (size_t i = index + 1; i < n; ++i) {                                           \
                T mult = lu[i * n + index] / piv_val;                                          \
                lu[i * n + index] = mult;                                                      \
                T *ri = lu + i * n;                                                            \
                const T *rk = lu + index * n;                                                  \
                for (size_t j = index + 1; j < n; ++j) {                                       \
                    T prod = mult * rk[j];                                                     \
                    ri[j] = ri[j] - prod;                                                      \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }                                                                                          \
    for (size_t i = 0; i < n; ++i) perm[p[i]] = i;         /* inverse(), :137-148 */           \
    free(p);                                                                                   \
    return ORC_OK;                                                                             \
}
DEFINE_GETRF(orc_dgetrf, double, DBL_EPSILON, fabs)
DEFINE_GETRF(orc_sgetrf, float, FLT_EPSILON, fabsf)

/* lu_forward_substitution (lu.rs:624-642): unit lower solve, plain sequential fold from zero,
 * no singularity check. */
#define DEFINE_FWD(NAME, T)                                                                    \
void NAME(size_t n, const T *lu, T *x)                                                         \
{                                                                                              \
    for (size_t i = 1; i < n; ++i) {                                                           \
        T sum = (T)0;                                                                          \
        const T *row = lu + i * n;                                                             \
        for (size_t k = 0; k < i; ++k) sum = sum + row[k] * x[k];                              \
        x[i] = x[i] - sum;                                                                     \
    }                                                                                          \
}
DEFINE_FWD(orc_dlu_forward_substitution, double)
DEFINE_FWD(orc_slu_forward_substitution, float)

/* back_substitution (src/matrix/mod.rs:318-357): for i descending; |u_ii| < eps -> DivByZero
 * ("Lower triangular matrix is singular to working precision.", :334-335 -- sic);
 * x_i = (x_i - utils::dot(u[i,i+1..n], x[i+1..n])) / u_ii.  `rs` = row stride. */
#define DEFINE_BACK(NAME, T, EPS, FABS, DOT)                                                   \
int NAME(size_t n, const T *u, size_t rs, T *x)                                                \
{                                                                                              \
    for (size_t ii = n; ii-- > 0;) {                                                           \
        T divisor = u[ii * rs + ii];                                                           \
        if (FABS(divisor) < EPS) return ORC_ERR_SINGULAR;                                      \
        T d = DOT(u + ii * rs + ii + 1, x + ii + 1, n - ii - 1);                               \
        x[ii] = (x[ii] - d) / divisor;                                                         \
    }                                                                                          \
    return ORC_OK;                                                                             \
}
DEFINE_BACK(orc_dback_substitution, double, DBL_EPSILON, fabs, orc_ddot)
DEFINE_BACK(orc_sback_substitution, float, FLT_EPSILON, fabsf, orc_sdot)

/* forward_substitution (src/matrix/mod.rs:363-398): general lower solve with diagonal + check
 * ("next" tier: solve_l_triangular). */
#define DEFINE_FWDG(NAME, T, EPS, FABS, DOT)                                                   \
int NAME(size_t n, const T *l, size_t rs, T *x)                                                \
{                                                                                              \
    for (size_t i = 0; i < n; ++i) {                                                           \
        T divisor = l[i * rs + i];                                                             \
        if (FABS(divisor) < EPS) return ORC_ERR_SINGULAR;                                      \
        T d = DOT(l + i * rs, x, i);                                                           \
        x[i] = (x[i] - d) / divisor;                                                           \
    }                                                                                          \
    return ORC_OK;                                                                             \
}
DEFINE_FWDG(orc_dforward_substitution, double, DBL_EPSILON, fabs, orc_ddot)
DEFINE_FWDG(orc_sforward_substitution, float, FLT_EPSILON, fabsf, orc_sdot)

/* PartialPivLu::solve (lu.rs:231-244): &p * b  (impl_permutation_mul.rs:21-41 ->
 * permute_vector_into_buffer, permutation_matrix.rs:369-382: buffer[perm[i]] = b[i]),
 * then lu_forward_substitution, then back_substitution.  b is overwritten with x. */
#define DEFINE_GETRS(NAME, T, FWD, BACK)                                                       \
int NAME(size_t n, const T *lu, const size_t *perm, T *b)                                      \
{                                                                                              \
    T *buf = (T *)malloc(sizeof(T) * (n ? n : 1));                                             \
    for (size_t i = 0; i < n; ++i) buf[perm[i]] = b[i];                                        \
    FWD(n, lu, buf);                                                                           \
    int rc = BACK(n, lu, n, buf);                                                              \
    if (rc == ORC_OK) memcpy(b, buf, sizeof(T) * n);                                           \
    free(buf);                                                                                 \
    return rc;                                                                                 \
}
DEFINE_GETRS(orc_dgetrs, double, orc_dlu_forward_substitution, orc_dback_substitution)
DEFINE_GETRS(orc_sgetrs, float, orc_slu_forward_substitution, orc_sback_substitution)

/* PartialPivLu::inverse (lu.rs:251-285): n solves of unit vectors; column i of inv = solve(e_i). */
int orc_dgetri(size_t n, const double *lu, const size_t *perm, double *inv) {
    double *e = (double *)calloc(n ? n : 1, sizeof(double));
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < n; ++j) e[j] = 0.0;
        e[i] = 1.0;
        int rc = orc_dgetrs(n, lu, perm, e);
        if (rc != ORC_OK) { free(e); return rc; }
        for (size_t j = 0; j < n; ++j) inv[j * n + i] = e[j];
    }
    free(e);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Cholesky: src/matrix/decomposition/cholesky.rs.
 *   decompose (:116-170): "gaxpy-rich" left-looking, column j:  for k in j..n:
 *       a_kj = a_kj - utils::dot(a[k,0..j], a[j,0..j]);  d = a_jj;  |d| < eps -> DecompFailure
 *       ("Matrix is singular to working precision."), d < 0 -> DecompFailure ("Diagonal entries of
 *       matrix are not all positive.");  a_kj = a_kj / sqrt(d) for k in j..n.
 *       Only the lower triangle is read or written; the strict upper triangle keeps the input.
 *   returns 0, +(j+1) for the singular case at column j, -(j+1) for the negative case.
 *   solve (:194-203): forward_substitution (mod.rs:363-398) then transpose_back_substitution
 *       (:329-365): for i descending: |l_ii| < eps -> DivByZero; x_i = x_i / l_ii;
 *       x_j = x_j - x_i * l_ij for j < i (separate mul, sub).
 *   det (:175-180): fold(one, a*b) over the diagonal, squared.   inverse (:209-233): n solves.
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_POTRF(NAME, T, EPS, FABS, SQRT, DOT)                                            \
long NAME(size_t n, T *a)                                                                      \
{                                                                                              \
    for (size_t j = 0; j < n; ++j) {                                                           \
        if (j > 0) {                                                                           \
            for (size_t k = j; k < n; ++k) {                                                   \
                T kj_dot = DOT(a + k * n, a + j * n, j);                                       \
                a[k * n + j] = a[k * n + j] - kj_dot;                                          \
            }                                                                                  \
        }                                                                                      \
        T diagonal = a[j * n + j];                                                             \
        if (FABS(diagonal) < EPS) return (long)(j + 1);                                        \
        else if (diagonal < (T)0) return -(long)(j + 1);                                       \
        T divisor = SQRT(diagonal);                                                            \
        for (size_t k = j; k < n; ++k) a[k * n + j] = a[k * n + j] / divisor;                  \
    }                                                                                          \
    return 0;                                                                                  \
}
DEFINE_POTRF(orc_dpotrf, double, DBL_EPSILON, fabs, sqrt, orc_ddot)
DEFINE_POTRF(orc_spotrf, float, FLT_EPSILON, fabsf, sqrtf, orc_sdot)

#define DEFINE_TBACK(NAME, T, EPS, FABS)                                                       \
int NAME(size_t n, const T *l, size_t rs, T *x)                                                \
{                                                                                              \
    for (size_t i = n; i-- > 0;) {                                                             \
        const T *row = l + i * rs;                                                             \
        T diagonal = row[i];                                                                   \
        if (FABS(diagonal) < EPS) return ORC_ERR_SINGULAR;                                     \
        x[i] = x[i] / diagonal;                                                                \
        for (size_t j = 0; j < i; ++j) {                                                       \
            T prod = x[i] * row[j];                                                            \
            x[j] = x[j] - prod;                                                                \
        }                                                                                      \
    }                                                                                          \
    return ORC_OK;                                                                             \
}
DEFINE_TBACK(orc_dtranspose_back_substitution, double, DBL_EPSILON, fabs)
DEFINE_TBACK(orc_stranspose_back_substitution, float, FLT_EPSILON, fabsf)

#define DEFINE_POTRS(NAME, T, FWD, TBACK)                                                      \
int NAME(size_t n, const T *l, T *b)                                                           \
{                                                                                              \
    int st = FWD(n, l, n, b);                                                                  \
    if (st != ORC_OK) return st;                                                               \
    return TBACK(n, l, n, b);                                                                  \
}
DEFINE_POTRS(orc_dpotrs, double, orc_dforward_substitution, orc_dtranspose_back_substitution)
DEFINE_POTRS(orc_spotrs, float, orc_sforward_substitution, orc_stranspose_back_substitution)

/* Cholesky::inverse: column i of inv = solve(e_i) */
int orc_dpotri(size_t n, const double *l, double *inv) {
    double *e = (double *)malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n; ++i) {
        for (size_t j = 0; j < n; ++j) e[j] = 0.0;
        e[i] = 1.0;
        int st = orc_dpotrs(n, l, e);
        if (st != ORC_OK) { free(e); return st; }
        for (size_t j = 0; j < n; ++j) inv[j * n + i] = e[j];
    }
    free(e);
    return ORC_OK;
}

/* Cholesky::det */
double orc_dpotrf_det(size_t n, const double *l) {
    double d = 1.0;
    for (size_t i = 0; i < n; ++i) d = d * l[i * n + i];
    return d * d;
}

/* Parity of a permutation via the reference's cycle-walking permute_by_swap
 * (permutation_matrix.rs:216-238, :479-509): +1 even, -1 odd.  Parity is a property of the
 * permutation, so any transposition decomposition gives the same sign. */
int orc_perm_sign(size_t n, const size_t *perm) {
    size_t *p = (size_t *)malloc(sizeof(size_t) * (n ? n : 1));
    memcpy(p, perm, sizeof(size_t) * n);
    int sign = 1;
    for (size_t i = 0; i < n; ++i) {
        while (p[i] != i) {
            size_t t = p[i];
            p[i] = p[t]; p[t] = t;
            sign = -sign;
        }
    }
    free(p);
    return sign;
}

/* PartialPivLu::det (lu.rs:291-300): fold(one, x*y) over the diagonal in order, times p.det(). */
double orc_ddet(size_t n, const double *lu, const size_t *perm) {
    double u_det = 1.0;
    for (size_t i = 0; i < n; ++i) u_det = u_det * lu[i * n + i];
    double p_det = (orc_perm_sign(n, perm) > 0) ? 1.0 : (0.0 - 1.0);
    return p_det * u_det;
}

/* Decomposition::unpack (lu.rs:138-149): l = unit_lower_triangular_part (lu.rs:644-663),
 * u = lu with strict lower part zeroed (internal_utils.rs:5-13). */
void orc_dunpack(size_t n, const double *lu, double *l, double *u) {
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j) {
            l[i * n + j] = (j < i) ? lu[i * n + j] : (j == i ? 1.0 : 0.0);
            u[i * n + j] = (j < i) ? 0.0 : lu[i * n + j];
        }
}

/* ------------------------------------------------------------------------------------------
 * Comparators  (src/ulp.rs:41-65, src/macros/comparison.rs:46-197,
 * src/macros/assert_matrix_eq.rs:106-146: element-wise over all (i,j), collecting mismatches).
 * Each returns the number of mismatching elements; *first_bad = index of the first one (or n).
 * ---------------------------------------------------------------------------------------- */
/* ulp_diff result codes */
#define ULP_EXACT 0
#define ULP_DIFF 1
#define ULP_SIGNS 2
#define ULP_NAN 3
int orc_ulp_diff_f64(double a, double b, uint64_t *diff) {
    *diff = 0;
    if (a == b) return ULP_EXACT;
    if (isnan(a) || isnan(b)) return ULP_NAN;
    if ((signbit(a) != 0) != (signbit(b) != 0)) return ULP_SIGNS;
    int64_t ai, bi;
    memcpy(&ai, &a, 8); memcpy(&bi, &b, 8);
    int64_t d = bi - ai;
    *diff = (uint64_t)(d < 0 ? -d : d);
    return ULP_DIFF;
}
int orc_ulp_diff_f32(float a, float b, uint64_t *diff) {
    *diff = 0;
    if (a == b) return ULP_EXACT;
    if (isnan(a) || isnan(b)) return ULP_NAN;
    if ((signbit(a) != 0) != (signbit(b) != 0)) return ULP_SIGNS;
    int32_t ai, bi;
    memcpy(&ai, &a, 4); memcpy(&bi, &b, 4);
    int32_t d = bi - ai;
    *diff = (uint64_t)(d < 0 ? -(int64_t)d : (int64_t)d);
    return ULP_DIFF;
}

#define DEFINE_CMP(SUF, T, ULPDIFF)                                                            \
size_t orc_cmp_exact_##SUF(size_t n, const T *a, const T *b, size_t *first_bad) {              \
    size_t bad = 0; *first_bad = n;                                                            \
    for (size_t i = 0; i < n; ++i)                                                             \
        if (!(a[i] == b[i])) { if (!bad) *first_bad = i; ++bad; }                              \
    return bad;                                                                                \
}                                                                                              \
static inline int abs_ok_##SUF(T a, T b, T tol) {                                              \
    if (a == b) return 1;                                                                      \
    T d = (a > b) ? a - b : b - a;                                                             \
    return d <= tol;                   /* NaN distance -> false */                             \
}                                                                                              \
size_t orc_cmp_abs_##SUF(size_t n, const T *a, const T *b, T tol, size_t *first_bad,           \
                         double *max_abs) {                                                    \
    size_t bad = 0; *first_bad = n; *max_abs = 0.0;                                            \
    for (size_t i = 0; i < n; ++i) {                                                           \
        double d = fabs((double)a[i] - (double)b[i]);                                          \
        if (d > *max_abs || d != d) *max_abs = d;                                              \
        if (!abs_ok_##SUF(a[i], b[i], tol)) { if (!bad) *first_bad = i; ++bad; }               \
    }                                                                                          \
    return bad;                                                                                \
}                                                                                              \
static inline int ulp_ok_##SUF(T a, T b, uint64_t tol, uint64_t *d) {                          \
    int r = ULPDIFF(a, b, d);                                                                  \
    return r == ULP_EXACT || (r == ULP_DIFF && *d <= tol);                                     \
}                                                                                              \
size_t orc_cmp_ulp_##SUF(size_t n, const T *a, const T *b, uint64_t tol, size_t *first_bad,    \
                         uint64_t *max_ulp) {                                                  \
    size_t bad = 0; *first_bad = n; *max_ulp = 0;                                              \
    for (size_t i = 0; i < n; ++i) {                                                           \
        uint64_t d;                                                                            \
        int ok = ulp_ok_##SUF(a[i], b[i], tol, &d);                                            \
        if (d > *max_ulp) *max_ulp = d;                                                        \
        if (!ok) { if (!bad) *first_bad = i; ++bad; }                                          \
    }                                                                                          \
    return bad;                                                                                \
}                                                                                              \
/* comp = float: abs(eps) first, then ulp(max_ulp) (comparison.rs:145-197) */                  \
size_t orc_cmp_float_##SUF(size_t n, const T *a, const T *b, T eps, uint64_t ulp_tol,          \
                           size_t *first_bad) {                                                \
    size_t bad = 0; *first_bad = n;                                                            \
    for (size_t i = 0; i < n; ++i) {                                                           \
        uint64_t d;                                                                            \
        if (!abs_ok_##SUF(a[i], b[i], eps) && !ulp_ok_##SUF(a[i], b[i], ulp_tol, &d)) {        \
            if (!bad) *first_bad = i; ++bad;                                                   \
        }                                                                                      \
    }                                                                                          \
    return bad;                                                                                \
}
DEFINE_CMP(f64, double, orc_ulp_diff_f64)
DEFINE_CMP(f32, float, orc_ulp_diff_f32)

/* testsupport::{is_lower_triangular,is_upper_triangular} (src/testsupport/constraints.rs:10-32) */
int orc_is_lower_triangular(size_t rows, size_t cols, const double *m) {
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = i + 1; j < cols; ++j)
            if (!(m[i * cols + j] == 0.0)) return 0;
    return 1;
}
int orc_is_upper_triangular(size_t rows, size_t cols, const double *m) {
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < i && j < cols; ++j)
            if (!(m[i * cols + j] == 0.0)) return 0;
    return 1;
}
