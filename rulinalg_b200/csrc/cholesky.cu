// cholesky.cu -- SURVEY 8f rank 4: Cholesky::{decompose, solve, inverse} (src/matrix/decomposition/cholesky.rs).
//
// Reference semantics kept (cholesky.rs:116-170): A = L L^T for a symmetric positive definite A, only the LOWER triangle
// is read and written (the strict upper triangle of the packed result is unspecified here; the reference leaves the
// input there and `unpack` zeroes it).  Column j:  a_kj -= utils::dot(a[k,0..j], a[j,0..j]) for k >= j;  d = a_jj;
// |d| < epsilon -> DecompFailure("Matrix is singular to working precision.") (info = j+1);  d < 0 -> DecompFailure
// ("Diagonal entries of matrix are not all positive.") (info = -(j+1));  a_kj /= sqrt(d).
//
// Blocked right-looking form, two levels like the LU driver (outer block 256, inner panels 64), no pivoting:
//   chol_diag_kernel   one CTA factors the 64x64 diagonal block in the reference's exact order (utils::dot's 8 partial
//                      sums, unfused multiply/add, IEEE sqrt and division) => n <= 64 is bit-identical to the reference;
//   chol_panel_kernel  L21 = A21 L11^-T, one thread per row, the row in registers, rows staged through shared memory;
//   transpose_kernel   L21^T for the update (the GEMM kernels take row-major B);
//   dgemm/sgemm        A22 -= L21 L21^T: rank-64 inside the outer block, rank-256 on the trailing matrix, issued per
//                      2048-column chunk from the chunk's diagonal down (the upper blocks are never computed).
// solve (cholesky.rs:194-203): L y = b with the non-unit lower trsv (forward_substitution, mod.rs:363-398), L^T x = y with
// the upper trsv on an explicit transpose (transpose_back_substitution, cholesky.rs:329-365); n <= 64: one thread per
// right-hand side in the reference's exact order.  inverse (cholesky.rs:209-233, n solves in the reference): the packed
// factor is rewritten as a unit-lower / upper pair (L D^-1, D L^T) and handed to the blocked multi-RHS getri.
#include <cfloat>
#include <climits>

#include "common.cuh"

namespace rla {
namespace {

constexpr int CB = 64;       // diagonal block / inner panel width
constexpr int CW = 256;      // outer block width
constexpr int CLD = CB + 1;  // padded shared row
constexpr int CH = 2048;     // column chunk of the trailing update
constexpr int PANEL_ROWS = 128;

template <typename T> struct CEps;
template <> struct CEps<double> { static __device__ __forceinline__ double v() { return DBL_EPSILON; } };
template <> struct CEps<float> { static __device__ __forceinline__ float v() { return FLT_EPSILON; } };

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }

// utils::dot (src/utils.rs:20-51): 8 partial sums over chunks of 8, combined (s+p0+p4), (+p1+p5), (+p2+p6), (+p3+p7),
// then the scalar tail; every multiply and add rounded separately.
template <typename T>
__device__ __forceinline__ T dot8(const T *xs, const T *ys, int len) {
    T p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    int i = 0;
    for (; i + 8 <= len; i += 8) {
        p0 = add_rn(p0, mul_rn(xs[i + 0], ys[i + 0]));
        p1 = add_rn(p1, mul_rn(xs[i + 1], ys[i + 1]));
        p2 = add_rn(p2, mul_rn(xs[i + 2], ys[i + 2]));
        p3 = add_rn(p3, mul_rn(xs[i + 3], ys[i + 3]));
        p4 = add_rn(p4, mul_rn(xs[i + 4], ys[i + 4]));
        p5 = add_rn(p5, mul_rn(xs[i + 5], ys[i + 5]));
        p6 = add_rn(p6, mul_rn(xs[i + 6], ys[i + 6]));
        p7 = add_rn(p7, mul_rn(xs[i + 7], ys[i + 7]));
    }
    T s = 0;
    s = add_rn(add_rn(s, p0), p4);
    s = add_rn(add_rn(s, p1), p5);
    s = add_rn(add_rn(s, p2), p6);
    s = add_rn(add_rn(s, p3), p7);
    for (; i < len; ++i) s = add_rn(s, mul_rn(xs[i], ys[i]));
    return s;
}

// Diagonal block (jb <= 64) at A[0..jb, 0..jb) of the pointer passed in: one thread per row, the reference's column loop.
template <typename T>
__global__ void __launch_bounds__(CB)
chol_diag_kernel(T *__restrict__ A, size_t ld, int jb, int col0, int32_t *__restrict__ info) {
    if (*info != 0) return;
    __shared__ T s[CB * CLD];
    const int k = threadIdx.x;
    if (k < jb)
        for (int c = 0; c <= k; ++c) s[k * CLD + c] = A[size_t(k) * ld + c];
    __syncthreads();
    for (int j = 0; j < jb; ++j) {
        if (j > 0 && k >= j && k < jb) s[k * CLD + j] = sub_rn(s[k * CLD + j], dot8(s + k * CLD, s + j * CLD, j));
        __syncthreads();
        const T d = s[j * CLD + j];
        if (fabs(d) < CEps<T>::v()) {
            if (k == 0) *info = col0 + j + 1;            // "Matrix is singular to working precision."
            return;
        } else if (d < T(0)) {
            if (k == 0) *info = -(col0 + j + 1);         // "Diagonal entries of matrix are not all positive."
            return;
        }
        const T divisor = sqrt_rn(d);
        __syncthreads();                                  // everyone has read the diagonal before it is overwritten
        if (k >= j && k < jb) s[k * CLD + j] = div_rn(s[k * CLD + j], divisor);
        __syncthreads();
    }
    if (k < jb)
        for (int c = 0; c <= k; ++c) A[size_t(k) * ld + c] = s[k * CLD + c];
}

// Rows below the diagonal block: x L11^T = a, one thread per row (x in registers), 128 rows per CTA staged through
// shared memory so that global accesses are 512-byte row segments.
template <typename T>
__global__ void __launch_bounds__(PANEL_ROWS)
chol_panel_kernel(T *__restrict__ A21, size_t ld, int nrows, const T *__restrict__ L11, int jb, const int32_t *__restrict__ info) {
    if (*info != 0) return;
    extern __shared__ __align__(16) unsigned char chol_smem[];
    T *l = reinterpret_cast<T *>(chol_smem);              // [CB][CLD]
    T *rows = l + CB * CLD;                                // [PANEL_ROWS][CLD]
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * PANEL_ROWS;
    const int nr = min(PANEL_ROWS, nrows - r0);
    for (int idx = tid; idx < jb * jb; idx += PANEL_ROWS) {
        const int r = idx / jb, c = idx - r * jb;
        l[r * CLD + c] = (c <= r) ? L11[size_t(r) * ld + c] : T(0);
    }
    for (int idx = tid; idx < nr * jb; idx += PANEL_ROWS) {
        const int r = idx / jb, c = idx - r * jb;
        rows[r * CLD + c] = A21[size_t(r0 + r) * ld + c];
    }
    __syncthreads();
    if (tid < nr) {
        T x[CB];
#pragma unroll
        for (int c = 0; c < CB; ++c) x[c] = (c < jb) ? rows[tid * CLD + c] : T(0);
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            if (c < jb) {
                T acc = x[c];
#pragma unroll
                for (int i = 0; i < c; ++i) acc = sub_rn(acc, mul_rn(x[i], l[c * CLD + i]));
                x[c] = div_rn(acc, l[c * CLD + c]);
            }
        }
#pragma unroll
        for (int c = 0; c < CB; ++c)
            if (c < jb) rows[tid * CLD + c] = x[c];
    }
    __syncthreads();
    for (int idx = tid; idx < nr * jb; idx += PANEL_ROWS) {
        const int r = idx / jb, c = idx - r * jb;
        A21[size_t(r0 + r) * ld + c] = rows[r * CLD + c];
    }
}

// dst[c][r] = src[r][c]  (rows x cols -> cols x rows)
template <typename T>
__global__ void transpose_kernel(const T *__restrict__ src, size_t lds, T *__restrict__ dst, size_t ldd, int rows, int cols,
                                 const int32_t *__restrict__ info) {
    if (info && *info != 0) return;
    __shared__ T tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[size_t(r) * lds + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[size_t(c) * ldd + r] = tile[threadIdx.x][i];
    }
}

template <typename T>
int transpose_launch(const T *src, size_t lds, T *dst, size_t ldd, int rows, int cols, const int32_t *info, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return RLA_OK;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<T><<<grid, block, 0, st>>>(src, lds, dst, ldd, rows, cols, info);
    RLA_LAUNCHED();
    return RLA_OK;
}

template <typename T>
int gemm_sub(size_t m, size_t k, size_t n, const T *a, size_t lda, const T *b, size_t ldb, T *c, size_t ldc, cudaStream_t st);
template <>
int gemm_sub<double>(size_t m, size_t k, size_t n, const double *a, size_t lda, const double *b, size_t ldb, double *c,
                     size_t ldc, cudaStream_t st) {
    return dgemm_launch(m, k, n, -1.0, a, lda, b, ldb, 1.0, c, ldc, st);
}
template <>
int gemm_sub<float>(size_t m, size_t k, size_t n, const float *a, size_t lda, const float *b, size_t ldb, float *c,
                    size_t ldc, cudaStream_t st) {
    return sgemm_launch(m, k, n, -1.0f, a, lda, b, ldb, 1.0f, c, ldc, st);
}

// Exact-order solve for n <= 64 (forward_substitution mod.rs:363-398, transpose_back_substitution cholesky.rs:329-365):
// thread t solves right-hand side t; rhs_is_identity: RHS t is e_t and the solution becomes COLUMN t of X (inverse).
template <typename T>
__global__ void chol_solve_small_kernel(int n, const T *__restrict__ L, size_t ld, T *__restrict__ X, size_t ldx, int nrhs,
                                        int rhs_is_identity, int32_t *__restrict__ info) {
    __shared__ T l[CB * CLD];
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int r = idx / n, c = idx - r * n;
        l[r * CLD + c] = (c <= r) ? L[size_t(r) * ld + c] : T(0);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= nrhs) return;
    T x[CB];
    for (int i = 0; i < n; ++i) x[i] = rhs_is_identity ? (i == t ? T(1) : T(0)) : X[i];
    for (int i = 0; i < n; ++i) {                           // L y = b
        const T d = l[i * CLD + i];
        if (fabs(d) < CEps<T>::v()) { *info = i + 1; return; }
        x[i] = div_rn(sub_rn(x[i], dot8(l + i * CLD, x, i)), d);
    }
    for (int i = n - 1; i >= 0; --i) {                      // L^T x = y
        const T d = l[i * CLD + i];
        if (fabs(d) < CEps<T>::v()) { *info = i + 1; return; }
        x[i] = div_rn(x[i], d);
        for (int j = 0; j < i; ++j) x[j] = sub_rn(x[j], mul_rn(x[i], l[i * CLD + j]));
    }
    if (rhs_is_identity) {
        for (int i = 0; i < n; ++i) X[size_t(i) * ldx + t] = x[i];
    } else {
        for (int i = 0; i < n; ++i) X[i] = x[i];
    }
}

// packed Cholesky factor -> packed LU factors of the same matrix: strict lower L D^-1, upper D L^T
template <typename T>
__global__ void chol_to_lu_kernel(int n, const T *__restrict__ L, size_t ld, T *__restrict__ M, size_t ldm) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= size_t(n) * n) return;
    const int i = int(idx / n), j = int(idx - size_t(i) * n);
    M[size_t(i) * ldm + j] = (i > j) ? L[size_t(i) * ld + j] / L[size_t(j) * ld + j]
                                     : L[size_t(i) * ld + i] * L[size_t(j) * ld + i];
}
__global__ void iota64_kernel(int64_t *p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void first_nonzero_kernel(int32_t *out, const int32_t *a, const int32_t *b) {
    *out = *a ? *a : *b;
}

}  // namespace

size_t potrf_workspace_elems(size_t n) { return size_t(CW) * ((n + 1) / 2 * 2) + size_t(CB) * CW; }

// In-place factorisation of the lower triangle of the n x n row-major matrix `a`.  *d_info: 0, j+1 (singular at column j)
// or -(j+1) (negative diagonal at column j).  ws: potrf_workspace_elems(n) elements.
template <typename T>
int potrf_launch(size_t n_, T *a, size_t ld, T *ws, int32_t *d_info, cudaStream_t st) {
    if (n_ > 0x7fffffffull / 2) return RLA_ERR_INVALID;
    const int n = int(n_);
    RLA_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    if (n == 0) return RLA_OK;
    static DeviceOnce attr_once;
    const size_t panel_smem = size_t(CB + PANEL_ROWS) * CLD * sizeof(T);
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(chol_panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(panel_smem)));
        attr_once.done(od_);
    }
    const size_t ldt = (n_ + 1) / 2 * 2;                    // transposed outer panel: CW x ldt
    T *t_outer = ws, *t_inner = ws + size_t(CW) * ldt;      // t_inner: CB x CW
    for (int J0 = 0; J0 < n; J0 += CW) {
        const int w = min(CW, n - J0);
        for (int j = J0; j < J0 + w; j += CB) {
            const int jb = min(CB, J0 + w - j);
            T *ajj = a + size_t(j) * ld + j;
            chol_diag_kernel<T><<<1, CB, 0, st>>>(ajj, ld, jb, j, d_info);
            RLA_LAUNCHED();
            const int below = n - j - jb;
            if (below <= 0) continue;
            T *a21 = a + size_t(j + jb) * ld + j;
            chol_panel_kernel<T><<<(below + PANEL_ROWS - 1) / PANEL_ROWS, PANEL_ROWS, panel_smem, st>>>(a21, ld, below, ajj, jb, d_info);
            RLA_LAUNCHED();
            const int nc = J0 + w - j - jb;                 // remaining columns of the outer block
            if (nc > 0) {
                RLA_TRY(transpose_launch<T>(a21, ld, t_inner, size_t(CW), nc, jb, d_info, st));
                RLA_TRY(gemm_sub<T>(size_t(below), size_t(jb), size_t(nc), a21, ld, t_inner, size_t(CW),
                                    a + size_t(j + jb) * ld + j + jb, ld, st));
            }
        }
        const int next = J0 + w, nrem = n - next;
        if (nrem <= 0) break;
        const T *l21 = a + size_t(next) * ld + J0;
        RLA_TRY(transpose_launch<T>(l21, ld, t_outer, ldt, nrem, w, d_info, st));
        for (int c0 = 0; c0 < nrem; c0 += CH) {            // lower block triangle only
            const int cw = min(CH, nrem - c0);
            RLA_TRY(gemm_sub<T>(size_t(nrem - c0), size_t(w), size_t(cw), l21 + size_t(c0) * ld, ld, t_outer + c0, ldt,
                                a + size_t(next + c0) * ld + next + c0, ld, st));
        }
    }
    return RLA_OK;
}
template int potrf_launch<double>(size_t, double *, size_t, double *, int32_t *, cudaStream_t);
template int potrf_launch<float>(size_t, float *, size_t, float *, int32_t *, cudaStream_t);

// b <- A^-1 b given the packed factor.  lt: n x ld scratch for L^T (n > 64), ws2: 2n elements, d_info2: two extra words.
template <typename T>
int potrs_launch(size_t n_, const T *l, size_t ld, T *d_b, T *lt, T *ws2, int32_t *d_info, int32_t *d_info2, int32_t *d_sync,
                 cudaStream_t st) {
    const int n = int(n_);
    RLA_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    if (n == 0) return RLA_OK;
    if (n <= CB) {
        chol_solve_small_kernel<T><<<1, CB, 0, st>>>(n, l, ld, d_b, 1, 1, 0, d_info);
        RLA_LAUNCHED();
        return RLA_OK;
    }
    RLA_TRY(trsv_launch<T>(true, n_, l, ld, d_b, ws2, d_info2, d_sync, st));
    RLA_TRY(transpose_launch<T>(l, ld, lt, ld, n, n, nullptr, st));
    RLA_TRY(trsv_launch<T>(false, n_, lt, ld, d_b, ws2, d_info2 + 1, d_sync, st));
    first_nonzero_kernel<<<1, 1, 0, st>>>(d_info, d_info2, d_info2 + 1);
    RLA_LAUNCHED();
    return RLA_OK;
}
template int potrs_launch<double>(size_t, const double *, size_t, double *, double *, double *, int32_t *, int32_t *, int32_t *, cudaStream_t);
template int potrs_launch<float>(size_t, const float *, size_t, float *, float *, float *, int32_t *, int32_t *, int32_t *, cudaStream_t);

// x <- A^-1 given the packed factor.  m: n x ld scratch, d_perm: n int64 scratch.
template <typename T>
int potri_launch(size_t n_, const T *l, size_t ld, T *x, size_t ldx, T *m, int64_t *d_perm, int32_t *d_info, cudaStream_t st) {
    const int n = int(n_);
    RLA_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    if (n == 0) return RLA_OK;
    if (n <= CB) {
        chol_solve_small_kernel<T><<<1, CB, 0, st>>>(n, l, ld, x, ldx, n, 1, d_info);
        RLA_LAUNCHED();
        return RLA_OK;
    }
    const size_t total = n_ * n_;
    chol_to_lu_kernel<T><<<unsigned((total + 255) / 256), 256, 0, st>>>(n, l, ld, m, ld);
    RLA_LAUNCHED();
    iota64_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_perm, n);
    RLA_LAUNCHED();
    return getri_launch<T>(n_, m, ld, d_perm, x, ldx, d_info, st);
}
template int potri_launch<double>(size_t, const double *, size_t, double *, size_t, double *, int64_t *, int32_t *, cudaStream_t);
template int potri_launch<float>(size_t, const float *, size_t, float *, size_t, float *, int64_t *, int32_t *, cudaStream_t);

}  // namespace rla
