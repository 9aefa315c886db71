/*
 * rla_b200.h -- C ABI of librla_b200.so: the B200 (sm_100a) implementation of rulinalg's
 * dense hot path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * Every entry point names the reference interface (file:line relative to the rulinalg tree)
 * that it replaces or serves.  INTEGRATION.md shows the Rust binding a maintainer adds.
 *
 * Conventions
 *   - return value: RLA_OK (0), a positive numerical status (RLA_ERR_SINGULAR -> Rust
 *     ErrorKind::DivByZero, src/error.rs:10-34), or a negative environment error (CUDA, no
 *     device, out of memory).  Nothing throws or unwinds across the boundary.
 *   - shape violations are asserted on the Rust side before the call (mat_mul.rs:21,
 *     lu.rs:165,232); the C layer re-checks only what would make it read out of bounds.
 *   - host pointers are borrowed for the duration of the call; nothing is retained.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute entry point
 *     returns RLA_ERR_NO_DEVICE.
 *   - all matrices are row-major; "ld*" / "rs*" are row strides in ELEMENTS.
 */
#ifndef RLA_B200_H
#define RLA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define RLA_API __attribute__((visibility("default")))
#else
#define RLA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RLA_OK 0
#define RLA_ERR_SINGULAR 1      /* |pivot| < epsilon  -> ErrorKind::DivByZero                */
#define RLA_ERR_INVALID 2       /* bad argument (negative stride on an output, null pointer) */
#define RLA_ERR_NOT_POSITIVE 3  /* Cholesky: a diagonal entry is negative (ErrorKind::DecompFailure)      */
#define RLA_ERR_CUDA (-1)       /* a CUDA call failed; see rla_last_cuda_error()             */
#define RLA_ERR_NOMEM (-2)      /* device or pinned-host allocation failed                   */
#define RLA_ERR_NO_DEVICE (-3)  /* no sm_100 device visible (there is no CPU fallback)       */

/* ---------------------------------------------------------------------------------------
 * Host-pointer entry points: the drop-in boundary.
 * ------------------------------------------------------------------------------------- */

/* Replaces matrixmultiply::dgemm at src/matrix/mat_mul.rs:57-67 (same argument order and
 * meaning): C <- alpha*A*B + beta*C, A is m x k, B is k x n, C is m x n, element (i,j) of X
 * at x[i*rsx + j*csx].  rulinalg always passes csa=csb=csc=1, alpha=1, beta=0, and C
 * uninitialised (mat_mul.rs:52-55): with beta == 0 C is never read and every element is
 * written.  k == 0 with beta == 0 zero-fills C.  Unit column strides are the fast path;
 * other strides are packed on the host first. */
RLA_API int rla_dgemm(size_t m, size_t k, size_t n, double alpha,
              const double *a, ptrdiff_t rsa, ptrdiff_t csa,
              const double *b, ptrdiff_t rsb, ptrdiff_t csb,
              double beta, double *c, ptrdiff_t rsc, ptrdiff_t csc);

/* Replaces matrixmultiply::sgemm at src/matrix/mat_mul.rs:33-43.  FP32 FFMA arithmetic, no
 * TF32 / tensor cores, so accuracy class matches the reference. */
RLA_API int rla_sgemm(size_t m, size_t k, size_t n, float alpha,
              const float *a, ptrdiff_t rsa, ptrdiff_t csa,
              const float *b, ptrdiff_t rsb, ptrdiff_t csb,
              float beta, float *c, ptrdiff_t rsc, ptrdiff_t csc);

/* Replaces the body of PartialPivLu::decompose (src/matrix/decomposition/lu.rs:163-195 with
 * gaussian_elimination :603-616).  `lu` is the n x n row-major contiguous matrix, factorised
 * in place into packed L\U (unit diagonal of L implicit).  `perm` (length n) receives
 * PartialPivLu.p.perm, i.e. the permutation AFTER p.inverse() (lu.rs:192):
 * perm[original_row] = final_position, P*A = L*U.  Pivot rule: first row attaining
 * max |a_ik| (lu.rs:173-178).  Returns RLA_ERR_SINGULAR when a pivot has |p| < DBL_EPSILON
 * (lu.rs:179-183); `lu` content is then unspecified (the reference drops the matrix). */
RLA_API int rla_dgetrf(size_t n, double *lu, size_t *perm);
RLA_API int rla_sgetrf(size_t n, float *lu, size_t *perm);      /* T = f32: FLT_EPSILON */

/* Replaces the body of PartialPivLu::solve (lu.rs:231-244): b <- U^-1 L^-1 P b with
 * P b = permute_vector_into_buffer (permutation_matrix.rs:369-382), unit-lower forward
 * substitution (lu.rs:624-642) and back_substitution (src/matrix/mod.rs:318-357), whose
 * |u_ii| < epsilon test returns RLA_ERR_SINGULAR ("Lower triangular matrix is singular to
 * working precision.", mod.rs:334-335).  On error b is left untouched. */
RLA_API int rla_dgetrs(size_t n, const double *lu, const size_t *perm, double *b);
RLA_API int rla_sgetrs(size_t n, const float *lu, const size_t *perm, float *b);

/* PartialPivLu::inverse (lu.rs:251-285) -- SURVEY 8f "next" row.  The reference performs n solves of unit
 * vectors; here X = U^-1 L^-1 P is a blocked multi-RHS substitution on the GEMM kernels (2n^3 flops).  n <= 64
 * keeps the reference's exact per-column operation order (bit-identical).  inv: n x n row-major contiguous,
 * written in full.  RLA_ERR_SINGULAR when some |u_ii| < epsilon (back_substitution's test, mod.rs:333-336). */
RLA_API int rla_dgetri(size_t n, const double *lu, const size_t *perm, double *inv);
RLA_API int rla_sgetri(size_t n, const float *lu, const size_t *perm, float *inv);

/* Cholesky::{decompose, solve, inverse} (src/matrix/decomposition/cholesky.rs:116-233) -- SURVEY 8f "next" row.
 * rla_?potrf: `a` is n x n row-major contiguous; on return its LOWER triangle holds L with A = L L^T (the reference's
 * `Cholesky.l`); only the lower triangle of the input is read; the strict upper triangle is unspecified on return
 * (`unpack` zeroes it, cholesky.rs:237-245).  RLA_ERR_SINGULAR <-> DecompFailure("Matrix is singular to working
 * precision."), RLA_ERR_NOT_POSITIVE <-> DecompFailure("Diagonal entries of matrix are not all positive.")
 * (cholesky.rs:151-158).  n <= 64 is bit-identical to the reference; larger n is a blocked right-looking factorisation
 * on the GEMM kernels.  rla_?potrs: b <- A^-1 b (forward_substitution + transpose_back_substitution, :194-203, :329-365);
 * RLA_ERR_SINGULAR when some |l_ii| < epsilon.  rla_?potri: inv (n x n contiguous) <- A^-1 (:209-233). */
RLA_API int rla_dpotrf(size_t n, double *a);
RLA_API int rla_spotrf(size_t n, float *a);
RLA_API int rla_dpotrs(size_t n, const double *l, double *b);
RLA_API int rla_spotrs(size_t n, const float *l, float *b);
RLA_API int rla_dpotri(size_t n, const double *l, double *inv);
RLA_API int rla_spotri(size_t n, const float *l, float *inv);
/* device-resident factorisation: a (row stride ld) in place; ws: rla_potrf_workspace_bytes(n, sizeof(T)) bytes of device
 * scratch; *d_info <- 0, j+1 (singular at column j) or -(j+1) (negative diagonal at column j); asynchronous on `stream` */
RLA_API size_t rla_potrf_workspace_bytes(size_t n, size_t elem_size);
RLA_API int rla_dpotrf_dev(size_t n, double *a, size_t ld, void *ws, int32_t *d_info, void *stream);
RLA_API int rla_spotrf_dev(size_t n, float *a, size_t ld, void *ws, int32_t *d_info, void *stream);

/* solve_l_triangular / solve_u_triangular (src/matrix/base/mod.rs:1015-1067 -> forward_substitution /
 * back_substitution, src/matrix/mod.rs:318-398) -- SURVEY 8f.  `a` is n x n row-major with row stride rs (only the
 * lower / upper triangle incl. the diagonal is read); x holds y on entry and the solution on exit.
 * RLA_ERR_SINGULAR when some |a_ii| < epsilon (x is then left untouched). */
RLA_API int rla_dtrsv(int lower, size_t n, const double *a, ptrdiff_t rs, double *x);
RLA_API int rla_strsv(int lower, size_t n, const float *a, ptrdiff_t rs, float *x);

/* `&Matrix * &Vector` (src/matrix/impl_ops.rs:298-314) -- SURVEY 8f.  y (length m) = A (m x n, row stride rs) * x. */
RLA_API int rla_dgemv(size_t m, size_t n, const double *a, ptrdiff_t rs, const double *x, double *y);
RLA_API int rla_sgemv(size_t m, size_t n, const float *a, ptrdiff_t rs, const float *x, float *y);
RLA_API int rla_dgemv_dev(size_t m, size_t n, const double *a, size_t lda, const double *x, double *y, void *stream);
RLA_API int rla_sgemv_dev(size_t m, size_t n, const float *a, size_t lda, const float *x, float *y, void *stream);

/* Factorisation kept resident in HBM for repeated solves (PartialPivLu is built for "multiple
 * such linear systems involving the same A", lu.rs:203-206).  rla_dgetrf_keep = rla_dgetrf
 * that also returns a handle; rla_lu_solve = rla_dgetrs without re-uploading lu. */
typedef struct rla_lu_handle rla_lu_handle;
RLA_API int rla_dgetrf_keep(size_t n, double *lu, size_t *perm, rla_lu_handle **out);
RLA_API int rla_dlu_solve(const rla_lu_handle *h, double *b);
RLA_API void rla_lu_free(rla_lu_handle *h);

/* Held operands: device-resident GEMM operands across host-API calls (the reference multiplies by the
 * same matrix repeatedly, e.g. lu.rs:789,907, eigen.rs:114-148).  matrixmultiply's signature cannot say "this operand has
 * not changed since the last call", so the caller says it: between rla_operand_hold(host, bytes) and
 * rla_operand_release(host) the byte range [host, host + bytes) is promised immutable (in Rust: a guard object that
 * borrows `&Matrix<T>` for its lifetime -- the borrow checker enforces the promise; INTEGRATION.md shows it).  While a
 * range is held, the first rla_dgemm / rla_sgemm that reads an A or B operand lying inside it (unit column stride, not
 * the small-call or multi-GPU path) keeps that operand's device copy, and later calls with the same (pointer, rows,
 * cols, row stride, element size) on the same device skip its host-to-device copy.  The same holds for the matrix
 * operand of rla_?getrs (the factors: PartialPivLu is built for "multiple such linear systems involving the same A",
 * lu.rs:203-206), rla_?trsv and rla_?gemv above the small-call size -- calls whose whole cost is that upload (f64 solve,
 * n = 4096: 2.8 ms with the factors re-uploaded, 0.4 ms with them held).  Results are bit-identical either way.  Holds nest (one release per hold of the same base; the byte count must match).  release frees the device
 * copies; it must not run concurrently with a product that reads the range.  rla_shutdown drops all resident copies.
 * If HBM has no room for a resident copy the call proceeds as if the range were not held.
 * rla_operand_resident_bytes: device bytes currently kept for held operands (all devices). */
RLA_API int rla_operand_hold(const void *host, size_t bytes);
RLA_API int rla_operand_release(const void *host);
RLA_API size_t rla_operand_resident_bytes(void);

/* ---------------------------------------------------------------------------------------
 * Device-resident twins (kernel timing, multi-GPU sharding, LU -> solve reuse).
 * All pointers are device pointers on the current device; `stream` is a cudaStream_t passed
 * as void* (NULL = the CUDA legacy default stream, as in the runtime API).  Calls are asynchronous on that
 * stream unless stated.
 * Workspaces: the LU / solve twins (rla_?getrf_dev, rla_?getrs_dev, rla_?getri_dev, rla_dlu_*_dev) use workspaces owned by
 * the calling host thread's context on the current device (pivot packets and tags, row-origin vector, solve tickets).
 * Keep at most ONE of them in flight per (host thread, device): enqueue the next one on the same stream, or synchronise
 * in between.  The GEMM / GEMV / fill twins have no workspace and may overlap freely; rla_?potrf_dev takes its workspace
 * from the caller.
 * ------------------------------------------------------------------------------------- */
RLA_API int rla_init(int device);                 /* select device + create context; idempotent      */
RLA_API int rla_device_count(void);
/* One host process, N GPUs (north_star: "Large GEMMs are sharded as row panels across 2/4/8 GPUs ...
 * large LU uses a 1D block-cyclic column layout").  After rla_set_devices(N) the host-pointer entry points rla_dgemm /
 * rla_sgemm / rla_dgetrf -- the only calls the reference sites mat_mul.rs:33-43,57-67 and lu.rs:163-195 can reach --
 * shard large problems over GPUs 0..N-1: GEMM as row panels of A and C with every column chunk of B uploaded once
 * (chunks dealt round-robin over the N PCIe links) and fanned out over NVLink peer memory; LU as 256-column blocks
 * dealt round-robin with the factored panel fanned out the same way.  Results are bit-identical to N = 1.  Small
 * problems stay on one GPU.  N = 1 (default) restores single-GPU behaviour.  RLA_ERR_INVALID when N exceeds the
 * visible devices or the GPUs lack peer access.  Process-wide; multi-GPU calls from several host threads serialise. */
RLA_API int rla_set_devices(int n_gpus);
RLA_API int rla_get_devices(void);
/* Returns everything the library holds for the calling host thread (streams, events, device and pinned buffers, LU
 * workspaces, staging rings) and the multi-GPU contexts.  Also runs when a host thread exits.  Later calls
 * re-initialise lazily.  (SURVEY 8b "Ownership": device memory is owned by the library, freed at rla_shutdown.) */
RLA_API int rla_shutdown(void);
RLA_API int rla_dev_alloc(void **p, size_t bytes);
RLA_API int rla_dev_free(void *p);
RLA_API int rla_host_alloc_pinned(void **p, size_t bytes);
RLA_API int rla_host_free_pinned(void *p);
RLA_API int rla_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
RLA_API int rla_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
RLA_API int rla_stream_sync(void *stream);

/* C <- alpha*A*B + beta*C on device, unit column strides (the mat_mul.rs fast path).
 * f64: FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) kernel; f32: register-blocked FFMA. */
RLA_API int rla_dgemm_dev(size_t m, size_t k, size_t n, double alpha,
                  const double *a, size_t lda, const double *b, size_t ldb,
                  double beta, double *c, size_t ldc, void *stream);
RLA_API int rla_sgemm_dev(size_t m, size_t k, size_t n, float alpha,
                  const float *a, size_t lda, const float *b, size_t ldb,
                  float beta, float *c, size_t ldc, void *stream);

/* Blocked right-looking LU with partial pivoting on an n x n matrix with row stride ld.
 * d_perm: int64[n] device, same meaning as rla_dgetrf's perm.  d_info: int32 device scalar,
 * 0 = ok, j+1 = |pivot| < epsilon at column j.  Workspace is owned by the library. */
RLA_API int rla_dgetrf_dev(size_t n, double *a, size_t ld, int64_t *d_perm, int32_t *d_info, void *stream);
RLA_API int rla_sgetrf_dev(size_t n, float *a, size_t ld, int64_t *d_perm, int32_t *d_info, void *stream);

/* Solve with a device-resident factorisation: d_b (length n) is overwritten with x.
 * d_info: 0 = ok, i+1 = |u_ii| < epsilon (then d_b is unspecified). */
RLA_API int rla_dgetrs_dev(size_t n, const double *lu, size_t ld, const int64_t *d_perm,
                   double *d_b, int32_t *d_info, void *stream);
RLA_API int rla_sgetrs_dev(size_t n, const float *lu, size_t ld, const int64_t *d_perm,
                   float *d_b, int32_t *d_info, void *stream);

/* inverse from a device-resident factorisation: x (n x n, row stride ldx) <- A^-1; d_info as rla_dgetrs_dev */
RLA_API int rla_dgetri_dev(size_t n, const double *lu, size_t ld, const int64_t *d_perm, double *x, size_t ldx,
                   int32_t *d_info, void *stream);
RLA_API int rla_sgetri_dev(size_t n, const float *lu, size_t ld, const int64_t *d_perm, float *x, size_t ldx,
                   int32_t *d_info, void *stream);

/* Building blocks of the 1D block-cyclic multi-GPU LU (driven by rulinalg_b200/sharded_lu.py, one
 * process per GPU, panel broadcast over NCCL).  Local matrix: n rows x ncols_loc columns, row stride ld; the
 * 256-wide global column block J of rank J mod g sits at local columns [(J/g)*256, ...).  Row indices are
 * global.  d_plan is an opaque device buffer of rla_lu_plan_bytes() bytes holding a block's net row
 * permutation; d_info must be zeroed by the caller before the first block.
 *   factor_block : owner factors the block with diagonal at global row row0, stored at local column lcol0
 *   laswp        : apply the block's interchanges to local columns [c0a,c1a) U [c0b,c1b)
 *   update       : with the broadcast panel ((n-row0) x w, row stride ldp, L11 on top of L21) compute
 *                  U12 = L11^-1 A12 and A22 -= L21 U12 on local columns [c0,c1)
 *   rowid_*      : the row-origin vector every rank carries; perm_from_rowid gives PartialPivLu.p.perm */
RLA_API size_t rla_lu_plan_bytes(void);
/* Largest n rla_?getrf / rla_?getrf_dev accept on the current device (a 64-column panel of n rows must fit the shared
 * memory of the panel kernel's row CTAs): 57 771 for f64, 115 689 for f32 on a 148-SM B200.  Larger n returns
 * RLA_ERR_INVALID before anything is enqueued (the reference has no limit; an n = 57 771 f64 matrix is 26.7 GB).
 * 0 when no device is usable. */
RLA_API size_t rla_lu_max_n(size_t elem_size);
/* development aid: with rla_set_tuning("lu_dbg", 8) the four inner panels of the last outer block record 64 x 8 words
 * each (globaltimer stamps in slots 0..6, the pivot row in slot 7); copies 4 x 512 words to host2048 */
RLA_API int rla_debug_lu_trace(unsigned long long *host2048);
/* test aid: the panel kernel forms its multipliers a / pivot from a correctly rounded reciprocal (five FMAs instead of an
 * IEEE division per row); this compares that against the IEEE division over `count` pseudo-random operand pairs
 * (mode 0: arbitrary bit patterns incl. NaN / Inf / subnormals; mode 1: LU-like, |a| <= |b|) and returns the number
 * of results that differ in any bit.  Must be 0. */
RLA_API int rla_debug_divcheck(int f32, int mode, unsigned long long seed, unsigned long long count,
                               unsigned long long *mismatches);
RLA_API int rla_dlu_factor_block_dev(size_t n, double *a_loc, size_t ld, size_t row0, size_t lcol0, size_t w,
                             int32_t *d_info, void *d_plan, void *stream);
RLA_API int rla_dlu_laswp_dev(double *a_loc, size_t ld, size_t w, const void *d_plan, const int32_t *d_info,
                      size_t c0a, size_t c1a, size_t c0b, size_t c1b, void *stream);
RLA_API int rla_dlu_update_dev(size_t n, double *a_loc, size_t ld, size_t row0, size_t w, const double *panel,
                       size_t ldp, size_t c0, size_t c1, const int32_t *d_info, void *stream);
RLA_API int rla_lu_rowid_init_dev(int32_t *rowid, size_t n, void *stream);
RLA_API int rla_lu_rowid_apply_dev(const void *d_plan, int32_t *rowid, const int32_t *d_info, void *stream);
RLA_API int rla_lu_perm_from_rowid_dev(const int32_t *rowid, int64_t *d_perm, size_t n, const int32_t *d_info,
                               void *stream);

/* Seeded U[lo, lo+scale) fill, bit-identical to the test oracle's generator, so that every
 * rank can synthesise its shard in HBM without a transfer.  Element (i,j) of the rows x cols
 * matrix uses counter offset + (row0+i)*cols_total + (col0+j). */
RLA_API int rla_fill_uniform_f64_dev(double *dst, size_t rows, size_t cols, size_t ld, uint64_t seed,
                             uint64_t offset, double lo, double scale, void *stream);
RLA_API int rla_fill_uniform_f32_dev(float *dst, size_t rows, size_t cols, size_t ld, uint64_t seed,
                             uint64_t offset, float lo, float scale, void *stream);

/* Tuning knobs (development / benchmarking); defaults are the measured best.  Unknown keys: RLA_ERR_INVALID.
 *   "dgemm_cfg"        -1 = auto (default), 0..7 = a fixed CTA shape (csrc/dgemm.cu); "sgemm_cfg": -1 auto, 0 = 128x128, 1 = 64x128
 *   "dgemm_streamk"    0 (default) = tiled kernels only; 2 / 3 = the 128x128 / 64x64 stream-K variant whenever operands are aligned
 *   "lu_cluster"       LU panel kernel: 1 (default) = automatic (column-slab kernel up to "lu_slab_rows" rows, cluster kernel up
 *                      to 4096, grid kernel above); 0 = grid kernel only; 2 = pushed-row cluster kernel; 3 = column-slab kernel
 *                      wherever it fits; 4 = cluster + grid kernels only; 5 = grid kernel in cluster mode wherever it fits
 *   "lu_slab_rows"     tallest panel the automatic rule gives to the column-slab kernel (default 1920)
 *   "lu_k3e_rows"      tallest panel the automatic rule gives to the grid kernel's cluster mode (default 0 = never)
 *   "lu_gmax"          cap on the grid panel kernel's row CTAs (default 32)
 *   "host_gemm_2d"     1 (default) = 2-D wavefront pipeline for large host-pointer products, 0 = row panels
 *   "host_gemm_s"      strips per dimension of that pipeline, 0 (default) = auto
 *   "host_gemm_kprefix" sixteenths of the k range uploaded and multiplied as rank-kc updates before the wavefront starts;
 *                      -1 (default) = k/4 when k >= 4096, 0 = off;  "host_gemm_kchunk": kc (default 256)
 *   "host_gemm_grade"  1 = graded first / last strips in that pipeline (default 0: no measurable gain)
 *   "host_stage"       1 (default) = pageable host operands of large calls travel through the library's pinned staging ring
 *                      (host.cu), 0 = plain cudaMemcpyAsync on them
 *   "lu_dbg"           timing experiments only */
RLA_API int rla_set_tuning(const char *key, int value);

/* Roofline denominators measured on the current device: issue-bound register loops on every SM.
 * kind 0: FP64 tensor pipe (DMMA.8x8x4); kind 1: FP32 FMA pipe (FFMA).  ~30 ms. */
RLA_API int rla_measure_peak(int kind, double *tflops);

/* Diagnostics */
RLA_API const char *rla_strerror(int status);
RLA_API int rla_last_cuda_error(void);            /* cudaError_t of the last failing CUDA call (per thread) */
RLA_API const char *rla_version(void);
/* number of kernels this library has launched in the calling thread since the last reset */
RLA_API uint64_t rla_launch_count(void);
RLA_API void rla_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* RLA_B200_H */
