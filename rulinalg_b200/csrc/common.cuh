// common.cuh -- shared declarations for librla_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <functional>

#include "../../include/rla_b200.h"

namespace rla {

// ---- per-thread error / launch accounting (api.cu) ---------------------------------------
void note_cuda_error(cudaError_t e);
void note_launch(unsigned n = 1);
uint64_t launch_count_take();        // calling thread's launch count since its last take / reset (worker threads hand it to their parent)

#define RLA_CUDA(expr)                                  \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) {                        \
            ::rla::note_cuda_error(_e);                 \
            return (_e == cudaErrorMemoryAllocation) ? RLA_ERR_NOMEM : RLA_ERR_CUDA; \
        }                                               \
    } while (0)

#define RLA_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != RLA_OK) return _s;  \
    } while (0)

// Checks the launch just issued and counts it.
#define RLA_LAUNCHED()                                  \
    do {                                                \
        cudaError_t _e = cudaGetLastError();            \
        if (_e != cudaSuccess) {                        \
            ::rla::note_cuda_error(_e);                 \
            return RLA_ERR_CUDA;                        \
        }                                               \
        ::rla::note_launch();                           \
    } while (0)

// Per-device one-time setup.  cudaFuncSetAttribute (dynamic shared memory opt-in, cluster opt-in) and
// occupancy answers are PER DEVICE: a process that drives several GPUs (rla_set_devices, or host threads
// that rla_init different devices) must repeat them on each one.  Usage:
//     static DeviceOnce once;  if (const int d = once.pending(); d >= 0) { ...cudaFuncSetAttribute...; once.done(d); }
// Two threads racing on the same device both run the (idempotent) setup; nobody skips it.
constexpr int RLA_MAX_DEVICES = 32;
inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { (void)cudaGetLastError(); d = 0; }
    return d & (RLA_MAX_DEVICES - 1);
}
struct DeviceOnce {
    std::atomic<uint32_t> mask{0};
    int pending() {              // current device id if its setup has not run yet, else -1
        const int d = current_device();
        return (mask.load(std::memory_order_acquire) & (1u << d)) ? -1 : d;
    }
    void done(int d) { mask.fetch_or(1u << d, std::memory_order_release); }
};
int device_num_sms();            // lu.cu: SM count of the current device (cached per device)

// ---- kernel launchers (one per .cu) ---------------------------------------------------------
// acc_from_c: C <- alpha * (C + A*B) with the sum continued in the accumulators (beta ignored): consecutive k-chunks of one
// product issued this way are bit-identical to the one-call product (dgemm.cu)
int dgemm_launch(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda,
                 const double *b, size_t ldb, double beta, double *c, size_t ldc,
                 cudaStream_t st, bool acc_from_c = false);
int sgemm_launch(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda,
                 const float *b, size_t ldb, float beta, float *c, size_t ldc, cudaStream_t st, bool acc_from_c = false);

// LU workspace: owned by the per-thread context, grown on demand.
struct LuWorkspace {
    int32_t *ipiv = nullptr;      // [n] pivot row chosen at each column (LAPACK-style, 0-based)
    void *scratch = nullptr;      // panel scratch (candidates, barrier words, row buffers)
    size_t ipiv_cap = 0, scratch_cap = 0;
    unsigned tag = 0;             // next free packet tag (unique per panel column while scratch lives)
    void *slab = nullptr;         // column-slab panel kernel: [64 headers | ticket | multiplier columns 64 x 4096 x 8 B]
    unsigned slab_epoch = 0, slab_tickets = 0;   // launch counter (header validity) and CTAs launched so far (ticket base)
    cudaStream_t side = nullptr;  // high-priority stream for the look-ahead panel factorisation
    cudaEvent_t ev_head = nullptr, ev_fact = nullptr;
};
// optional hook of getrf_launch: called (at enqueue time) when rows [row0, row0 + nrows) of the packed factors can no
// longer change -- everything queued on `st` so far produces them; the host API starts their download there
using LuRowsFinal = std::function<int(int row0, int nrows, cudaStream_t st)>;
// optional hook of getrf_launch: called once, right after the first outer block's factorisation has been enqueued -- the
// only part of the algorithm that touches nothing but the first 256 columns; the host API makes `st` wait there for the
// upload of the remaining columns, which so runs under that block's panel kernels
using LuAfterFirstBlock = std::function<int(cudaStream_t st)>;
template <typename T>
int getrf_launch(size_t n, T *a, size_t ld, int64_t *d_perm, int32_t *d_info, LuWorkspace &ws,
                 cudaStream_t st, const LuRowsFinal *rows_final = nullptr, const LuAfterFirstBlock *after_first = nullptr);
template <typename T>
int getrs_launch(size_t n, const T *lu, size_t ld, const int64_t *d_perm, T *d_b, T *ws3,
                 int32_t *d_info, int32_t *d_flags, cudaStream_t st);

template <typename T>
int getri_launch(size_t n, const T *lu, size_t ld, const int64_t *d_perm, T *x, size_t ldx, int32_t *d_info,
                 cudaStream_t st);
template <typename T>
int getri_small_launch(int n, const T *lu, size_t ld, const int64_t *d_perm, T *x, size_t ldx, int32_t *d_info,
                       cudaStream_t st);
template <typename T>
int trsv_launch(bool lower, size_t n, const T *a, size_t ld, T *d_x, T *ws2, int32_t *d_info, int32_t *d_sync, cudaStream_t st);
// Cholesky (cholesky.cu): in-place lower factor, solve, inverse.  info: 0 / j+1 (singular) / -(j+1) (negative diagonal).
size_t potrf_workspace_elems(size_t n);
template <typename T>
int potrf_launch(size_t n, T *a, size_t ld, T *ws, int32_t *d_info, cudaStream_t st);
template <typename T>
int potrs_launch(size_t n, const T *l, size_t ld, T *d_b, T *lt, T *ws2, int32_t *d_info, int32_t *d_info2, int32_t *d_sync,
                 cudaStream_t st);
template <typename T>
int potri_launch(size_t n, const T *l, size_t ld, T *x, size_t ldx, T *m, int64_t *d_perm, int32_t *d_info, cudaStream_t st);
template <typename T>
int gemv_launch(size_t m, size_t n, const T *a, size_t lda, const T *x, T *y, cudaStream_t st);
size_t lu_plan_bytes();
size_t lu_max_n(size_t elem_size);
int lu_divcheck(int f32, int mode, unsigned long long seed, unsigned long long count, unsigned long long *mismatches);
int lu_trace_fetch(unsigned long long *host512);
template <typename T>
int lu_factor_block_dev(int n, T *a_loc, size_t ld, int row0, int lcol0, int w, int32_t *d_info, void *d_plan,
                        LuWorkspace &ws, cudaStream_t st);
template <typename T>
int lu_laswp_dev(T *a_loc, size_t ld, int w, const void *d_plan, const int32_t *d_info, int c0a, int c1a, int c0b,
                 int c1b, cudaStream_t st);
template <typename T>
int lu_update_dev(int n, T *a_loc, size_t ld, int row0, int w, const T *panel, size_t ldp, int c0, int c1,
                  const int32_t *d_info, cudaStream_t st);
int lu_rowid_init_dev(int32_t *rowid, int n, cudaStream_t st);
int lu_rowid_apply_dev(const void *d_plan, int32_t *rowid, const int32_t *d_info, cudaStream_t st);
int lu_perm_from_rowid_dev(const int32_t *rowid, int64_t *perm, int n, const int32_t *d_info, cudaStream_t st);

template <typename T>
int fill_uniform_launch(T *dst, size_t rows, size_t cols, size_t ld, uint64_t seed,
                        uint64_t offset, T lo, T scale, cudaStream_t st);

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte async copy global->shared; bytes beyond src_bytes are zero-filled (src_bytes in {0,8,16}).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = 4*g + t): A[g][t], B[t][g], C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace rla
