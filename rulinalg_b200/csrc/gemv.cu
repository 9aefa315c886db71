// gemv.cu -- y = A x for a row-major matrix (`&Matrix * &Vector`,
// src/matrix/impl_ops.rs:298-314, row-wise utils::dot).  HBM-bound: algorithmic bytes = sizeof(T)*(m*n + n + m).
// One warp per row; 16-byte loads of the row, x re-read through L1/L2 (it is tiny next to A); lane partial
// sums combined with xor shuffles.  Summation order differs from utils::dot's 8 partial sums => tolerance-based parity.
#include "common.cuh"

namespace rla {
namespace {

constexpr int GEMV_THREADS = 256;

template <typename T> struct V128;
template <> struct V128<double> { using type = double2; static constexpr int N = 2; };
template <> struct V128<float> { using type = float4; static constexpr int N = 4; };

__device__ __forceinline__ double dot128(const double2 &a, const double2 &b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float dot128(const float4 &a, const float4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

template <typename T>
__global__ void __launch_bounds__(GEMV_THREADS)
gemv_kernel(int m, int n, const T *__restrict__ A, size_t lda, const T *__restrict__ x, T *__restrict__ y) {
    using V = typename V128<T>::type;
    constexpr int VN = V128<T>::N;
    const int warp = (blockIdx.x * GEMV_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * GEMV_THREADS) >> 5;
    const bool vec_ok = (lda % VN == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int row = warp; row < m; row += nwarps) {
        const T *a = A + size_t(row) * lda;
        T acc0 = T(0), acc1 = T(0);
        int j = 0;
        if (vec_ok) {
            const int nv = n / VN;
            const V *av = reinterpret_cast<const V *>(a);
            const V *xv = reinterpret_cast<const V *>(x);
            int q = lane;
            T acc2 = T(0), acc3 = T(0);
            for (; q + 96 < nv; q += 128) {         // four independent 16-byte loads in flight per lane
                const V a0 = av[q], a1 = av[q + 32], a2 = av[q + 64], a3 = av[q + 96];
                acc0 += dot128(a0, xv[q]);
                acc1 += dot128(a1, xv[q + 32]);
                acc2 += dot128(a2, xv[q + 64]);
                acc3 += dot128(a3, xv[q + 96]);
            }
            for (; q < nv; q += 32) acc0 += dot128(av[q], xv[q]);
            acc0 += acc2;
            acc1 += acc3;
            j = nv * VN;
        }
        for (int c = j + lane; c < n; c += 32) acc0 += a[c] * x[c];
        T acc = acc0 + acc1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) y[row] = acc;
    }
}

}  // namespace

template <typename T>
int gemv_launch(size_t m, size_t n, const T *a, size_t lda, const T *x, T *y, cudaStream_t st) {
    if (m == 0) return RLA_OK;
    if (m > 0x7fffffffull || n > 0x7fffffffull) return RLA_ERR_INVALID;
    size_t blocks = (m * 32 + GEMV_THREADS - 1) / GEMV_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;      // 8 resident CTAs of 256 threads per SM, rows strided over the grid
    gemv_kernel<T><<<unsigned(blocks), GEMV_THREADS, 0, st>>>(int(m), int(n), a, lda, x, y);
    RLA_LAUNCHED();
    return RLA_OK;
}
template int gemv_launch<double>(size_t, size_t, const double *, size_t, const double *, double *, cudaStream_t);
template int gemv_launch<float>(size_t, size_t, const float *, size_t, const float *, float *, cudaStream_t);

}  // namespace rla
