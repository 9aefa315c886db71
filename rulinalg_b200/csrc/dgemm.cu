// dgemm.cu -- K1: FP64 GEMM on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// Serves rla_dgemm / rla_dgemm_dev, i.e. the matrixmultiply::dgemm call at
// src/matrix/mat_mul.rs:57-67 (C = alpha*A*B + beta*C, row-major, unit column strides), and the
// trailing update of the blocked LU (lu.cu).
//
// Design (DESIGN.md "K1"):
//   CTA tile 128x128, k-slab 16, 256 threads = 8 warps as 2(m) x 4(n), warp tile 64x32 =
//   8x4 DMMA fragments -> 64 accumulator doubles (128 registers) per thread.  A (k contiguous)
//   and B (n contiguous) slabs are staged global->shared with 16-byte cp.async (LDGSTS) in a
//   4-stage ring, one __syncthreads per slab.  Shared rows are padded by 4 doubles so that the
//   8-byte fragment loads of a half-warp hit 32 distinct banks for both operands:
//     A frag  As[row=g][k=t]   : word = row*40 + 2k   -> bank 8g+2t (+0/1)
//     B frag  Bs[k=t][col=g]   : word = k*264 + 2col  -> bank 8t+2g (+0/1)
//   Tiles are rasterised in bands of 16 tile-rows so that the 148 co-resident CTAs share A and B
//   slabs through the 126 MB L2 (HBM traffic ~ compulsory; see profiles/).
//   beta == 0 never reads C (mat_mul.rs:52-55 hands over uninitialised memory).
#include "common.cuh"

namespace rla {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int LDAS = BK + 4;             // padded A row (doubles)
constexpr int LDBS = BN + 4;             // padded B row (doubles)
constexpr int A_STAGE = BM * LDAS;       // doubles per stage
constexpr int B_STAGE = BK * LDBS;
constexpr size_t SMEM_BYTES = size_t(STAGES) * (A_STAGE + B_STAGE) * sizeof(double);
constexpr int BAND = 16;                 // tile-rows per raster band

template <bool ALIGNED>
__device__ __forceinline__ void load_slab(double *As, double *Bs, const double *__restrict__ A,
                                          size_t lda, const double *__restrict__ B, size_t ldb,
                                          int M, int N, int K, int m0, int n0, int k0, int tid) {
    if (ALIGNED) {
        // A: 128 rows x 8 chunks of 2 doubles
        const int ca = tid & 7, ra = tid >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = ra + 32 * i;
            const int gk = k0 + 2 * ca;
            int bytes = 0;
            const double *src = A;
            if (m0 + row < M && gk < K) {
                bytes = (K - gk >= 2) ? 16 : 8;
                src = A + size_t(m0 + row) * lda + gk;
            }
            cp_async16(smem_u32(As + row * LDAS + 2 * ca), src, bytes);
        }
        // B: 16 rows x 64 chunks
        const int cb = tid & 63, rb = tid >> 6;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = rb + 4 * i;
            const int gn = n0 + 2 * cb;
            int bytes = 0;
            const double *src = B;
            if (k0 + row < K && gn < N) {
                bytes = (N - gn >= 2) ? 16 : 8;
                src = B + size_t(k0 + row) * ldb + gn;
            }
            cp_async16(smem_u32(Bs + row * LDBS + 2 * cb), src, bytes);
        }
    } else {
        // 8-byte path for odd leading dimensions / unaligned bases
        const int ca = tid & 15, ra = tid >> 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = ra + 16 * i;
            const int gk = k0 + ca;
            const bool ok = (m0 + row < M) && (gk < K);
            const double *src = ok ? A + size_t(m0 + row) * lda + gk : A;
            cp_async8(smem_u32(As + row * LDAS + ca), src, ok ? 8 : 0);
        }
        const int cb = tid & 127, rb = tid >> 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = rb + 2 * i;
            const int gn = n0 + cb;
            const bool ok = (k0 + row < K) && (gn < N);
            const double *src = ok ? B + size_t(k0 + row) * ldb + gn : B;
            cp_async8(smem_u32(Bs + row * LDBS + cb), src, ok ? 8 : 0);
        }
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(THREADS, 1)
dgemm_dmma_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, size_t lda,
                  const double *__restrict__ B, size_t ldb, double beta, double *__restrict__ C,
                  size_t ldc, int tiles_m, int tiles_n) {
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + STAGES * A_STAGE;

    // band-rasterised tile order
    const int bid = blockIdx.x;
    const int per_band = BAND * tiles_n;
    const int band = bid / per_band;
    const int rem = bid - band * per_band;
    const int band_rows = min(BAND, tiles_m - band * BAND);
    const int tm = band * BAND + rem % band_rows;
    const int tn = rem / band_rows;
    const int m0 = tm * BM, n0 = tn * BN;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int KT = (K + BK - 1) / BK;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_slab<ALIGNED>(As + s * A_STAGE, Bs + s * B_STAGE, A, lda, B, ldb, M, N, K, m0, n0, s * BK, tid);
        cp_async_commit();
    }

    const double *a_frag_base = As + (wm * 64 + g) * LDAS + t;
    const double *b_frag_base = Bs + t * LDBS + wn * 32 + g;

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) {
                const int s = nk % STAGES;
                load_slab<ALIGNED>(As + s * A_STAGE, Bs + s * B_STAGE, A, lda, B, ldb, M, N, K, m0, n0, nk * BK, tid);
            }
            cp_async_commit();
        }
        const int s = kt % STAGES;
        const double *ap = a_frag_base + s * A_STAGE;
        const double *bp = b_frag_base + s * B_STAGE;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = ap[i * 8 * LDAS + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = bp[kk * LDBS + j * 8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[row g][cols 2t,2t+1] of each 8x8 fragment
    const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + wm * 64 + i * 8 + g;
        if (row >= M) continue;
        double *crow = C + size_t(row) * ldc;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + 2 * t;
            if (col >= N) continue;
            double v0 = alpha * acc[i][j][0];
            double v1 = alpha * acc[i][j][1];
            if (col + 1 < N && vec_ok) {
                if (beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2 *>(crow + col);
                    v0 += beta * old.x;
                    v1 += beta * old.y;
                }
                *reinterpret_cast<double2 *>(crow + col) = make_double2(v0, v1);
            } else {
                if (beta != 0.0) v0 += beta * crow[col];
                crow[col] = v0;
                if (col + 1 < N) {
                    if (beta != 0.0) v1 += beta * crow[col + 1];
                    crow[col + 1] = v1;
                }
            }
        }
    }
}

// k == 0 (or alpha == 0 shortcut not used): C <- beta*C, zero-fill when beta == 0 without reading C.
template <typename T>
__global__ void scale_c_kernel(size_t M, size_t N, T beta, T *C, size_t ldc) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const size_t r = idx / N, c = idx - r * N;
    T *p = C + r * ldc + c;
    *p = (beta == T(0)) ? T(0) : (*p) * beta;
}

}  // namespace

template <typename T>
int scale_c_launch(size_t m, size_t n, T beta, T *c, size_t ldc, cudaStream_t st) {
    const size_t total = m * n;
    if (total == 0) return RLA_OK;
    const unsigned blocks = unsigned((total + 255) / 256);
    scale_c_kernel<T><<<blocks, 256, 0, st>>>(m, n, beta, c, ldc);
    RLA_LAUNCHED();
    return RLA_OK;
}
template int scale_c_launch<double>(size_t, size_t, double, double *, size_t, cudaStream_t);
template int scale_c_launch<float>(size_t, size_t, float, float *, size_t, cudaStream_t);

int dgemm_launch(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda,
                 const double *b, size_t ldb, double beta, double *c, size_t ldc,
                 cudaStream_t st) {
    if (m == 0 || n == 0) return RLA_OK;
    if (k == 0) return scale_c_launch<double>(m, n, beta, c, ldc, st);
    if (m > 0x7fffffffull || n > 0x7fffffffull || k > 0x7fffffffull) return RLA_ERR_INVALID;

    static bool attr_set[2] = {false, false};
    const bool aligned = ((lda & 1) == 0) && ((ldb & 1) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    const int tiles_m = int((m + BM - 1) / BM), tiles_n = int((n + BN - 1) / BN);
    const size_t tiles = size_t(tiles_m) * tiles_n;
    if (tiles > 0x7fffffffull) return RLA_ERR_INVALID;
    if (aligned) {
        if (!attr_set[1]) {
            RLA_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
            attr_set[1] = true;
        }
        dgemm_dmma_kernel<true><<<unsigned(tiles), THREADS, SMEM_BYTES, st>>>(
            int(m), int(n), int(k), alpha, a, lda, b, ldb, beta, c, ldc, tiles_m, tiles_n);
    } else {
        if (!attr_set[0]) {
            RLA_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
            attr_set[0] = true;
        }
        dgemm_dmma_kernel<false><<<unsigned(tiles), THREADS, SMEM_BYTES, st>>>(
            int(m), int(n), int(k), alpha, a, lda, b, ldb, beta, c, ldc, tiles_m, tiles_n);
    }
    RLA_LAUNCHED();
    return RLA_OK;
}

}  // namespace rla
