// sgemm.cu -- K2: FP32 GEMM on the FP32 FMA pipe.  No TF32, no tensor cores: the reference's
// matrixmultiply::sgemm (src/matrix/mat_mul.rs:33-43) is plain IEEE binary32 arithmetic and the
// accuracy class must match it (BASELINE.json north_star).
//
// Design (DESIGN.md "K2"):
//   CTA tile 128x128, k-slab 16, 256 threads, 8x8 register tile per thread laid out as 2x2 blocks
//   of 4x4 (rows ty*4+{0..3}, 64+ty*4+{0..3}; cols tx*4.., 64+tx*4..) so shared reads are LDS.128
//   and C stores are 256 contiguous bytes per half-row.  A and B slabs go global->shared with
//   16-byte cp.async in a 3-stage ring, A kept in its native [m][k] orientation (row padded to
//   20 floats so the two row groups of a warp land in different bank quads), B as [k][n].
//   The 8x8 tile is held as 32 packed float2 accumulators (pairs along n) and updated with the
//   Blackwell packed FP32 FMA  fma.rn.f32x2 (SASS FFMA2, scalar-broadcast A operand): two IEEE
//   binary32 FMAs per lane per instruction, which halves the issue-slot and register-port pressure
//   that capped the scalar-FFMA version at 58 % of the FMA pipe (ncu: dispatch stalls).  Per 4
//   k-steps a thread issues 8+8 LDS.128 for 128 FFMA2.  Two CTAs per SM (<=128 registers).
#include "common.cuh"

namespace rla {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, THREADS = 256;
constexpr int LDAS = BK + 4;   // floats
constexpr int LDBS = BN + 4;
constexpr int A_STAGE = BM * LDAS;
constexpr int B_STAGE = BK * LDBS;
constexpr size_t SMEM_BYTES = size_t(STAGES) * (A_STAGE + B_STAGE) * sizeof(float);
constexpr int BAND = 16;

// two independent IEEE fp32 FMAs: c.{x,y} = a.{x,y} * b.{x,y} + c.{x,y}
__device__ __forceinline__ void ffma2(unsigned long long &c, unsigned long long a, unsigned long long b) {
    // volatile: keeps the j-outer / i-inner issue order below, so the 64-bit B pair sits in the operand
    // reuse cache for 8 consecutive FFMA2 and the register file only supplies the accumulator pair + A scalar
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

template <bool ALIGNED>
__device__ __forceinline__ void load_slab(float *As, float *Bs, const float *__restrict__ A, size_t lda,
                                          const float *__restrict__ B, size_t ldb, int M, int N, int K,
                                          int m0, int n0, int k0, int tid) {
    if (ALIGNED) {
        const int ca = tid & 3, ra = tid >> 2;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = ra + 64 * i;
            const int gk = k0 + 4 * ca;
            int bytes = 0;
            const float *src = A;
            if (m0 + row < M && gk < K) {
                bytes = min(K - gk, 4) * 4;
                src = A + size_t(m0 + row) * lda + gk;
            }
            cp_async16(smem_u32(As + row * LDAS + 4 * ca), src, bytes);
        }
        const int cb = tid & 31, rb = tid >> 5;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = rb + 8 * i;
            const int gn = n0 + 4 * cb;
            int bytes = 0;
            const float *src = B;
            if (k0 + row < K && gn < N) {
                bytes = min(N - gn, 4) * 4;
                src = B + size_t(k0 + row) * ldb + gn;
            }
            cp_async16(smem_u32(Bs + row * LDBS + 4 * cb), src, bytes);
        }
    } else {
        const int ca = tid & 15, ra = tid >> 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = ra + 16 * i;
            const int gk = k0 + ca;
            const bool ok = (m0 + row < M) && (gk < K);
            cp_async4(smem_u32(As + row * LDAS + ca), ok ? A + size_t(m0 + row) * lda + gk : A, ok ? 4 : 0);
        }
        const int cb = tid & 127, rb = tid >> 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = rb + 2 * i;
            const int gn = n0 + cb;
            const bool ok = (k0 + row < K) && (gn < N);
            cp_async4(smem_u32(Bs + row * LDBS + cb), ok ? B + size_t(k0 + row) * ldb + gn : B, ok ? 4 : 0);
        }
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(THREADS, 2)
sgemm_ffma_kernel(int M, int N, int K, float alpha, const float *__restrict__ A, size_t lda,
                  const float *__restrict__ B, size_t ldb, float beta, float *__restrict__ C, size_t ldc,
                  int tiles_m, int tiles_n) {
    extern __shared__ __align__(16) float smem_f[];
    float *As = smem_f;
    float *Bs = smem_f + STAGES * A_STAGE;

    const int bid = blockIdx.x;
    const int per_band = BAND * tiles_n;
    const int band = bid / per_band;
    const int rem = bid - band * per_band;
    const int band_rows = min(BAND, tiles_m - band * BAND);
    const int tm = band * BAND + rem % band_rows;
    const int tn = rem / band_rows;
    const int m0 = tm * BM, n0 = tn * BN;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    unsigned long long acc[8][4];          // acc[i][j2] = (C[i][2*j2], C[i][2*j2+1])
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

    const int KT = (K + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_slab<ALIGNED>(As + s * A_STAGE, Bs + s * B_STAGE, A, lda, B, ldb, M, N, K, m0, n0, s * BK, tid);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int s = kt % STAGES;
        const float *ap = As + s * A_STAGE + (ty * 4) * LDAS;
        const float *bp = Bs + s * B_STAGE + tx * 4;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 a4[8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a4[r] = *reinterpret_cast<const float4 *>(ap + r * LDAS + kk);
                a4[4 + r] = *reinterpret_cast<const float4 *>(ap + (64 + r) * LDAS + kk);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(bp + (kk + q) * LDBS);
                const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(bp + (kk + q) * LDBS + 64);
                const unsigned long long bv[4] = {b0.x, b0.y, b1.x, b1.y};
                // j outer / i inner: the 64-bit B pair stays in the operand-reuse cache across the 8 rows, so each
                // FFMA2 fetches one scalar (A) and one pair (accumulator) from the register file
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float av = (q == 0) ? a4[i].x : (q == 1) ? a4[i].y : (q == 2) ? a4[i].z : a4[i].w;
                        ffma2(acc[i][j], pack2(av, av), bv[j]);   // ptxas folds the pack into the .F32 broadcast operand
                    }
                }
            }
            if (kk == 0) {
                // queue the copies for slab kt+STAGES-1 once the FMA pipe has work (same idea as dgemm.cu)
                const int nk = kt + STAGES - 1;
                if (nk < KT) {
                    const int ns = nk % STAGES;
                    load_slab<ALIGNED>(As + ns * A_STAGE, Bs + ns * B_STAGE, A, lda, B, ldb, M, N, K, m0, n0, nk * BK, tid);
                }
                cp_async_commit();
            }
        }
    }
    cp_async_wait<0>();

    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (row >= M) continue;
        float *crow = C + size_t(row) * ldc;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = n0 + h * 64 + tx * 4;
            if (col >= N) continue;
            const float2 p0 = unpack2(acc[i][h * 2]), p1 = unpack2(acc[i][h * 2 + 1]);
            float v[4] = {alpha * p0.x, alpha * p0.y, alpha * p1.x, alpha * p1.y};
            if (col + 3 < N && vec_ok) {
                if (beta != 0.f) {
                    const float4 old = *reinterpret_cast<const float4 *>(crow + col);
                    v[0] += beta * old.x; v[1] += beta * old.y; v[2] += beta * old.z; v[3] += beta * old.w;
                }
                *reinterpret_cast<float4 *>(crow + col) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (col + q < N) {
                        if (beta != 0.f) v[q] += beta * crow[col + q];
                        crow[col + q] = v[q];
                    }
                }
            }
        }
    }
}

}  // namespace

template <typename T>
int scale_c_launch(size_t m, size_t n, T beta, T *c, size_t ldc, cudaStream_t st);

int sgemm_launch(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda, const float *b,
                 size_t ldb, float beta, float *c, size_t ldc, cudaStream_t st) {
    if (m == 0 || n == 0) return RLA_OK;
    if (k == 0) return scale_c_launch<float>(m, n, beta, c, ldc, st);
    if (m > 0x7fffffffull || n > 0x7fffffffull || k > 0x7fffffffull) return RLA_ERR_INVALID;
    static bool attr_set[2] = {false, false};
    const bool aligned = ((lda & 3) == 0) && ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    const int tiles_m = int((m + BM - 1) / BM), tiles_n = int((n + BN - 1) / BN);
    const size_t tiles = size_t(tiles_m) * tiles_n;
    if (tiles > 0x7fffffffull) return RLA_ERR_INVALID;
    if (aligned) {
        if (!attr_set[1]) {
            RLA_CUDA(cudaFuncSetAttribute(sgemm_ffma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
            attr_set[1] = true;
        }
        sgemm_ffma_kernel<true><<<unsigned(tiles), THREADS, SMEM_BYTES, st>>>(int(m), int(n), int(k), alpha, a, lda, b, ldb,
                                                                              beta, c, ldc, tiles_m, tiles_n);
    } else {
        if (!attr_set[0]) {
            RLA_CUDA(cudaFuncSetAttribute(sgemm_ffma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
            attr_set[0] = true;
        }
        sgemm_ffma_kernel<false><<<unsigned(tiles), THREADS, SMEM_BYTES, st>>>(int(m), int(n), int(k), alpha, a, lda, b, ldb,
                                                                               beta, c, ldc, tiles_m, tiles_n);
    }
    RLA_LAUNCHED();
    return RLA_OK;
}

}  // namespace rla
