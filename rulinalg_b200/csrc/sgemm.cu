// sgemm.cu -- K2: FP32 GEMM on the FP32 FMA pipe.  No TF32, no tensor cores: the reference's
// matrixmultiply::sgemm (src/matrix/mat_mul.rs:33-43) is plain IEEE binary32 arithmetic and the
// accuracy class must match it (BASELINE.json north_star).
//
// Design (DESIGN.md "K2"):
//   CTA tile 128x128, k-slab 16, 256 threads, 8x8 register tile per thread laid out as 2x2 blocks
//   of 4x4 (rows ty*4+{0..3}, 64+ty*4+{0..3}; cols tx*4.., 64+tx*4..) so shared reads are LDS.128
//   and C stores are 256 contiguous bytes per half-row.  Two CTAs per SM (<= 128 registers).
//   The 8x8 tile is held as 32 packed float2 accumulators (pairs along n) and updated with the
//   Blackwell packed FP32 FMA  fma.rn.f32x2 (SASS FFMA2, scalar-broadcast A operand that ptxas folds
//   from mov.b64 {x,x}): two IEEE binary32 FMAs per lane per instruction, halving issue-slot and
//   register-port pressure (the scalar-FFMA version sat at 58 % of the FMA pipe with 1.2 dispatch-stall
//   cycles per issue).
//   A is stored k-major in shared memory (As[k][m], transposed on the way in: LDG.128 along k -> 4
//   STS.32) so one k-step needs only 2+2 LDS.128 = 16 fragment registers; that leaves room to
//   DOUBLE-BUFFER the fragments (load k+1 while the FFMA2s of k issue) inside the 128-register budget.
//   The first FFMA2 version kept A as [m][k] float4 fragments (32 live registers, no double buffering):
//   every warp alternated LDS bursts and FFMA2 bursts and the FMA pipe idled 31 % of the time.
//   Global loads for slab kt+1 are issued before slab kt's math and parked in registers (A) / cp.async (B).
#include "common.cuh"

namespace rla {
namespace {

// Two CTA shapes from one template: BM = 128 (256 threads, the asymptotic shape) and BM = 64 (128 threads: twice the
// tiles for products that cannot fill 148 SMs with 128 x 128 tiles -- 1024^3 is 64 of those).  BN = 128, BK = 16, 8x8 per
// thread in both.
constexpr int BN = 128, BK = 16;
constexpr int LDT = 128 + 4;   // padded row of the k-major A tile and of the B tile (floats)
constexpr int TILE = BK * LDT; // floats per operand per stage
constexpr int STAGES = 2;
constexpr size_t SMEM_BYTES = size_t(STAGES) * 2 * TILE * sizeof(float);
constexpr int BAND = 16;

// two independent IEEE fp32 FMAs: c.{x,y} = a.{x,y} * b.{x,y} + c.{x,y}
__device__ __forceinline__ void ffma2(unsigned long long &c, unsigned long long a, unsigned long long b) {
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

struct Frag {
    float4 a0, a1;            // A[k][ty*4..+3], A[k][64+ty*4..+3]
    ulonglong2 b0, b1;        // B[k][tx*4..+3] as two float2 pairs, B[k][64+tx*4..+3]
};
template <int HALF_M>
__device__ __forceinline__ void load_frag(Frag &f, const float *ap, const float *bp, int kk) {
    f.a0 = *reinterpret_cast<const float4 *>(ap + kk * LDT);
    f.a1 = *reinterpret_cast<const float4 *>(ap + kk * LDT + HALF_M);
    f.b0 = *reinterpret_cast<const ulonglong2 *>(bp + kk * LDT);
    f.b1 = *reinterpret_cast<const ulonglong2 *>(bp + kk * LDT + 64);
}
__device__ __forceinline__ void mma_frag(unsigned long long (&acc)[8][4], const Frag &f) {
    const float av[8] = {f.a0.x, f.a0.y, f.a0.z, f.a0.w, f.a1.x, f.a1.y, f.a1.z, f.a1.w};
    const unsigned long long bv[4] = {f.b0.x, f.b0.y, f.b1.x, f.b1.y};
    // j outer / i inner, kept in this order (volatile asm): the 64-bit B pair stays in the operand-reuse cache for 8
    // consecutive FFMA2, so each one fetches only the A scalar and the accumulator pair from the register file
    // (<= 2 reads per bank = the pipe's 2 cycles).  ptxas' own order re-read B pairs: 2.25-2.5 cycles per FFMA2.
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(acc[i][j], pack2(av[i], av[i]), bv[j]);
    }
}

// ACC_C: the accumulators start from C and the epilogue stores alpha * acc (C_out = alpha * (C_in + A*B)): consecutive k-chunks
// of one product issued this way repeat, per element, the FMA chain of the one-call product bit for bit (see dgemm.cu; used by
// the host pipeline's k-prefix).
template <bool ALIGNED, int BM, bool ACC_C = false>
__global__ void __launch_bounds__(2 * BM, BM == 128 ? 2 : 4)
sgemm_ffma_kernel(int M, int N, int K, float alpha, const float *__restrict__ A, size_t lda,
                  const float *__restrict__ B, size_t ldb, float beta, float *__restrict__ C, size_t ldc,
                  int tiles_m, int tiles_n) {
    constexpr int THREADS = 2 * BM, HALF_M = BM / 2;
    extern __shared__ __align__(16) float smem_f[];
    float *As = smem_f;                       // [STAGES][BK][LDT]  k-major
    float *Bs = smem_f + STAGES * TILE;       // [STAGES][BK][LDT]

    const int bid = blockIdx.x;
    const int per_band = BAND * tiles_n;
    const int band = bid / per_band;
    const int rem = bid - band * per_band;
    const int band_rows = min(BAND, tiles_m - band * BAND);
    const int tm = band * BAND + rem % band_rows;
    const int tn = rem / band_rows;
    const int m0 = tm * BM, n0 = tn * BN;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;      // ty < BM / 8: rows ty*4.. and BM/2 + ty*4..
    // (measured and dropped: lanes as 4 ty x 8 tx -- a 32 x 64 warp tile, 4 instead of 6-8 shared-memory wavefronts per
    // k-step -- 58.5 vs 59.5 TFLOP/s at 8192^3: shared-memory traffic is not what holds the FMA pipe at 80 %)

    // global->shared copy roles.  A: thread owns k-quad (tid&3) of rows (tid>>2) and (tid>>2)+BM/2.
    const int a_kq = (tid & 3) * 4, a_row = tid >> 2;
    // B: thread owns float4 column chunk (tid&31) of k rows (tid>>5) + i * THREADS/32.
    constexpr int B_ROWS = THREADS / 32, B_PASSES = BK / B_ROWS;
    const int b_c4 = (tid & 31) * 4, b_row = tid >> 5;
    const bool a_ok0 = m0 + a_row < M, a_ok1 = m0 + a_row + HALF_M < M;
    const float *a_src0 = A + size_t(a_ok0 ? m0 + a_row : 0) * lda + a_kq;
    const float *a_src1 = A + size_t(a_ok1 ? m0 + a_row + HALF_M : 0) * lda + a_kq;
    const int b_gn = n0 + b_c4;
    const float *b_src = B + size_t(b_row) * ldb + (b_gn < N ? b_gn : 0);
    const int b_bytes_full = b_gn + 3 < N ? 16 : (b_gn < N ? (N - b_gn) * 4 : 0);

    // The slab pointers advance by constant strides (64-bit adds): base + k0 arithmetic compiles to IMAD.WIDE / IMAD,
    // which execute on the same FMA pipe the FFMA2s need (ncu: ~15 per slab per warp, with pipe-throttle stalls).
    const float *pa0 = a_src0, *pa1 = a_src1;                 // slab being loaded: A columns k0 + a_kq ..
    const float *pb = b_src;                                  //                    B rows k0 + b_row (+8)
    const size_t b_step = size_t(BK) * ldb, b_pass = size_t(B_ROWS) * ldb;
    auto load_a = [&](int k0, float4 &r0, float4 &r1) {
        r0 = make_float4(0.f, 0.f, 0.f, 0.f);
        r1 = r0;
        const int gk = k0 + a_kq;
        if (ALIGNED && gk + 3 < K) {
            if (a_ok0) r0 = *reinterpret_cast<const float4 *>(pa0);
            if (a_ok1) r1 = *reinterpret_cast<const float4 *>(pa1);
        } else {
            float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (gk + q < K) {
                    if (a_ok0) t0[q] = pa0[q];
                    if (a_ok1) t1[q] = pa1[q];
                }
            r0 = make_float4(t0[0], t0[1], t0[2], t0[3]);
            r1 = make_float4(t1[0], t1[1], t1[2], t1[3]);
        }
        pa0 += BK;
        pa1 += BK;
    };
    // Transposing stores, conflict-free: k rows 8..15 keep their m index XOR 8 (a swizzle at the granularity of two
    // float4 fragments, so the LDS.128 fragment loads below stay aligned and only pick one of two base pointers).  With
    // the plain layout the four k-quads of a warp sat 16*kq banks apart -- quads 0/2 and 1/3 on the SAME banks: every
    // STS.32 was a 2-way conflict (ncu r1: 136 M conflicts at n = 8192, shared-memory wavefronts at 67 % of peak).
    auto store_a = [&](float *as, const float4 &r0, const float4 &r1) {
        float *p = as + a_kq * LDT + (a_row ^ (a_kq & 8));
        p[0] = r0.x; p[LDT] = r0.y; p[2 * LDT] = r0.z; p[3 * LDT] = r0.w;
        p[HALF_M] = r1.x; p[LDT + HALF_M] = r1.y; p[2 * LDT + HALF_M] = r1.z; p[3 * LDT + HALF_M] = r1.w;
    };
    auto copy_b = [&](float *bs, int k0) {
#pragma unroll
        for (int i = 0; i < B_PASSES; ++i) {
            const int row = b_row + B_ROWS * i;
            const bool ok = k0 + row < K;
            if (ALIGNED) {
                const int bytes = ok ? b_bytes_full : 0;
                cp_async16(smem_u32(bs + row * LDT + b_c4), bytes ? pb + i * b_pass : B, bytes);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool okq = ok && (b_gn + q < N);
                    cp_async4(smem_u32(bs + row * LDT + b_c4 + q), okq ? B + size_t(k0 + row) * ldb + b_gn + q : B, okq ? 4 : 0);
                }
            }
        }
        pb += b_step;
    };

    unsigned long long acc[8][4];          // acc[i][j2] = (C[i][2*j2], C[i][2*j2+1])
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

    const int KT = (K + BK - 1) / BK;
    {   // prologue: slab 0 into stage 0
        float4 r0, r1;
        load_a(0, r0, r1);
        copy_b(Bs, 0);
        cp_async_commit();
        if (ACC_C) {
            const bool cvec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = m0 + ((i < 4) ? ty * 4 + i : HALF_M + ty * 4 + (i - 4));
                if (row >= M) continue;
                const float *crow = C + size_t(row) * ldc;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = n0 + h * 64 + tx * 4;
                    if (col >= N) continue;
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    if (col + 3 < N && cvec) {
                        const float4 old = *reinterpret_cast<const float4 *>(crow + col);
                        v[0] = old.x; v[1] = old.y; v[2] = old.z; v[3] = old.w;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (col + q < N) v[q] = crow[col + q];
                    }
                    acc[i][h * 2] = pack2(v[0], v[1]);
                    acc[i][h * 2 + 1] = pack2(v[2], v[3]);
                }
            }
        }
        store_a(As, r0, r1);
        cp_async_wait<0>();
        __syncthreads();
    }

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt & 1;
        const float *ap = (s ? As + TILE : As) + ty * 4;
        const float *ap8 = (s ? As + TILE : As) + ((ty * 4) ^ 8);     // k rows 8..15 (see store_a)
        const float *bp = (s ? Bs + TILE : Bs) + tx * 4;
        const bool more = kt + 1 < KT;
        float4 r0, r1;
        Frag f[2];
        load_frag<HALF_M>(f[0], ap, bp, 0);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            if (kk + 1 < BK) load_frag<HALF_M>(f[(kk + 1) & 1], (kk + 1) & 8 ? ap8 : ap, bp, kk + 1);
            mma_frag(acc, f[kk & 1]);
            // slab kt+1 (A parked in registers, B by cp.async into the other stage) is requested AFTER the first k-step's
            // FFMA2s are in the pipe: right after the barrier every warp of the CTA would otherwise run ~100 address /
            // predicate / copy-issue instructions with the FMA pipe idle (same fix as in the DGEMM kernel)
            if (kk == 0 && more) {
                load_a((kt + 1) * BK, r0, r1);
                copy_b(s ? Bs : Bs + TILE, (kt + 1) * BK);
                cp_async_commit();
            }
        }
        if (more) {
            store_a(s ? As : As + TILE, r0, r1);
            cp_async_wait<0>();
        }
        __syncthreads();
    }

    const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + ((i < 4) ? ty * 4 + i : HALF_M + ty * 4 + (i - 4));
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

template <bool ALIGNED, int BM, bool ACC_C = false>
int sgemm_launch_cfg(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda, const float *b, size_t ldb,
                     float beta, float *c, size_t ldc, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int tiles_m = int((m + BM - 1) / BM), tiles_n = int((n + BN - 1) / BN);
    const size_t tiles = size_t(tiles_m) * tiles_n;
    if (tiles > 0x7fffffffull) return RLA_ERR_INVALID;
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(sgemm_ffma_kernel<ALIGNED, BM, ACC_C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        attr_once.done(od_);
    }
    sgemm_ffma_kernel<ALIGNED, BM, ACC_C><<<unsigned(tiles), 2 * BM, SMEM_BYTES, st>>>(int(m), int(n), int(k), alpha, a, lda, b, ldb, beta,
                                                                                       c, ldc, tiles_m, tiles_n);
    RLA_LAUNCHED();
    return RLA_OK;
}

int g_sgemm_cfg = -1;    // rla_set_tuning("sgemm_cfg", v): -1 auto, 0 = 128 x 128 tiles, 1 = 64 x 128 tiles

int sgemm_launch(size_t m, size_t k, size_t n, float alpha, const float *a, size_t lda, const float *b,
                 size_t ldb, float beta, float *c, size_t ldc, cudaStream_t st, bool acc_from_c) {
    if (m == 0 || n == 0) return RLA_OK;
    if (acc_from_c) {
        if (k == 0) return scale_c_launch<float>(m, n, alpha, c, ldc, st);
        if (m > 0x7fffffffull || n > 0x7fffffffull || k > 0x7fffffffull) return RLA_ERR_INVALID;
        if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) return RLA_ERR_INVALID;
        const size_t t128 = ((m + 127) / 128) * ((n + 127) / 128);
        return t128 * 10 < size_t(device_num_sms()) * 9 ? sgemm_launch_cfg<true, 64, true>(m, k, n, alpha, a, lda, b, ldb, 0.f, c, ldc, st)
                                                          : sgemm_launch_cfg<true, 128, true>(m, k, n, alpha, a, lda, b, ldb, 0.f, c, ldc, st);
    }
    if (k == 0) return scale_c_launch<float>(m, n, beta, c, ldc, st);
    if (m > 0x7fffffffull || n > 0x7fffffffull || k > 0x7fffffffull) return RLA_ERR_INVALID;
    const bool aligned = ((lda & 3) == 0) && ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    // fewer 128 x 128 tiles than SMs: the 64 x 128 shape (twice the tiles) wins -- measured (tools/gemm_sweep.py s):
    // 1024^3 33.2 vs 20.7 TFLOP/s, 512^3 7.0 vs 4.7; from 144 tiles (1536^3) on the big tile is ahead (49.2 vs 48.1)
    int cfg = g_sgemm_cfg;
    if (cfg < 0) {
        const size_t t128 = ((m + 127) / 128) * ((n + 127) / 128);
        cfg = t128 * 10 < size_t(device_num_sms()) * 9 ? 1 : 0;
    }
    if (cfg == 1)
        return aligned ? sgemm_launch_cfg<true, 64>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st)
                       : sgemm_launch_cfg<false, 64>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    return aligned ? sgemm_launch_cfg<true, 128>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st)
                   : sgemm_launch_cfg<false, 128>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
}

}  // namespace rla
