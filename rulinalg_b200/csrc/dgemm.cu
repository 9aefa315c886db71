// dgemm.cu -- K1: FP64 GEMM on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// Serves rla_dgemm / rla_dgemm_dev, i.e. the matrixmultiply::dgemm call at
// src/matrix/mat_mul.rs:57-67 (C = alpha*A*B + beta*C, row-major, unit column strides), and the
// trailing update of the blocked LU (lu.cu).
//
// Design (DESIGN.md "K1"):
//   Warp tile 64x32 = 8x4 DMMA fragments -> 64 accumulator doubles (128 registers) per thread.
//   Two CTA shapes, same code:
//     cfg 0  128x64  tile, 4 warps (2x2), 3-stage ring, TWO CTAs per SM: while one CTA sits at its
//            slab barrier, issues its cp.async or runs its epilogue, the other keeps the DMMA pipe fed.
//     cfg 1  128x128 tile, 8 warps (2x4), 4-stage ring, one CTA per SM.
//     cfg 2  as cfg 1 with k-slab 32 and a 3-stage ring (half the barriers; +0.4 % on long-k products).
//     cfg 5/6 128x128 tile, 16 warps (4x4) with 32x32 warp tiles (4 warps per scheduler, 120 registers), k-slab
//            16 / 32.  cfg 6 is the fastest long-k shape by a small margin (34.5 vs 34.3 TFLOP/s at n = 8192):
//            every tiling lands at 92-93 % of the DMMA issue peak.
//   A (k contiguous) and B (n contiguous) k-slabs of 16 are staged global->shared with 16-byte
//   cp.async (LDGSTS), one __syncthreads per slab; the copies for slab kt+S-1 are issued AFTER the
//   first quarter of slab kt's DMMAs so the tensor pipe never waits on address arithmetic.  Source
//   pointers/predicates are hoisted: full slabs advance by a constant stride, only the last
//   (ragged) slab re-evaluates k bounds.
//   Shared rows are padded by 4 doubles so the 8-byte fragment loads of a half-warp hit 32 distinct
//   banks for both operands (ncu: 0 bank conflicts):
//     A frag  As[row=g][k=t]   : word = row*40 + 2k        -> bank 8g+2t (+0/1)
//     B frag  Bs[k=t][col=g]   : word = k*(2*BN+8) + 2col  -> bank 8t+2g (+0/1)
//   Tiles are rasterised in bands of 16 tile-rows so co-resident CTAs share A and B slabs in L2.
//   beta == 0 never reads C (mat_mul.rs:52-55 hands over uninitialised memory); beta != 0 loads C in
//   batches of 8 independent 16-byte loads before the FMAs (rank-k updates are epilogue-heavy).
#include <mutex>

#include "common.cuh"

namespace rla {
namespace {

constexpr int BAND = 16;                 // tile-rows per raster band

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_, int BK_ = 16>
struct Cfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_, BK = BK_;
    static constexpr int THREADS = WM * WN * 32;
    static constexpr int LDAS = BK + 4;  // padded A row (doubles): word stride 2*BK+8 == 8 (mod 32) for BK in {16,32}
    static constexpr int LDBS = BN + 4;
    static constexpr int A_STAGE = BM * LDAS;
    static constexpr int B_STAGE = BK * LDBS;
    static constexpr size_t SMEM = size_t(STAGES) * (A_STAGE + B_STAGE) * sizeof(double);
    static constexpr int A_CHUNKS = BM * (BK / 2) / THREADS;       // 16-byte chunks per thread per slab
    static constexpr int B_CHUNKS = BK * (BN / 2) / THREADS;
    static constexpr int MT = BM / WM / 8, NT = BN / WN / 8;   // 8x8 DMMA fragments per warp tile (8x4 = 64x32 or 4x4 = 32x32)
    static_assert(BM == WM * MT * 8 && BN == WN * NT * 8 && MT * NT <= 32, "warp tile");
    static_assert(A_CHUNKS * THREADS == BM * (BK / 2) && B_CHUNKS * THREADS == BK * (BN / 2), "chunking");
};
using CfgSmall = Cfg<128, 64, 2, 2, 3, 2>;
using CfgLarge = Cfg<128, 128, 2, 4, 4, 1>;
using CfgLargeK32 = Cfg<128, 128, 2, 4, 3, 1, 32>;
using CfgSmallK32 = Cfg<128, 64, 2, 2, 2, 2, 32>;
using CfgW16 = Cfg<128, 128, 4, 4, 4, 1, 16>;       // 16 warps, 32x32 warp tiles: 4 warps per scheduler
using CfgW16K32 = Cfg<128, 128, 4, 4, 3, 1, 32>;
using CfgTiny = Cfg<64, 64, 2, 2, 3, 4, 16>;         // 64x64 tiles for products too small to fill the SMs with 128x64

// Per-thread copy plan for the aligned (16-byte) path.  Thread `tid` always copies chunk column
// `tid % chunks_per_row` of rows `tid / chunks_per_row + i * rows_per_pass`: one base pointer per
// operand plus a constant row stride, so the per-slab address arithmetic is a handful of IMADs.
template <class C>
struct CopyPlan {
    static constexpr int ACH = C::BK / 2, A_ROWS = C::THREADS / ACH;        // A: chunks per row, rows per pass
    static constexpr int BCH = C::BN / 2, B_ROWS = C::THREADS / BCH;     // B
    static_assert(A_ROWS * C::A_CHUNKS == C::BM && B_ROWS * C::B_CHUNKS == C::BK, "copy plan");
    const double *a_src;     // A + (m0 + row0)*lda + 2*ch
    const double *b_src;     // B + row0*ldb + n0 + 2*ch
    size_t a_step, b_step;   // elements between passes (A_ROWS*lda, B_ROWS*ldb)
    uint32_t a_dst, b_dst;   // shared byte offsets inside a stage
    uint32_t a_valid;        // bit i: row of pass i is inside M
    int a_k, b_k, b_bytes;   // k offset inside a slab (A chunk / B row of pass 0); B bytes (16/8/0)
};

template <class C, bool ALIGNED>
__device__ __forceinline__ void issue_slab(const CopyPlan<C> &p, uint32_t as_base, uint32_t bs_base, int k0, int K,
                                           size_t ldb, bool full, const double *A, const double *B, size_t lda,
                                           int M, int N, int m0, int n0, int tid) {
    if (ALIGNED) {
        using P = CopyPlan<C>;
        int abytes = 16;
        if (!full) {
            const int rem = K - (k0 + p.a_k);
            abytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        }
        const double *asrc = p.a_src + k0;
#pragma unroll
        for (int i = 0; i < C::A_CHUNKS; ++i) {
            const int bytes = ((p.a_valid >> i) & 1u) ? abytes : 0;
            cp_async16(as_base + p.a_dst + i * (P::A_ROWS * C::LDAS * 8), bytes ? asrc + i * p.a_step : A, bytes);
        }
        const double *bsrc = p.b_src + size_t(k0) * ldb;
#pragma unroll
        for (int i = 0; i < C::B_CHUNKS; ++i) {
            int bytes = p.b_bytes;
            if (!full && k0 + p.b_k + i * P::B_ROWS >= K) bytes = 0;
            cp_async16(bs_base + p.b_dst + i * (P::B_ROWS * C::LDBS * 8), bytes ? bsrc + i * p.b_step : B, bytes);
        }
    } else {
        // 8-byte path for odd leading dimensions / unaligned bases (bounds evaluated per element)
        constexpr int AE = C::BM * C::BK / C::THREADS, BE = C::BK * C::BN / C::THREADS;
#pragma unroll
        for (int i = 0; i < AE; ++i) {
            const int e = tid + i * C::THREADS;
            const int row = e / C::BK, kk = e % C::BK;
            const bool ok = (m0 + row < M) && (k0 + kk < K);
            cp_async8(as_base + (row * C::LDAS + kk) * 8, ok ? A + size_t(m0 + row) * lda + k0 + kk : A, ok ? 8 : 0);
        }
#pragma unroll
        for (int i = 0; i < BE; ++i) {
            const int e = tid + i * C::THREADS;
            const int row = e / C::BN, cc = e % C::BN;
            const bool ok = (k0 + row < K) && (n0 + cc < N);
            cp_async8(bs_base + (row * C::LDBS + cc) * 8, ok ? B + size_t(k0 + row) * ldb + n0 + cc : B, ok ? 8 : 0);
        }
    }
}

template <class C>
__device__ __forceinline__ void mma_k4(double (&acc)[C::MT][C::NT][2], const double *ap, const double *bp, int kk) {
    double af[C::MT], bf[C::NT];
#pragma unroll
    for (int i = 0; i < C::MT; ++i) af[i] = ap[i * 8 * C::LDAS + kk];
#pragma unroll
    for (int j = 0; j < C::NT; ++j) bf[j] = bp[kk * C::LDBS + j * 8];
#pragma unroll
    for (int i = 0; i < C::MT; ++i)
#pragma unroll
        for (int j = 0; j < C::NT; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
}

// ACC_C: the accumulators START from C instead of zero and the epilogue stores alpha * acc (C_out = alpha * (C_in + A*B),
// beta is not used).  A product split along k into consecutive calls -- chunk 0 plain with alpha = 1, every later chunk
// ACC_C, the last one with the caller's alpha -- then performs, per C element, exactly the DMMA accumulation chain of
// the one-call product (chunk boundaries are multiples of 4), i.e. it is BIT-IDENTICAL to it.  The host pipeline uses
// this to make work available linearly in the uploaded bytes during its first milliseconds (api.cu, gemm_host).
template <class C, bool ALIGNED, bool ACC_C = false>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
dgemm_dmma_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, size_t lda,
                  const double *__restrict__ B, size_t ldb, double beta, double *__restrict__ Cmat,
                  size_t ldc, int tiles_m, int tiles_n) {
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + C::STAGES * C::A_STAGE;

    // band-rasterised tile order
    const int bid = blockIdx.x;
    const int per_band = BAND * tiles_n;
    const int band = bid / per_band;
    const int rem = bid - band * per_band;
    const int band_rows = min(BAND, tiles_m - band * BAND);
    const int tm = band * BAND + rem % band_rows;
    const int tn = rem / band_rows;
    const int m0 = tm * C::BM, n0 = tn * C::BN;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % C::WM, wn = warp / C::WM;

    CopyPlan<C> plan;
    if (ALIGNED) {
        using P = CopyPlan<C>;
        const int arow = tid / P::ACH, ach = tid % P::ACH;
        plan.a_k = 2 * ach;
        plan.a_dst = (arow * C::LDAS + 2 * ach) * 8;
        plan.a_step = size_t(P::A_ROWS) * lda;
        plan.a_valid = 0;
#pragma unroll
        for (int i = 0; i < C::A_CHUNKS; ++i)
            if (m0 + arow + i * P::A_ROWS < M) plan.a_valid |= 1u << i;
        plan.a_src = A + size_t(m0 + arow) * lda + 2 * ach;
        const int brow = tid / P::BCH, bch = tid % P::BCH;
        plan.b_k = brow;
        plan.b_dst = (brow * C::LDBS + 2 * bch) * 8;
        plan.b_step = size_t(P::B_ROWS) * ldb;
        const int gn = n0 + 2 * bch;
        plan.b_bytes = gn + 1 < N ? 16 : (gn < N ? 8 : 0);
        plan.b_src = B + size_t(brow) * ldb + (gn < N ? gn : 0);
    }

    double acc[C::MT][C::NT][2];
#pragma unroll
    for (int i = 0; i < C::MT; ++i)
#pragma unroll
        for (int j = 0; j < C::NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int KT = (K + C::BK - 1) / C::BK;
    const int KT_FULL = K / C::BK;                  // slabs needing no k-bound checks
    const uint32_t as_u32 = smem_u32(As), bs_u32 = smem_u32(Bs);

#pragma unroll
    for (int s = 0; s < C::STAGES - 1; ++s) {
        if (s < KT)
            issue_slab<C, ALIGNED>(plan, as_u32 + s * C::A_STAGE * 8, bs_u32 + s * C::B_STAGE * 8, s * C::BK, K, ldb,
                                   s < KT_FULL, A, B, lda, M, N, m0, n0, tid);
        cp_async_commit();
    }

    const double *a_frag_base = As + (wm * C::MT * 8 + g) * C::LDAS + t;
    const double *b_frag_base = Bs + t * C::LDBS + wn * C::NT * 8 + g;

    if (ACC_C) {                                 // (while the first slabs are in flight)
        const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cmat) & 15) == 0);
#pragma unroll
        for (int i = 0; i < C::MT; ++i) {
            const int row = m0 + wm * C::MT * 8 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < C::NT; ++j) {
                const int col = n0 + wn * C::NT * 8 + j * 8 + 2 * t;
                if (row < M && col < N) {
                    const double *cp = Cmat + size_t(row) * ldc + col;
                    if (col + 1 < N && vec_ok) {
                        const double2 v = *reinterpret_cast<const double2 *>(cp);
                        acc[i][j][0] = v.x;
                        acc[i][j][1] = v.y;
                    } else {
                        acc[i][j][0] = cp[0];
                        if (col + 1 < N) acc[i][j][1] = cp[1];
                    }
                }
            }
        }
    }

    int stage = 0;
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<C::STAGES - 2>();
        __syncthreads();
        const double *ap = a_frag_base + stage * C::A_STAGE;
        const double *bp = b_frag_base + stage * C::B_STAGE;
        mma_k4<C>(acc, ap, bp, 0);               // feed the tensor pipe first ...
        {                                        // ... then queue the copies for slab kt+STAGES-1
            const int nk = kt + C::STAGES - 1;
            if (nk < KT) {
                int ns = stage + C::STAGES - 1;
                if (ns >= C::STAGES) ns -= C::STAGES;
                issue_slab<C, ALIGNED>(plan, as_u32 + ns * C::A_STAGE * 8, bs_u32 + ns * C::B_STAGE * 8, nk * C::BK, K, ldb,
                                       nk < KT_FULL, A, B, lda, M, N, m0, n0, tid);
            }
            cp_async_commit();
        }
#pragma unroll
        for (int kk = 4; kk < C::BK; kk += 4) mma_k4<C>(acc, ap, bp, kk);
        if (++stage == C::STAGES) stage = 0;
    }
    cp_async_wait<0>();

    // epilogue: thread owns C[row g][cols 2t,2t+1] of each 8x8 fragment
    const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cmat) & 15) == 0);
    const bool interior = vec_ok && (m0 + C::BM <= M) && (n0 + C::BN <= N);
    double *cbase = Cmat + size_t(m0 + wm * C::MT * 8 + g) * ldc + n0 + wn * C::NT * 8 + 2 * t;
    if (interior) {
        if (beta == 0.0) {
#pragma unroll
            for (int i = 0; i < C::MT; ++i)
#pragma unroll
                for (int j = 0; j < C::NT; ++j)
                    *reinterpret_cast<double2 *>(cbase + size_t(i * 8) * ldc + j * 8) =
                        make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        } else {
#pragma unroll
            for (int ib = 0; ib < C::MT; ib += 2) {
                double2 old[2][C::NT];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < C::NT; ++j)
                        old[i][j] = *reinterpret_cast<const double2 *>(cbase + size_t((ib + i) * 8) * ldc + j * 8);
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < C::NT; ++j)
                        *reinterpret_cast<double2 *>(cbase + size_t((ib + i) * 8) * ldc + j * 8) =
                            make_double2(alpha * acc[ib + i][j][0] + beta * old[i][j].x,
                                         alpha * acc[ib + i][j][1] + beta * old[i][j].y);
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < C::MT; ++i) {
        const int row = m0 + wm * C::MT * 8 + i * 8 + g;
        if (row >= M) continue;
        double *crow = Cmat + size_t(row) * ldc;
#pragma unroll
        for (int j = 0; j < C::NT; ++j) {
            const int col = n0 + wn * C::NT * 8 + j * 8 + 2 * t;
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

// -------------------------------------------------------------------------------------------
// Stream-K variant for products whose tile count does not fill the machine evenly (BASELINE config C1, 1024^3: 256 tiles
// of 64x64 put two tiles on 108 SMs and one on 40 -- 86.5 % at best, 65 % measured).  Work unit = (tile, k-slab); CTA b of g
// takes the contiguous units [b*U/g, (b+1)*U/g), i.e. every SM gets the same number of k-slabs whatever the tile count.
// A segment that covers a whole tile is written straight to C; a partial segment stores its accumulators to a workspace
// slot (a CTA has at most two: slot 2b = its first segment, 2b+1 = its last), and a fix-up kernel (one CTA per split tile)
// adds the slots of the tile in ascending CTA = ascending k order and applies alpha / beta.  The summation order is
// fixed by (shape, grid), so results are deterministic; they differ from the one-tile-one-CTA kernel in the last bits
// (partial sums over k ranges), inside the same error bound.
// -------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ long long sk_u0(long long b, long long U, long long g) { return b * U / g; }

template <class C>
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int &m0, int &n0) {
    const int per_band = BAND * tiles_n;
    const int band = tile / per_band;
    const int rem = tile - band * per_band;
    const int band_rows = min(BAND, tiles_m - band * BAND);
    m0 = (band * BAND + rem % band_rows) * C::BM;
    n0 = (rem / band_rows) * C::BN;
}

// C tile <- alpha*acc + beta*C (the epilogue of dgemm_dmma_kernel as a function; beta == 0 never reads C)
template <class C>
__device__ __forceinline__ void write_tile(const double (&acc)[C::MT][C::NT][2], double alpha, double beta, double *Cmat,
                                           size_t ldc, int M, int N, int m0, int n0, int wm, int wn, int g, int t) {
    const bool vec_ok = ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cmat) & 15) == 0);
#pragma unroll
    for (int i = 0; i < C::MT; ++i) {
        const int row = m0 + wm * C::MT * 8 + i * 8 + g;
        if (row >= M) continue;
        double *crow = Cmat + size_t(row) * ldc;
#pragma unroll
        for (int j = 0; j < C::NT; ++j) {
            const int col = n0 + wn * C::NT * 8 + j * 8 + 2 * t;
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

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
dgemm_streamk_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, size_t lda,
                     const double *__restrict__ B, size_t ldb, double beta, double *__restrict__ Cmat, size_t ldc,
                     int tiles_m, int tiles_n, double *__restrict__ ws) {
    extern __shared__ __align__(16) double smem[];
    double *As = smem;
    double *Bs = smem + C::STAGES * C::A_STAGE;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % C::WM, wn = warp / C::WM;
    const int KT = (K + C::BK - 1) / C::BK;
    const int KT_FULL = K / C::BK;
    const uint32_t as_u32 = smem_u32(As), bs_u32 = smem_u32(Bs);
    const double *a_frag_base = As + (wm * C::MT * 8 + g) * C::LDAS + t;
    const double *b_frag_base = Bs + t * C::LDBS + wn * C::NT * 8 + g;

    const long long U = (long long)tiles_m * tiles_n * KT, G = gridDim.x, b = blockIdx.x;
    long long u = sk_u0(b, U, G);
    const long long uend = sk_u0(b + 1, U, G);
    bool first = true;
    while (u < uend) {
        const int tile = int(u / KT);
        const int kt0 = int(u - (long long)tile * KT);
        const int kt1 = int(min((long long)KT, kt0 + (uend - u)));
        int m0, n0;
        tile_coords<C>(tile, tiles_m, tiles_n, m0, n0);

        CopyPlan<C> plan;
        {
            using P = CopyPlan<C>;
            const int arow = tid / P::ACH, ach = tid % P::ACH;
            plan.a_k = 2 * ach;
            plan.a_dst = (arow * C::LDAS + 2 * ach) * 8;
            plan.a_step = size_t(P::A_ROWS) * lda;
            plan.a_valid = 0;
#pragma unroll
            for (int i = 0; i < C::A_CHUNKS; ++i)
                if (m0 + arow + i * P::A_ROWS < M) plan.a_valid |= 1u << i;
            plan.a_src = A + size_t(m0 + arow) * lda + 2 * ach;
            const int brow = tid / P::BCH, bch = tid % P::BCH;
            plan.b_k = brow;
            plan.b_dst = (brow * C::LDBS + 2 * bch) * 8;
            plan.b_step = size_t(P::B_ROWS) * ldb;
            const int gn = n0 + 2 * bch;
            plan.b_bytes = gn + 1 < N ? 16 : (gn < N ? 8 : 0);
            plan.b_src = B + size_t(brow) * ldb + (gn < N ? gn : 0);
        }
        double acc[C::MT][C::NT][2];
#pragma unroll
        for (int i = 0; i < C::MT; ++i)
#pragma unroll
            for (int j = 0; j < C::NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        const int nseg = kt1 - kt0;
#pragma unroll
        for (int s = 0; s < C::STAGES - 1; ++s) {
            if (s < nseg)
                issue_slab<C, true>(plan, as_u32 + s * C::A_STAGE * 8, bs_u32 + s * C::B_STAGE * 8, (kt0 + s) * C::BK, K, ldb,
                                    kt0 + s < KT_FULL, A, B, lda, M, N, m0, n0, tid);
            cp_async_commit();
        }
        int stage = 0;
        for (int i = 0; i < nseg; ++i) {
            cp_async_wait<C::STAGES - 2>();
            __syncthreads();
            const double *ap = a_frag_base + stage * C::A_STAGE;
            const double *bp = b_frag_base + stage * C::B_STAGE;
            mma_k4<C>(acc, ap, bp, 0);
            {
                const int ni = i + C::STAGES - 1;
                if (ni < nseg) {
                    int ns = stage + C::STAGES - 1;
                    if (ns >= C::STAGES) ns -= C::STAGES;
                    issue_slab<C, true>(plan, as_u32 + ns * C::A_STAGE * 8, bs_u32 + ns * C::B_STAGE * 8, (kt0 + ni) * C::BK, K,
                                        ldb, kt0 + ni < KT_FULL, A, B, lda, M, N, m0, n0, tid);
                }
                cp_async_commit();
            }
#pragma unroll
            for (int kk = 4; kk < C::BK; kk += 4) mma_k4<C>(acc, ap, bp, kk);
            if (++stage == C::STAGES) stage = 0;
        }
        cp_async_wait<0>();
        __syncthreads();                          // the next segment's prologue overwrites the stages

        if (kt0 == 0 && kt1 == KT) {
            write_tile<C>(acc, alpha, beta, Cmat, ldc, M, N, m0, n0, wm, wn, g, t);
        } else {
            double *slot = ws + size_t(2 * b + (first ? 0 : 1)) * (C::BM * C::BN);
#pragma unroll
            for (int i = 0; i < C::MT; ++i)
#pragma unroll
                for (int j = 0; j < C::NT; ++j)
                    *reinterpret_cast<double2 *>(slot + (size_t(i * C::NT + j) * C::THREADS + tid) * 2) =
                        make_double2(acc[i][j][0], acc[i][j][1]);
        }
        first = false;
        u += nseg;
    }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS)
dgemm_streamk_fixup_kernel(int M, int N, int K, double alpha, double beta, double *__restrict__ Cmat, size_t ldc, int tiles_m,
                           int tiles_n, const double *__restrict__ ws, int G) {
    const int tile = blockIdx.x;
    const int KT = (K + C::BK - 1) / C::BK;
    const long long U = (long long)tiles_m * tiles_n * KT;
    const long long ua = (long long)tile * KT, ub = ua + KT;
    long long bf = ua * G / U, bl = (ub - 1) * G / U;             // CTAs that hold the tile's first and last unit
    while (sk_u0(bf + 1, U, G) <= ua) ++bf;
    while (sk_u0(bf, U, G) > ua) --bf;
    while (sk_u0(bl + 1, U, G) <= ub - 1) ++bl;
    while (sk_u0(bl, U, G) > ub - 1) --bl;
    if (bf == bl) return;                                          // one CTA had the whole tile and wrote it itself
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp % C::WM, wn = warp / C::WM;
    double acc[C::MT][C::NT][2];
#pragma unroll
    for (int i = 0; i < C::MT; ++i)
#pragma unroll
        for (int j = 0; j < C::NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (long long b = bf; b <= bl; ++b) {                         // ascending CTA = ascending k: a fixed summation order
        const double *slot = ws + size_t(2 * b + (sk_u0(b, U, G) >= ua ? 0 : 1)) * (C::BM * C::BN);
#pragma unroll
        for (int i = 0; i < C::MT; ++i)
#pragma unroll
            for (int j = 0; j < C::NT; ++j) {
                const double2 p = *reinterpret_cast<const double2 *>(slot + (size_t(i * C::NT + j) * C::THREADS + tid) * 2);
                acc[i][j][0] += p.x;
                acc[i][j][1] += p.y;
            }
    }
    int m0, n0;
    tile_coords<C>(tile, tiles_m, tiles_n, m0, n0);
    write_tile<C>(acc, alpha, beta, Cmat, ldc, M, N, m0, n0, wm, wn, g, t);
}

// k == 0: C <- beta*C, zero-fill when beta == 0 without reading C.
template <typename T>
__global__ void scale_c_kernel(size_t M, size_t N, T beta, T *C, size_t ldc) {
    const size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const size_t r = idx / N, c = idx - r * N;
    T *p = C + r * ldc + c;
    *p = (beta == T(0)) ? T(0) : (*p) * beta;
}

template <class C, bool ALIGNED, bool ACC_C = false>
int launch_cfg(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda, const double *b, size_t ldb,
               double beta, double *c, size_t ldc, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int tiles_m = int((m + C::BM - 1) / C::BM), tiles_n = int((n + C::BN - 1) / C::BN);
    const size_t tiles = size_t(tiles_m) * tiles_n;
    if (tiles > 0x7fffffffull) return RLA_ERR_INVALID;
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<C, ALIGNED, ACC_C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::SMEM)));
        RLA_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<C, ALIGNED, ACC_C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_once.done(od_);
    }
    dgemm_dmma_kernel<C, ALIGNED, ACC_C><<<unsigned(tiles), C::THREADS, C::SMEM, st>>>(int(m), int(n), int(k), alpha, a, lda, b, ldb,
                                                                                       beta, c, ldc, tiles_m, tiles_n);
    RLA_LAUNCHED();
    return RLA_OK;
}

// The library's own stream-ordered memory pool, one per device, never trimmed: a freed workspace is handed to the next
// product without a trip to the driver (the device's default pool releases at every synchronisation: ~400 us per call).
int streamk_pool(cudaMemPool_t *out) {
    static std::mutex mu;
    static cudaMemPool_t pools[RLA_MAX_DEVICES] = {};
    const int dev = current_device();
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        RLA_CUDA(cudaMemPoolCreate(&pools[dev], &props));
        unsigned long long keep = ~0ull;
        RLA_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
    }
    *out = pools[dev];
    return RLA_OK;
}

template <class C>
int launch_streamk(size_t m, size_t k, size_t n, double alpha, const double *a, size_t lda, const double *b, size_t ldb,
                   double beta, double *c, size_t ldc, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int tiles_m = int((m + C::BM - 1) / C::BM), tiles_n = int((n + C::BN - 1) / C::BN);
    if (const int od_ = attr_once.pending(); od_ >= 0) {
        RLA_CUDA(cudaFuncSetAttribute(dgemm_streamk_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::SMEM)));
        RLA_CUDA(cudaFuncSetAttribute(dgemm_streamk_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_once.done(od_);
    }
    const long long KT = (long long)((k + C::BK - 1) / C::BK), U = (long long)tiles_m * tiles_n * KT;
    long long grid = (long long)device_num_sms() * C::MINB;
    if (grid > U) grid = U;
    // stream-ordered workspace: two accumulator slots per CTA; safe for concurrent products on different streams
    cudaMemPool_t pool = nullptr;
    RLA_TRY(streamk_pool(&pool));
    double *ws = nullptr;
    RLA_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), size_t(2) * size_t(grid) * C::BM * C::BN * sizeof(double), pool, st));
    dgemm_streamk_kernel<C><<<unsigned(grid), C::THREADS, C::SMEM, st>>>(int(m), int(n), int(k), alpha, a, lda, b, ldb, beta, c,
                                                                         ldc, tiles_m, tiles_n, ws);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        note_launch();
        dgemm_streamk_fixup_kernel<C><<<unsigned(tiles_m * tiles_n), C::THREADS, 0, st>>>(int(m), int(n), int(k), alpha, beta, c,
                                                                                          ldc, tiles_m, tiles_n, ws, int(grid));
        e = cudaGetLastError();
        if (e == cudaSuccess) note_launch();
    }
    cudaFreeAsync(ws, st);                      // stream-ordered: released after the fix-up kernel
    if (e != cudaSuccess) {
        note_cuda_error(e);
        return RLA_ERR_CUDA;
    }
    return RLA_OK;
}

}  // namespace

int g_dgemm_cfg = -1;  // -1 = auto (cost model below); rla_set_tuning("dgemm_cfg", 0..7) forces one
int g_dgemm_streamk = 0;   // rla_set_tuning("dgemm_streamk", v): 0 (default) = tiled kernels only; 2 = 128x128 stream-K whenever the
                           // operands are aligned, 3 = 64x64 stream-K.  Off by default on the evidence of
                           // profiles/r02_dgemm_streamk_sweep.jsonl: it wins 5-10 % where the tile count quantises badly AND k is long
                           // (1280^3, 1792^3, 1024x4096x1024), ties from 2048^3 up and LOSES at 512^3-1024^3 (1024^3: main kernel 77 us
                           // + fix-up 17 us against 85 us tiled under ncu) -- short k ranges per CTA refill the cp.async ring per segment
                           // and the 27 MB of partial tiles are reduced by only 64 CTAs.  It would also end the shape-independence of
                           // every C element's rounding that the pipelines and the multi-GPU paths rely on for bit-identity.

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
                 cudaStream_t st, bool acc_from_c) {
    if (m == 0 || n == 0) return RLA_OK;
    if (acc_from_c) {
        // C <- alpha * (C + A*B), the sum continued in the accumulators (see dgemm_dmma_kernel); aligned operands only
        if (k == 0) return scale_c_launch<double>(m, n, alpha, c, ldc, st);
        if (m > 0x7fffffffull || n > 0x7fffffffull || k > 0x7fffffffull) return RLA_ERR_INVALID;
        if ((lda & 1) || (ldb & 1) || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) return RLA_ERR_INVALID;
        const double t128 = double((m + 127) / 128) * double((n + 127) / 128);
        const double t64 = double((m + 63) / 64) * double((n + 63) / 64);
        const double kd = double(k);
        const double cost128 = ceil(t128 / 148.0) * 4.0 / (0.939 * kd / (kd + 28.0));
        const double cost64 = ceil(t64 / 148.0) / (0.918 * kd / (kd + 8.0));
        return cost128 < cost64 ? launch_cfg<CfgW16K32, true, true>(m, k, n, alpha, a, lda, b, ldb, 0.0, c, ldc, st)
                                : launch_cfg<CfgTiny, true, true>(m, k, n, alpha, a, lda, b, ldb, 0.0, c, ldc, st);
    }
    if (k == 0) return scale_c_launch<double>(m, n, beta, c, ldc, st);
    if (m > 0x7fffffffull || n > 0x7fffffffull || k > 0x7fffffffull) return RLA_ERR_INVALID;
    const bool aligned = ((lda & 1) == 0) && ((ldb & 1) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
    // -1 (default): pick between 128x128 tiles (cfg 6, one CTA per SM, best asymptote) and 64x64 tiles (cfg 7,
    // four CTAs per SM: 4x finer tail, shorter pipeline fill) from a two-term cost model fitted to the
    // measured table in profiles/r01_dgemm_cfg_table.md:  cost = waves * tile_work / (e_inf * k / (k + k0)).
    int cfg = g_dgemm_cfg;
    if (aligned && cfg < 0 && g_dgemm_streamk >= 2) {
        return g_dgemm_streamk == 2 ? launch_streamk<CfgW16K32>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st)
                                    : launch_streamk<CfgTiny>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    }
    if (cfg < 0) {
        if (aligned) {
            const double t128 = double((m + 127) / 128) * double((n + 127) / 128);
            const double t64 = double((m + 63) / 64) * double((n + 63) / 64);
            const double kd = double(k);
            const double cost128 = ceil(t128 / 148.0) * 4.0 / (0.939 * kd / (kd + 28.0));
            const double cost64 = ceil(t64 / 148.0) / (0.918 * kd / (kd + 8.0));
            cfg = cost128 < cost64 ? 6 : 7;
        } else {
            cfg = (k >= 2048 && (m / 128) * (n / 128) >= 4 * 148) ? 1 : 0;
        }
    }
    if (cfg == 1) {
        return aligned ? launch_cfg<CfgLarge, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st)
                       : launch_cfg<CfgLarge, false>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    }
    if (cfg == 2 && aligned) return launch_cfg<CfgLargeK32, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    if (cfg == 3 && aligned) return launch_cfg<CfgSmallK32, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    if (cfg == 5 && aligned) return launch_cfg<CfgW16, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    if (cfg == 6 && aligned) return launch_cfg<CfgW16K32, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    if (cfg == 7 && aligned) return launch_cfg<CfgTiny, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
    return aligned ? launch_cfg<CfgSmall, true>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st)
                   : launch_cfg<CfgSmall, false>(m, k, n, alpha, a, lda, b, ldb, beta, c, ldc, st);
}

}  // namespace rla
