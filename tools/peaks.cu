// peaks.cu -- measurement tool (NOT product code): pipe peaks and informative library ceilings
// on the B200 the job lands on.  SURVEY.md F10: MEASURED_PEAKS.json has no FP64/FP32 peak.
//   - FP64 tensor (DMMA.8x8x4) issue-bound peak, register-only loop
//   - FP64 DFMA peak, FP32 FFMA peak
//   - cuBLAS Dgemm / Sgemm (informative ceiling; never linked into librla_b200.so)
//   - cuSOLVER Dgetrf (informative ceiling)
//   - grid-wide barrier and 16-CTA cluster barrier latency (sizing input for the LU panel kernel)
// Output: JSON lines on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu -lcublas -lcusolver
#include <cooperative_groups.h>
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_peak(double *out, int iters, double a0, double b0) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void dfma_peak(double *out, int iters, double a0, double b0) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void ffma_peak(float *out, int iters, float a0, float b0) {
    float c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    float a = a0 + threadIdx.x * 1e-6f, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    if (s == 123.456f) out[0] = s;
}

// Blackwell packed fp32 FMA (fma.rn.f32x2 -> SASS FFMA2): two IEEE fp32 FMAs per lane per instruction
template <int ILP>
__global__ void ffma2_peak(float *out, int iters, float a0, float b0) {
    unsigned long long c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = 0ull;
    unsigned long long a, b;
    float af = a0 + threadIdx.x * 1e-6f;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(af));
    asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(b0));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(a), "l"(b));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= c[i];
    if (s == 12345ull) out[0] = 1.f;
}

// custom grid barrier: monotonically increasing counter
__global__ void grid_barrier_lat(unsigned *counter, int rounds, long long *cycles) {
    long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            const unsigned target = unsigned(r) * gridDim.x;
            while (*((volatile unsigned *)counter) < target) {
            }
            __threadfence();
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = clock64() - t0;
}

__global__ void coop_barrier_lat(int rounds, long long *cycles) {
    cg::grid_group grid = cg::this_grid();
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = clock64() - t0;
}

__global__ void cluster_barrier_lat(int rounds, long long *cycles) {
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::);
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = clock64() - t0;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

static bool want(const char *sel, const char *group) { return strcmp(sel, "all") == 0 || strcmp(sel, group) == 0; }

int main(int argc, char **argv) {
    // usage: peaks <big:0|1> <group: all|pipes|barriers|cublas|cusolver>
    const bool big = argc > 1 && atoi(argv[1]) > 0;
    const char *sel = argc > 2 ? argv[2] : "all";
    setvbuf(stdout, nullptr, _IOLBF, 0);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d,\"cc\":\"%d.%d\",\"clock_khz\":%d,\"l2_bytes\":%d}\n",
           prop.name, sms, prop.major, prop.minor, clk_khz, prop.l2CacheSize);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double *dout;
    CK(cudaMalloc(&dout, 1024));

    // ---- DMMA peak: sweep warps/SM ------------------------------------------------------
    if (want(sel, "pipes")) {
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 20000;
        constexpr int ILP = 8;
        dmma_peak<ILP><<<sms, threads>>>(dout, 100, 1.0, 1.0);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            dmma_peak<ILP><<<sms, threads>>>(dout, iters, 1.0, 1.0);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double flops = 2.0 * 256.0 * ILP * double(iters) * (threads / 32) * sms;
        printf("{\"probe\":\"dmma_peak\",\"threads_per_sm\":%d,\"ilp\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",
               threads, ILP, best, flops / best * 1e-9);
    }
    // sustained DMMA (2 s) to see the power-capped clock
    {
        const int threads = 512, iters = 20000;
        constexpr int ILP = 8;
        CK(cudaEventRecord(e0));
        int launches = 0;
        for (; launches < 400; ++launches) dmma_peak<ILP><<<sms, threads>>>(dout, iters, 1.0, 1.0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const float ms = time_ms(e0, e1);
        const double flops = 2.0 * 256.0 * ILP * double(iters) * (threads / 32) * sms * launches;
        printf("{\"probe\":\"dmma_sustained\",\"threads_per_sm\":%d,\"seconds\":%.3f,\"tflops\":%.3f}\n",
               threads, ms * 1e-3, flops / ms * 1e-9);
    }
    // ---- DFMA peak ---------------------------------------------------------------------------
    for (int threads : {256, 512, 1024}) {
        const int iters = 20000;
        constexpr int ILP = 8;
        dfma_peak<ILP><<<sms, threads>>>(dout, 100, 1.0, 1.0);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            dfma_peak<ILP><<<sms, threads>>>(dout, iters, 1.0000001, 1e-9);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double flops = 2.0 * ILP * double(iters) * threads * sms;
        printf("{\"probe\":\"dfma_peak\",\"threads_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", threads, best,
               flops / best * 1e-9);
    }
    // ---- FFMA peak ---------------------------------------------------------------------------
    for (int threads : {256, 512, 1024}) {
        const int iters = 40000;
        constexpr int ILP = 16;
        ffma_peak<ILP><<<sms, threads>>>((float *)dout, 100, 1.0f, 1.0f);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            ffma_peak<ILP><<<sms, threads>>>((float *)dout, iters, 1.0000001f, 1e-9f);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double flops = 2.0 * ILP * double(iters) * threads * sms;
        printf("{\"probe\":\"ffma_peak\",\"threads_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", threads, best,
               flops / best * 1e-9);
    }
    for (int threads : {256, 512, 1024}) {
        const int iters = 40000;
        constexpr int ILP = 16;
        ffma2_peak<ILP><<<sms, threads>>>((float *)dout, 100, 1.0f, 1.0f);
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            ffma2_peak<ILP><<<sms, threads>>>((float *)dout, iters, 1.0000001f, 1e-9f);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        const double flops = 4.0 * ILP * double(iters) * threads * sms;
        printf("{\"probe\":\"ffma2_peak\",\"threads_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", threads, best,
               flops / best * 1e-9);
    }
    {
        const int threads = 1024, iters = 40000;
        constexpr int ILP = 16;
        CK(cudaEventRecord(e0));
        int launches = 0;
        for (; launches < 300; ++launches) ffma_peak<ILP><<<sms, threads>>>((float *)dout, iters, 1.0000001f, 1e-9f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        const float ms = time_ms(e0, e1);
        const double flops = 2.0 * ILP * double(iters) * threads * sms * launches;
        printf("{\"probe\":\"ffma_sustained\",\"seconds\":%.3f,\"tflops\":%.3f}\n", ms * 1e-3, flops / ms * 1e-9);
    }

    }
    // ---- barrier latencies ---------------------------------------------------------------------
    if (want(sel, "barriers")) {
        unsigned *counter;
        long long *cycles;
        CK(cudaMalloc(&counter, 4));
        CK(cudaMalloc(&cycles, 8));
        const int rounds = 2000;
        for (int ctas : {8, 16, 32, 74, 148}) {
            for (int threads : {128, 512}) {
                CK(cudaMemset(counter, 0, 4));
                CK(cudaEventRecord(e0));
                grid_barrier_lat<<<ctas, threads>>>(counter, rounds, cycles);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                printf("{\"probe\":\"grid_barrier_atomic\",\"ctas\":%d,\"threads\":%d,\"us_per_barrier\":%.3f}\n", ctas,
                       threads, time_ms(e0, e1) * 1e3 / rounds);
            }
        }
        for (int ctas : {16, 148}) {
            int r = rounds;
            long long *cy = cycles;
            void *args[] = {&r, &cy};
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void *)coop_barrier_lat, dim3(ctas), dim3(256), args, 0, 0));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            printf("{\"probe\":\"grid_barrier_coop\",\"ctas\":%d,\"us_per_barrier\":%.3f}\n", ctas,
                   time_ms(e0, e1) * 1e3 / rounds);
        }
        for (int csize : {2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csize);
            cfg.blockDim = dim3(256);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = csize;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (csize > 8) CK(cudaFuncSetAttribute(cluster_barrier_lat, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            CK(cudaEventRecord(e0));
            cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_barrier_lat, rounds, cycles);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            if (e != cudaSuccess) {
                printf("{\"probe\":\"cluster_barrier\",\"cluster\":%d,\"error\":\"%s\"}\n", csize, cudaGetErrorString(e));
                cudaGetLastError();
            } else {
                printf("{\"probe\":\"cluster_barrier\",\"cluster\":%d,\"us_per_barrier\":%.3f}\n", csize,
                       time_ms(e0, e1) * 1e3 / rounds);
            }
        }
    }

    // ---- cuBLAS ceilings -------------------------------------------------------------------------
    if (want(sel, "cublas")) {
    cublasHandle_t h;
    cublasCreate(&h);
    cublasSetMathMode(h, CUBLAS_DEFAULT_MATH);  // default math: FP32 SGEMM stays FP32 (TF32 is opt-in)
    for (int n : {1024, 2048, 4096, 8192, 16384}) {
        if (n > 8192 && !big) continue;
        double *a, *b, *c;
        const size_t bytes = size_t(n) * n * 8;
        CK(cudaMalloc(&a, bytes));
        CK(cudaMalloc(&b, bytes));
        CK(cudaMalloc(&c, bytes));
        CK(cudaMemset(a, 0, bytes));
        CK(cudaMemset(b, 0, bytes));
        const double one = 1.0, zero = 0.0;
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, a, n, b, n, &zero, c, n);
        float best = 1e30f;
        const int reps = n >= 8192 ? 3 : 10;
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(e0));
            cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, a, n, b, n, &zero, c, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"probe\":\"cublas_dgemm\",\"n\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", n, best,
               2.0 * n * double(n) * n / best * 1e-9);
        float *fa = (float *)a, *fb = (float *)b, *fc = (float *)c;
        const float onef = 1.f, zerof = 0.f;
        cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &onef, fa, n, fb, n, &zerof, fc, n);
        best = 1e30f;
        for (int r = 0; r < reps; ++r) {
            CK(cudaEventRecord(e0));
            cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &onef, fa, n, fb, n, &zerof, fc, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"probe\":\"cublas_sgemm\",\"n\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", n, best,
               2.0 * n * double(n) * n / best * 1e-9);
        cudaFree(a);
        cudaFree(b);
        cudaFree(c);
    }
    // LU trailing-update shape: C(n x n) -= A(n x 128/256) B
    for (int kk : {128, 256, 512}) {
        const int n = 16384;
        double *a, *b, *c;
        CK(cudaMalloc(&a, size_t(n) * kk * 8));
        CK(cudaMalloc(&b, size_t(n) * kk * 8));
        CK(cudaMalloc(&c, size_t(n) * n * 8));
        CK(cudaMemset(a, 0, size_t(n) * kk * 8));
        CK(cudaMemset(b, 0, size_t(n) * kk * 8));
        CK(cudaMemset(c, 0, size_t(n) * n * 8));
        const double mone = -1.0, one = 1.0;
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, kk, &mone, a, n, b, kk, &one, c, n);
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, kk, &mone, a, n, b, kk, &one, c, n);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        printf("{\"probe\":\"cublas_dgemm_rank_k\",\"n\":%d,\"k\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", n, kk, best,
               2.0 * n * double(n) * kk / best * 1e-9);
        cudaFree(a);
        cudaFree(b);
        cudaFree(c);
    }
    }
    // ---- cuSOLVER getrf ceiling -------------------------------------------------------------------
    if (want(sel, "cusolver")) {
    cusolverDnHandle_t sh;
    cusolverDnCreate(&sh);
    for (int n : {1024, 4096, 8192, 16384, 32768}) {
        if (n > 8192 && !big) continue;
        double *a;
        int *ipiv, *info;
        const size_t bytes = size_t(n) * n * 8;
        CK(cudaMalloc(&a, bytes));
        CK(cudaMalloc(&ipiv, n * 4));
        CK(cudaMalloc(&info, 4));
        std::vector<double> hrow(size_t(n) * 64);
        int lwork = 0;
        cusolverDnDgetrf_bufferSize(sh, n, n, a, n, &lwork);
        double *work;
        CK(cudaMalloc(&work, size_t(lwork) * 8));
        float best = 1e30f;
        for (int r = 0; r < 2; ++r) {
            // refill with pseudo-random data (host LCG, tiled)
            unsigned long long s = 88172645463325252ull;
            for (auto &v : hrow) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                v = double(s >> 11) * (1.0 / 9007199254740992.0);
            }
            for (size_t off = 0; off < size_t(n) * n; off += hrow.size()) {
                size_t cnt = std::min(hrow.size(), size_t(n) * n - off);
                hrow[off % 61] += 1e-3;
                CK(cudaMemcpy(a + off, hrow.data(), cnt * 8, cudaMemcpyHostToDevice));
            }
            CK(cudaEventRecord(e0));
            cusolverDnDgetrf(sh, n, n, a, n, work, ipiv, info);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            best = fminf(best, time_ms(e0, e1));
        }
        int hinfo = -1;
        CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
        printf("{\"probe\":\"cusolver_dgetrf\",\"n\":%d,\"ms\":%.3f,\"tflops\":%.3f,\"info\":%d}\n", n, best,
               2.0 / 3.0 * n * double(n) * n / best * 1e-9, hinfo);
        cudaFree(a);
        cudaFree(ipiv);
        cudaFree(info);
        cudaFree(work);
    }
    }
    return 0;
}
