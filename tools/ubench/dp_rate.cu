// micro-benchmark: throughput of unfused f64 mul + sub (DMUL, DADD) vs DFMA at 512 threads per SM (the LU panel kernel's shape)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double *out, int iters, double m0) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i + threadIdx.x;
    const double m = m0 + threadIdx.x * 1e-9, u = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) c[i] = __dsub_rn(c[i], __dmul_rn(m, u + i));
            else if (MODE == 1) c[i] = __fma_rn(-m, u + i, c[i]);
            else if (MODE == 2) c[i] = __dmul_rn(c[i], m);
            else c[i] = __dadd_rn(c[i], m);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) out[0] = s;
}
int main() {
    double *d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 512>>>(d, iters, 1.0); else if (mode == 1) k<1><<<148, 512>>>(d, iters, 1.0);
            else if (mode == 2) k<2><<<148, 512>>>(d, iters, 1.0); else k<3><<<148, 512>>>(d, iters, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double instr = double(iters) * 16 * (mode == 0 ? 2 : 1) * 512 * 148;
        printf("{\"mode\":\"%s\",\"ms\":%.3f,\"dp_lane_instr_per_clk_per_sm\":%.2f}\n",
               mode == 0 ? "dmul+dsub" : mode == 1 ? "dfma" : mode == 2 ? "dmul" : "dadd", best, instr / (best * 1e-3) / 1.965e9 / 148);
    }
    return 0;
}
