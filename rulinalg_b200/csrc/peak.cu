// peak.cu -- rla_measure_peak: the roofline denominators, measured on the device the bench runs on.
// Register-only issue-bound loops on every SM (1024 threads per SM, 8 independent chains per thread):
//   kind 0  FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), 512 flop per warp instruction
//   kind 1  FP32 FMA pipe:    fma.rn.f32 (SASS FFMA), 2 flop per lane instruction
// bench.py calls this in-run instead of quoting a constant (MEASURED_PEAKS.json holds no FP64 / FP32 figure).
#include "context.cuh"

namespace rla {
namespace {

constexpr int ILP = 8;

__global__ void __launch_bounds__(1024, 1) dmma_peak_kernel(double *out, int iters, double a0, double b0) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(1024, 1) ffma_peak_kernel(float *out, int iters, float a0, float b0) {
    float c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = float(i);
    const float a = a0 + threadIdx.x * 1e-6f, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fmaf(c[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    if (s == 123.456f) out[0] = s;
}

}  // namespace

int measure_peak(int kind, double *tflops) {
    if (!tflops || kind < 0 || kind > 1) return RLA_ERR_INVALID;
    Context &cx = thread_ctx();
    RLA_TRY(cx.dInfo.ensure(64));
    const int sms = device_num_sms();
    const int iters = kind == 0 ? 20000 : 40000;          // ~5 ms per launch
    const double flop_per_thread = kind == 0 ? double(iters) * ILP * 512.0 / 32.0 : double(iters) * ILP * 2.0;
    cudaEvent_t e0, e1;
    RLA_CUDA(cudaEventCreate(&e0));
    RLA_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {                   // rep 0 = warm-up
        RLA_CUDA(cudaEventRecord(e0, cx.stream));
        if (kind == 0) dmma_peak_kernel<<<sms, 1024, 0, cx.stream>>>(static_cast<double *>(cx.dInfo.p), iters, 1.0, 1.0);
        else ffma_peak_kernel<<<sms, 1024, 0, cx.stream>>>(static_cast<float *>(cx.dInfo.p), iters, 1.0f, 1.0f);
        RLA_LAUNCHED();
        RLA_CUDA(cudaEventRecord(e1, cx.stream));
        RLA_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        RLA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = flop_per_thread * 1024.0 * sms / (double(best) * 1e-3) * 1e-12;
    return RLA_OK;
}

}  // namespace rla
