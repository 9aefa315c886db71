// fill.cu -- seeded synthetic data generator (test/bench support inside the product library so that
// every rank can synthesise its shard in HBM).  Bit-identical to the generator the test-side checker uses:
// value(i) = lo + scale * u(splitmix64(seed*K + counter)), counter = offset + row*cols + col.
#include "common.cuh"

namespace rla {
namespace {

__device__ __forceinline__ double to_unit(uint64_t r, double) { return double(r >> 11) * (1.0 / 9007199254740992.0); }
__device__ __forceinline__ float to_unit(uint64_t r, float) { return float(r >> 40) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double fmul_add(double lo, double scale, double u) { return __dadd_rn(lo, __dmul_rn(scale, u)); }
__device__ __forceinline__ float fmul_add(float lo, float scale, float u) { return __fadd_rn(lo, __fmul_rn(scale, u)); }

template <typename T>
__global__ void fill_uniform_kernel(T *dst, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset,
                                    T lo, T scale) {
    const size_t total = rows * cols;
    for (size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += size_t(gridDim.x) * blockDim.x) {
        const size_t r = idx / cols, c = idx - r * cols;
        const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + (offset + idx));
        dst[r * ld + c] = fmul_add(lo, scale, to_unit(h, T(0)));
    }
}

}  // namespace

template <typename T>
int fill_uniform_launch(T *dst, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset, T lo, T scale,
                        cudaStream_t st) {
    const size_t total = rows * cols;
    if (total == 0) return RLA_OK;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    fill_uniform_kernel<T><<<unsigned(blocks), 256, 0, st>>>(dst, rows, cols, ld, seed, offset, lo, scale);
    RLA_LAUNCHED();
    return RLA_OK;
}
template int fill_uniform_launch<double>(double *, size_t, size_t, size_t, uint64_t, uint64_t, double, double, cudaStream_t);
template int fill_uniform_launch<float>(float *, size_t, size_t, size_t, uint64_t, uint64_t, float, float, cudaStream_t);

}  // namespace rla
