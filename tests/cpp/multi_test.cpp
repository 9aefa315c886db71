// multi_test.cpp -- the multi-GPU path behind the C ABI, exercised the way a Rust host would: plain heap memory,
// rla_set_devices(N), then the same rla_dgemm / rla_dgetrf calls the reference sites make
// (src/matrix/mat_mul.rs:57-67, src/matrix/decomposition/lu.rs:163-195).  The N-GPU results must equal the 1-GPU
// results bit for bit.  Exit codes: 0 ok, 77 fewer than 2 usable GPUs (nothing to test), 1 failure.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "rla_b200.h"

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static void fill(std::vector<double> &v, uint64_t seed) {
    for (size_t i = 0; i < v.size(); ++i) v[i] = double(splitmix64(seed * 0x100000001B3ull + i) >> 11) * (1.0 / 9007199254740992.0);
}
#define CHECK(call)                                                                  \
    do {                                                                             \
        int st_ = (call);                                                            \
        if (st_ != RLA_OK) {                                                         \
            fprintf(stderr, "%s -> %d (%s)\n", #call, st_, rla_strerror(st_));       \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char **argv) {
    const int want = argc > 1 ? atoi(argv[1]) : 2;
    if (rla_device_count() < want || rla_init(0) != RLA_OK) {
        fprintf(stderr, "multi_test: needs %d sm_100 GPUs (there is no CPU fallback)\n", want);
        return 77;
    }
    // ---- GEMM: ragged row split (m not a multiple of 128 * N), pageable operands ----
    const size_t m = 9000, k = 4096, n = 6016;
    std::vector<double> a(m * k), b(k * n), c1(m * n), cN(m * n);
    fill(a, 12);
    fill(b, 2049);
    CHECK(rla_set_devices(1));
    CHECK(rla_dgemm(m, k, n, 1.0, a.data(), k, 1, b.data(), n, 1, 0.0, c1.data(), n, 1));
    CHECK(rla_set_devices(want));
    if (rla_get_devices() != want) { fprintf(stderr, "rla_get_devices\n"); return 1; }
    CHECK(rla_dgemm(m, k, n, 1.0, a.data(), k, 1, b.data(), n, 1, 0.0, cN.data(), n, 1));
    if (memcmp(c1.data(), cN.data(), m * n * sizeof(double)) != 0) {
        size_t bad = 0;
        for (size_t i = 0; i < m * n; ++i) bad += c1[i] != cN[i];
        fprintf(stderr, "dgemm: %zu of %zu elements differ between 1 and %d GPUs\n", bad, m * n, want);
        return 1;
    }
    // spot values against a long double dot
    for (size_t s = 0; s < 64; ++s) {
        const size_t i = splitmix64(s) % m, j = splitmix64(s + 1000) % n;
        long double t = 0;
        for (size_t q = 0; q < k; ++q) t += (long double)a[i * k + q] * b[q * n + j];
        const double err = double((long double)cN[i * n + j] - t);
        if (!(err < 1e-9 && err > -1e-9)) { fprintf(stderr, "dgemm: C[%zu,%zu] off by %g\n", i, j, err); return 1; }
    }
    // ---- LU: n = 8192 -> 32 column blocks dealt over the GPUs ----
    const size_t nn = 8192;
    std::vector<double> l1(nn * nn), lN;
    fill(l1, 12);
    lN = l1;
    std::vector<size_t> p1(nn), pN(nn);
    CHECK(rla_set_devices(1));
    CHECK(rla_dgetrf(nn, l1.data(), p1.data()));
    CHECK(rla_set_devices(want));
    CHECK(rla_dgetrf(nn, lN.data(), pN.data()));
    if (memcmp(p1.data(), pN.data(), nn * sizeof(size_t)) != 0) { fprintf(stderr, "dgetrf: perm differs\n"); return 1; }
    if (memcmp(l1.data(), lN.data(), nn * nn * sizeof(double)) != 0) {
        size_t bad = 0;
        for (size_t i = 0; i < nn * nn; ++i) bad += l1[i] != lN[i];
        fprintf(stderr, "dgetrf: %zu elements differ between 1 and %d GPUs\n", bad, want);
        return 1;
    }
    CHECK(rla_set_devices(1));
    CHECK(rla_shutdown());
    printf("multi_test ok (%d GPUs): dgemm %zux%zux%zu and dgetrf %zu bit-identical to 1 GPU\n", want, m, k, n, nn);
    return 0;
}
