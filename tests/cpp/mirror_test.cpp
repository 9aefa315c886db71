// C++ mirror smoke test over the C ABI: reproduces reference KATs (mat_mul.rs:318-341,398-412;
// tests/mat/mod.rs:100-124; lu.rs:848-862; lu.rs:759-772).  Exit 0 = pass, 77 = no device (the
// library refused to compute: there is no CPU fallback), anything else = failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "rulinalg_b200.hpp"

#define REQUIRE(x) do { if (!(x)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); return 1; } } while (0)

int main() {
    using namespace rla;
    try {
        Matrix<double> a(3, 2, {1, 2, 3, 4, 5, 6}), b(2, 3, {1, 2, 3, 4, 5, 6});
        Matrix<double> c = a * b;
        const double want[9] = {9, 12, 15, 19, 26, 33, 29, 40, 51};
        REQUIRE(c.rows() == 3 && c.cols() == 3);
        for (int i = 0; i < 9; ++i) REQUIRE(c.data()[i] == want[i]);
        Matrix<float> af(3, 2, {1, 2, 3, 4, 5, 6}), bf(2, 3, {1, 2, 3, 4, 5, 6});
        Matrix<float> cf = af * bf;
        for (int i = 0; i < 9; ++i) REQUIRE(cf.data()[i] == float(want[i]));
        // strided slice (row_stride 3, cols 2): mat_mul.rs:398-412
        Matrix<double> cc(2, 3, {1, 2, 3, 4, 5, 6});
        MatrixSlice<double> d = MatrixSlice<double>::from_matrix(cc, 0, 0, 2, 2);
        Matrix<double> e = d * Matrix<double>(2, 2, {1, 2, 3, 4});
        REQUIRE(e(0, 0) == 7 && e(0, 1) == 10 && e(1, 0) == 19 && e(1, 1) == 28);
        bool panicked = false;
        try { (void)(Matrix<double>::ones(2, 3) * Matrix<double>::ones(2, 3)); } catch (const Panic &) { panicked = true; }
        REQUIRE(panicked);
        // exact LU factors: tests/mat/mod.rs:100-124
        auto lu = PartialPivLu<double>::decompose(Matrix<double>(3, 3, {1, 3, 5, 2, 4, 7, 1, 1, 0}));
        const double packed[9] = {2, 4, 7, 0.5, 1, 1.5, 0.5, -1, -2};
        for (int i = 0; i < 9; ++i) REQUIRE(lu.lu().data()[i] == packed[i]);
        REQUIRE(lu.p().perm()[0] == 1 && lu.p().perm()[1] == 0 && lu.p().perm()[2] == 2);
        // solve KAT: lu.rs:848-862 (comp = ulp, tol = 100)
        auto lu4 = PartialPivLu<double>::decompose(Matrix<double>(4, 4, {5, 0, 0, 1, 2, 2, 2, 1, 4, 5, 5, 5, 1, 6, 4, 5}));
        Vector<double> y = lu4.solve(Vector<double>({9, 16, 49, 45}));
        for (int i = 0; i < 4; ++i) REQUIRE(std::fabs(y[i] - (i + 1)) <= 100 * 2.2e-16 * (i + 1));
        // singular -> DivByZero: lu.rs:759-772
        bool div0 = false;
        try { PartialPivLu<double>::decompose(Matrix<double>(4, 4, {1, 2, 3, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0})); }
        catch (const Error &er) { div0 = er.kind() == ErrorKind::DivByZero; }
        REQUIRE(div0);
        // Cholesky KATs: cholesky.rs:39-59 (unpack), :65-79 (solve, exact), :418-442 (singular), :502-539 (inverse)
        auto ch = Cholesky<double>::decompose(Matrix<double>(3, 3, {1, 3, 1, 3, 13, 11, 1, 11, 21}));
        const double lwant[9] = {1, 0, 0, 3, 2, 0, 1, 4, 2};
        Matrix<double> lch = ch.unpack();
        for (int i = 0; i < 9; ++i) REQUIRE(lch.data()[i] == lwant[i]);
        Vector<double> y1 = ch.solve(Vector<double>({3, 2, 1}));
        REQUIRE(y1[0] == 23.25 && y1[1] == -7.75 && y1[2] == 3.0);
        REQUIRE(ch.det() == 16.0);
        Matrix<double> inv2 = Cholesky<double>::decompose(Matrix<double>(2, 2, {4, 6, 6, 25})).inverse();
        const double iwant[4] = {0.390625, -0.09375, -0.09375, 0.0625};
        for (int i = 0; i < 4; ++i) REQUIRE(std::fabs(inv2.data()[i] - iwant[i]) <= 4 * 2.2e-16);
        bool decomp_fail = false;
        try { Cholesky<double>::decompose(Matrix<double>(3, 3, {1, 3, 5, 3, 9, 15, 5, 15, 65})); }
        catch (const Error &er) { decomp_fail = er.kind() == ErrorKind::DecompFailure; }
        REQUIRE(decomp_fail);
        bool negative = false;
        try { Cholesky<double>::decompose(Matrix<double>(2, 2, {1, 0, 0, -4})); }
        catch (const Error &er) { negative = er.kind() == ErrorKind::DecompFailure && std::string(er.what()).find("not all positive") != std::string::npos; }
        REQUIRE(negative);
    } catch (const RlaFailure &f) {
        std::fprintf(stderr, "%s\n", f.what());
        return f.status == RLA_ERR_NO_DEVICE ? 77 : 2;
    }
    std::puts("mirror_test ok");
    return 0;
}
