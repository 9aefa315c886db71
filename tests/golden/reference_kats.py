"""Known-answer vectors transcribed from the reference's own tests for the hot path.

Each entry cites the reference file:line (relative to /root/reference).  These are the
golden vectors that pin the oracle (tests/test_oracle_kat.py) and that the CUDA path must
also reproduce through the C ABI (tests/test_gpu_kat.py).  The Rust reference cannot be
run in this image (no rustc), so the vectors are the expected values written in the
reference's test sources, not outputs generated here.
"""
import numpy as np

# --- GEMM: exact small-integer KATs -------------------------------------------------------
# src/matrix/mat_mul.rs:293-341 (f32 and f64), README.md:62-82
GEMM_3x2_2x3 = dict(
    a=[[1., 2.], [3., 4.], [5., 6.]],
    b=[[1., 2., 3.], [4., 5., 6.]],
    c=[[9., 12., 15.], [19., 26., 33.], [29., 40., 51.]],
)
# src/matrix/mat_mul.rs:368-396: d = 2x2 slice at [1,1] of a 3x3 of 2s (row_stride 3)
GEMM_SLICE_BASIC = dict(
    parent=np.full((3, 3), 2.0), start=(1, 1), rows=2, cols=2,
    times_ones=np.full((2, 2), 4.0),      # &d * ones(2,2)
    times_self=np.full((2, 2), 8.0),      # &d * &d
)
# src/matrix/mat_mul.rs:398-412: slice [0,0] 2x2 of a 2x3 (row_stride 3, cols 2) times 2x2
GEMM_SLICE_UNEVEN = dict(
    parent=[[1., 2., 3.], [4., 5., 6.]], start=(0, 0), rows=2, cols=2,
    rhs=[[1., 2.], [3., 4.]],
    c=[[7., 10.], [19., 28.]],
)

# --- LU -----------------------------------------------------------------------------------
LAPLACIAN_9 = [
    [-4., 1., 0., 1., 0., 0., 0., 0., 0.],
    [1., -4., 1., 0., 1., 0., 0., 0., 0.],
    [0., 1., -4., 0., 0., 1., 0., 0., 0.],
    [1., 0., 0., -4., 1., 0., 1., 0., 0.],
    [0., 1., 0., 1., -4., 1., 0., 1., 0.],
    [0., 0., 1., 0., 1., -4., 0., 0., 1.],
    [0., 0., 0., 1., 0., 0., -4., 1., 0.],
    [0., 0., 0., 0., 1., 0., 1., -4., 1.],
    [0., 0., 0., 0., 0., 1., 0., 1., -4.],
]
# tests/mat/mod.rs:4-26  (comp = abs, tol = 1e-8)
SOLVE_LAPLACIAN = dict(
    a=LAPLACIAN_9,
    b=[-100., 0., 0., -100., 0., 0., -100., 0., 0.],
    x=[42.85714286, 18.75, 7.14285714, 52.67857143, 25.0, 9.82142857, 42.85714286, 18.75, 7.14285714],
    abs_tol=1e-8,
)
# tests/mat/mod.rs:100-124  (exact factors; comp = float)
LU_EXACT_3x3 = dict(
    a=[[1., 3., 5.], [2., 4., 7.], [1., 1., 0.]],
    l=[[1., 0., 0.], [0.5, 1., 0.], [0.5, -1., 1.]],
    u=[[2., 4., 7.], [0., 1., 1.5], [0., 0., -2.]],
    p=[[0., 1., 0.], [1., 0., 0.], [0., 0., 1.]],
)
# tests/mat/mod.rs:127-170 + src/matrix/decomposition/lu.rs:775-793: P^-1 L U == A (comp = float)
LU_RECONSTRUCT = [
    [[1., 2., 3., 4., 5.], [3., 0., 4., 5., 6.], [2., 1., 2., 3., 4.], [0., 0., 0., 6., 5.], [0., 0., 0., 5., 6.]],
    LAPLACIAN_9,
    [[1., 1., 0., 0.], [0., 0., 1., 0.], [-1., 0., 0., 0.], [0., 0., 0., 1.]],
    [[-3., 0., 4., 1.], [-12., 5., 17., 1.], [15., 0., -18., -5.], [6., 20., -10., -15.]],
]
# src/matrix/decomposition/lu.rs:808-822 (comp = float)
LU_INVERSE_4x4 = dict(
    a=[[5., 0., 0., 1.], [2., 2., 2., 1.], [4., 5., 5., 5.], [1., 6., 4., 5.]],
    inv=[[1.85185185185185203e-01, 1.85185185185185175e-01, -7.40740740740740561e-02, -1.02798428206033007e-17],
         [1.66666666666666630e-01, 6.66666666666666519e-01, -6.66666666666666519e-01, 4.99999999999999833e-01],
         [-3.88888888888888840e-01, 1.11111111111111174e-01, 5.55555555555555358e-01, -4.99999999999999833e-01],
         [7.40740740740740838e-02, -9.25925925925925819e-01, 3.70370370370370294e-01, 5.13992141030165006e-17]],
)
# src/matrix/decomposition/lu.rs:835-845 (assert_scalar_eq!, comp = float)
LU_DET_4x4 = dict(
    a=[[5., 0., 0., 1.], [0., 2., 2., 1.], [15., 4., 7., 10.], [5., 2., 17., 32.]],
    det=149.99999999999997,
)
# src/matrix/decomposition/lu.rs:848-862 (comp = ulp, tol = 100)
LU_SOLVE_4x4 = dict(
    a=[[5., 0., 0., 1.], [2., 2., 2., 1.], [4., 5., 5., 5.], [1., 6., 4., 5.]],
    b=[9., 16., 49., 45.],
    x=[1., 2., 3., 4.],
    ulp_tol=100,
)
# src/matrix/decomposition/lu.rs:759-772, src/matrix/impl_mat.rs:672-677: singular -> DivByZero / det 0
LU_SINGULAR = [[1., 2., 3., 4.], [0., 0., 0., 0.], [0., 0., 0., 0.], [0., 0., 0., 0.]]
# src/matrix/impl_mat.rs:648-670
DET_5x5 = dict(
    a=[[1., 2., 3., 4., 5.], [3., 0., 4., 5., 6.], [2., 1., 2., 3., 4.], [0., 0., 0., 6., 5.], [0., 0., 0., 5., 6.]],
    det=99.0,
)
# src/matrix/impl_mat.rs:681-693 (exact)
SOLVE_2x2 = dict(a=[[2., 3.], [1., 2.]], b=[8., 5.], x=[1., 2.])
# src/matrix/decomposition/lu.rs:865-889 (exact)
FORWARD_SUBST = [
    dict(lu=np.zeros((0, 0)), b=[], x=[]),
    dict(lu=[[3.0]], b=[1.0], x=[1.0]),
    dict(lu=[[3.0, 2.0], [2.0, 2.0]], b=[1.0, 2.0], x=[1.0, 0.0]),
]
# doc-test src/matrix/decomposition/lu.rs:218-230: identity(4) solve is the identity map
SOLVE_IDENTITY = dict(n=4, b=[3.0, 4.0, 2.0, 1.0])

# ----------------------------------------------------------------------------------------------
# Cholesky: src/matrix/decomposition/cholesky.rs
# ----------------------------------------------------------------------------------------------
# doc-test :39-59 (comp = float)
CHOL_DOC_3x3 = dict(a=[[1., 3., 1.], [3., 13., 11.], [1., 11., 21.]], l=[[1., 0., 0.], [3., 2., 0.], [1., 4., 2.]])
# doc-test :65-79 (assert_vector_eq, exact)
CHOL_DOC_SOLVE = dict(a=[[1., 3., 1.], [3., 13., 11.], [1., 11., 21.]],
                      b1=[3., 2., 1.], b2=[-2., 1., 0.], y1=[23.25, -7.75, 3.0], y2=[-22.25, 7.75, -3.0])
# tests :393-415 (comp = float)
CHOL_UNPACK = [
    dict(a=[[4.0]], l=[[2.0]]),
    dict(a=[[9.0, -6.0], [-6.0, 20.0]], l=[[3.0, 0.0], [-2.0, 4.0]]),
]
# tests :418-442: decompose(x).is_err()
CHOL_SINGULAR = [
    [[0.0]],
    [[0.0, 0.0], [0.0, 1.0]],
    [[1.0, 0.0], [0.0, 0.0]],
    [[1.0, 3.0, 5.0], [3.0, 9.0, 15.0], [5.0, 15.0, 65.0]],
]
# tests :452-466 (comp = float); empty -> 1.0 (:445-449)
CHOL_DET = [
    dict(a=[[1.0]], det=1.0),
    dict(a=[[1.0, 3.0, 5.0], [3.0, 18.0, 33.0], [5.0, 33.0, 65.0]], det=36.0),
]
# tests :469-499 (comp = float)
CHOL_SOLVE = [
    dict(a=[[1.0]], b=[4.0], x=[4.0]),
    dict(a=[[4.0, 6.0], [6.0, 25.0]], b=[2.0, 4.0], x=[0.40625, 0.0625]),
]
# tests :502-539 (comp = float)
CHOL_INVERSE = [
    dict(a=[[2.0]], inv=[[0.5]]),
    dict(a=[[4.0, 6.0], [6.0, 25.0]], inv=[[0.390625, -0.09375], [-0.09375, 0.0625]]),
    dict(a=[[9.0, 6.0, 3.0], [6.0, 20.0, 10.0], [3.0, 10.0, 14.0]],
         inv=[[0.1388888888888889, -0.0416666666666667, 0.0],
              [-0.0416666666666667, 0.0902777777777778, -0.0555555555555556],
              [0.0, -0.0555555555555556, 0.1111111111111111]]),
]
