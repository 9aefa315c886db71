"""rulinalg_b200 -- B200-native (sm_100a) implementation of rulinalg's dense hot path:
`&Matrix * &Matrix` (f32/f64) and `PartialPivLu::decompose/solve`, behind the reference's own
Matrix / MatrixSlice / Vector / PartialPivLu API.  All arithmetic runs in librla_b200.so
(hand-written CUDA; C ABI in include/rla_b200.h); there is no CPU fallback."""
from .error import Error, ErrorKind, Panic
from ._lib import RlaError, DeviceBuffer, lib, check, LIB_PATH, SYMBOLS
from .matrix import Matrix, MatrixSlice, MatrixSliceMut, Vector, PermutationMatrix, PartialPivLu, LUP, Cholesky

__all__ = ["Matrix", "MatrixSlice", "MatrixSliceMut", "Vector", "PermutationMatrix", "PartialPivLu", "LUP", "Cholesky",
           "Error", "ErrorKind", "Panic", "RlaError", "DeviceBuffer", "lib", "check", "LIB_PATH", "SYMBOLS"]
