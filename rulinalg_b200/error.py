"""Mirror of rulinalg's error type (src/error.rs:10-63)."""
from __future__ import annotations

import enum


class ErrorKind(enum.Enum):
    InvalidArg = "InvalidArg"
    DecompFailure = "DecompFailure"
    AlgebraFailure = "AlgebraFailure"
    DivByZero = "DivByZero"
    ScalarConversionFailure = "ScalarConversionFailure"
    InvalidPermutation = "InvalidPermutation"


class Error(Exception):
    """`Result::Err(Error)`: numerical failure reported by value in Rust, raised here."""

    def __init__(self, kind: ErrorKind, message: str):
        super().__init__(message)
        self._kind = kind

    def kind(self) -> ErrorKind:
        return self._kind


class Panic(AssertionError):
    """Rust `assert!`/`panic!` (shape violations: mat_mul.rs:21, lu.rs:165,232)."""
