"""ctypes binding of librla_b200.so (include/rla_b200.h).  No CPU fallback: if the library is
missing or no B200 is visible, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librla_b200.so")

RLA_OK = 0
RLA_ERR_SINGULAR = 1
RLA_ERR_INVALID = 2
RLA_ERR_NOT_POSITIVE = 3
RLA_ERR_CUDA = -1
RLA_ERR_NOMEM = -2
RLA_ERR_NO_DEVICE = -3

# every symbol include/rla_b200.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "rla_dgemm", "rla_sgemm", "rla_dgetrf", "rla_sgetrf", "rla_dgetrs", "rla_sgetrs",
    "rla_dgemv", "rla_sgemv", "rla_dgemv_dev", "rla_sgemv_dev", "rla_dtrsv", "rla_strsv", "rla_dgetri", "rla_sgetri", "rla_dgetri_dev", "rla_sgetri_dev",
    "rla_dgetrf_keep", "rla_dlu_solve", "rla_lu_free", "rla_operand_hold", "rla_operand_release", "rla_operand_resident_bytes",
    "rla_dpotrf", "rla_spotrf", "rla_dpotrs", "rla_spotrs", "rla_dpotri", "rla_spotri",
    "rla_potrf_workspace_bytes", "rla_dpotrf_dev", "rla_spotrf_dev",
    "rla_init", "rla_device_count", "rla_set_devices", "rla_get_devices", "rla_shutdown", "rla_dev_alloc", "rla_dev_free", "rla_host_alloc_pinned",
    "rla_host_free_pinned", "rla_memcpy_h2d", "rla_memcpy_d2h", "rla_stream_sync",
    "rla_dgemm_dev", "rla_sgemm_dev", "rla_dgetrf_dev", "rla_sgetrf_dev", "rla_dgetrs_dev", "rla_sgetrs_dev",
    "rla_fill_uniform_f64_dev", "rla_fill_uniform_f32_dev",
    "rla_lu_plan_bytes", "rla_lu_max_n", "rla_debug_lu_trace", "rla_debug_divcheck", "rla_dlu_factor_block_dev", "rla_dlu_laswp_dev", "rla_dlu_update_dev",
    "rla_lu_rowid_init_dev", "rla_lu_rowid_apply_dev", "rla_lu_perm_from_rowid_dev",
    "rla_measure_peak", "rla_set_tuning", "rla_strerror", "rla_last_cuda_error", "rla_version", "rla_launch_count", "rla_launch_count_reset",
]


class RlaError(RuntimeError):
    """Environment failure (CUDA error, no device, out of memory): the Rust shim panics on these."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"librla_b200 status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RlaError(RLA_ERR_NO_DEVICE, f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
    l = C.CDLL(LIB_PATH)
    sz, pd, dbl, flt, P, u64, i32 = C.c_size_t, C.c_ssize_t, C.c_double, C.c_float, C.c_void_p, C.c_uint64, C.c_int
    l.rla_dgemm.argtypes = [sz, sz, sz, dbl, P, pd, pd, P, pd, pd, dbl, P, pd, pd]
    l.rla_sgemm.argtypes = [sz, sz, sz, flt, P, pd, pd, P, pd, pd, flt, P, pd, pd]
    for f in ("rla_dgetrf", "rla_sgetrf"):
        getattr(l, f).argtypes = [sz, P, P]
    for f in ("rla_dgetrs", "rla_sgetrs"):
        getattr(l, f).argtypes = [sz, P, P, P]
    for f in ("rla_dgemv", "rla_sgemv"):
        getattr(l, f).argtypes = [sz, sz, P, pd, P, P]
    for f in ("rla_dgemv_dev", "rla_sgemv_dev"):
        getattr(l, f).argtypes = [sz, sz, P, sz, P, P, P]
    for f in ("rla_dtrsv", "rla_strsv"):
        getattr(l, f).argtypes = [i32, sz, P, pd, P]
    for f in ("rla_dgetri", "rla_sgetri"):
        getattr(l, f).argtypes = [sz, P, P, P]
    for f in ("rla_dgetri_dev", "rla_sgetri_dev"):
        getattr(l, f).argtypes = [sz, P, sz, P, P, sz, P, P]
    for f in ("rla_dpotrf", "rla_spotrf"):
        getattr(l, f).argtypes = [sz, P]
    for f in ("rla_dpotrs", "rla_spotrs", "rla_dpotri", "rla_spotri"):
        getattr(l, f).argtypes = [sz, P, P]
    l.rla_potrf_workspace_bytes.argtypes = [sz, sz]
    l.rla_potrf_workspace_bytes.restype = sz
    for f in ("rla_dpotrf_dev", "rla_spotrf_dev"):
        getattr(l, f).argtypes = [sz, P, sz, P, P, P]
    l.rla_dgetrf_keep.argtypes = [sz, P, P, C.POINTER(P)]
    l.rla_dlu_solve.argtypes = [P, P]
    l.rla_lu_free.argtypes = [P]
    l.rla_lu_free.restype = None
    l.rla_operand_hold.argtypes = [P, sz]
    l.rla_operand_release.argtypes = [P]
    l.rla_operand_resident_bytes.restype = sz
    l.rla_init.argtypes = [i32]
    l.rla_set_devices.argtypes = [i32]
    l.rla_dev_alloc.argtypes = [C.POINTER(P), sz]
    l.rla_dev_free.argtypes = [P]
    l.rla_host_alloc_pinned.argtypes = [C.POINTER(P), sz]
    l.rla_host_free_pinned.argtypes = [P]
    l.rla_memcpy_h2d.argtypes = [P, P, sz, P]
    l.rla_memcpy_d2h.argtypes = [P, P, sz, P]
    l.rla_stream_sync.argtypes = [P]
    l.rla_dgemm_dev.argtypes = [sz, sz, sz, dbl, P, sz, P, sz, dbl, P, sz, P]
    l.rla_sgemm_dev.argtypes = [sz, sz, sz, flt, P, sz, P, sz, flt, P, sz, P]
    for f in ("rla_dgetrf_dev", "rla_sgetrf_dev"):
        getattr(l, f).argtypes = [sz, P, sz, P, P, P]
    for f in ("rla_dgetrs_dev", "rla_sgetrs_dev"):
        getattr(l, f).argtypes = [sz, P, sz, P, P, P, P]
    l.rla_fill_uniform_f64_dev.argtypes = [P, sz, sz, sz, u64, u64, dbl, dbl, P]
    l.rla_fill_uniform_f32_dev.argtypes = [P, sz, sz, sz, u64, u64, flt, flt, P]
    l.rla_lu_plan_bytes.restype = sz
    l.rla_lu_max_n.argtypes = [sz]
    l.rla_lu_max_n.restype = sz
    l.rla_debug_lu_trace.argtypes = [P]
    l.rla_debug_divcheck.argtypes = [i32, i32, u64, u64, C.POINTER(u64)]
    l.rla_dlu_factor_block_dev.argtypes = [sz, P, sz, sz, sz, sz, P, P, P]
    l.rla_dlu_laswp_dev.argtypes = [P, sz, sz, P, P, sz, sz, sz, sz, P]
    l.rla_dlu_update_dev.argtypes = [sz, P, sz, sz, sz, P, sz, sz, sz, P, P]
    l.rla_lu_rowid_init_dev.argtypes = [P, sz, P]
    l.rla_lu_rowid_apply_dev.argtypes = [P, P, P, P]
    l.rla_lu_perm_from_rowid_dev.argtypes = [P, P, sz, P, P]
    l.rla_measure_peak.argtypes = [i32, C.POINTER(dbl)]
    l.rla_set_tuning.argtypes = [C.c_char_p, i32]
    l.rla_strerror.argtypes = [i32]
    l.rla_strerror.restype = C.c_char_p
    l.rla_version.restype = C.c_char_p
    l.rla_launch_count.restype = u64
    l.rla_launch_count_reset.restype = None
    _lib = l
    return l


def check(status: int) -> int:
    """Raise on environment errors; return numerical statuses (0 / RLA_ERR_SINGULAR / RLA_ERR_NOT_POSITIVE) to the caller."""
    if status in (RLA_OK, RLA_ERR_SINGULAR, RLA_ERR_NOT_POSITIVE):
        return status
    l = lib()
    msg = l.rla_strerror(status).decode()
    if status == RLA_ERR_CUDA:
        msg += f" (cudaError {l.rla_last_cuda_error()})"
    raise RlaError(status, msg)


class DeviceBuffer:
    """Raw HBM allocation through the C ABI (used by tests/bench for the device-resident twins)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().rla_dev_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value

    def upload(self, arr, stream=None):
        import numpy as np
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(lib().rla_memcpy_h2d(self.ptr, arr.ctypes.data, arr.nbytes, stream))
        check(lib().rla_stream_sync(stream))

    def download(self, shape, dtype, stream=None):
        import numpy as np
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().rla_memcpy_d2h(out.ctypes.data, self.ptr, out.nbytes, stream))
        check(lib().rla_stream_sync(stream))
        return out

    def free(self):
        if self.ptr:
            lib().rla_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
