"""Row-panel sharded GEMM across the GPUs of one NVSwitch box (BASELINE config C4).

One process per GPU (torchrun); rank r owns rows [r*m_local, (r+1)*m_local) of A and C; B (k x n)
originates on rank 0 and is the path's one real exchange step.  torch.distributed is plumbing
only: B travels as k-panel chunks (rows of B are contiguous in row-major) broadcast over NCCL on a
side stream, and the DMMA kernel consumes chunk j as a rank-kc update (beta = 1 accumulation over
k chunks) while chunk j+1 is still in flight -- event-chained, no host synchronisation.

The plan (which rows / which k chunks) is pure host logic and is what the gloo CPU tests cover; the
arithmetic always goes through librla_b200 (`gemm_fn` is injectable only so those tests can run
the plumbing without a GPU).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Tuple


@dataclass(frozen=True)
class PanelPlan:
    """Static schedule of one sharded product."""
    world_size: int
    rank: int
    m_local: int
    k: int
    n: int
    k_chunks: Tuple[Tuple[int, int], ...]      # (k0, kc) per broadcast chunk, ascending

    @property
    def m_global(self) -> int:
        return self.m_local * self.world_size

    @property
    def row_range(self) -> Tuple[int, int]:
        return self.rank * self.m_local, (self.rank + 1) * self.m_local

    @property
    def flops_local(self) -> float:
        return 2.0 * self.m_local * self.k * self.n

    @property
    def flops_global(self) -> float:
        return self.flops_local * self.world_size


def make_plan(world_size: int, rank: int, m_local: int, k: int, n: int, chunk_rows: int = 2048) -> PanelPlan:
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    if min(m_local, k, n) < 0:
        raise ValueError("negative dimension")
    if world_size == 1 or k == 0:
        chunks: List[Tuple[int, int]] = [(0, k)] if k else []
    else:
        chunk_rows = max(16, (chunk_rows // 16) * 16)       # whole k-slabs of the DMMA kernel
        chunks = [(k0, min(chunk_rows, k - k0)) for k0 in range(0, k, chunk_rows)]
    return PanelPlan(world_size, rank, m_local, k, n, tuple(chunks))


def _device_gemm(dtype_char: str) -> Callable:
    from . import _lib
    fn = _lib.lib().rla_dgemm_dev if dtype_char == "d" else _lib.lib().rla_sgemm_dev

    def call(m, k, n, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, stream):
        _lib.check(fn(m, k, n, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, stream))
    return call


class RowPanelGemm:
    """C_local = A_local * B with B broadcast from rank 0 (device-resident operands, torch tensors)."""

    def __init__(self, plan: PanelPlan, dtype, group=None, gemm_fn: Callable | None = None):
        import torch
        self.plan = plan
        self.group = group
        self.dtype = dtype
        self._char = "d" if dtype == torch.float64 else "s"
        self._gemm = gemm_fn
        self._comm_stream = None
        self._events = []

    def _setup_cuda(self):
        import torch
        if self._gemm is None:
            self._gemm = _device_gemm(self._char)
        if self._comm_stream is None and self.plan.world_size > 1:
            self._comm_stream = torch.cuda.Stream()
            self._events = [torch.cuda.Event() for _ in self.plan.k_chunks]

    def run(self, a_local, b, c_local):
        """a_local (m_local x k), b (k x n; contents valid on rank 0, overwritten elsewhere),
        c_local (m_local x n).  Asynchronous on torch's current stream."""
        import torch
        import torch.distributed as dist
        p = self.plan
        if a_local.is_cuda:
            self._setup_cuda()
            stream = torch.cuda.current_stream()
            sptr = stream.cuda_stream
            if p.world_size == 1:
                self._gemm(p.m_local, p.k, p.n, 1.0, a_local.data_ptr(), a_local.stride(0), b.data_ptr(), b.stride(0),
                           0.0, c_local.data_ptr(), c_local.stride(0), sptr)
                return
            if not p.k_chunks:                            # k == 0: the reference yields zeros (empty sum)
                c_local.zero_()
                return
            comm = self._comm_stream
            comm.wait_stream(stream)                      # B on rank 0 was produced on the compute stream
            with torch.cuda.stream(comm):
                for j, (k0, kc) in enumerate(p.k_chunks):
                    dist.broadcast(b[k0:k0 + kc], src=0, group=self.group)
                    self._events[j].record(comm)
            es = a_local.element_size()
            for j, (k0, kc) in enumerate(p.k_chunks):
                stream.wait_event(self._events[j])
                self._gemm(p.m_local, kc, p.n, 1.0, a_local.data_ptr() + k0 * es, a_local.stride(0),
                           b.data_ptr() + k0 * b.stride(0) * es, b.stride(0), 0.0 if j == 0 else 1.0,
                           c_local.data_ptr(), c_local.stride(0), sptr)
            return
        # CPU tensors: plumbing-only path for the gloo tests (gemm_fn must be injected)
        if self._gemm is None:
            raise RuntimeError("RowPanelGemm on CPU tensors needs an injected gemm_fn; the product path is CUDA-only")
        if not p.k_chunks:
            c_local.zero_()
        for j, (k0, kc) in enumerate(p.k_chunks):
            if p.world_size > 1:
                chunk = b[k0:k0 + kc]
                dist.broadcast(chunk, src=0, group=self.group)
            self._gemm(a_local[:, k0:k0 + kc], b[k0:k0 + kc], c_local, accumulate=(j > 0))
