"""1D block-cyclic multi-GPU PartialPivLu (BASELINE config C5).

One process per GPU.  Global column block J (256 columns) lives on rank J mod g; every rank stores its blocks
side by side as an n x ncols_loc row-major matrix, so a panel is entirely local to its owner (all rows
present) and the pivot search needs no communication.  Per column block:

    owner   : factor the block in place (librla_b200 panel kernels), pack [header | L11 over L21]
    all     : ONE NCCL broadcast of that buffer (header = net row permutation of the block + info word)
    all     : apply the interchanges to every local column outside the block, U12 = L11^-1 A12,
              A22 -= L21 U12 on the local blocks right of J (DMMA kernel)

With look-ahead the owner of block J+1 updates and factors that block first and its broadcast runs on a
side stream while everybody finishes the tail update of block J.

Measured and switched off (`p2p_first=True` re-enables it): the factored panel J+1 is on the critical path of exactly
ONE other rank -- the owner of block J+2 -- so it can go there first as a point-to-point send, the broadcast to
everybody else following off the critical path.  Through torch.distributed this is SLOWER (8 GPUs, n = 32768:
234.7 ms vs 195.3 ms): ProcessGroupNCCL serialises unbatched send/recv with every other operation of the group, so
the extra 67 MB transfer lands in front of the broadcast instead of beside it.

torch.distributed is plumbing (rendezvous + the broadcast); all arithmetic is librla_b200's CUDA.  The layout
arithmetic is pure host logic and is what the gloo CPU test covers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

BLOCK = 256
HEADER_BYTES = 16384          # >= rla_lu_plan_bytes() + 8, multiple of 256 so the panel stays 16-byte aligned


@dataclass(frozen=True)
class BlockCyclicLayout:
    n: int
    world_size: int
    rank: int
    block: int = BLOCK

    @property
    def nblocks(self) -> int:
        return (self.n + self.block - 1) // self.block

    def owner(self, J: int) -> int:
        return J % self.world_size

    def width(self, J: int) -> int:
        return min(self.block, self.n - J * self.block)

    def local_blocks(self, rank: int | None = None) -> List[int]:
        r = self.rank if rank is None else rank
        return list(range(r, self.nblocks, self.world_size))

    def ncols_local(self, rank: int | None = None) -> int:
        return sum(self.width(J) for J in self.local_blocks(rank))

    def local_col0(self, J: int) -> int:
        """first local column of global block J on its owner"""
        return (J // self.world_size) * self.block

    def first_local_col_after(self, J: int, rank: int | None = None) -> int:
        """first local column (on `rank`) belonging to a global block with index > J"""
        r = self.rank if rank is None else rank
        nxt = [K for K in self.local_blocks(r) if K > J]
        return self.local_col0(nxt[0]) if nxt else self.ncols_local(r)

    def global_cols(self, rank: int | None = None) -> List[Tuple[int, int]]:
        """[(global col0, width)] of the local blocks, in local order"""
        return [(J * self.block, self.width(J)) for J in self.local_blocks(rank)]


def scatter_columns(a_global, layout: BlockCyclicLayout, rank: int | None = None):
    """local n x ncols_loc matrix of `rank` cut out of a global matrix (numpy or torch)"""
    parts = [a_global[:, c0:c0 + w] for (c0, w) in layout.global_cols(rank)]
    if hasattr(a_global, "numpy") or type(a_global).__module__.startswith("torch"):
        import torch
        return torch.cat(parts, dim=1).contiguous() if parts else a_global[:, :0].contiguous()
    import numpy as np
    return np.ascontiguousarray(np.concatenate(parts, axis=1)) if parts else a_global[:, :0].copy()


def gather_columns(locals_, layout: BlockCyclicLayout):
    """inverse of scatter_columns over all ranks (numpy)"""
    import numpy as np
    out = np.empty((layout.n, layout.n), dtype=locals_[0].dtype)
    for r, loc in enumerate(locals_):
        lc = 0
        for (c0, w) in layout.global_cols(r):
            out[:, c0:c0 + w] = loc[:, lc:lc + w]
            lc += w
    return out


class BlockCyclicLu:
    """In-place distributed LU of the local column blocks (f64).  Returns (perm, info) device tensors."""

    def __init__(self, layout: BlockCyclicLayout, group=None, lookahead: bool = True, p2p_first: bool = False):
        import torch
        from . import _lib
        self.layout = layout
        self.group = group
        self.lookahead = lookahead and layout.world_size > 1
        self._l = _lib.lib()
        self._check = _lib.check
        n = layout.n
        dev = torch.device("cuda", torch.cuda.current_device())
        plan_bytes = int(self._l.rla_lu_plan_bytes())
        assert plan_bytes + 8 <= HEADER_BYTES
        self.plan_bytes = plan_bytes
        nbuf = 2 if self.lookahead else 1
        self.bufs = [torch.empty(HEADER_BYTES + n * layout.block * 8, dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        # the next-next owner receives the panel twice (point-to-point first, then as a member of the broadcast): the second
        # copy lands in a scratch buffer so it never overwrites bytes its head update may be reading
        self.p2p_first = bool(p2p_first) and self.lookahead and layout.world_size > 2
        self.scratch = torch.empty_like(self.bufs[0]) if self.p2p_first else None
        self.comm2 = torch.cuda.Stream() if self.p2p_first else None     # carries the redundant broadcast copy only
        self.rowid = torch.empty(n, dtype=torch.int32, device=dev)
        self.perm = torch.empty(n, dtype=torch.int64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        self.side = torch.cuda.Stream(priority=-1) if self.lookahead else None
        self.ev = [torch.cuda.Event() for _ in range(4)]

    # -- pieces ------------------------------------------------------------------------------------
    def _info_view(self, buf):
        """int32 view of the header's info word (header layout: plan | pad | info)"""
        import torch
        off = (self.plan_bytes + 7) // 8 * 8
        return buf[off:off + 4].view(torch.int32)

    def _panel_ptr(self, buf):
        return buf.data_ptr() + HEADER_BYTES

    def _factor_and_pack(self, a_loc, J, buf, prev_info, stream):
        """owner side: factor block J in place, pack [plan | info | panel] into buf (all on `stream`)"""
        import torch
        lay = self.layout
        row0, w, lc0 = J * lay.block, lay.width(J), lay.local_col0(J)
        hinfo = self._info_view(buf)
        with torch.cuda.stream(stream):
            hinfo.copy_(prev_info)                      # the running status travels in the header
            self._check(self._l.rla_dlu_factor_block_dev(lay.n, a_loc.data_ptr(), a_loc.stride(0), row0, lc0, w,
                                                         hinfo.data_ptr(), buf.data_ptr(), stream.cuda_stream))
            rows = lay.n - row0
            panel = buf[HEADER_BYTES:HEADER_BYTES + rows * w * 8].view(torch.float64).view(rows, w)
            panel.copy_(a_loc[row0:, lc0:lc0 + w])

    def _broadcast(self, J, buf, stream, first_to=None):
        """panel J from its owner to everybody; with `first_to` (a rank) that rank gets it point-to-point first and takes
        the broadcast copy into the scratch buffer.  Returns an event recorded when THIS rank's copy in `buf` is usable."""
        import torch
        import torch.distributed as dist
        lay = self.layout
        if lay.world_size == 1:
            return None
        rows, w = lay.n - J * lay.block, lay.width(J)
        nbytes = HEADER_BYTES + rows * w * 8
        owner = lay.owner(J)
        ev = torch.cuda.Event()
        with torch.cuda.stream(stream):
            if first_to is not None and first_to != owner:
                if lay.rank == owner:
                    dist.send(buf[:nbytes], dst=first_to, group=self.group)
                    dist.broadcast(buf[:nbytes], src=owner, group=self.group)
                elif lay.rank == first_to:
                    dist.recv(buf[:nbytes], src=owner, group=self.group)
                    ev.record(stream)                                   # usable here: before the broadcast has even started
                    with torch.cuda.stream(self.comm2):                 # (its own stream: `stream` factors the next block next)
                        dist.broadcast(self.scratch[:nbytes], src=owner, group=self.group)
                    return ev
                else:
                    dist.broadcast(buf[:nbytes], src=owner, group=self.group)
            else:
                dist.broadcast(buf[:nbytes], src=owner, group=self.group)
            ev.record(stream)
        return ev

    def _swaps(self, a_loc, J, buf, stream):
        """everyone: block J's row interchanges on the row-origin vector and on the local columns outside it"""
        import torch
        lay = self.layout
        w, ncl = lay.width(J), a_loc.shape[1]
        info_ptr = self._info_view(buf).data_ptr()
        s = stream.cuda_stream
        with torch.cuda.stream(stream):
            self._check(self._l.rla_lu_rowid_apply_dev(buf.data_ptr(), self.rowid.data_ptr(), info_ptr, s))
            if lay.rank == lay.owner(J):
                lc0 = lay.local_col0(J)
                self._check(self._l.rla_dlu_laswp_dev(a_loc.data_ptr(), a_loc.stride(0), w, buf.data_ptr(), info_ptr,
                                                      0, lc0, lc0 + w, ncl, s))
            else:
                self._check(self._l.rla_dlu_laswp_dev(a_loc.data_ptr(), a_loc.stride(0), w, buf.data_ptr(), info_ptr,
                                                      0, ncl, 0, 0, s))

    def _update(self, a_loc, J, buf, stream, c0, c1):
        """everyone: U12 = L11^-1 A12 and A22 -= L21 U12 on local columns [c0, c1) with panel J"""
        import torch
        lay = self.layout
        if c1 <= c0:
            return
        with torch.cuda.stream(stream):
            self._check(self._l.rla_dlu_update_dev(lay.n, a_loc.data_ptr(), a_loc.stride(0), J * lay.block, lay.width(J),
                                                   self._panel_ptr(buf), lay.width(J), c0, c1,
                                                   self._info_view(buf).data_ptr(), stream.cuda_stream))

    # -- driver ------------------------------------------------------------------------------------
    def decompose(self, a_loc):
        import torch
        lay = self.layout
        assert a_loc.is_cuda and a_loc.dtype == torch.float64 and tuple(a_loc.shape) == (lay.n, lay.ncols_local())
        main = torch.cuda.current_stream()
        self.info.zero_()
        self._check(self._l.rla_lu_rowid_init_dev(self.rowid.data_ptr(), lay.n, main.cuda_stream))
        nb, ncl = lay.nblocks, a_loc.shape[1]
        if not self.lookahead:
            buf = self.bufs[0]
            for J in range(nb):
                if lay.rank == lay.owner(J):
                    self._factor_and_pack(a_loc, J, buf, self.info, main)
                self._broadcast(J, buf, main)
                self._swaps(a_loc, J, buf, main)
                self._update(a_loc, J, buf, main, lay.first_local_col_after(J), ncl)
                self.info.copy_(self._info_view(buf))
        else:
            side = self.side
            ev_ready, ev_head = self.ev[0], self.ev[1]
            # block 0 has no predecessor: factor + broadcast on the main stream
            if lay.rank == lay.owner(0):
                self._factor_and_pack(a_loc, 0, self.bufs[0], self.info, main)
            self._broadcast(0, self.bufs[0], main)
            for J in range(nb):
                cur, nxt = self.bufs[J % 2], self.bufs[(J + 1) % 2]
                have_next = J + 1 < nb
                lo = lay.first_local_col_after(J)
                # panel J+1 is urgent for the owner of block J+2 only
                first = lay.owner(J + 2) if (self.p2p_first and J + 2 < nb) else None
                self._swaps(a_loc, J, cur, main)
                if have_next and lay.rank == lay.owner(J + 1):
                    wn = lay.width(J + 1)                          # lo is block J+1's first local column
                    self._update(a_loc, J, cur, main, lo, lo + wn)                  # head: its columns first
                    ev_head.record(main)
                    side.wait_event(ev_head)                       # also orders `nxt` after its last readers
                    self._factor_and_pack(a_loc, J + 1, nxt, self._info_view(cur), side)
                    ev_next = self._broadcast(J + 1, nxt, side, first)
                    self._update(a_loc, J, cur, main, lo + wn, ncl)                 # tail, overlapped
                else:
                    ev_next = None
                    if have_next:
                        ev_ready.record(main)                      # everything that read `nxt` (panel J-1) is queued
                        side.wait_event(ev_ready)
                        ev_next = self._broadcast(J + 1, nxt, side, first)   # receive panel J+1 under the update below
                    self._update(a_loc, J, cur, main, lo, ncl)
                self.info.copy_(self._info_view(cur))
                if ev_next is not None:
                    main.wait_event(ev_next)
        self._check(self._l.rla_lu_perm_from_rowid_dev(self.rowid.data_ptr(), self.perm.data_ptr(), lay.n,
                                                       self.info.data_ptr(), main.cuda_stream))
        return self.perm, self.info
