"""N>1 host logic on CPU: world_size-2 gloo run of the row-panel sharding plumbing (plan, k-chunk
broadcast of B from rank 0, per-chunk accumulation).  The arithmetic stand-in is the oracle (tests
may use it); the product path's arithmetic is CUDA-only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rulinalg_b200.sharded import RowPanelGemm, make_plan


def test_plan_shapes():
    p = make_plan(8, 3, 4096, 32768, 32768, chunk_rows=2048)
    assert p.m_global == 32768 and p.row_range == (3 * 4096, 4 * 4096)
    assert len(p.k_chunks) == 16 and p.k_chunks[0] == (0, 2048) and p.k_chunks[-1] == (30720, 2048)
    assert sum(kc for _, kc in p.k_chunks) == 32768
    assert p.flops_global == 2.0 * 32768 ** 3
    p1 = make_plan(1, 0, 8192, 8192, 8192)
    assert p1.k_chunks == ((0, 8192),)
    ragged = make_plan(2, 1, 10, 100, 7, chunk_rows=48)
    assert ragged.k_chunks == ((0, 48), (48, 48), (96, 4))
    assert make_plan(2, 0, 4, 0, 4).k_chunks == ()
    with pytest.raises(ValueError):
        make_plan(2, 2, 1, 1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, m_local, k, n, out_dir):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = make_plan(world, rank, m_local, k, n, chunk_rows=32)
    a_full = oracle.fill_uniform((m_local * world, k), 12)
    r0, r1 = plan.row_range
    a_local = torch.from_numpy(a_full[r0:r1].copy())
    b = torch.from_numpy(oracle.fill_uniform((k, n), 2049)) if rank == 0 else torch.zeros(k, n, dtype=torch.float64)
    c_local = torch.full((m_local, n), float("nan"), dtype=torch.float64)

    def gemm_fn(a_chunk, b_chunk, c, accumulate):
        prod = oracle.gemm(np.ascontiguousarray(a_chunk.numpy()), np.ascontiguousarray(b_chunk.numpy()))
        if accumulate:
            c += torch.from_numpy(prod)
        else:
            c.copy_(torch.from_numpy(prod))

    RowPanelGemm(plan, torch.float64, gemm_fn=gemm_fn).run(a_local, b, c_local)
    np.save(os.path.join(out_dir, f"c{rank}.npy"), c_local.numpy())
    np.save(os.path.join(out_dir, f"b{rank}.npy"), b.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_panel_gemm_gloo_world2(tmp_path, oracle):
    world, m_local, k, n = 2, 24, 100, 36
    mp.spawn(_worker, args=(world, _free_port(), m_local, k, n, str(tmp_path)), nprocs=world, join=True)
    a = oracle.fill_uniform((m_local * world, k), 12)
    b = oracle.fill_uniform((k, n), 2049)
    c = np.concatenate([np.load(tmp_path / f"c{r}.npy") for r in range(world)])
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"b{r}.npy"), b)       # B arrived everywhere, bit-exact
    ref = oracle.gemm(a, b)
    assert not np.isnan(c).any()
    np.testing.assert_allclose(c, ref, rtol=1e-13)
