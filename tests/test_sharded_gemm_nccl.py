"""Row-panel sharded GEMM over NCCL (one process per GPU, BASELINE config C4) on real GPUs: every rank's C panel must
equal the single-GPU kernel's result for the same k chunking bit for bit, and agree with the CPU oracle within the GEMM
gate (ulp <= ceil(4 sqrt k) on U[0,1) data).  Needs two B200s; skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m_local, k, n, chunk, dtype_name, out_dir):
    import oracle
    import rulinalg_b200 as rla
    from rulinalg_b200.sharded import RowPanelGemm, make_plan
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    assert rla.lib().rla_init(rank) == 0
    npdt = np.dtype(dtype_name)
    tdt = torch.float64 if npdt == np.float64 else torch.float32
    a = oracle.fill_uniform((m_local * world, k), 12, npdt)
    b = oracle.fill_uniform((k, n), 2049, npdt)
    a_loc = torch.from_numpy(a[rank * m_local:(rank + 1) * m_local].copy()).cuda()
    b_dev = torch.from_numpy(b).cuda() if rank == 0 else torch.zeros(k, n, dtype=tdt, device="cuda")
    c_loc = torch.full((m_local, n), float("nan"), dtype=tdt, device="cuda")
    plan = make_plan(world, rank, m_local, k, n, chunk_rows=chunk)
    op = RowPanelGemm(plan, tdt)
    for _ in range(2):                       # twice: the second run re-broadcasts over a B that is already in place
        op.run(a_loc, b_dev, c_loc)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"c{rank}.npy"), c_loc.cpu().numpy())
    np.save(os.path.join(out_dir, f"b{rank}.npy"), b_dev.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dtype_name,m_local,k,n,chunk", [("float64", 384, 1000, 520, 256), ("float64", 1024, 4096, 1024, 2048),
                                                          ("float32", 512, 2048, 768, 512)])
def test_row_panel_gemm_world2_nccl(tmp_path, oracle, dtype_name, m_local, k, n, chunk):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import rulinalg_b200 as rla
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), m_local, k, n, chunk, dtype_name, str(tmp_path)), nprocs=world, join=True)
    npdt = np.dtype(dtype_name)
    a = oracle.fill_uniform((m_local * world, k), 12, npdt)
    b = oracle.fill_uniform((k, n), 2049, npdt)
    got = np.concatenate([np.load(tmp_path / f"c{r}.npy") for r in range(world)], axis=0)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"b{r}.npy"), b), "B did not arrive intact"
    # (1) the single-GPU kernel on the same k chunking (beta = 1 accumulation over chunks): bit-identical
    assert rla.lib().rla_init(0) == 0
    l = rla.lib()
    fn = l.rla_dgemm_dev if npdt == np.float64 else l.rla_sgemm_dev
    m = m_local * world
    da, db, dc = rla.DeviceBuffer(a.nbytes), rla.DeviceBuffer(b.nbytes), rla.DeviceBuffer(m * n * npdt.itemsize)
    da.upload(a); db.upload(b)
    it = npdt.itemsize
    for j, k0 in enumerate(range(0, k, chunk)):
        kc = min(chunk, k - k0)
        assert rla.check(fn(m, kc, n, 1.0, da.ptr + k0 * it, k, db.ptr + k0 * n * it, n, 0.0 if j == 0 else 1.0, dc.ptr, n, None)) == 0
    ref_gpu = dc.download((m, n), npdt)
    for d in (da, db, dc):
        d.free()
    assert np.array_equal(got, ref_gpu)
    # (2) the CPU oracle
    oracle.assert_matrix_eq(got, oracle.gemm(a, b), comp="ulp", tol=int(np.ceil(4 * np.sqrt(k))))
