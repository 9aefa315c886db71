"""1D block-cyclic LU (rulinalg_b200/sharded_lu.py).
CPU: layout arithmetic + a world-size-2 gloo round trip of the scatter/gather plumbing.
GPU: world 1 through the same driver (bit-identical to rla_dgetrf_dev), and world 2 over NCCL when the box
has two GPUs (bit-identical again: every element sees the same operations in the same order)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rulinalg_b200.sharded_lu import BlockCyclicLayout, gather_columns, scatter_columns


def test_layout_arithmetic():
    lay = BlockCyclicLayout(32768, 8, 3)
    assert lay.nblocks == 128 and lay.owner(11) == 3 and lay.local_blocks()[:3] == [3, 11, 19]
    assert lay.ncols_local() == 4096 and lay.local_col0(19) == 512
    assert lay.first_local_col_after(3) == 256 and lay.first_local_col_after(2) == 0 and lay.first_local_col_after(127) == 4096
    ragged = [BlockCyclicLayout(1000, 3, r) for r in range(3)]
    assert [l.ncols_local() for l in ragged] == [488, 256, 256] and ragged[0].width(3) == 232
    a = np.arange(1000 * 1000, dtype=np.float64).reshape(1000, 1000)
    locs = [scatter_columns(a, ragged[r]) for r in range(3)]
    assert np.array_equal(gather_columns(locs, ragged[0]), a)
    one = BlockCyclicLayout(700, 1, 0)
    assert np.array_equal(scatter_columns(a[:700, :700], one), a[:700, :700])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay = BlockCyclicLayout(n, world, rank)
    a = torch.arange(n * n, dtype=torch.float64).reshape(n, n) if rank == 0 else torch.zeros(n, n, dtype=torch.float64)
    dist.broadcast(a, src=0)
    loc = scatter_columns(a, lay)
    # panel-broadcast plumbing stand-in: every owner sends its first block's rows below the diagonal
    for J in range(lay.nblocks):
        w, row0 = lay.width(J), J * lay.block
        buf = loc[row0:, lay.local_col0(J):lay.local_col0(J) + w].contiguous() if rank == lay.owner(J) \
            else torch.empty(n - row0, w, dtype=torch.float64)
        dist.broadcast(buf, src=lay.owner(J))
        assert torch.equal(buf, a[row0:, row0:row0 + w])
    np.save(os.path.join(out_dir, f"loc{rank}.npy"), loc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_scatter_broadcast_gather_gloo_world2(tmp_path):
    n, world = 700, 2
    mp.spawn(_gloo_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    lay = BlockCyclicLayout(n, world, 0)
    got = gather_columns([np.load(tmp_path / f"loc{r}.npy") for r in range(world)], lay)
    assert np.array_equal(got, np.arange(n * n, dtype=np.float64).reshape(n, n))


# ------------------------------------------------------------------------------------------- GPU
def _single_gpu_lu(a_np):
    import rulinalg_b200 as rla
    f = rla.PartialPivLu.decompose(rla.Matrix.from_numpy(a_np))
    return f.lu.to_numpy(), f.p.perm()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [200, 256, 700, 1280])
def test_block_cyclic_lu_world1_matches_getrf(oracle, n):
    import rulinalg_b200 as rla
    from rulinalg_b200.sharded_lu import BlockCyclicLu
    assert rla.lib().rla_init(0) == 0
    torch.cuda.set_device(0)
    a = oracle.fill_uniform((n, n), 12)
    ref_lu, ref_perm = _single_gpu_lu(a)
    lay = BlockCyclicLayout(n, 1, 0)
    a_loc = torch.from_numpy(a).cuda()
    perm, info = BlockCyclicLu(lay).decompose(a_loc)
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    assert np.array_equal(a_loc.cpu().numpy(), ref_lu)
    assert perm.cpu().numpy().tolist() == ref_perm.tolist()
    # singular input -> info != 0
    z = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    z[0, :] = 1.0
    _, info = BlockCyclicLu(lay).decompose(z)
    assert int(info.item()) != 0


def _nccl_worker(rank, world, port, n, lookahead, out_dir, p2p_first=False):
    import oracle
    import rulinalg_b200 as rla
    from rulinalg_b200.sharded_lu import BlockCyclicLu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    assert rla.lib().rla_init(rank) == 0
    lay = BlockCyclicLayout(n, world, rank)
    a = oracle.fill_uniform((n, n), 12)
    a_loc = torch.from_numpy(scatter_columns(a, lay)).cuda()
    perm, info = BlockCyclicLu(lay, lookahead=lookahead, p2p_first=p2p_first).decompose(a_loc)
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"lu{rank}.npy"), a_loc.cpu().numpy())
    np.save(os.path.join(out_dir, f"perm{rank}.npy"), perm.cpu().numpy())
    np.save(os.path.join(out_dir, f"info{rank}.npy"), info.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world,n,lookahead,p2p_first", [(2, 700, False, False), (2, 1280, True, False), (2, 2048, True, False),
                                                         (4, 2048, True, False), (4, 3000, True, True), (8, 4608, True, False),
                                                         (8, 4608, True, True)])
def test_block_cyclic_lu_world2_nccl(tmp_path, oracle, world, n, lookahead, p2p_first):
    """world > 2 with p2p_first also exercises the optional point-to-point-first path"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_nccl_worker, args=(world, _free_port(), n, lookahead, str(tmp_path), p2p_first), nprocs=world, join=True)
    lay = BlockCyclicLayout(n, world, 0)
    got = gather_columns([np.load(tmp_path / f"lu{r}.npy") for r in range(world)], lay)
    perms = [np.load(tmp_path / f"perm{r}.npy") for r in range(world)]
    assert all(int(np.load(tmp_path / f"info{r}.npy")[0]) == 0 for r in range(world))
    assert all(np.array_equal(perms[0], q) for q in perms[1:])
    a = oracle.fill_uniform((n, n), 12)
    ref_lu, ref_perm = _single_gpu_lu(a)
    assert perms[0].tolist() == ref_perm.tolist()
    assert np.array_equal(got, ref_lu)
