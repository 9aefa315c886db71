"""torchrun tool: time the 1D block-cyclic LU at WORLD_SIZE GPUs.  python -m torch.distributed.run ... tools/lu_dist_bench.py n [lookahead]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import rulinalg_b200 as rla
from rulinalg_b200.sharded_lu import BlockCyclicLayout, BlockCyclicLu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lookahead = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
l = rla.lib(); rla.check(l.rla_init(lr))
lay = BlockCyclicLayout(n, world, rank)
ncl = lay.ncols_local()
a0 = torch.empty(n, ncl, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, ncl, ncl, 12 + rank, 0, 0.0, 1.0, s))
a = torch.empty_like(a0)
lu = BlockCyclicLu(lay, lookahead=lookahead)
best = 1e30
for rep in range(3):
    a.copy_(a0)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); perm, info = lu.decompose(a); e1.record(); e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    best = min(best, float(ms.item()))
ok = int(info.item()) == 0 and sorted(perm.cpu().tolist()) == list(range(n))
if rank == 0:
    print(json.dumps(dict(op="dist_dgetrf", n=n, gpus=world, lookahead=lookahead, ms=best, tflops=2 / 3 * n ** 3 / best * 1e-9, ok=ok)), flush=True)
if world > 1: dist.destroy_process_group()
