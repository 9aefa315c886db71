"""One 64-column panel through the column-slab kernel (for ncu): python tools/lu_slab_once.py [n] [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 3
l.rla_set_tuning(b"lu_cluster", mode)
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
torch.manual_seed(1)
a0 = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
for rep in range(3):
    a = a0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), n, 0, 0, 64, info.data_ptr(), plan.data_ptr(), s))
    torch.cuda.synchronize()
print("info", int(info.item()))
