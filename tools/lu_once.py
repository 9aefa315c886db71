"""One device-resident LU (+solve) of size n for ncu launch-list captures: python tools/lu_once.py n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
a = torch.empty(n, n, dtype=torch.float64, device="cuda")
rla.check(l.rla_fill_uniform_f64_dev(a.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, s))
perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s))
b = torch.ones(n, dtype=torch.float64, device="cuda")
rla.check(l.rla_dgetrs_dev(n, a.data_ptr(), n, perm.data_ptr(), b.data_ptr(), info.data_ptr(), s))
torch.cuda.synchronize(); print("info", int(info.item()))
