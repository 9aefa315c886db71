"""A few device-resident GEMM launches for ncu --set full captures: python tools/gemm_once.py d|s n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
which = sys.argv[1] if len(sys.argv) > 1 else "d"; n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
dt = torch.float64 if which == "d" else torch.float32
a = torch.rand(n, n, dtype=dt, device="cuda"); b = torch.rand(n, n, dtype=dt, device="cuda"); c = torch.empty(n, n, dtype=dt, device="cuda")
fn = l.rla_dgemm_dev if which == "d" else l.rla_sgemm_dev
for _ in range(3):
    rla.check(fn(n, n, n, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0, c.data_ptr(), n, s))
torch.cuda.synchronize()
