"""One device-resident Cholesky factorisation of size n for ncu launch-list captures: python tools/chol_once.py n"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
g = torch.rand(n, min(n, 2048), dtype=torch.float64, device="cuda")
a = g @ g.T / g.shape[1] + torch.eye(n, dtype=torch.float64, device="cuda") * 4
ws = torch.empty(int(l.rla_potrf_workspace_bytes(n, 8)), dtype=torch.uint8, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
rla.check(l.rla_dpotrf_dev(n, a.data_ptr(), n, ws.data_ptr(), info.data_ptr(), s))
torch.cuda.synchronize(); print("info", int(info.item()))
