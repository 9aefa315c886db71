"""One DGEMM for ncu: python tools/dgemm_once.py m k n [streamk mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
m, k, n = (int(x) for x in sys.argv[1:4])
rla.check(l.rla_set_tuning(b"dgemm_streamk", int(sys.argv[4]) if len(sys.argv) > 4 else 0))
a = torch.rand(m, k, dtype=torch.float64, device="cuda"); b = torch.rand(k, n, dtype=torch.float64, device="cuda")
c = torch.empty(m, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    rla.check(l.rla_dgemm_dev(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, s))
torch.cuda.synchronize()
