"""dgetrf time vs the cap on the grid panel kernel's row CTAs (lu_gmax): how many SMs the latency-bound panel may take
from the overlapped Schur update."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
def run(n, reps=2):
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda"); a = torch.empty_like(a0)
    perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    best = 1e30
    for _ in range(reps):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s)); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in [int(x) for x in os.environ.get("NS", "8192,16384,32768").split(",")]:
    for g in [int(x) for x in os.environ.get("GS", "32,48,64,80,96,112,128,147").split(",")]:
        l.rla_set_tuning(b"lu_gmax", g)
        print(json.dumps(dict(n=n, lu_gmax=g, ms=round(run(n), 2))), flush=True)
