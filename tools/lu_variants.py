"""timing experiments for the LU panel kernel (development tool)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
def run(n, reps=3):
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda"); a = torch.empty_like(a0)
    perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    best = 1e30
    for _ in range(reps):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s)); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (8192, 16384, 32768):
    for thr in (64, 96, 128, 147):
        l.rla_set_tuning(b"lu_gmax", 147 if thr > 112 else 112); l.rla_set_tuning(b"lu_dbg", thr << 8)
        ms = run(n)
        print(json.dumps(dict(n=n, direct_g=thr, ms=ms)), flush=True)
