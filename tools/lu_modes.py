"""LU time per panel-kernel mode (0 grid exchange, 1 cluster pull, 2 cluster push): python tools/lu_modes.py [n ...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096, 8192]
for n in sizes:
    a0 = torch.empty(n, n, dtype=torch.float64, device="cuda")
    rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, s))
    a = torch.empty_like(a0)
    perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    for mode in (0, 1):
        rla.check(l.rla_set_tuning(b"lu_cluster", mode))
        best = 1e30
        for _ in range(4):
            a.copy_(a0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s)); e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(json.dumps(dict(n=n, lu_cluster=mode, ms=best, tflops=2 / 3 * n ** 3 / best * 1e-9, info=int(info.item()))), flush=True)
