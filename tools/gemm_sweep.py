"""Device-resident GEMM sweep: python tools/gemm_sweep.py [s|d] -> JSON lines (best of 5, CUDA events)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else "s"
dt, fn = (torch.float32, l.rla_sgemm_dev) if which == "s" else (torch.float64, l.rla_dgemm_dev)
key = b"sgemm_cfg" if which == "s" else b"dgemm_cfg"
cfgs = [-1, 0, 1] if which == "s" else [-1, 6, 7]
shapes = [(512,)*3, (1024,)*3, (1536,)*3, (2048,)*3, (3072,)*3, (4096,)*3, (8192,)*3, (65536, 256, 256)]
for (m, k, n) in shapes:
    a = torch.rand(m, k, dtype=dt, device="cuda"); b = torch.rand(k, n, dtype=dt, device="cuda"); c = torch.empty(m, n, dtype=dt, device="cuda")
    for cfg in cfgs:
        rla.check(l.rla_set_tuning(key, cfg))
        best = 1e30
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rla.check(fn(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, s)); e1.record(); e1.synchronize()
            if it: best = min(best, e0.elapsed_time(e1))
        print(json.dumps(dict(op=which + "gemm", m=m, k=k, n=n, cfg=cfg, ms=best, tflops=2.0 * m * k * n / best * 1e-9)), flush=True)
    rla.check(l.rla_set_tuning(key, -1))
