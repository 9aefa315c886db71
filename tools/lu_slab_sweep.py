"""LU time against the tallest panel given to the column-slab kernel (lu_slab_rows): python tools/lu_slab_sweep.py [n ...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]
for dt, fn, fill in ((torch.float64, l.rla_dgetrf_dev, l.rla_fill_uniform_f64_dev), (torch.float32, l.rla_sgetrf_dev, l.rla_fill_uniform_f32_dev)):
    for n in sizes:
        a0 = torch.empty(n, n, dtype=dt, device="cuda")
        rla.check(fill(a0.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, s))
        a = torch.empty_like(a0)
        perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
        out = dict(dtype=str(dt)[6:], n=n)
        rla.check(l.rla_set_tuning(b"lu_cluster", 1))
        for thr in (0, 960, 1440, 1920):
            rla.check(l.rla_set_tuning(b"lu_slab_rows", thr))
            best = 1e30
            for _ in range(5):
                a.copy_(a0); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); rla.check(fn(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s)); e1.record(); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[f"ms_slab_rows_{thr}"] = round(best, 4)
        out["info"] = int(info.item())
        print(json.dumps(out), flush=True)
rla.check(l.rla_set_tuning(b"lu_slab_rows", 1920))
