"""Time one outer block factorisation (4 panels + inner trsm/gemm) with both panel kernels, and single panels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
for n in [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096]:
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
    for w in (64, 256):
        row = []
        for mode in (0, 1):
            l.rla_set_tuning(b"lu_cluster", mode)
            best = 1e9
            for it in range(5):
                a = a0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), n, 0, 0, w, info.data_ptr(), plan.data_ptr(), s))
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            row.append(best)
        print(f"n={n} w={w}: grid {row[0]*1e3:8.1f} us  cluster {row[1]*1e3:8.1f} us   per column {row[0]*1e3/w:6.2f} / {row[1]*1e3/w:6.2f} us", flush=True)
