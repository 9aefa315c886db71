"""A/B of the column-slab panel kernel (lu_cluster = 3) against the cluster pull kernel (1): bit-identical results, timings.
python tools/lu_slab_check.py [n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [5, 64, 65, 100, 256, 300, 481, 777, 1024, 2048, 3000, 3840, 4096, 8192]
ok = True
for dt, fn in ((torch.float64, l.rla_dgetrf_dev), (torch.float32, l.rla_sgetrf_dev)):
    for n in sizes:
        torch.manual_seed(n)
        a0 = torch.rand(n, n, dtype=dt, device="cuda") - 0.5
        res = {}
        for mode in (1, 3):
            l.rla_set_tuning(b"lu_cluster", mode)
            perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
            best = 1e9
            for it in range(4):
                a = a0.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                rla.check(fn(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s))
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[mode] = (a, perm.clone(), int(info.item()), best)
        same = torch.equal(res[1][0], res[3][0]) and torch.equal(res[1][1], res[3][1]) and res[1][2] == res[3][2]
        ok = ok and same
        print(f"{str(dt)[6:]} n={n:6d} cluster {res[1][3]:8.3f} ms  slab {res[3][3]:8.3f} ms  info {res[3][2]}  identical={same}", flush=True)
        if not same:
            d = (res[1][0] != res[3][0])
            rows = d.any(dim=1).nonzero().flatten(); cols = d.any(dim=0).nonzero().flatten()
            print("   mismatch rows", rows[:8].tolist(), "n", rows.numel(), "cols", cols[:8].tolist(), "n", cols.numel(),
                  "perm equal", torch.equal(res[1][1], res[3][1]), flush=True)
l.rla_set_tuning(b"lu_cluster", 1)
print("ALL IDENTICAL" if ok else "MISMATCH")
