"""A/B of the two panel kernels (grid-wide vs thread-block cluster): results must be bit-identical; prints timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
sizes = [int(x) for x in sys.argv[1:]] or [65, 100, 256, 300, 777, 1024, 2048, 3000, 4096, 6144, 8192]
for dt, fn in ((torch.float64, l.rla_dgetrf_dev), (torch.float32, l.rla_sgetrf_dev)):
    for n in sizes:
        torch.manual_seed(n)
        a0 = torch.rand(n, n, dtype=dt, device="cuda") - 0.5
        res = {}
        for mode in (0, 1):
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
        same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
        print(f"{str(dt)[6:]} n={n:6d} grid {res[0][3]:8.3f} ms  cluster {res[1][3]:8.3f} ms  info {res[1][2]}  identical={same}", flush=True)
        if not same:
            d = (res[0][0] != res[1][0])
            rows = d.any(dim=1).nonzero().flatten(); cols = d.any(dim=0).nonzero().flatten()
            print("   mismatch rows", rows[:8].tolist(), "n", rows.numel(), "cols", cols[:8].tolist(), "n", cols.numel(),
                  "perm equal", torch.equal(res[0][1], res[1][1]), flush=True)
l.rla_set_tuning(b"lu_cluster", 1)
