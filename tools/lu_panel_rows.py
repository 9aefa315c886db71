"""Time of ONE 64-column panel against the panel height, per panel kernel (lu_cluster mode 4 = cluster pull K3b, 3 = column slab K3d,
0 = grid K3): python tools/lu_panel_rows.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
for n in (64, 128, 256, 480, 700, 960, 1400, 1920, 2800, 3840, 4096, 5000, 6200):
    torch.manual_seed(n)
    a0 = torch.rand(n, 64, dtype=torch.float64, device="cuda") - 0.5
    out = dict(rows=n)
    for mode, name in ((0, "grid_K3"), (5, "grid_cluster_K3e"), (4, "cluster_K3b"), (3, "slab_K3d")):
        l.rla_set_tuning(b"lu_cluster", mode)
        best = 1e30
        for rep in range(6):
            a = a0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), 64, 0, 0, 64, info.data_ptr(), plan.data_ptr(), s))
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_us"] = round(best * 1e3, 1)
    print(json.dumps(out), flush=True)
l.rla_set_tuning(b"lu_cluster", 1)
