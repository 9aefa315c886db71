"""Panel-level A/B of the two panel kernels: factor ONE 64-column panel of an n x n matrix with both and locate the first difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w = int(sys.argv[2]) if len(sys.argv) > 2 else 64
trials = int(sys.argv[3]) if len(sys.argv) > 3 else 20
l.rla_set_tuning(b"lu_dbg", 8)
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
bad = 0
for t in range(trials):
    torch.manual_seed(t)
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
    out = []
    for mode in (0, 1):
        l.rla_set_tuning(b"lu_cluster", mode)
        a = a0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
        rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), n, 0, 0, w, info.data_ptr(), plan.data_ptr(), s))
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 2048)()
        rla.check(l.rla_debug_lu_trace(buf))
        piv = np.array(buf, dtype=np.int64).reshape(4, 64, 8)[:, :, 7].reshape(-1)[:w].copy()
        out.append((a, plan.clone(), piv))
    d = out[0][0] != out[1][0]
    if d.any() or not torch.equal(out[0][1][:4 + 8 * 512], out[1][1][:4 + 8 * 512]):
        bad += 1
        cols = d.any(dim=0).nonzero().flatten(); rows = d.any(dim=1).nonzero().flatten()
        print(f"trial {t}: first bad col {cols[:4].tolist()} ncols {cols.numel()} rows {rows[:6].tolist()} nrows {rows.numel()}", flush=True)
        dp = np.nonzero(out[0][2] != out[1][2])[0]
        print("   first differing pivot at column", dp[:4].tolist(), "grid", out[0][2][dp[:4]].tolist(), "cluster", out[1][2][dp[:4]].tolist())
        c0 = int(cols[0]) if cols.numel() else -1
        if c0 >= 0:
            rr = d[:, c0].nonzero().flatten()
            print("   rows differing in that column:", rr[:10].tolist(), "count", rr.numel())
            r = int(rr[0]); print("   grid", out[0][0][r, c0].item(), "cluster", out[1][0][r, c0].item())
            for b0 in range(0, w, 64):
                blk = d[:, b0:b0 + 64]
                g, c = out[0][0][:, b0:b0 + 64], out[1][0][:, b0:b0 + 64]
                same_set = torch.equal(torch.sort(g[:, 0]).values, torch.sort(c[:, 0]).values)
                fc = blk.any(dim=0).nonzero().flatten()
                ms = [bool(torch.equal(torch.sort(g[:, j]).values, torch.sort(c[:, j]).values)) for j in range(g.shape[1])]
                first_ms = ms.index(False) if False in ms else -1
                print(f"   block cols {b0}..{b0+63}: differing cols {fc.numel()} first {fc[:3].tolist()} rows differing {int(blk.any(dim=1).sum())} first col whose multiset differs {first_ms}")
            pg = out[0][1][4:4 + 4 * 1024].view(torch.int32); pc = out[1][1][4:4 + 4 * 1024].view(torch.int32)
            print("   plan nt", out[0][1][:4].view(torch.int32).item(), out[1][1][:4].view(torch.int32).item(), "plan rows/origin equal", torch.equal(pg, pc))
print("bad trials:", bad, "of", trials)
