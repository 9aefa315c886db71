"""%globaltimer trace of ONE 64-column panel factored by the column-slab kernel: python tools/lu_slab_trace.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
l.rla_set_tuning(b"lu_dbg", 8)
l.rla_set_tuning(b"lu_cluster", 3)
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
torch.manual_seed(1)
a0 = torch.rand(n, n, dtype=torch.float64, device="cuda") - 0.5
for rep in range(3):
    a = a0.clone(); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), n, 0, 0, 64, info.data_ptr(), plan.data_ptr(), s))
    e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 2048)()
rla.check(l.rla_debug_lu_trace(buf))
t = np.array(buf, dtype=np.int64)
col = t[:512].reshape(64, 8)
cta = t[1536:1536 + 128].reshape(16, 8)
t0 = cta[0, 0]
print(f"n={n} panel kernel (event time incl. launch) {e0.elapsed_time(e1)*1e3:.1f} us")
print("CTA: start, loops done, joined, written back (us since CTA 0 start)")
for k in range(16):
    if cta[k, 0] == 0: continue
    print(f"  cta {k:2d}: " + "  ".join(f"{(cta[k, i] - t0) / 1e3:8.2f}" for i in range(4)))
print("col: own-step start | verdict | updated || all warps done | published | next owner saw | next owner applies   (us since start; deltas to previous column's publish)")
prev = t0
for c in range(64):
    r = col[c]
    f = lambda v: f"{(v - t0) / 1e3:8.2f}" if v else "     -  "
    print(f"  c{c:2d}: {f(r[0])} {f(r[1])} {f(r[6])} | {f(r[2])} {f(r[3])} {f(r[4])} {f(r[5])}   step {(r[3] - prev) / 1e3:6.2f}")
    prev = r[3]
