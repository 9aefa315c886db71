"""ONE rla_dgetrf call on a pinned host matrix over N GPUs of this process, with phase times (RLA_MULTI_TRACE=1):
python tools/multi_lu_probe.py N [n]"""
import os, sys, time
os.environ["RLA_MULTI_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rulinalg_b200 as rla
N = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
buf = torch.empty(n, n, dtype=torch.float64).pin_memory()
tmp = torch.empty(2048, n, dtype=torch.float64, device="cuda")
lu = buf.numpy(); perm = np.empty(n, dtype=np.uint64)
rla.check(l.rla_set_devices(N))
for rep in range(3):
    for r0 in range(0, n, 2048):
        rla.check(l.rla_fill_uniform_f64_dev(tmp.data_ptr(), 2048, n, n, 12, r0 * n, 0.0, 1.0, s)); buf[r0:r0 + 2048].copy_(tmp)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); st = l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data); print("status", st, "ms", (time.perf_counter() - t0) * 1e3, flush=True)
