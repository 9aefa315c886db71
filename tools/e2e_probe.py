"""host-API (pinned) DGEMM end-to-end timing vs pipeline shape"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
n = 8192
a = torch.rand(n, n, dtype=torch.float64).pin_memory(); b = torch.rand(n, n, dtype=torch.float64).pin_memory()
c = torch.empty(n, n, dtype=torch.float64).pin_memory()
def run():
    rla.check(l.rla_dgemm(n, n, n, 1.0, a.data_ptr(), n, 1, b.data_ptr(), n, 1, 0.0, c.data_ptr(), n, 1))
for mode, S in ((0, 8), (1, 2), (1, 4), (1, 6), (1, 8), (1, 12), (1, 16), (1, 32)):
    l.rla_set_tuning(b"host_gemm_2d", mode); l.rla_set_tuning(b"host_gemm_s", S)
    run()
    ts = []
    for _ in range(4):
        t = time.perf_counter(); run(); ts.append(time.perf_counter() - t)
    print(json.dumps(dict(mode="2d" if mode else "1d", S=S, ms=min(ts) * 1e3, tflops=2 * n ** 3 / min(ts) * 1e-12)), flush=True)
