import os, sys, json, time
sys.path.insert(0, "/root/repo")
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
n = 8192
a = torch.rand(n, n, dtype=torch.float64).pin_memory(); b = torch.rand(n, n, dtype=torch.float64).pin_memory()
c = torch.empty(n, n, dtype=torch.float64).pin_memory()
def run():
    rla.check(l.rla_dgemm(n, n, n, 1.0, a.data_ptr(), n, 1, b.data_ptr(), n, 1, 0.0, c.data_ptr(), n, 1))
for pre, kc, S, grade in ((4, 256, 0, 0), (4, 256, 0, 1), (4, 256, 8, 0), (4, 256, 12, 0), (4, 256, 24, 0), (4, 256, 32, 0), (4, 128, 0, 0), (3, 256, 0, 0), (5, 256, 0, 0), (4, 384, 0, 0), (2, 256, 0, 0), (6, 256, 0, 0)):
    l.rla_set_tuning(b"host_gemm_kprefix", pre); l.rla_set_tuning(b"host_gemm_kchunk", kc); l.rla_set_tuning(b"host_gemm_s", S); l.rla_set_tuning(b"host_gemm_grade", grade)
    run(); ts = []
    for _ in range(5):
        t = time.perf_counter(); run(); ts.append(time.perf_counter() - t)
    print(json.dumps(dict(kprefix=pre, kchunk=kc, S=S, grade=grade, ms=round(min(ts) * 1e3, 2), tflops=round(2 * n ** 3 / min(ts) * 1e-12, 2))), flush=True)
for k, v in ((b"host_gemm_kprefix", -1), (b"host_gemm_kchunk", 256), (b"host_gemm_s", 0), (b"host_gemm_grade", 0)):
    l.rla_set_tuning(k, v)
