import os, sys, json, time
sys.path.insert(0, "/root/repo")
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
n = 8192
a = torch.rand(n, n, dtype=torch.float32).pin_memory(); b = torch.rand(n, n, dtype=torch.float32).pin_memory()
c = torch.empty(n, n, dtype=torch.float32).pin_memory()
def run():
    rla.check(l.rla_sgemm(n, n, n, 1.0, a.data_ptr(), n, 1, b.data_ptr(), n, 1, 0.0, c.data_ptr(), n, 1))
ref = None
for pre in (0, -1, 3, 6):
    l.rla_set_tuning(b"host_gemm_kprefix", pre)
    c.fill_(float("nan")); run()
    if ref is None: ref = c.clone()
    ts = []
    for _ in range(5):
        t = time.perf_counter(); run(); ts.append(time.perf_counter() - t)
    print(json.dumps(dict(dtype="f32", n=n, kprefix_16ths=pre, ms=round(min(ts) * 1e3, 2), tflops=round(2 * n ** 3 / min(ts) * 1e-12, 2), bit_identical_to_plain=bool(torch.equal(c, ref)))), flush=True)
l.rla_set_tuning(b"host_gemm_kprefix", -1)
