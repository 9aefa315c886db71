"""host-API DGEMM end to end against the k-prefix of the 2-D pipeline: python tools/e2e_kprefix_probe.py [n]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.rand(n, n, dtype=torch.float64).pin_memory(); b = torch.rand(n, n, dtype=torch.float64).pin_memory()
c = torch.empty(n, n, dtype=torch.float64).pin_memory()
def run():
    rla.check(l.rla_dgemm(n, n, n, 1.0, a.data_ptr(), n, 1, b.data_ptr(), n, 1, 0.0, c.data_ptr(), n, 1))
ref = None
for pre, kc in ((0, 512), (2, 512), (3, 512), (4, 256), (4, 512), (4, 1024), (6, 512), (8, 512), (-1, 512)):
    l.rla_set_tuning(b"host_gemm_kprefix", pre); l.rla_set_tuning(b"host_gemm_kchunk", kc)
    c.fill_(float("nan")); run()
    if ref is None: ref = c.clone()
    same = bool(torch.equal(c, ref))
    ts = []
    for _ in range(5):
        t = time.perf_counter(); run(); ts.append(time.perf_counter() - t)
    print(json.dumps(dict(n=n, kprefix_16ths=pre, kchunk=kc, ms=round(min(ts) * 1e3, 2), tflops=round(2 * n ** 3 / min(ts) * 1e-12, 2), bit_identical_to_plain=same)), flush=True)
l.rla_set_tuning(b"host_gemm_kprefix", -1); l.rla_set_tuning(b"host_gemm_kchunk", 256)
# pageable
an, bn = a.numpy().copy(), b.numpy().copy(); cn = np.empty((n, n))
def runp():
    rla.check(l.rla_dgemm(n, n, n, 1.0, an.ctypes.data, n, 1, bn.ctypes.data, n, 1, 0.0, cn.ctypes.data, n, 1))
for pre in (0, -1):
    l.rla_set_tuning(b"host_gemm_kprefix", pre)
    runp(); ts = []
    for _ in range(4):
        t = time.perf_counter(); runp(); ts.append(time.perf_counter() - t)
    print(json.dumps(dict(n=n, pageable=True, kprefix_16ths=pre, ms=round(min(ts) * 1e3, 2), bit_identical_to_plain=bool(np.array_equal(cn, ref.numpy())))), flush=True)
l.rla_set_tuning(b"host_gemm_kprefix", -1)
