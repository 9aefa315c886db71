"""Informative library ceilings on the same box, measured through torch (cuBLAS matmul with TF32 off; cuSOLVER getrf via
torch.linalg.lu_factor_ex) on the SAME seeded U[0,1) matrices the bench uses (so info == 0 and real pivoting happens).
These are context for BASELINE.md, not part of the product path.  JSON lines on stdout."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream

def best_ms(fn, reps=4):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for dt, name in ((torch.float64, "cublas_dgemm"), (torch.float32, "cublas_sgemm")):
    for n in (1024, 2048, 4096, 8192):
        a = torch.rand(n, n, dtype=dt, device="cuda"); b = torch.rand(n, n, dtype=dt, device="cuda"); c = torch.empty(n, n, dtype=dt, device="cuda")
        ms = best_ms(lambda: torch.matmul(a, b, out=c))
        print(json.dumps(dict(probe=name, n=n, ms=ms, tflops=2.0 * n ** 3 / ms * 1e-9)), flush=True)
for n in (4096, 8192, 32768):
    a = torch.empty(n, n, dtype=torch.float64, device="cuda")
    rla.check(l.rla_fill_uniform_f64_dev(a.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, s))
    at = a.t().contiguous().t()          # column-major copy: what getrf factors in place (rows of A = rows of the problem)
    torch.cuda.synchronize()
    res = {}
    def run():
        res["out"] = torch.linalg.lu_factor_ex(at, pivot=True, check_errors=False)
    ms = best_ms(run, reps=2 if n > 8192 else 4)
    info = int(res["out"].info.item())
    print(json.dumps(dict(probe="cusolver_dgetrf_via_torch", n=n, ms=ms, tflops=2.0 / 3.0 * n ** 3 / ms * 1e-9, info=info,
                          note="includes torch's own copy of the input (lu_factor_ex is out of place)")), flush=True)
    del a, at, res
    torch.cuda.empty_cache()
