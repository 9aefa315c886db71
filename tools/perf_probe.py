"""Quick device-resident perf probe (development tool, not the bench contract): GFLOP/s of the
DGEMM / SGEMM / LU / solve kernels through the C ABI device twins, CUDA-event timed on torch's stream."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rulinalg_b200 as rla


def timed(fn, reps, warm=2):
    st = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    l = rla.lib()
    assert l.rla_init(0) == 0
    torch.cuda.set_device(0)
    s = torch.cuda.current_stream().cuda_stream
    which = sys.argv[1:] or ["dgemm", "sgemm", "lu", "solve"]
    out = []
    if "dgemm" in which or "sgemm" in which:
        shapes = [(n, n, n) for n in (256, 512, 1024, 1536, 2048, 4096, 8192, 16384)] + [(65536, 256, 256), (4096, 32768, 32768), (16384, 256, 16384), (16384, 64, 192)]
        for name, dt, fn, cfg in (("dgemm", torch.float64, l.rla_dgemm_dev, 0), ("dgemm", torch.float64, l.rla_dgemm_dev, 1),
                                  ("dgemm", torch.float64, l.rla_dgemm_dev, 7), ("dgemm", torch.float64, l.rla_dgemm_dev, 5),
                                  ("dgemm", torch.float64, l.rla_dgemm_dev, 6),
                                  ("sgemm", torch.float32, l.rla_sgemm_dev, 0)):
            if name not in which:
                continue
            l.rla_set_tuning(b"dgemm_cfg", cfg)
            for (m, k, n) in shapes:
                if name == "sgemm" and k == 32768:
                    continue
                a = torch.rand(m, k, dtype=dt, device="cuda")
                b = torch.rand(k, n, dtype=dt, device="cuda")
                c = torch.empty(m, n, dtype=dt, device="cuda")
                reps = 3 if m * k * n > 2e11 else 10
                best, med = timed(lambda: rla.check(fn(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, s)), reps)
                rec = dict(op=name, cfg=cfg, m=m, k=k, n=n, ms_best=best, ms_med=med, tflops_best=2 * m * k * n / best * 1e-9, tflops_med=2 * m * k * n / med * 1e-9)
                print(json.dumps(rec), flush=True)
                out.append(rec)
                del a, b, c
    if "chol" in which:
        # Cholesky: flops n^3/3
        for dt, fn, name, es in ((torch.float64, l.rla_dpotrf_dev, "dpotrf", 8), (torch.float32, l.rla_spotrf_dev, "spotrf", 4)):
            for n in (1024, 4096, 8192, 16384, 32768):
                if name == "spotrf" and n > 8192:
                    continue
                m = torch.rand(n, min(n, 2048), dtype=dt, device="cuda")
                a0 = m @ m.T / m.shape[1] + torch.eye(n, dtype=dt, device="cuda") * 4     # SPD, checker-side product (torch)
                del m
                a = torch.empty_like(a0)
                ws = torch.empty(int(l.rla_potrf_workspace_bytes(n, es)), dtype=torch.uint8, device="cuda")
                info = torch.zeros(1, dtype=torch.int32, device="cuda")

                def run():
                    a.copy_(a0)
                    rla.check(fn(n, a.data_ptr(), n, ws.data_ptr(), info.data_ptr(), s))

                def copy_only():
                    a.copy_(a0)

                reps = 2 if n >= 16384 else 5
                best, med = timed(run, reps, warm=1)
                cbest, _ = timed(copy_only, reps, warm=1)
                ms = best - cbest
                print(json.dumps(dict(op=name, n=n, ms=ms, tflops=n ** 3 / 3 / ms * 1e-9, info=int(info.item()))), flush=True)
                del a0, a, ws
                torch.cuda.empty_cache()
    if "gemv" in which:
        for n in (4096, 16384, 32768):
            a = torch.rand(n, n, dtype=torch.float64, device="cuda"); x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty(n, dtype=torch.float64, device="cuda")
            best, med = timed(lambda: rla.check(l.rla_dgemv_dev(n, n, a.data_ptr(), n, x.data_ptr(), y.data_ptr(), s)), 10)
            print(json.dumps(dict(op="dgemv", n=n, ms=best, gbs=8 * n * n / best * 1e-6)), flush=True)
            del a
    if "pcie" in which:
        nbytes = 1 << 30
        h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        s2 = torch.cuda.Stream()
        def h2d(): d.copy_(h, non_blocking=True)
        def d2h(): h.copy_(d, non_blocking=True)
        def both():
            d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
        for name, fn, mult in (("h2d", h2d, 1), ("d2h", d2h, 1), ("bidir", both, 2)):
            best, med = timed(fn, 5)
            print(json.dumps(dict(op="pcie_" + name, gib=1, ms=best, gbs=mult * nbytes / best * 1e-6)), flush=True)
    l.rla_set_tuning(b"dgemm_cfg", -1)
    if "lu" in which:
        for dt, fn, name in ((torch.float64, l.rla_dgetrf_dev, "dgetrf"), (torch.float32, l.rla_sgetrf_dev, "sgetrf")):
            for n in (256, 1024, 2048, 4096, 8192, 16384, 32768):
                if name == "sgetrf" and n > 8192:
                    continue
                a0 = torch.rand(n, n, dtype=dt, device="cuda")
                a = torch.empty_like(a0)
                perm = torch.empty(n, dtype=torch.int64, device="cuda")
                info = torch.zeros(1, dtype=torch.int32, device="cuda")

                def run():
                    a.copy_(a0)
                    rla.check(fn(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), s))

                def copy_only():
                    a.copy_(a0)

                reps = 2 if n >= 16384 else 5
                best, med = timed(run, reps, warm=1)
                cbest, _ = timed(copy_only, reps, warm=1)
                ms = best - cbest
                rec = dict(op=name, n=n, ms=ms, tflops=2 / 3 * n ** 3 / ms * 1e-9, info=int(info.item()))
                print(json.dumps(rec), flush=True)
                if name == "dgetrf" and n in (4096, 32768):
                    b = torch.ones(n, dtype=dt, device="cuda")
                    b0 = b.clone()

                    def solve():
                        b.copy_(b0)
                        rla.check(l.rla_dgetrs_dev(n, a.data_ptr(), n, perm.data_ptr(), b.data_ptr(), info.data_ptr(), s))

                    sbest, smed = timed(solve, 5, warm=1)
                    rec = dict(op="dgetrs", n=n, ms=sbest, gbs=8 * n * n / sbest * 1e-6, info=int(info.item()))
                    print(json.dumps(rec), flush=True)
                if name == "dgetrf" and n in (1024, 4096, 8192):
                    x = torch.empty(n, n, dtype=dt, device="cuda")
                    ibest, _ = timed(lambda: rla.check(l.rla_dgetri_dev(n, a.data_ptr(), n, perm.data_ptr(), x.data_ptr(), n, info.data_ptr(), s)), 3, warm=1)
                    print(json.dumps(dict(op="dgetri", n=n, ms=ibest, tflops=2 * n ** 3 / ibest * 1e-9, info=int(info.item()))), flush=True)
                    del x
                del a0, a
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
