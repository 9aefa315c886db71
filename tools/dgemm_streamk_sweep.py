"""DGEMM: tiled kernels (auto rule) against the stream-K variants: python tools/dgemm_streamk_sweep.py [n ...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
s = torch.cuda.current_stream().cuda_stream
shapes = [tuple(int(v) for v in x.split("x")) for x in sys.argv[1:]] or [(n, n, n) for n in (512, 768, 1024, 1280, 1536, 1792, 2048, 2560, 3072, 4096)] + [(1024, 4096, 1024), (2048, 512, 2048), (4096, 256, 4096), (65536, 256, 256)]
for (m, k, n) in shapes:
    a = torch.rand(m, k, dtype=torch.float64, device="cuda"); b = torch.rand(k, n, dtype=torch.float64, device="cuda")
    out = dict(m=m, k=k, n=n)
    ref = None
    for mode, name in ((0, "tiled"), (2, "sk128"), (3, "sk64")):
        rla.check(l.rla_set_tuning(b"dgemm_streamk", mode))
        c = torch.full((m, n), float("nan"), dtype=torch.float64, device="cuda")
        best = 1e30
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            rla.check(l.rla_dgemm_dev(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, s))
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_us"] = round(best * 1e3, 1)
        # back-to-back: 20 products queued, host submission hidden behind the device
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            rla.check(l.rla_dgemm_dev(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, s))
        e1.record(); torch.cuda.synchronize()
        out[name + "_us_queued"] = round(e0.elapsed_time(e1) * 1e3 / 20, 1)
        out[name + "_tflops"] = round(2.0 * m * k * n / best * 1e-9, 2)
        if ref is None: ref = c
        else: out[name + "_maxrel_vs_tiled"] = float(((c - ref).abs() / ref.abs()).max().item())
    print(json.dumps(out), flush=True)
rla.check(l.rla_set_tuning(b"dgemm_streamk", 0))
