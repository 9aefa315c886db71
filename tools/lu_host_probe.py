"""rla_dgetrf end to end from host memory (pinned / pageable): python tools/lu_host_probe.py [n ...]"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rulinalg_b200 as rla
import oracle
l = rla.lib(); rla.check(l.rla_init(0))
for n in [int(x) for x in sys.argv[1:]] or [2048, 4096, 8192]:
    a0 = oracle.fill_uniform((n, n), 12)
    for kind in ("pinned", "pageable"):
        lu = torch.empty(n, n, dtype=torch.float64).pin_memory().numpy() if kind == "pinned" else np.empty((n, n))
        perm = np.empty(n, dtype=np.uint64)
        best = 1e30
        for _ in range(4):
            lu[...] = a0
            t0 = time.perf_counter(); st = rla.check(l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data)); best = min(best, time.perf_counter() - t0)
        # device-resident reference for bit-identity
        s = torch.cuda.current_stream().cuda_stream
        ad = torch.from_numpy(a0).cuda(); pd = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
        rla.check(l.rla_dgetrf_dev(n, ad.data_ptr(), n, pd.data_ptr(), info.data_ptr(), s)); torch.cuda.synchronize()
        same = bool(np.array_equal(ad.cpu().numpy(), lu)) and bool(np.array_equal(pd.cpu().numpy().astype(np.uint64), perm))
        print(json.dumps(dict(n=n, host=kind, decompose_ms=round(best * 1e3, 3), status=int(st), identical_to_device_path=same)), flush=True)
