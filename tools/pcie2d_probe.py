"""PCIe efficiency of strided (2-D) pinned copies, as used by the host GEMM pipeline's B column chunks."""
import json, time
import torch
from cuda.bindings import runtime as rt
n = 8192
h = torch.empty(n, n, dtype=torch.float64).pin_memory()
d = torch.empty(n, n, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
def run(width_cols, rows, kind):
    wbytes = width_cols * 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(4):
        e0.record(st)
        if kind == "h2d":
            err, = rt.cudaMemcpy2DAsync(d.data_ptr(), n * 8, h.data_ptr(), n * 8, wbytes, rows, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, st.cuda_stream)
        else:
            err, = rt.cudaMemcpy2DAsync(h.data_ptr(), n * 8, d.data_ptr(), n * 8, wbytes, rows, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, st.cuda_stream)
        assert err == rt.cudaError_t.cudaSuccess, err
        e1.record(st); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps(dict(kind=kind, width_cols=width_cols, rows=rows, mb=wbytes * rows / 1e6, ms=best, gbs=wbytes * rows / best * 1e-6)), flush=True)
for kind in ("h2d", "d2h"):
    for w in (8192, 4096, 2048, 1024, 512, 256):
        run(w, 8192, kind)
