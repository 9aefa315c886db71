"""Per-column phase timing of the pushed-row cluster panel kernel (globaltimer stamps of warp 0, CTA rank 0).
For column c: slot 6 = candidates of c posted (a-phase done), 1 = packet sent, 2 = own update + staging done,
3 = packets arrived, 4 = everyone staged, 5 = row pushed, 0 = row c arrived."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
l.rla_set_tuning(b"lu_dbg", 8 | 4)
l.rla_set_tuning(b"lu_cluster", 2)
s = torch.cuda.current_stream().cuda_stream
a = torch.rand(n, n, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
for _ in range(2):
    rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), n, 0, 0, 64, info.data_ptr(), plan.data_ptr(), s))
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 2048)()
rla.check(l.rla_debug_lu_trace(buf))
t = np.array(buf, dtype=np.int64)[:512].reshape(64, 8)
print("n =", n, "info", int(info.item()))
c = np.arange(2, 63)
def show(name, v):
    print(f"{name:46s} median {np.median(v):7.0f} ns  min {v.min():6d}  max {v.max():6d}")
show("row(c-1) arrived -> candidates(c) posted", t[c, 6] - t[c - 1, 0])
show("candidates posted -> packet sent (B1+reduce)", t[c, 1] - t[c, 6])
show("packet sent -> own update + stage done", t[c, 2] - t[c, 1])
show("update done -> packets arrived", t[c, 3] - t[c, 2])
show("packets arrived -> all staged (verdict+B2)", t[c, 4] - t[c, 3])
show("all staged -> push issued", t[c, 5] - t[c, 4])
show("push issued -> row(c) arrived", t[c, 0] - t[c, 5])
show("column total", t[c, 0] - t[c - 1, 0])

full = np.array(buf, dtype=np.int64)
for wi, wname in ((1, "warp 5"), (2, "warp 10"), (3, "warp 15")):
    tw = full[wi * 512:(wi + 1) * 512].reshape(64, 8)
    show(f"{wname}: row(c-1) seen - warp0 saw it", tw[c - 1, 0] - t[c - 1, 0])
    show(f"{wname}: row(c-1) seen -> candidates(c) posted", tw[c, 1] - tw[c - 1, 0])
    show(f"{wname}: posted -> update + stage done", tw[c, 2] - tw[c, 1])
    show(f"{wname}: stage done -> next row seen", tw[c, 0] - tw[c, 2])

print("absolute timeline (ns) relative to warp 5 seeing row(c-1):")
for cc in (20, 21, 40):
    base = full[512:1024].reshape(64, 8)[cc - 1, 0]
    w0 = {k: int(t[cc, k] - base) for k in (6, 1, 2, 3, 4, 5, 0)}
    print(f" c={cc}: warp0 row(c-1) seen {int(t[cc-1,0]-base)}, cand posted {w0[6]}, packet sent {w0[1]}, own update+stage {w0[2]}, packets arrived {w0[3]}, "
          f"B2 issued {w0[4]}, push done {w0[5]}, row(c) seen {w0[0]}")
    for wi, wname in ((1, "w5"), (2, "w10"), (3, "w15")):
        tw = full[wi * 512:(wi + 1) * 512].reshape(64, 8)
        print(f"        {wname}: row(c-1) seen {int(tw[cc-1,0]-base)}, posted {int(tw[cc,1]-base)}, staged {int(tw[cc,2]-base)}, row(c) seen {int(tw[cc,0]-base)}")
