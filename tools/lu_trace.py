"""Per-column phase timing of the LU panel kernel (globaltimer stamps of row CTA 0 and the hub)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rulinalg_b200 as rla
l = rla.lib(); rla.check(l.rla_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
l.rla_set_tuning(b"lu_dbg", 8 | 4)
if len(sys.argv) > 2: l.rla_set_tuning(b"lu_cluster", int(sys.argv[2]))
s = torch.cuda.current_stream().cuda_stream
a = torch.rand(n, n, dtype=torch.float64, device="cuda")
perm = torch.empty(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
# factor only the first 64 columns' worth: run full getrf on a 64-wide... simplest: n x n LU, trace holds the LAST panel; use n where last panel is big: not possible,
# so instead factor a tall problem via the block API: first block only
plan = torch.empty(int(l.rla_lu_plan_bytes()), dtype=torch.uint8, device="cuda")
rla.check(l.rla_dlu_factor_block_dev(n, a.data_ptr(), n, 0, 0, 64, info.data_ptr(), plan.data_ptr(), s))
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 2048)()
rla.check(l.rla_debug_lu_trace(buf))
t = np.array(buf, dtype=np.int64)[:512].reshape(64, 8)
names = ["start->packet (local argmax)", "unused", "unused", "packet->verdict", "row fetch+swap", "div+update", "column total"]
d = np.stack([t[:, 1] - t[:, 0], 0 * t[:, 0], 0 * t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3], t[:, 4] - t[:, 0]], axis=1)
print("n =", n, " info", int(info.item()))
for i, nm in enumerate(names):
    print(f"{nm:34s} median {np.median(d[4:, i]):8.0f} ns   mean {d[4:, i].mean():8.0f}   min {d[4:, i].min():6d} max {d[4:, i].max():6d}")

if t[:, 5].any():
    for nm, v in (("  1 -> 5 (sync A + staging)", t[:, 5] - t[:, 1]), ("  5 -> 6 (wait for packets)", t[:, 6] - t[:, 5]), ("  6 -> 2 (verdict + sync B)", t[:, 2] - t[:, 6])):
        print(f"{nm:34s} median {np.median(v[4:]):8.0f} ns   min {v[4:].min():6d} max {v[4:].max():6d}")

tt = np.array(buf, dtype=np.int64)[1536:1544]
if tt[7] > 0:
    nm = ["entry->init+cluster sync", "load panel + first candidates", "column loop", "registers -> smem", "last fold + publish + write back", "gather outside columns", "final cluster sync"]
    for i in range(7):
        print(f"  kernel phase {nm[i]:34s} {tt[i+1]-tt[i]:8d} ns")
    print(f"  kernel total {tt[7]-tt[0]} ns")
