#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for rulinalg_b200.

  python bench.py --gpus N --steps K --warmup W              (our arm; torchrun launches N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (reference arm: CPU port)

Metric (BASELINE.json): DGEMM GFLOP/s, with % of the measured FP64 tensor (DMMA) peak in `roofline`.
Workload at N=1: f64 `&A * &B`, A, B 8192 x 8192 (the point of BASELINE configs[1]'s sweep that the
>= 80 %-of-peak target is quoted on).  A "step" is one full product.
Workload at N>1: BASELINE configs[3] (C4) as SURVEY 8d states it -- the weak series m = 4096*N, k = n = 32768:
every GPU owns 4096 rows of A and C, B (8 GiB) is broadcast from rank 0 over NCCL in 16 k chunks of 2048 rows
overlapped with the DMMA kernel.  The strong 32768^3 point and C5 (LU n = 32768, 1D block-cyclic) are in `extras`.
Every N>1 run verifies itself: Freivalds + sampled extended-precision entries per rank for the GEMM, sampled-row
reconstruction P A = L U for the distributed LU (`parity_ok`).

  value  = whole-job GFLOP/s with operands resident in HBM (CUDA events, max over ranks)
  e2e    = same metric through the reference-facing host call: rla_dgemm with HOST buffers, H2D of A and B and
           D2H of C inside the timed region; at N>1 rank 0 makes that one call after rla_set_devices(N)
           (one host process driving N GPUs -- what a Rust caller gets), the other ranks idle
  extras = the other BASELINE configs measured the same way (SGEMM, LU, solve, sweep points, host-API LU,
           pageable-memory e2e, the reference's own bench shapes)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SQUARE = 8192
N_WIDE = 32768                    # C4 / C5 width
M_PER_GPU_WEAK = 4096             # C4 weak series: rows of A per GPU
# fallbacks only: the peaks are measured in-run by rla_measure_peak (register-only DMMA / FFMA loops on every SM)
FP64_DMMA_PEAK_TFLOPS = 37.13     # profiles/peaks_r01.jsonl
FP32_FFMA_PEAK_TFLOPS = 72.4
HBM_PEAK_GBS_FALLBACK = 6650.0


def measured_peaks(l):
    """(fp64 DMMA TFLOP/s, fp32 FFMA TFLOP/s, source) measured on the current device, constants if that fails"""
    import ctypes as C
    out = []
    for kind, fb in ((0, FP64_DMMA_PEAK_TFLOPS), (1, FP32_FFMA_PEAK_TFLOPS)):
        v = C.c_double(0.0)
        st = l.rla_measure_peak(kind, C.byref(v))
        out.append(v.value if st == 0 and v.value > 0 else fb)
    return out[0], out[1], "measured in this run by rla_measure_peak (issue-bound DMMA.8x8x4 / FFMA register loops, 1024 threads x 148 SMs)"


def hbm_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return HBM_PEAK_GBS_FALLBACK, "fallback (B200_PROFILING.md)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--n", type=int, default=N_SQUARE, help="development override of the square size")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe line), runs during the timed regions
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def window(self, t0, t1):
        return [ln for (t, ln) in self.lines if t0 <= t <= t1]

    def stop(self):
        if self.proc:
            self.proc.terminate()

    @staticmethod
    def summarise(lines):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s > 0.5 * max(mx)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU restatement (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_dgemm_sample(n: int, rows: int, reps: int = 1):
    """Time `rows` rows of the n x n x n product with the single-threaded port (the reference is
    single-threaded: SURVEY.md 2).  Returns (GFLOP/s, seconds per rep)."""
    import numpy as np
    import oracle
    oracle.build()
    a = oracle.fill_uniform((rows, n), 12)
    b = oracle.fill_uniform((n, n), 2049)
    c = np.empty((rows, n))
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        oracle.gemm(a, b, c=c, fast=True)
        ts.append(time.perf_counter() - t)
    best = min(ts)
    return 2.0 * rows * n * n / best * 1e-9, best


def run_reference(args):
    """The reference's CPU path (C restatement of rulinalg mat_mul + matrixmultiply, one thread: the reference is
    single-threaded) on a bounded sample of OUR arm's workload: N=1 the 8192^3 product, N>1 the C4 product (k = 32768)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import oracle
    oracle.build()
    if args.gpus == 1:
        n = k = args.n
        rows, cols = 256, n
        workload = f"f64 DGEMM {n}x{n}x{n} (&A * &B), reference CPU path"
        sample = f"{rows} of {n} rows of A per step ({rows}x{k}x{cols} product), single thread"
    else:
        n = k = args.n if args.n != N_SQUARE else N_WIDE
        rows, cols = 64, min(n, 4096)
        workload = (f"f64 DGEMM ({M_PER_GPU_WEAK}*{args.gpus})x{k} times {k}x{n} (C4 weak series, &A * &B), reference CPU path")
        sample = f"{rows} rows of A x the first {cols} of {n} columns of B per step ({rows}x{k}x{cols} product, full k), single thread"
    a = oracle.fill_uniform((rows, k), 12)
    b = oracle.fill_uniform((k, cols), 2049)
    c = np.empty((rows, cols))
    for _ in range(max(args.warmup, 1)):
        oracle.gemm(a, b, c=c, fast=True)
    per = []
    for _ in range(args.steps):
        t = time.perf_counter()
        oracle.gemm(a, b, c=c, fast=True)
        per.append(time.perf_counter() - t)
    ms = statistics.mean(per) * 1e3
    gflops = 2.0 * rows * k * cols / (ms * 1e-3) * 1e-9
    line = {
        "impl": "reference", "metric": "dgemm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "C restatement of rulinalg mat_mul + matrixmultiply 0.1.x order (oracle/oracle.c); "
                                 "the Rust reference cannot be built in this image (no rustc)"},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


# ----------------------------------------------------------------------------------------------
# in-run checkers (the verdict on parity is printed in the JSON line; nothing here is timed)
# ----------------------------------------------------------------------------------------------
def gamma(k, u=2.0 ** -53):
    return k * u / (1 - k * u)


def gemm_parity(rla, l, torch, a_loc, b, c_loc, sptr, samples=512, seed=7):
    """C_loc = A_loc * B on U[0,1) data: (i) Freivalds  C_loc x  vs  A_loc (B x)  within 3*gamma_k*|A||B||x| (all data
    non-negative, so |A||B||x| is the product itself), through the library's own HBM-bound gemv kernel -- a different
    kernel from the one under test; (ii) `samples` entries against extended-precision dots (oracle, host) within the
    Higham bound gamma_k * sum|a||b| and the expected random-walk growth 8 sqrt(k) u."""
    import numpy as np
    import oracle
    m, k = a_loc.shape
    n = b.shape[1]
    dev = a_loc.device
    x = torch.empty(n, dtype=torch.float64, device=dev)
    rla.check(l.rla_fill_uniform_f64_dev(x.data_ptr(), 1, n, n, 4000, 0, 0.0, 1.0, sptr))
    y = torch.empty(k, dtype=torch.float64, device=dev)
    z = torch.empty(m, dtype=torch.float64, device=dev)
    w = torch.empty(m, dtype=torch.float64, device=dev)
    rla.check(l.rla_dgemv_dev(k, n, b.data_ptr(), b.stride(0), x.data_ptr(), y.data_ptr(), sptr))
    rla.check(l.rla_dgemv_dev(m, k, a_loc.data_ptr(), a_loc.stride(0), y.data_ptr(), z.data_ptr(), sptr))
    rla.check(l.rla_dgemv_dev(m, n, c_loc.data_ptr(), c_loc.stride(0), x.data_ptr(), w.data_ptr(), sptr))
    torch.cuda.synchronize()
    excess = float(((w - z).abs() - 3 * gamma(k) * z.abs() * 2).max().item())   # x2: z itself carries a gamma_k error
    rng = np.random.default_rng(seed)
    ii = rng.integers(0, m, samples)
    jj = rng.integers(0, n, samples)
    it, jt = torch.as_tensor(ii, device=dev), torch.as_tensor(jj, device=dev)
    a_rows = a_loc.index_select(0, it).cpu().numpy()                       # samples x k
    b_cols = b.index_select(1, jt).t().contiguous().cpu().numpy()          # samples x k
    got = c_loc[it, jt].cpu().numpy()
    idx = np.arange(samples)
    truth, absd = oracle.gemm_truth_samples(a_rows, b_cols.T, idx, idx)
    err = np.abs(got - truth)
    higham_ok = bool(np.all(err <= gamma(k) * absd))
    rel = float(np.max(err / np.abs(truth)))
    ok = bool(excess <= 0.0 and higham_ok and rel < 8 * (k ** 0.5) * 2.0 ** -53)
    return {"ok": ok, "freivalds_excess": excess, "samples": int(samples), "sample_max_rel_err": rel,
            "sample_within_higham_bound": higham_ok, "rel_gate": 8 * (k ** 0.5) * 2.0 ** -53}


def lu_residual_single(rla, l, torch, a0, lu, perm, info, sptr):
    """HPL scaled residual ||A x - b||_inf / (||A||_inf ||x||_inf n eps) of a solve with the factors (b = ones)."""
    n = a0.shape[0]
    dev = a0.device
    x = torch.ones(n, dtype=torch.float64, device=dev)
    rla.check(l.rla_dgetrs_dev(n, lu.data_ptr(), lu.stride(0), perm.data_ptr(), x.data_ptr(), info.data_ptr(), sptr))
    r = torch.empty(n, dtype=torch.float64, device=dev)
    rla.check(l.rla_dgemv_dev(n, n, a0.data_ptr(), a0.stride(0), x.data_ptr(), r.data_ptr(), sptr))
    torch.cuda.synchronize()
    res = float((r - 1.0).abs().max().item())
    anorm = 0.0
    for r0 in range(0, n, 4096):                                   # row sums in slabs (no 8 GiB temporary)
        anorm = max(anorm, float(a0[r0:r0 + 4096].abs().sum(dim=1).max().item()))
    xn = float(x.abs().max().item())
    eps = 2.0 ** -52
    return res / (anorm * xn * n * eps)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rulinalg_b200 as rla
    from rulinalg_b200.sharded import RowPanelGemm, make_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")     # host-side rendezvous while rank 0 drives every GPU (e2e)
    l = rla.lib()
    rla.check(l.rla_init(local_rank))
    dev = torch.device("cuda", local_rank)
    if world == 1:
        n = args.n
        m_local, k = n, n
    else:
        n = k = args.n if args.n != N_SQUARE else N_WIDE
        m_local = M_PER_GPU_WEAK if n == N_WIDE else n
    K, W = args.steps, max(args.warmup, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def min_over_ranks(x: float) -> float:
        return -max_over_ranks(-x)

    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    peak64, peak32, peak_src = measured_peaks(l)

    # ---- synthetic operands, generated in HBM by the library's seeded generator --------------
    a = torch.empty(m_local, k, dtype=torch.float64, device=dev)
    b = torch.empty(k, n, dtype=torch.float64, device=dev)
    c = torch.empty(m_local, n, dtype=torch.float64, device=dev)
    rla.check(l.rla_fill_uniform_f64_dev(a.data_ptr(), m_local, k, k, 12, rank * m_local * k, 0.0, 1.0, sptr))
    if rank == 0:
        rla.check(l.rla_fill_uniform_f64_dev(b.data_ptr(), k, n, n, 2049, 0, 0.0, 1.0, sptr))
    else:
        b.zero_()
    chunk_rows = 2048
    plan = make_plan(world, rank, m_local, k, n, chunk_rows=chunk_rows)
    op = RowPanelGemm(plan, torch.float64)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    # ---- value: device-resident, CUDA events, max over ranks -------------------------------------
    for _ in range(max(W, 3)):
        op.run(a, b, c)
    barrier()
    l.rla_launch_count_reset()
    t_clk0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        op.run(a, b, c)
    e1.record(stream)
    barrier()
    t_clk1 = time.perf_counter()
    launches_dev = int(l.rla_launch_count())
    ms_step = max_over_ranks(e0.elapsed_time(e1) / K)
    gflops = plan.flops_global / (ms_step * 1e-3) * 1e-9

    # ---- parity of what was just timed (per rank; every rank must pass) ---------------------------
    parity = gemm_parity(rla, l, torch, a, b, c, sptr)
    parity_ok = min_over_ranks(1.0 if parity["ok"] else 0.0) > 0.5
    parity["ranks_checked"] = world

    # ---- roofline of the dominant kernel (dgemm_dmma_kernel), timed alone on its launch stream ----
    for _ in range(2):
        rla.check(l.rla_dgemm_dev(m_local, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, sptr))
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(stream)
    reps = max(3, min(K, 10))
    for _ in range(reps):
        rla.check(l.rla_dgemm_dev(m_local, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, sptr))
    r1.record(stream)
    torch.cuda.synchronize()
    kern_ms = r0.elapsed_time(r1) / reps
    achieved_tf = 2.0 * m_local * k * n / (kern_ms * 1e-3) * 1e-12
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "dgemm_ncu_summary.json")
    if os.path.exists(tpath) and n == N_SQUARE and m_local == N_SQUARE:
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch_n8192")
            traffic_src = tj.get("source", "profiles/dgemm_ncu_summary.json (one ncu --set full capture of this kernel at n = 8192)")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "dgemm_dmma_kernel", "achieved": achieved_tf, "peak": peak64,
                "unit": "TFLOP/s", "frac": achieved_tf / peak64, "traffic": traffic, "traffic_source": traffic_src,
                "kernel_ms": kern_ms, "flops_per_launch": 2.0 * m_local * k * n,
                "algorithmic_bytes_per_launch": 8.0 * (m_local * k + k * n + m_local * n),
                "peak_source": "FP64 tensor pipe (DMMA.8x8x4): " + peak_src + "; MEASURED_PEAKS.json holds only HBM and bf16 "
                               "figures, neither bounds an FP64 GEMM (nominal: 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2)",
                "peak_fp32_ffma_tflops": peak32}

    clocks = None
    if sampler:
        time.sleep(0.15)
        clocks = ClockSampler.summarise(sampler.window(t_clk0, t_clk1))

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -----------------------
    if world == 1:
        e2e, launches_e2e = e2e_single(rla, l, torch, np, a, b, m_local, k, n, K, W)
    else:
        a_keep = None
        del a, b, c, op
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        e2e, launches_e2e = None, 0
        if rank == 0:
            e2e, launches_e2e = e2e_multi(rla, l, torch, np, world, m_local, k, n, K, W, dev, sptr)
        dist.barrier(group=host_group)

    line = {
        "metric": "dgemm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": (f"f64 DGEMM &A*&B {m_local}x{k} times {k}x{n} (BASELINE configs[1] point n={n})" if world == 1 else
                                f"f64 DGEMM &A*&B, BASELINE configs[3] (C4) weak series: A ({m_local}*{world})x{k} row-panel sharded "
                                f"({m_local} rows per GPU), k = n = {n}, B ({k * n * 8 / 2**30:.0f} GiB) broadcast from rank 0 over NCCL in "
                                f"{len(plan.k_chunks)} k chunks of {chunk_rows} rows overlapped with the kernel"),
                   "m_global": plan.m_global, "k": k, "n": n, "seeds": {"A": 12, "B": 2049}, "distribution": "U[0,1)",
                   "l2_policy": f"inputs larger than L2 ({(m_local * k + k * n + m_local * n) * 8 / 2**20:.0f} MiB per GPU vs 126 MiB)",
                   "parallelism": f"row-panel x{world}"},
        "parity_ok": bool(parity_ok), "parity": parity,
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": int(sum_over_ranks(launches_dev)),
        "gpu_launches_e2e": int(sum_over_ranks(launches_e2e)),
        "clocks": clocks,
    }

    # ---- cpu_baseline (rank 0, N = 1 only): bounded sample of the same product on one host core ---
    if world == 1 and not args.no_cpu_baseline:
        rows = 2048 if n >= 8192 else min(n, 2048)
        g, secs = cpu_dgemm_sample(n, rows)
        line["cpu_baseline"] = {"value": g, "unit": "GFLOP/s", "cores": 1, "kind": "port", "seconds": secs,
                                "sample": f"first {rows} of {n} rows of A ({rows}x{n}x{n} product), single thread, "
                                          f"host has {os.cpu_count()} cores"}
    elif world == 1:
        line["cpu_baseline"] = None

    # ---- extras: the other BASELINE configs ----------------------------------------------------------
    if world == 1 and not args.no_extras:
        del a, b, c
        torch.cuda.empty_cache()
        line["extras"] = extras(rla, l, torch, np, dev, sptr, peak64, peak32)
    if world > 1 and not args.no_extras:
        ex = {}
        ex.update(strong_gemm_extra(rla, l, torch, dist, world, rank, dev, sptr, peak64, max_over_ranks, min_over_ranks))
        ex.update(dist_lu_extra(rla, l, torch, dist, world, rank, dev, peak64))
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            ex.update(multi_lu_extra(rla, l, torch, np, world, dev, sptr, peak64))
        dist.barrier(group=host_group)
        line["extras"] = ex
        line["parity_ok"] = bool(line["parity_ok"] and all(v.get("parity_ok", True) for v in ex.values() if isinstance(v, dict)))

    if sampler:
        sampler.stop()
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def pinned_like(torch, shape):
    return torch.empty(*shape, dtype=torch.float64).pin_memory()


def e2e_single(rla, l, torch, np, a, b, m, k, n, K, W):
    """rla_dgemm (matrixmultiply::dgemm's signature, mat_mul.rs:57-67) with pinned and with pageable host buffers"""
    a_h, b_h, c_h = pinned_like(torch, (m, k)), pinned_like(torch, (k, n)), pinned_like(torch, (m, n))
    a_h.copy_(a)
    b_h.copy_(b)
    torch.cuda.synchronize()

    def timed(fn):
        for _ in range(max(1 if W > 0 else 0, 1)):
            fn()
        torch.cuda.synchronize()
        l.rla_launch_count_reset()
        t0 = time.perf_counter()
        for _ in range(K):
            fn()
        ms = (time.perf_counter() - t0) * 1e3 / K
        return ms, int(l.rla_launch_count())

    ms, launches = timed(lambda: rla.check(l.rla_dgemm(m, k, n, 1.0, a_h.data_ptr(), k, 1, b_h.data_ptr(), n, 1, 0.0, c_h.data_ptr(), n, 1)))
    chk = float(c_h[0, :8].sum().item())
    # what `&a * &b` really hands over: Vec<T> storage, i.e. pageable memory (numpy arrays here)
    a_p, b_p, c_p = a_h.numpy().copy(), b_h.numpy().copy(), np.empty((m, n))
    ms_p, _ = timed(lambda: rla.check(l.rla_dgemm(m, k, n, 1.0, a_p.ctypes.data, k, 1, b_p.ctypes.data, n, 1, 0.0, c_p.ctypes.data, n, 1)))
    same = bool(np.array_equal(c_p, c_h.numpy()))
    flops = 2.0 * m * k * n
    e2e = {"value": flops / (ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": (m * k + k * n) * 8,
           "d2h_bytes_per_step": m * n * 8, "ms_per_step": ms, "api": "rla_dgemm (host pointers, pinned)", "checksum": chk,
           "pageable": {"value": flops / (ms_p * 1e-3) * 1e-9, "ms_per_step": ms_p, "bit_identical_to_pinned": same,
                        "api": "rla_dgemm (host pointers, pageable numpy buffers through the library's pinned staging ring)"}}
    return e2e, launches


def e2e_multi(rla, l, torch, np, world, m_local, k, n, K, W, dev, sptr):
    """Rank 0 alone: ONE rla_dgemm call on host buffers after rla_set_devices(world) -- the multi-GPU path a Rust
    caller reaches through mat_mul.rs:57-67.  A is (m_local*world) x k, pinned; H2D of A and B, the NVLink fan-out of
    B and the D2H of C are inside the timed region."""
    import oracle
    m = m_local * world
    a_h, b_h, c_h = pinned_like(torch, (m, k)), pinned_like(torch, (k, n)), pinned_like(torch, (m, n))
    tmp = torch.empty(max(m_local, 2048), k, dtype=torch.float64, device=dev)
    for r in range(world):                                   # same synthetic A as the device-resident run
        rla.check(l.rla_fill_uniform_f64_dev(tmp.data_ptr(), m_local, k, k, 12, r * m_local * k, 0.0, 1.0, sptr))
        a_h[r * m_local:(r + 1) * m_local].copy_(tmp[:m_local])
    for k0 in range(0, k, 2048):
        rla.check(l.rla_fill_uniform_f64_dev(tmp.data_ptr(), 2048, n, n, 2049, k0 * n, 0.0, 1.0, sptr))
        b_h[k0:k0 + 2048].copy_(tmp[:2048])
    torch.cuda.synchronize()
    del tmp
    torch.cuda.empty_cache()
    rla.check(l.rla_set_devices(world))

    def step():
        rla.check(l.rla_dgemm(m, k, n, 1.0, a_h.data_ptr(), k, 1, b_h.data_ptr(), n, 1, 0.0, c_h.data_ptr(), n, 1))
    step()
    l.rla_launch_count_reset()
    Ke = max(2, min(K, 5))
    t0 = time.perf_counter()
    for _ in range(Ke):
        step()
    ms = (time.perf_counter() - t0) * 1e3 / Ke
    launches = int(l.rla_launch_count())
    rla.check(l.rla_set_devices(1))
    # sampled entries of the host result against extended-precision dots
    rng = np.random.default_rng(11)
    S = 128
    ii, jj = rng.integers(0, m, S), rng.integers(0, n, S)
    an, bn, cn = a_h.numpy(), b_h.numpy(), c_h.numpy()
    truth, absd = oracle.gemm_truth_samples(an, bn, ii, jj)
    err = np.abs(cn[ii, jj] - truth)
    ok = bool(np.all(err <= gamma(k) * absd) and np.max(err / np.abs(truth)) < 8 * (k ** 0.5) * 2.0 ** -53)
    flops = 2.0 * m * k * n
    e2e = {"value": flops / (ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": (m * k + k * n) * 8,
           "d2h_bytes_per_step": m * n * 8, "ms_per_step": ms, "steps": Ke,
           "api": f"rla_set_devices({world}) + ONE rla_dgemm call on pinned host buffers from rank 0's process "
                  f"(row panels over {world} GPUs, B chunks uploaded round-robin and fanned out over NVLink peer memory)",
           "parity_ok": ok, "checksum": float(c_h[0, :8].sum().item())}
    return e2e, launches


def extras(rla, l, torch, np, dev, sptr, peak64, peak32):
    """SGEMM / sweep / LU / solve figures (device-resident, CUDA events, best of a few) + host-API figures."""
    out = {}
    stream = torch.cuda.current_stream()
    hbm, hbm_src = hbm_peak_gbs()

    def timed(fn, reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    for name, dt, fn, peak in (("dgemm", torch.float64, l.rla_dgemm_dev, peak64),
                               ("sgemm", torch.float32, l.rla_sgemm_dev, peak32)):
        for (m, k, n) in ((512, 512, 512), (1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 8192),
                          (16384, 16384, 16384), (65536, 256, 256)):
            if name == "dgemm" and m == 8192:
                continue
            a = torch.rand(m, k, dtype=dt, device=dev); b = torch.rand(k, n, dtype=dt, device=dev)
            c = torch.empty(m, n, dtype=dt, device=dev)
            ms = timed(lambda: rla.check(fn(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, sptr)), 3)
            tf = 2.0 * m * k * n / ms * 1e-9
            out[f"{name}_{m}x{k}x{n}"] = {"ms": ms, "tflops": tf, "frac_of_peak": tf / peak}
            del a, b, c
    for n in (1024, 2048, 4096, 8192, 32768):
        a0 = torch.empty(n, n, dtype=torch.float64, device=dev)
        rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, sptr))
        a = torch.empty_like(a0)
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        best = 1e30
        for _ in range(3 if n <= 8192 else 2):
            a.copy_(a0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), sptr))
            e1.record(stream); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf = 2.0 / 3.0 * n ** 3 / best * 1e-9
        out[f"dgetrf_{n}"] = {"ms": best, "tflops": tf, "frac_of_peak": tf / peak64, "info": int(info.item())}
        resid = lu_residual_single(rla, l, torch, a0, a, perm, info, sptr)
        out[f"dgetrf_{n}"]["hpl_scaled_residual"] = resid
        out[f"dgetrf_{n}"]["parity_ok"] = bool(int(info.item()) == 0 and resid <= 16.0)
        bvec = torch.ones(n, dtype=torch.float64, device=dev)
        b0 = bvec.clone()

        def solve():
            bvec.copy_(b0)
            rla.check(l.rla_dgetrs_dev(n, a.data_ptr(), n, perm.data_ptr(), bvec.data_ptr(), info.data_ptr(), sptr))
        ms = timed(solve, 3)
        gbs = 8.0 * n * n / ms * 1e-6
        out[f"dgetrs_{n}"] = {"ms": ms, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm, "hbm_peak_source": hbm_src}
        del a0, a
        torch.cuda.empty_cache()
    # Cholesky: flops = n^3 / 3; SPD input = G G^T / k + 4 I built with torch (checker side, untimed)
    for n in (4096, 16384):
        g = torch.rand(n, 2048, dtype=torch.float64, device=dev)
        a0 = g @ g.T / 2048 + torch.eye(n, dtype=torch.float64, device=dev) * 4
        del g
        a = torch.empty_like(a0)
        ws = torch.empty(int(l.rla_potrf_workspace_bytes(n, 8)), dtype=torch.uint8, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        best = 1e30
        for _ in range(3):
            a.copy_(a0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rla.check(l.rla_dpotrf_dev(n, a.data_ptr(), n, ws.data_ptr(), info.data_ptr(), sptr))
            e1.record(stream); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf = n ** 3 / 3.0 / best * 1e-9
        out[f"dpotrf_{n}"] = {"ms": best, "tflops": tf, "frac_of_peak": tf / peak64, "info": int(info.item())}
        del a0, a, ws
        torch.cuda.empty_cache()
    out["host_api_lu"] = host_lu_extra(rla, l, torch, np)
    out["reference_bench_shapes"] = reference_shapes_extra(rla, l, np)
    return out


def host_lu_extra(rla, l, torch, np):
    """BASELINE configs[2] (C3) end to end: PartialPivLu::decompose + solve through rla_dgetrf / rla_dgetrs with HOST
    buffers (lu.rs:163-195 consumes a Vec, :231-244), copies inside the timed region; pinned and pageable."""
    out = {}
    import oracle
    for n in (4096, 8192):
        a0 = oracle.fill_uniform((n, n), 12)
        res = {}
        for kind in ("pinned", "pageable"):
            if kind == "pinned":
                buf = torch.empty(n, n, dtype=torch.float64).pin_memory()
                lu = buf.numpy()
            else:
                lu = np.empty((n, n))
            perm = np.empty(n, dtype=np.uint64)
            best_f, best_s = 1e30, 1e30
            for _ in range(3):
                lu[...] = a0
                t0 = time.perf_counter()
                st = rla.check(l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data))
                t1 = time.perf_counter()
                x = np.ones(n)
                st2 = rla.check(l.rla_dgetrs(n, lu.ctypes.data, perm.ctypes.data, x.ctypes.data))
                t2 = time.perf_counter()
                best_f, best_s = min(best_f, (t1 - t0) * 1e3), min(best_s, (t2 - t1) * 1e3)
            r = a0 @ x - 1.0
            eps = 2.0 ** -52
            resid = float(np.max(np.abs(r)) / (np.max(np.sum(np.abs(a0), axis=1)) * np.max(np.abs(x)) * n * eps))
            # the same solve with the factors held (rla_operand_hold: resident in HBM after the first call)
            rla.check(l.rla_operand_hold(lu.ctypes.data, lu.nbytes))
            xh = np.ones(n)
            rla.check(l.rla_dgetrs(n, lu.ctypes.data, perm.ctypes.data, xh.ctypes.data))
            best_h = 1e30
            for _ in range(3):
                xh = np.ones(n)
                t0 = time.perf_counter()
                rla.check(l.rla_dgetrs(n, lu.ctypes.data, perm.ctypes.data, xh.ctypes.data))
                best_h = min(best_h, (time.perf_counter() - t0) * 1e3)
            rla.check(l.rla_operand_release(lu.ctypes.data))
            res[kind] = {"decompose_ms": best_f, "solve_ms": best_s, "solve_factors_held_ms": best_h,
                         "held_bit_identical": bool(np.array_equal(x, xh)), "status": [int(st), int(st2)], "hpl_scaled_residual": resid,
                         "parity_ok": bool(st == 0 and st2 == 0 and resid <= 16.0)}
        res["pcie_floor_ms_one_way"] = n * n * 8 / 55e9 * 1e3
        res["note"] = ("decompose = H2D + factor + D2H of the factors (block rows downloaded as they become final); solve re-uploads the "
                       "factors (rla_dgetrs signature) unless their host range is held (rla_operand_hold; Rust: a guard inside PartialPivLu)")
        out[f"n{n}"] = res
    return out


def reference_shapes_extra(rla, l, np):
    """The only shapes the reference itself measures (benches/linalg/matrix.rs:40-65, lu.rs:52-139, triangular.rs:6-58),
    through the host API, beside the single-thread CPU port: these are launch / PCIe-latency bound on a GPU."""
    import oracle
    out = {}

    def best_us(fn, reps):
        fn()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best * 1e6

    for (m, k, n) in ((10, 10, 10), (128, 100, 128), (128, 1000, 128)):
        a = oracle.fill_uniform((m, k), 12, np.float32)
        b = oracle.fill_uniform((k, n), 2049, np.float32)
        c = np.empty((m, n), np.float32)
        gpu = best_us(lambda: rla.check(l.rla_sgemm(m, k, n, 1.0, a.ctypes.data, k, 1, b.ctypes.data, n, 1, 0.0, c.ctypes.data, n, 1)), 50)
        cpu = best_us(lambda: oracle.gemm(a, b, fast=True), 50)
        out[f"sgemm_{m}x{k}x{n}"] = {"gpu_api_us": gpu, "cpu_port_us": cpu}
    for n in (10, 100):
        a0 = oracle.fill_uniform((n, n), 12) + n * np.eye(n)
        perm = np.empty(n, dtype=np.uint64)
        lu = a0.copy()

        def dec():
            lu[...] = a0
            rla.check(l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data))
        gpu = best_us(dec, 30)
        cpu = best_us(lambda: oracle.lu_decompose(a0, fast=True), 30)
        x = np.ones(n)
        gpu_s = best_us(lambda: rla.check(l.rla_dgetrs(n, lu.ctypes.data, perm.ctypes.data, x.ctypes.data)), 30)
        rl, rp = oracle.lu_decompose(a0, fast=True)
        cpu_s = best_us(lambda: oracle.lu_solve(rl, rp, np.ones(n), fast=True), 30)
        out[f"lu_decompose_{n}"] = {"gpu_api_us": gpu, "cpu_port_us": cpu}
        out[f"lu_solve_{n}"] = {"gpu_api_us": gpu_s, "cpu_port_us": cpu_s}
    for n in (100, 1000, 10000):
        eye = np.eye(n)
        x = np.ones(n)
        for lower, nm in ((0, "u"), (1, "l")):
            gpu = best_us(lambda: rla.check(l.rla_dtrsv(lower, n, eye.ctypes.data, n, x.ctypes.data)), 5 if n >= 10000 else 20)
            f = oracle.forward_substitution if lower else oracle.back_substitution
            cpu = best_us(lambda: f(eye, np.ones(n)), 3 if n >= 10000 else 10)
            out[f"solve_{nm}_triangular_{n}"] = {"gpu_api_us": gpu, "cpu_port_us": cpu}
            if n >= 1000:                      # the triangle held in HBM (rla_operand_hold): the call is no longer its upload
                rla.check(l.rla_operand_hold(eye.ctypes.data, eye.nbytes))
                out[f"solve_{nm}_triangular_{n}"]["gpu_api_held_us"] = best_us(
                    lambda: rla.check(l.rla_dtrsv(lower, n, eye.ctypes.data, n, x.ctypes.data)), 10)
                rla.check(l.rla_operand_release(eye.ctypes.data))
    out["note"] = ("best-of wall time per call incl. host<->device copies; cpu_port = oracle (C restatement, one thread). "
                   "Crossover: see BASELINE.md")
    return out


def strong_gemm_extra(rla, l, torch, dist, world, rank, dev, sptr, peak64, max_over_ranks, min_over_ranks, n=N_WIDE):
    """C4 strong point: f64 32768^3, A and C in row panels of n/N rows, B broadcast from rank 0 in 16 chunks."""
    from rulinalg_b200.sharded import RowPanelGemm, make_plan
    m_local = n // world
    a = torch.empty(m_local, n, dtype=torch.float64, device=dev)
    b = torch.empty(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(m_local, n, dtype=torch.float64, device=dev)
    rla.check(l.rla_fill_uniform_f64_dev(a.data_ptr(), m_local, n, n, 12, rank * m_local * n, 0.0, 1.0, sptr))
    if rank == 0:
        rla.check(l.rla_fill_uniform_f64_dev(b.data_ptr(), n, n, n, 2049, 0, 0.0, 1.0, sptr))
    else:
        b.zero_()
    plan = make_plan(world, rank, m_local, n, n, chunk_rows=2048)
    op = RowPanelGemm(plan, torch.float64)
    stream = torch.cuda.current_stream()
    op.run(a, b, c)
    torch.cuda.synchronize()
    dist.barrier()
    reps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        op.run(a, b, c)
    e1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / reps)
    par = gemm_parity(rla, l, torch, a, b, c, sptr, samples=256, seed=5)
    ok = min_over_ranks(1.0 if par["ok"] else 0.0) > 0.5
    tf = 2.0 * n ** 3 / ms * 1e-9
    del a, b, c
    torch.cuda.empty_cache()
    return {f"dgemm_strong_{n}": {"ms": ms, "tflops": tf, "frac_of_peak": tf / (peak64 * world), "gpus": world,
                                  "workload": f"f64 {n}^3, row panels of {m_local} rows, B (8 GiB) broadcast over NCCL in {len(plan.k_chunks)} chunks",
                                  "parity_ok": bool(ok), "parity": par}}


def dist_lu_extra(rla, l, torch, dist, world, rank, dev, peak64, n=N_WIDE):
    """C5: PartialPivLu f64 n x n, column blocks of 256 dealt round-robin, NCCL panel broadcast, look-ahead; verified by
    reconstructing sampled rows of P A = L U across the ranks."""
    from rulinalg_b200.sharded_lu import BlockCyclicLayout, BlockCyclicLu
    lay = BlockCyclicLayout(n, world, rank)
    ncl = lay.ncols_local()
    a0 = torch.empty(n, ncl, dtype=torch.float64, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, ncl, ncl, 12 + rank, 0, 0.0, 1.0, sp))
    a = torch.empty_like(a0)
    lu = BlockCyclicLu(lay, lookahead=True)
    best, info, perm = 1e30, None, None
    for _ in range(3):
        a.copy_(a0)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        perm, info = lu.decompose(a)
        e1.record()
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = min(best, float(ms.item()))
    tf = 2.0 / 3.0 * n ** 3 / best * 1e-9
    # ---- sampled-row reconstruction: rows i of L U must equal rows perm^-1(i) of A ----
    S = 64
    gen = torch.Generator().manual_seed(3)
    fin = torch.sort(torch.cat([torch.randint(0, n, (S - 8,), generator=gen), torch.arange(n - 8, n)])).values.to(dev)
    orig_of_final = torch.empty(n, dtype=torch.int64, device=dev)
    orig_of_final[perm] = torch.arange(n, device=dev)
    # L rows: every rank contributes its local columns; reassemble in global column order
    mine = a.index_select(0, fin)                                            # S x ncl
    pieces = [torch.empty(S, lay.ncols_local(r), dtype=torch.float64, device=dev) for r in range(world)]
    dist.all_gather(pieces, mine)
    lrows = torch.zeros(S, n, dtype=torch.float64, device=dev)
    for r in range(world):
        lc = 0
        for (c0, w) in lay.global_cols(r):
            lrows[:, c0:c0 + w] = pieces[r][:, lc:lc + w]
            lc += w
    cols = torch.arange(n, device=dev)
    lrows = torch.where(cols[None, :] < fin[:, None], lrows, torch.zeros((), dtype=torch.float64, device=dev))
    lrows[torch.arange(S, device=dev), fin] = 1.0
    # local U: keep a[r, lc] where r <= global column of lc
    gcol = torch.cat([torch.arange(c0, c0 + w) for (c0, w) in lay.global_cols()]).to(dev)
    umax = 0.0
    for r0 in range(0, n, 4096):                                             # mask in slabs, in place on `a`
        rows = torch.arange(r0, min(n, r0 + 4096), device=dev)
        blk = a[r0:r0 + 4096]
        blk.mul_((rows[:, None] <= gcol[None, :]).to(torch.float64))
        umax = max(umax, float(blk.abs().max().item()))
    rec = torch.empty(S, ncl, dtype=torch.float64, device=dev)
    rla.check(l.rla_dgemm_dev(S, n, ncl, 1.0, lrows.data_ptr(), n, a.data_ptr(), a.stride(0), 0.0, rec.data_ptr(), ncl, sp))
    torch.cuda.synchronize()
    want = a0.index_select(0, orig_of_final[fin])
    amax = float(a0.abs().max().item())
    t = torch.tensor([float((rec - want).abs().max().item()), umax, amax], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err, umax, amax = (float(v) for v in t.tolist())
    rho = umax / amax
    tol = 8 * n * 2.0 ** -53 * rho * amax
    ok = bool(int(info.item()) == 0 and err <= tol)
    return {f"dist_dgetrf_{n}": {"ms": best, "tflops": tf, "frac_of_peak": tf / (peak64 * world),
                                  "gpus": world, "layout": "1D block-cyclic columns, block 256, NCCL panel broadcast, look-ahead 1",
                                  "info": int(info.item()), "parity_ok": ok,
                                  "parity": {"check": f"{S} sampled rows of L U vs the same rows of P A (abs, tol = 8 n u rho max|A|)",
                                             "max_abs_err": err, "tol": tol, "growth_rho": rho}}}


def multi_lu_extra(rla, l, torch, np, world, dev, sptr, peak64, n=N_WIDE):
    """C5 through the drop-in boundary: ONE rla_dgetrf call on a pinned host matrix after rla_set_devices(world)
    (rank 0's process drives every GPU; H2D, peer-memory panel fan-out and D2H inside the timed region)."""
    buf = torch.empty(n, n, dtype=torch.float64).pin_memory()
    tmp = torch.empty(2048, n, dtype=torch.float64, device=dev)
    lu = buf.numpy()
    perm = np.empty(n, dtype=np.uint64)

    def refill():
        for r0 in range(0, n, 2048):
            rla.check(l.rla_fill_uniform_f64_dev(tmp.data_ptr(), 2048, n, n, 12, r0 * n, 0.0, 1.0, sptr))
            buf[r0:r0 + 2048].copy_(tmp)
        torch.cuda.synchronize()
    rla.check(l.rla_set_devices(world))
    best, st = 1e30, -99
    for _ in range(2):
        refill()
        t0 = time.perf_counter()
        st = rla.check(l.rla_dgetrf(n, lu.ctypes.data, perm.ctypes.data))
        best = min(best, (time.perf_counter() - t0) * 1e3)
    rla.check(l.rla_set_devices(1))
    # residual of a solve with the factors: device-resident check on GPU 0 (A regenerated from its seed)
    a0 = torch.empty(n, n, dtype=torch.float64, device=dev)
    rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, sptr))
    lud = torch.empty(n, n, dtype=torch.float64, device=dev)
    lud.copy_(buf)
    permd = torch.as_tensor(perm.astype(np.int64)).to(dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    resid = lu_residual_single(rla, l, torch, a0, lud, permd, info, sptr)
    tf = 2.0 / 3.0 * n ** 3 / best * 1e-9
    del a0, lud, tmp
    torch.cuda.empty_cache()
    return {f"capi_multi_dgetrf_{n}": {"ms": best, "tflops": tf, "frac_of_peak": tf / (peak64 * world), "gpus": world, "status": int(st),
                                        "api": f"rla_set_devices({world}) + ONE rla_dgetrf call on a pinned host matrix (e2e incl. 8 GiB H2D + D2H)",
                                        "hpl_scaled_residual": resid, "parity_ok": bool(st == 0 and resid <= 16.0)}}


class _StdoutGuard:
    """Everything any library writes to fd 1 during the run (e.g. NCCL's version banner) goes to stderr;
    only the final JSON line is written to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        self.real = os.fdopen(self.saved, "w")
        return self

    def emit(self, text):
        self.real.write(text + "\n")
        self.real.flush()

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        return False


_GUARD = None


def emit_line(line):
    text = json.dumps(line)
    if _GUARD is not None:
        _GUARD.emit(text)
    else:
        print(text, flush=True)


def main():
    global _GUARD
    args = parse()
    with _StdoutGuard() as g:
        _GUARD = g
        if args.impl == "reference":
            return run_reference(args)
        return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
