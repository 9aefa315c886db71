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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.n
    rows = 256
    import numpy as np
    import oracle
    oracle.build()
    a = oracle.fill_uniform((rows, n), 12)
    b = oracle.fill_uniform((n, n), 2049)
    c = np.empty((rows, n))
    for _ in range(max(args.warmup, 1)):
        oracle.gemm(a, b, c=c, fast=True)
    per = []
    for _ in range(args.steps):
        t = time.perf_counter()
        oracle.gemm(a, b, c=c, fast=True)
        per.append(time.perf_counter() - t)
    ms = statistics.mean(per) * 1e3
    gflops = 2.0 * rows * n * n / (ms * 1e-3) * 1e-9
    sample = f"{rows} of {n} rows of A per step ({rows}x{n}x{n} product), single thread"
    line = {
        "impl": "reference", "metric": "dgemm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"f64 DGEMM {n}x{n}x{n} (&A * &B), reference CPU path", "sample": sample},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "C restatement of rulinalg mat_mul + matrixmultiply 0.1.x order (oracle/oracle.c); "
                                 "the Rust reference cannot be built in this image (no rustc)"},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    l = rla.lib()
    rla.check(l.rla_init(local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.n
    m_local, k = n, n
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

    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    # ---- synthetic operands, generated in HBM by the library's seeded generator --------------
    a = torch.empty(m_local, k, dtype=torch.float64, device=dev)
    b = torch.empty(k, n, dtype=torch.float64, device=dev)
    c = torch.empty(m_local, n, dtype=torch.float64, device=dev)
    rla.check(l.rla_fill_uniform_f64_dev(a.data_ptr(), m_local, k, k, 12, rank * m_local * k, 0.0, 1.0, sptr))
    if rank == 0:
        rla.check(l.rla_fill_uniform_f64_dev(b.data_ptr(), k, n, n, 2049, 0, 0.0, 1.0, sptr))
    else:
        b.zero_()
    plan = make_plan(world, rank, m_local, k, n, chunk_rows=2048)
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
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dgemm_ncu_summary.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch_n8192")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "dgemm_dmma_kernel<true>", "achieved": achieved_tf, "peak": FP64_DMMA_PEAK_TFLOPS,
                "unit": "TFLOP/s", "frac": achieved_tf / FP64_DMMA_PEAK_TFLOPS, "traffic": traffic,
                "kernel_ms": kern_ms, "flops_per_launch": 2.0 * m_local * k * n,
                "peak_source": "FP64 tensor pipe (DMMA.8x8x4) issue-bound peak measured on this pool with tools/peaks.cu "
                               "(profiles/peaks_r01.jsonl); MEASURED_PEAKS.json holds only HBM and bf16 figures, neither "
                               "bounds an FP64 GEMM (nominal: 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2)"}

    # ---- e2e: host buffers through the public API, copies inside the timed region ------------------
    a_h = torch.empty(m_local, k, dtype=torch.float64).pin_memory()
    c_h = torch.empty(m_local, n, dtype=torch.float64).pin_memory()
    b_h = torch.empty(k, n, dtype=torch.float64).pin_memory() if rank == 0 else None
    a_h.copy_(a)
    if rank == 0:
        b_h.copy_(b)
    torch.cuda.synchronize()

    def e2e_step():
        if world == 1:
            # the reference-facing call: matrixmultiply::dgemm's signature (mat_mul.rs:57-67)
            rla.check(l.rla_dgemm(m_local, k, n, 1.0, a_h.data_ptr(), k, 1, b_h.data_ptr(), n, 1, 0.0, c_h.data_ptr(), n, 1))
        else:
            if rank == 0:
                b.copy_(b_h, non_blocking=True)
            a.copy_(a_h, non_blocking=True)
            op.run(a, b, c)
            c_h.copy_(c, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    e2e_warm = 1 if W > 0 else 0
    for _ in range(max(e2e_warm, 1)):
        e2e_step()
    barrier()
    l.rla_launch_count_reset()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / K)
    launches_e2e = int(l.rla_launch_count())
    e2e_gflops = plan.flops_global / (e2e_ms * 1e-3) * 1e-9
    h2d = m_local * k * 8 * world + k * n * 8
    d2h = m_local * n * 8 * world
    # a cheap checksum of the e2e result so the D2H read is real
    chk = float(c_h[0, :8].sum().item())

    clocks = None
    if sampler:
        time.sleep(0.15)
        clocks = ClockSampler.summarise(sampler.window(t_clk0, time.perf_counter()))

    line = {
        "metric": "dgemm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"f64 DGEMM &A*&B, per-GPU row panel {m_local}x{k} times {k}x{n} (BASELINE configs[1] point n={n}; "
                               f"N>1: A ({m_local}*N)x{k} row-panel sharded, B broadcast from rank 0 over NCCL in 2048-row k chunks)",
                   "m_global": plan.m_global, "k": k, "n": n, "seeds": {"A": 12, "B": 2049}, "distribution": "U[0,1)",
                   "l2_policy": f"inputs larger than L2 ({3 * n * n * 8 / 2**20:.0f} MiB per GPU vs 126 MiB)",
                   "parallelism": f"row-panel x{world}"},
        "roofline": roofline,
        "e2e": {"value": e2e_gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "api": "rla_dgemm (host pointers, pinned)" if world == 1 else "sharded.RowPanelGemm + pinned H2D/D2H",
                "checksum": chk},
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

    # ---- extras: the other BASELINE configs, N = 1 only ---------------------------------------------
    if world == 1 and not args.no_extras:
        del a_h, b_h, c_h
        line["extras"] = extras(rla, l, torch, dev, sptr)

    # ---- N > 1 extra: BASELINE config C5, f64 LU n = 32768 1D block-cyclic over the N GPUs ---------------
    if world > 1 and not args.no_extras:
        del a_h, c_h, a, b, c
        torch.cuda.empty_cache()
        line["extras"] = dist_lu_extra(rla, l, torch, dist, world, rank, dev)

    if sampler:
        sampler.stop()
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def extras(rla, l, torch, dev, sptr):
    """SGEMM / sweep / LU / solve figures (device-resident, CUDA events, best of a few)."""
    out = {}
    stream = torch.cuda.current_stream()

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

    for name, dt, fn, peak in (("dgemm", torch.float64, l.rla_dgemm_dev, FP64_DMMA_PEAK_TFLOPS),
                               ("sgemm", torch.float32, l.rla_sgemm_dev, FP32_FFMA_PEAK_TFLOPS)):
        for (m, k, n) in ((1024, 1024, 1024), (4096, 4096, 4096), (8192, 8192, 8192), (16384, 16384, 16384), (65536, 256, 256)):
            if name == "dgemm" and m == 8192:
                continue
            a = torch.rand(m, k, dtype=dt, device=dev); b = torch.rand(k, n, dtype=dt, device=dev)
            c = torch.empty(m, n, dtype=dt, device=dev)
            ms = timed(lambda: rla.check(fn(m, k, n, 1.0, a.data_ptr(), k, b.data_ptr(), n, 0.0, c.data_ptr(), n, sptr)), 3)
            tf = 2.0 * m * k * n / ms * 1e-9
            out[f"{name}_{m}x{k}x{n}"] = {"ms": ms, "tflops": tf, "frac_of_peak": tf / peak}
            del a, b, c
    for n in (4096, 32768):
        a0 = torch.empty(n, n, dtype=torch.float64, device=dev)
        rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, n, n, 12, 0, 0.0, 1.0, sptr))
        a = torch.empty_like(a0)
        perm = torch.empty(n, dtype=torch.int64, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        best = 1e30
        for _ in range(3 if n <= 4096 else 2):
            a.copy_(a0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            rla.check(l.rla_dgetrf_dev(n, a.data_ptr(), n, perm.data_ptr(), info.data_ptr(), sptr))
            e1.record(stream); e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf = 2.0 / 3.0 * n ** 3 / best * 1e-9
        out[f"dgetrf_{n}"] = {"ms": best, "tflops": tf, "frac_of_peak": tf / FP64_DMMA_PEAK_TFLOPS, "info": int(info.item())}
        bvec = torch.ones(n, dtype=torch.float64, device=dev)
        b0 = bvec.clone()

        def solve():
            bvec.copy_(b0)
            rla.check(l.rla_dgetrs_dev(n, a.data_ptr(), n, perm.data_ptr(), bvec.data_ptr(), info.data_ptr(), sptr))
        ms = timed(solve, 3)
        out[f"dgetrs_{n}"] = {"ms": ms, "gbs": 8.0 * n * n / ms * 1e-6}
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
        out[f"dpotrf_{n}"] = {"ms": best, "tflops": tf, "frac_of_peak": tf / FP64_DMMA_PEAK_TFLOPS, "info": int(info.item())}
        del a0, a, ws
        torch.cuda.empty_cache()
    return out


def dist_lu_extra(rla, l, torch, dist, world, rank, dev, n=32768):
    """PartialPivLu f64 n x n, column blocks of 256 dealt round-robin, NCCL panel broadcast, look-ahead."""
    from rulinalg_b200.sharded_lu import BlockCyclicLayout, BlockCyclicLu
    lay = BlockCyclicLayout(n, world, rank)
    ncl = lay.ncols_local()
    a0 = torch.empty(n, ncl, dtype=torch.float64, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    rla.check(l.rla_fill_uniform_f64_dev(a0.data_ptr(), n, ncl, ncl, 12 + rank, 0, 0.0, 1.0, sp))
    a = torch.empty_like(a0)
    lu = BlockCyclicLu(lay, lookahead=True)
    best, info = 1e30, None
    for _ in range(3):
        a.copy_(a0)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, info = lu.decompose(a)
        e1.record()
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = min(best, float(ms.item()))
    tf = 2.0 / 3.0 * n ** 3 / best * 1e-9
    return {f"dist_dgetrf_{n}": {"ms": best, "tflops": tf, "frac_of_peak": tf / (FP64_DMMA_PEAK_TFLOPS * world),
                                  "gpus": world, "layout": "1D block-cyclic columns, block 256, NCCL panel broadcast, look-ahead 1",
                                  "info": int(info.item())}}


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
