"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into a small JSON + markdown for profiles/.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/name [flops_per_launch] [peak_tflops]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed": "dram_read_pct_peak",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "fma_pipe_active_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid_size",
    "launch__block_size": "block_size",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
}
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    flops = float(sys.argv[3]) if len(sys.argv) > 3 else None
    peak = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    summ = []
    for vals in rows[2:]:
        d = {}
        rec = dict(zip(hdr, zip(units, vals)))
        d["kernel"] = rec.get("Kernel Name", ("", ""))[1]
        for k, name in KEYS.items():
            if k in rec and rec[k][1] != "":
                u, v = rec[k]
                try:
                    fv = float(v.replace(",", ""))
                except ValueError:
                    continue
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u)
                d[name] = fv * scale if scale is not None and name in ("dram_read", "dram_write", "l2_bytes", "duration") else fv
        stalls = {}
        for h in hdr:
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len(STALLS):-len("_per_issue_active.ratio")]] = float(rec[h][1])
                except ValueError:
                    pass
        d["stall_cycles_per_issue"] = {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes_per_launch"] = d["dram_read"] + d["dram_write"]
        if flops and "duration" in d:
            d["tflops_under_profiler"] = flops / d["duration"] * 1e-12
            if peak:
                d["frac_of_peak_under_profiler"] = d["tflops_under_profiler"] / peak
        summ.append(d)
    json.dump({"source": rep, "command": "ncu --set full --clock-control none --import-source on", "kernels": summ},
              open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary: {rep}\n\n")
        for d in summ:
            f.write(f"## {d['kernel'][:120]}\n\n| metric | value |\n|---|---|\n")
            for k, v in d.items():
                if k in ("kernel",):
                    continue
                f.write(f"| {k} | {v} |\n")
            f.write("\n")
    print(json.dumps(summ[0] if summ else {}, indent=1)[:1500])


if __name__ == "__main__":
    main()
