#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python tools/lu_trace.py 4096 | tail -8
timeout 600 python tools/perf_probe.py lu > gpurun_out/perf_probe_lu.jsonl 2>&1; grep -E '"dgetrf"|dgetrs|dgetri' gpurun_out/perf_probe_lu.jsonl
