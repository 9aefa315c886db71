#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/perf_probe.py lu > gpurun_out/perf_probe_lu.jsonl 2> gpurun_out/perf_probe.err
echo "probe rc=$?"; tail -5 gpurun_out/perf_probe.err
cat gpurun_out/perf_probe_lu.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_lu8192.csv python tools/lu_once.py 8192 > gpurun_out/ncu_lu.log 2>&1
echo "ncu lu rc=$?"
