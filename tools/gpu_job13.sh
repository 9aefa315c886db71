#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/perf_probe.py pcie 2>&1 | grep pcie
timeout 300 python tools/perf_probe.py sgemm 2>&1 | grep -E '"m": (1024|8192|16384|4096|65536), "k": (1024|8192|16384|4096|256)'
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_ffma -s 1 -c 1 -o gpurun_out/prof_sgemm3 python tools/gemm_once.py s 8192 > gpurun_out/ncu_sgemm.log 2>&1
