#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 -k "lu or LU or solve or inverse or block_cyclic" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/perf_probe.py lu 2>&1 | grep -E '"dgetrf"|"sgetrf"'
