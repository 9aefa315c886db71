#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/perf_probe.py sgemm 2>&1 | grep -E '"m": (8192|16384|4096|65536), "k": (8192|16384|4096|256)'
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu bench rc=$?"
