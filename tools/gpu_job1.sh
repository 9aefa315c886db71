#!/bin/bash
# first GPU job: pipe peaks + library ceilings, smoke, gpu tests, perf probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/gpu_info.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_job1.csv &
SMI=$!
timeout 600 ./tools/peaks 1 > gpurun_out/peaks.jsonl 2> gpurun_out/peaks.err
echo "peaks rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/perf_probe.py > gpurun_out/perf_probe.jsonl 2> gpurun_out/perf_probe.err
echo "probe rc=$?"; tail -5 gpurun_out/perf_probe.err
kill $SMI
cat gpurun_out/peaks.jsonl | head -60
cat gpurun_out/perf_probe.jsonl
