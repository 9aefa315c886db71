#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_job2.csv &
SMI=$!
: > gpurun_out/peaks.jsonl
for g in pipes barriers cublas cusolver; do
  timeout 240 ./tools/peaks 1 $g >> gpurun_out/peaks.jsonl 2>> gpurun_out/peaks.err
  echo "peaks $g rc=$?"
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tools/perf_probe.py > gpurun_out/perf_probe.jsonl 2> gpurun_out/perf_probe.err
echo "probe rc=$?"; tail -5 gpurun_out/perf_probe.err
kill $SMI
cat gpurun_out/peaks.jsonl
cat gpurun_out/perf_probe.jsonl
