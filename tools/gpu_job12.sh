#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "2d or strides or oracle_positive" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_n1_2d.json 2> gpurun_out/bench_n1_2d.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1_2d.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_2d.json')); print('value',d['value'],'e2e',d['e2e'])"
