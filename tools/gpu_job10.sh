#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_lu.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu2.log
for la in 0 1; do
for n in 8192 16384 32768; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/lu_dist_bench.py $n $la 2>&1 | grep -E 'dist_dgetrf|Error|error' | tail -3
done; done
timeout 300 python tools/lu_dist_bench.py 16384 2>&1 | grep -E 'dist_dgetrf|Error|error' | tail -3
timeout 300 python tools/perf_probe.py sgemm 2>&1 | grep -E '"m": (8192|16384|4096), "k": (8192|16384|4096)'
