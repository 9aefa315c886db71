#!/bin/bash
# 2-GPU box: N=2 bench (NCCL row-panel sharding), then N=1 for the same box
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 rc=$?"; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err
echo "n1 rc=$?"; tail -3 gpurun_out/bench_n1_short.err; cat gpurun_out/bench_n1_short.json
