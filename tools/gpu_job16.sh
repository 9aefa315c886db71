#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "n$N rc=$?"; grep -v -E "^\*|OMP_NUM|^$|NCCL version" gpurun_out/bench_n$N.err | tail -5; cat gpurun_out/bench_n$N.json | cut -c1-4000
