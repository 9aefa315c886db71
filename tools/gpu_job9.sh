#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_ffma -s 1 -c 1 -o gpurun_out/prof_sgemm2 python tools/gemm_once.py s 8192 > gpurun_out/ncu_sgemm.log 2>&1
echo "ncu sgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 1 -c 1 -o gpurun_out/prof_dgemm2 python tools/gemm_once.py d 8192 > gpurun_out/ncu_dgemm.log 2>&1
echo "ncu dgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel -s 8 -c 1 -o gpurun_out/prof_panel python tools/lu_once.py 8192 > gpurun_out/ncu_panel.log 2>&1
echo "ncu panel rc=$?"
ls -la gpurun_out/*.ncu-rep
