#!/bin/bash
# round-end evidence on one B200: tests, smoke, bench (both arms), launch lists, ncu full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/perf_probe.py lu > gpurun_out/perf_probe_lu.jsonl 2>&1; echo "probe lu rc=$?"
timeout 300 python tools/e2e_probe.py > gpurun_out/e2e_probe.jsonl 2>&1; echo "e2e probe rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_lu8192.csv python tools/lu_once.py 8192 > gpurun_out/ncu_lu.log 2>&1; echo "ncu lu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_lu4096.csv python tools/lu_once.py 4096 > gpurun_out/ncu_lu4096.log 2>&1; echo "ncu lu4096 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 1 -c 1 -o gpurun_out/prof_dgemm_final python tools/gemm_once.py d 8192 > gpurun_out/ncu_dgemm.log 2>&1; echo "ncu dgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgemm_ffma -s 1 -c 1 -o gpurun_out/prof_sgemm_final python tools/gemm_once.py s 8192 > gpurun_out/ncu_sgemm.log 2>&1; echo "ncu sgemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel_kernel -s 8 -c 1 -o gpurun_out/prof_panel_final python tools/lu_once.py 8192 > gpurun_out/ncu_panel.log 2>&1; echo "ncu panel rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel_cluster -s 8 -c 1 -o gpurun_out/prof_clpanel_final python tools/lu_once.py 4096 > gpurun_out/ncu_clpanel.log 2>&1; echo "ncu cluster panel rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trsv_kernel -c 2 -o gpurun_out/prof_trsv_final python tools/lu_once.py 8192 > gpurun_out/ncu_trsv.log 2>&1; echo "ncu trsv rc=$?"
