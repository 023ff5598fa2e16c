#!/bin/bash
# round-2 profile set: launch list + full capture of the default bench, DRAM traffic of the three hot kernels on the three
# workloads, launch list of the C4 traversal tool
mkdir -p gpurun_out
timeout -k 10 600 bash profiles/run_ncu.sh r2a > gpurun_out/prof_r2a.log 2>&1
timeout -k 10 600 bash tools/run_traffic.sh > gpurun_out/prof_traffic.log 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a_c4.csv python tools/bench_traversal.py --no-check > gpurun_out/ncu_c4_r2a.log 2>&1
ls -la gpurun_out | tail -20
