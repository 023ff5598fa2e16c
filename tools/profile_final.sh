#!/bin/bash
# final profile set of the round on the final kernels: traffic (hash-stamped), launch list + full capture, default bench line
mkdir -p gpurun_out
timeout -k 10 600 bash tools/run_traffic.sh > gpurun_out/prof_traffic.log 2>&1
timeout -k 10 600 bash profiles/run_ncu.sh r2c > gpurun_out/prof_r2c.log 2>&1
timeout -k 10 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 1500 gpurun_out/final_bench.json
ls gpurun_out | tail -30
