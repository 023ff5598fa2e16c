#!/bin/bash
# Profiling recipe of this repo (run under gpurun, one GPU).  Outputs land in gpurun_out/; summaries are copied to profiles/.
#   bash profiles/run_ncu.sh <tag> [extra bench.py args]
set -u
TAG=${1:-r1}; shift || true
CMD="python bench.py --steps 1 --warmup 3 --spp 8 --no-e2e --no-cpu-baseline --no-sub $*"
mkdir -p gpurun_out
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_bench_${TAG}.log 2>&1
# 2. full capture of the three hot kernels: rounds 0..2 of the first timed batch
ncu --set full --clock-control none --import-source on -k regex:'k_(shade|extend|shadow|bin)' -s 4 -c 8 -f -o gpurun_out/prof_${TAG} $CMD >> gpurun_out/ncu_bench_${TAG}.log 2>&1
ls -la gpurun_out/
