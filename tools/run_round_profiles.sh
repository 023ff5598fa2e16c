#!/bin/bash
# the round's ncu evidence in one GPU call: bash tools/run_round_profiles.sh <tag>
#   launch list + full capture of the default bench command (Cornell), DRAM traffic of the three workloads, launch list of the
#   C4 tool (build kernels) and a full capture of k_trace_closest on the incoherent C4 batch
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout -k 10 500 bash profiles/run_ncu.sh $TAG > /dev/null 2>&1; echo "cornell capture rc=$?"
timeout -k 10 600 bash tools/run_traffic.sh > /dev/null 2>&1; echo "traffic rc=$?"
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}_c4.csv python tools/bench_traversal.py --no-check --reps 1 > gpurun_out/ncu_c4_${TAG}.log 2>&1; echo "c4 launches rc=$?"
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 10 -c 1 -f -o gpurun_out/prof_${TAG}_c4 python tools/bench_traversal.py --no-check >> gpurun_out/ncu_c4_${TAG}.log 2>&1; echo "c4 capture rc=$?"
ls -la gpurun_out | grep -E "${TAG}|traffic_"
