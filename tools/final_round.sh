#!/bin/bash
# the round's closing GPU call: suite, DRAM traffic of the three workloads, default bench line, reference arm, C4 tool with the parity check
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout -k 10 500 bash tools/run_traffic.sh > /dev/null 2>&1; echo "traffic rc=$?"
timeout -k 10 300 python tools/bench_traversal.py --big 16777216 > gpurun_out/final_c4.jsonl 2> gpurun_out/final_c4.err; echo "c4 rc=$?"
python tools/fmt_traversal.py < gpurun_out/final_c4.jsonl
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_c4.csv python tools/bench_traversal.py --no-check --reps 1 > gpurun_out/ncu_c4_final.log 2>&1; echo "c4 launches rc=$?"
