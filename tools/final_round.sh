#!/bin/bash
# the round's closing GPU call: suite, DRAM traffic of the three workloads, C4 tool with the parity check, C4 launch list,
# default bench line (after the traffic figures, so that it can report them), reference arm, smoke
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout -k 10 500 bash tools/run_traffic.sh > /dev/null 2>&1; echo "traffic rc=$?"
python profiles/traffic_table.py > gpurun_out/final_traffic_table.log 2>&1; cp profiles/roofline_traffic.json gpurun_out/roofline_traffic.json
timeout -k 10 300 python tools/bench_traversal.py --big 16777216 > gpurun_out/final_c4.jsonl 2> gpurun_out/final_c4.err; echo "c4 rc=$?"
python tools/fmt_traversal.py < gpurun_out/final_c4.jsonl
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_c4.csv python tools/bench_traversal.py --no-check --reps 1 > gpurun_out/ncu_c4_final.log 2>&1; echo "c4 launches rc=$?"
bash tools/final_lines.sh 1
timeout -k 10 300 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
