#!/bin/bash
# A/B of bench_traversal.py argument sets on C4: bash tools/ab_args.sh "<args>" "<args>" ...
mkdir -p gpurun_out
i=0
for a in "$@"; do
  i=$((i+1)); echo "== $a"
  timeout -k 10 200 python tools/bench_traversal.py --no-check $a > gpurun_out/ab_args_$i.jsonl 2> gpurun_out/ab_args_$i.err; echo "  rc=$?"
  python tools/fmt_traversal.py < gpurun_out/ab_args_$i.jsonl; tail -2 gpurun_out/ab_args_$i.err
done
