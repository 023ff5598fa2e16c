#!/bin/bash
# A/B of the BVH8 collapse on C4: greedy largest-area expansion against the SAH-optimal cut at several primitive-test costs
#   bash tools/ab_collapse.sh "0:0 1:30 1:60" [extra args of bench_traversal.py]
mkdir -p gpurun_out
variants=${1:-0:0 1:30 1:60 1:100}
shift
for v in $variants; do
  c=${v%%:*}; p=${v##*:}
  echo "== collapse $c prim cost $p"
  timeout -k 10 200 python tools/bench_traversal.py --collapse $c --prim-cost $p --no-check "$@" > gpurun_out/ab_collapse_${c}_${p}.jsonl 2> gpurun_out/ab_collapse_${c}_${p}.err; echo "  rc=$?"
  python tools/fmt_traversal.py < gpurun_out/ab_collapse_${c}_${p}.jsonl; tail -3 gpurun_out/ab_collapse_${c}_${p}.err
done
