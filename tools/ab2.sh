#!/bin/bash
# A/B of variant build directories (PB2_BUILD_DIR) on the C4 traversal batches and on whole renders:
#   bash tools/ab2.sh "<dir> <dir> ..." "<workload> ..." [traversal: 1|0]
mkdir -p gpurun_out
for v in $1; do
  echo "== $v"
  if [ "${3:-1}" = "1" ]; then
    PB2_BUILD_DIR=$v timeout -k 10 200 python tools/bench_traversal.py --no-check > gpurun_out/ab2_$v.jsonl 2> gpurun_out/ab2_$v.err; echo "  traversal rc=$?"
    python tools/fmt_traversal.py < gpurun_out/ab2_$v.jsonl
  fi
  for w in $2; do
    PB2_BUILD_DIR=$v timeout -k 10 200 python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline --no-sub > gpurun_out/ab2_${v}_$w.json 2> gpurun_out/ab2_${v}_$w.err; echo "  $w rc=$?"
    python tools/bench_summary.py gpurun_out/ab2_${v}_$w.json 2>/dev/null | sed -n 1,2p
  done
done
