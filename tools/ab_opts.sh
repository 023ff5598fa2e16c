#!/bin/bash
# A/B of pb2 scene options on whole renders: bash tools/ab_opts.sh "<name=value[,name=value] ...>" "<workload> ..."   ("-" = no option)
mkdir -p gpurun_out
for o in $1; do
  echo "== $o"
  opts=""; if [ "$o" != "-" ]; then for kv in ${o//,/ }; do opts="$opts --opt $kv"; done; fi
  for w in $2; do
    timeout -k 10 200 python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline --no-sub $opts > gpurun_out/abo_${o}_$w.json 2> gpurun_out/abo_${o}_$w.err; echo "  $w rc=$?"
    python tools/bench_summary.py gpurun_out/abo_${o}_$w.json 2>/dev/null | sed -n 1,2p
  done
done
