#!/bin/bash
for pif in 33554432 67108864 134217728; do for w in cornell material_grid; do
echo "== $w paths_in_flight=$pif"; python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline --no-sub --opt paths_in_flight=$pif 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,2p; done; done
