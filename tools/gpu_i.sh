#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_instancing.py -m gpu -q -x 2>&1 | tail -4
for srt in 0 1; do echo "== sorter $srt"; timeout 200 python tools/bench_traversal.py --sorter $srt --no-check 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if not l.startswith('{'):
        print('  !!', l.rstrip()[:300]); continue
    d = json.loads(l)
    if 'mrays_per_s' in d: print(f\"  {d['what'][:40]:40s} {d['mrays_per_s']:8.1f} Mrays/s  nodes {d['nodes_per_ray']:.2f} prims {d['prims_per_ray']:.2f}\")
    elif d['what']=='bvh_build': print('  build ms', [round(x,2) for x in d['build_ms_all']], 'sah', round(d['sah_cost'],2), 'nodes', d['n_nodes'])
"; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_sort|DeviceRadixSort' -c 60 --csv python tools/bench_traversal.py --no-check --rays 1024 2>/dev/null | grep -E "k_sort|RadixSort" | awk -F'","' '{print $5, $(NF)}' | sed 's/"//g' | sort | uniq -c | sort -rn | head -12
