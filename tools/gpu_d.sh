#!/bin/bash
for r in 4 16; do
  echo "== builder 2 radius $r"; python tools/bench_traversal.py --builder 2 --ploc-radius $r --no-check 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if not l.startswith('{'):
        print('  !!', l.rstrip()[:300]); continue
    d = json.loads(l)
    if 'mrays_per_s' in d: print(f\"  {d['what'][:40]:40s} {d['mrays_per_s']:8.1f} Mrays/s  frac {d['roofline']['frac']:.3f}  nodes {d['nodes_per_ray']:.2f} prims {d['prims_per_ray']:.2f}\")
    elif d['what']=='bvh_build': print('  build ms', [round(x,2) for x in d['build_ms_all']], 'sah', round(d['sah_cost'],2), 'nodes', d['n_nodes'], 'depth', d['depth'])
"
done
for b in 0 2; do echo "== terrain render builder $b"; python bench.py --workload terrain --steps 3 --no-e2e --no-cpu-baseline --no-sub --builder $b 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json | sed -n 1,2p; done
