python -m pytest tests -m gpu -q 2>&1 | tail -3
for c in 0 1; do echo "== coop $c"; python tools/bench_traversal.py --coop $c 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['what'], round(d.get('mrays_per_s',d.get('build_ms',0)),1), round(d['roofline']['frac'],3))
"; for w in terrain material_grid; do python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline --opt coop_prims=$c 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,2p; done; done
