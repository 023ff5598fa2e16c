for v in _build_st6 _build_st12; do echo "== $v"; PB2_BUILD_DIR=$v timeout 200 python tools/bench_traversal.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['what'], round(d.get('mrays_per_s',d.get('build_ms',0)),1), round(d['roofline']['frac'],3))
"; for w in terrain; do PB2_BUILD_DIR=$v timeout 200 python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,2p; done; done
