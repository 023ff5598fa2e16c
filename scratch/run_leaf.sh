for v in _build _build_leaf2 _build_leaf1; do echo "== $v"; PB2_BUILD_DIR=$v python tools/bench_traversal.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['what'], round(d.get('mrays_per_s',d.get('build_ms',0)),1), round(d['roofline']['frac'],3), d.get('n_nodes',''), round(d.get('nodes_per_ray',0),1), round(d.get('prims_per_ray',0),1))
"; for w in cornell material_grid; do PB2_BUILD_DIR=$v python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,2p; done; done
PB2_BUILD_DIR=_build_leaf1 python -m pytest tests/test_gpu_traversal.py -m gpu -q 2>&1 | tail -2
