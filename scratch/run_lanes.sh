timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in cornell material_grid terrain; do for t in 0 1; do echo "== $w two_lanes=$t"; timeout 200 python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline --opt two_lanes=$t 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,2p; done; done
