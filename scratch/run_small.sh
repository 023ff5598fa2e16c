timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
for w in cornell material_grid; do for o in 0 -1; do echo "== $w small_scene=$o"; timeout 200 python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline --opt small_scene=$o 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,3p; done; done
