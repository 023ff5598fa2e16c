timeout 120 python -m pytest tests/test_gpu_traversal.py -m gpu -q 2>&1 | tail -2
timeout 200 python tools/bench_traversal.py 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['what'], round(d.get('mrays_per_s',d.get('build_ms',0)),2), round(d['roofline']['frac'],3), d.get('build_ms_all',''), d.get('n_nodes',''), d.get('sah_cost',''))
"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_collapse|k_scatter|k_refit|k_emit|k_radix|k_morton|RadixSort' -c 40 --csv --log-file gpurun_out/launches_build.csv python tools/bench_traversal.py --rays 1024 > /dev/null 2>&1
