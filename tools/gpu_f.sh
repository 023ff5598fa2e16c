#!/bin/bash
python bench.py --workload terrain --steps 3 --no-cpu-baseline --no-sub 2> /tmp/err.txt > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null| sed -n 1,2p; grep e2e /tmp/err.txt | tail -3
python -m pytest tests/test_gpu_render.py -m gpu -q 2>&1 | tail -3
