#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/j_pytest.log 2>&1
tail -14 gpurun_out/j_pytest.log
python bench.py --no-sub --no-cpu-baseline 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json | head -3
