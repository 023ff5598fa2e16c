#!/bin/bash
# A/B of variant build directories (PB2_BUILD_DIR): bash tools/ab.sh "<dir> <dir> ..." "<workload> ..." [tests]
for v in $1; do for w in $2; do echo "== $v $w"; PB2_BUILD_DIR=$v python bench.py --workload $w --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null > /tmp/b.json; python tools/bench_summary.py /tmp/b.json 2>/dev/null | sed -n 1,2p; done
if [ -n "$3" ]; then echo "-- tests $v"; PB2_BUILD_DIR=$v python -m pytest tests -m gpu -q 2>&1 | grep -E "^E       Assertion|^E       assert|FAILED|passed|failed" | cut -c1-200; fi; done
