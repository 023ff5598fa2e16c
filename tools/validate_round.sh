#!/bin/bash
# first GPU pass of the round: the whole -m gpu suite, the default bench line (with sub-records), the C4 traversal tool
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/validate_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/validate_pytest.log 2>&1
tail -30 gpurun_out/validate_pytest.log
( time python bench.py ) > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
tail -c 6000 gpurun_out/validate_bench.json
tail -5 gpurun_out/validate_bench.err
( time python tools/bench_traversal.py --big 16777216 ) > gpurun_out/validate_trav.jsonl 2> gpurun_out/validate_trav.err
cat gpurun_out/validate_trav.jsonl | cut -c 1-600
tail -5 gpurun_out/validate_trav.err
