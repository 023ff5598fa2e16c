#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/b_pytest.log 2>&1
tail -25 gpurun_out/b_pytest.log
# source-level capture of the incoherent closest-hit batch of C4 (launch 10 of k_trace_closest: 9 launches belong to the primary batch)
ncu --set full --clock-control none --import-source on -k regex:k_trace_closest -s 10 -c 1 -f -o gpurun_out/b_c4_inco python tools/bench_traversal.py --no-check > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
ls -la gpurun_out/*.ncu-rep
