#!/bin/bash
# compute-sanitizer over the kernels written this round (small cases): memcheck, then racecheck (shared-memory hazards)
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_traversal.py tests/test_gpu_kat.py tests/test_gpu_multi.py tests/test_gpu_collapse.py "tests/test_gpu_instancing.py::test_instanced_scene_matches_the_flattened_oracle" -m gpu -q -x -k "not 3_000_001" > gpurun_out/sanitize_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|at pb2" gpurun_out/sanitize_memcheck.log | sort | uniq -c | head -12
timeout -k 10 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_kat.py::test_radix_sort_is_a_stable_sort "tests/test_gpu_traversal.py::test_closest_hit_matches_brute_force" tests/test_gpu_render.py::test_one_frame_same_seed_parity -m gpu -q -x -k "not 3_000_001" > gpurun_out/sanitize_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|hazard|at pb2" gpurun_out/sanitize_racecheck.log | sort | uniq -c | head -12
