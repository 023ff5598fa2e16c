timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r1k_bench_cornell.json 2>gpurun_out/r1k_bench_cornell.err
timeout 300 python bench.py --workload material_grid --steps 3 > gpurun_out/r1k_bench_grid.json 2>/dev/null
timeout 300 python bench.py --workload terrain --steps 3 > gpurun_out/r1k_bench_terrain.json 2>/dev/null
timeout 300 python tools/bench_traversal.py --big 16777216 > gpurun_out/r1k_c4.jsonl 2>/dev/null
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1k_bench_ref.json 2>/dev/null
for f in gpurun_out/r1k_bench_cornell.json gpurun_out/r1k_bench_grid.json gpurun_out/r1k_bench_terrain.json; do echo $f; python tools/bench_summary.py $f 2>/dev/null; done
cut -c1-330 gpurun_out/r1k_c4.jsonl
bash profiles/run_ncu.sh r1k > /dev/null 2>&1; ls gpurun_out | grep r1k
