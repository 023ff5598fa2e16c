#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for red in all root; do for sc in weak strong; do
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --reduce $red --scaling $sc --no-sub --no-e2e > gpurun_out/e2_n${N}_${red}_$sc.json 2> gpurun_out/e2_n${N}_${red}_$sc.err
echo "== N=$N reduce=$red scaling=$sc rc=$?"; python tools/bench_summary.py gpurun_out/e2_n${N}_${red}_$sc.json 2>/dev/null | head -2
done; done
