#!/bin/bash
# N GPUs: bench.py under torchrun, weak and strong, both reductions (every launch under its own timeout)
N=${1:-2}
mkdir -p gpurun_out
for sc in weak strong; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --scaling $sc > gpurun_out/e_n${N}_$sc.json 2> gpurun_out/e_n${N}_$sc.err
echo "rc=$?"; tail -c 2500 gpurun_out/e_n${N}_$sc.json; echo; grep -v "^\[e2e\|NCCL INFO\|OMP_NUM\|\*\*\*\*" gpurun_out/e_n${N}_$sc.err | tail -5
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --reduce root --no-sub --no-e2e > gpurun_out/e_n${N}_root.json 2> gpurun_out/e_n${N}_root.err
python tools/bench_summary.py gpurun_out/e_n${N}_root.json | head -2
