#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/n${N}_weak.json 2> gpurun_out/n${N}_weak.err
echo "rc=$?"; python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/n${N}_weak.json') if l.startswith('{')][-1])
print('weak value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'c5', d.get('c5'), 'c2_strong', d.get('c2_strong'))
PY
grep -v "NCCL INFO\|^\[e2e\|OMP_NUM\|\*\*\*" gpurun_out/n${N}_weak.err | tail -5
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-sub --no-e2e > gpurun_out/n${N}_strong.json 2> gpurun_out/n${N}_strong.err
echo "rc=$?"; python tools/bench_summary.py gpurun_out/n${N}_strong.json 2>/dev/null | head -2
