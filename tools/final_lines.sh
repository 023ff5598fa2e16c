#!/bin/bash
# the round's final bench lines: N = 1 default (with sub-records), N = 2 weak with the reduce record
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
timeout -k 10 400 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo rc=$?
else
timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err; echo rc=$?
fi
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/final_n$N.json') if l.startswith('{')][-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'reduce', d.get('reduce'), 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
for k in ('c3','c4','c5','c2_strong'):
    if k in d: print(k, json.dumps(d[k])[:400])
PY
