#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/h_pytest.log 2>&1
tail -16 gpurun_out/h_pytest.log
python bench.py 2> gpurun_out/h_bench.err > gpurun_out/h_bench.json
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/h_bench.json') if l.startswith('{')][-1])
r = d['roofline']
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'stage', {k: round(v,2) for k,v in r['stage_ms_per_step'].items()}, 'frac', round(r['frac'],3))
print('c3', round(d['c3']['value'],1), 'k_shade frac', round(d['c3']['k_shade']['frac'],3))
c4 = d['c4']; print('c4 build cold/warm', round(c4['build']['cold_ms'],2), round(c4['build']['warm_ms'],2), {k: (round(c4[k]['mrays_per_s']), round(c4[k]['roofline']['frac'],3)) for k in ('primary','incoherent','any_hit')})
print('c5', round(d['c5']['value'],1), d['c5']['ms_per_step'])
PY
