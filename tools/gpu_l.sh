#!/bin/bash
for flags in "--no-cpu-baseline --no-e2e" "--no-e2e" "--no-cpu-baseline"; do
python bench.py $flags 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1]); c4 = d['c4']
print('$flags', 'value', round(d['value'],1), 'c4 load_s', round(c4['scene_load_s'],3), 'cold', round(c4['build']['cold_ms'],1), 'warm', round(c4['build']['warm_ms'],2), 'traffic', d['roofline']['traffic'])
"; done
