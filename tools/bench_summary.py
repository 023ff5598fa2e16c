import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
r = d["roofline"]
print(f'value {d["value"]:.1f} Msamples/s  {d["mrays_per_s"]:.0f} Mrays/s  ms/step {d["ms_per_step"]:.1f}  e2e {d["e2e"]["value"] if d["e2e"] else None}')
print('stage ms/step', {k: round(v, 2) for k, v in r["stage_ms_per_step"].items()})
print('dominant', r["kernel"], 'achieved GB/s', round(r["achieved"]), 'frac', round(r["frac"], 3), 'all', {k: round(v) for k, v in r["all_stage_gbs"].items()})
print('clocks', d["clocks"], 'launches', d["gpu_launches"], 'cpu', d.get("cpu_baseline"))
