import json, sys
for l in sys.stdin:
    try:
        d = json.loads(l)
    except Exception:
        print(l.rstrip()); continue
    if "ms" in d:
        print(f'{d["what"]:45s} {d["ms"]:8.3f} ms {d["mrays_per_s"]:8.1f} Mrays/s  frac {d["roofline"]["frac"]:.3f}  nodes/ray {d["nodes_per_ray"]:.1f} prims/ray {d["prims_per_ray"]:.1f}')
    else:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k not in ("roofline", "build_ms_all")})
