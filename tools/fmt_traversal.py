"""one line per record of tools/bench_traversal.py (stdin: its JSON lines)"""
import json
import sys

for l in sys.stdin:
    if not l.startswith("{"):
        print("  !!", l.rstrip()[:200])
        continue
    d = json.loads(l)
    if "mrays_per_s" in d:
        print(f"  {d['what'][:40]:40s} {d['mrays_per_s']:8.1f} Mrays/s  frac {d['roofline']['frac']:.3f}  nodes {d['nodes_per_ray']:.2f} "
              f"prims {d['prims_per_ray']:.2f} parity {d.get('parity_vs_oracle_bvh')}")
    elif d["what"] == "bvh_build":
        print("  build ms", [round(x, 2) for x in d["build_ms_all"]], "sah", round(d["sah_cost"], 2), "nodes", d["n_nodes"], "bytes", d["bvh_bytes"])
