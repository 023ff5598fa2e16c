#!/usr/bin/env python
"""gpurun_out/traffic_<workload>.csv (tools/run_traffic.sh: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum over the trace / shade kernels of one bench.py run) -> profiles/roofline_traffic.json, the
measured DRAM bytes per launch that bench.py reports as roofline.traffic."""
import collections
import sys
import csv
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from source_hash import kernel_source_sha  # noqa: E402
out = {}
for w in ["cornell", "material_grid", "terrain"]:
    lines = [l for l in open(ROOT / "gpurun_out" / f"traffic_{w}.csv") if not l.startswith("==")]
    L = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1])
        d = L.setdefault(r["ID"], {"name": name})
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(r["Metric Unit"], 1)
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * scale
    per = {}
    for k in ["k_shade", "k_extend", "k_shadow"]:
        ls = [l for l in L.values() if l["name"] == k]
        n = len(ls)
        sel = ls[3 * n // 5:4 * n // 5]  # 3 warm-up passes, the timed step, the counting pass: take the timed step
        if sel:
            per[k.replace("k_", "")] = round(sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in sel) / len(sel))
            print(w, k, len(sel), "launches/step,", round(per[k.replace("k_", "")] / 1e6, 1), "MB/launch,",
                  round(1e3 * sum(l["gpu__time_duration.sum"] for l in sel) / len(sel), 3), "ms/launch under ncu")
    out[w] = per
json.dump({"kernel_source_sha": kernel_source_sha(), "_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the launches of one timed bench.py step "
                       "(tools/run_traffic.sh, profiles/traffic_table.py); bench.py copies the dominant kernel's figure into roofline.traffic while kernel_source_sha (tools/source_hash.py) still matches the sources it runs", **out},
          open(ROOT / "profiles" / "roofline_traffic.json", "w"), indent=1)
