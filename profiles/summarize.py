#!/usr/bin/env python
"""Turn the raw ncu outputs of profiles/run_ncu.sh (gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep)
into the small tracked summaries under profiles/:  <tag>_launches.md (per-kernel share of the step) and
<tag>_ncu.md (the metrics of each fully captured launch).   python profiles/summarize.py <tag>"""
import collections
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
out_dir = ROOT / "profiles"

lines = [l for l in open(ROOT / "gpurun_out" / f"launches_{tag}.csv") if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("pb2::<unnamed>::", "").replace("void ", "")
    if name.startswith("cub::"):
        name = name.split("<")[0]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(out_dir / f"{tag}_launches.md", "w") as f:
    f.write(f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)\n\n")
    f.write("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / tot * 100:.1f}% | {v[1] / v[0]:.1f} |\n")
print(open(out_dir / f"{tag}_launches.md").read())

rep = ROOT / "gpurun_out" / f"prof_{tag}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
    idx = [hdr.index(w) for w in want if w in hdr]
    kn = hdr.index("Kernel Name")
    with open(out_dir / f"{tag}_ncu.md", "w") as f:
        f.write(f"# ncu --set full capture `{tag}` (per launch)\n\n| metric | unit | " + " | ".join(
            re.sub(r"\(.*", "", r[kn]).split("::")[-1] + f" #{i}" for i, r in enumerate(rows[2:])) + " |\n")
        f.write("|---|---|" + "---|" * len(rows[2:]) + "\n")
        for i in idx:
            vals = []
            for r in rows[2:]:
                try:
                    vals.append(f"{float(r[i].replace(',', '')):.4g}")
                except ValueError:
                    vals.append(r[i])
            f.write(f"| {hdr[i]} | {units[i]} | " + " | ".join(vals) + " |\n")
    print(open(out_dir / f"{tag}_ncu.md").read()[:1500])
