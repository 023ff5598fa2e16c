#!/usr/bin/env python
"""Converged-image parity (north_star: "converged 4096-spp images must match within relMSE < 1e-3").
Renders the same scene with the same seeds 0..spp-1 on the GPU (through the host surface) and with the CPU oracle
(reference headers when oracle/_ref is present), at a reduced resolution so the CPU side finishes in about a
minute, and prints relMSE = mean((gpu - ref)^2 / (ref^2 + 1e-2)) plus the match statistics of the first frame.

    python tools/converged_parity.py [--spp 4096] [--width 240] [--height 135]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=4096)
    ap.add_argument("--width", type=int, default=240)
    ap.add_argument("--height", type=int, default=135)
    args = ap.parse_args()
    import orc
    from pupiloptixlab_b200 import pupil, scenes
    lib, kind = orc.ref(), "reference"
    if lib is None:
        lib, kind = orc.port(), "port"
    pupil.init(0)
    for name, desc in (("cornell", scenes.cornell_box(args.width, args.width, 8)), ("material_grid", scenes.material_grid(args.width, args.height, 8))):
        pupil.load_scene(desc)
        pupil.pass_config(frames_per_run=1)
        pupil.run(1)
        g1 = pupil.buffer("final result")[..., :3].reshape(-1, 3).astype(np.float64)
        t0 = time.perf_counter()
        pupil.pass_config(frames_per_run=min(args.spp, 256))
        pupil.run(max(1, args.spp // 256))
        g = pupil.buffer("pt accum buffer")[..., :3].reshape(-1, 3).astype(np.float64)
        t_gpu = time.perf_counter() - t0
        o = orc.OracleScene(lib, desc)
        r1 = o.render(1)["frame"][:, :3].astype(np.float64)
        t0 = time.perf_counter()
        r = o.render(args.spp, threads=os.cpu_count())["accum"][:, :3].astype(np.float64)
        t_cpu = time.perf_counter() - t0
        ok1 = (np.abs(g1 - r1) <= 1e-4 * np.maximum(1.0, np.abs(r1))).all(1)
        print(json.dumps({"what": "converged_parity", "scene": name, "width": desc.sensor.width, "height": desc.sensor.height, "spp": args.spp,
                          "oracle": kind, "relmse": float(np.mean((g - r) ** 2 / (r ** 2 + 1e-2))),
                          "max_abs_diff": float(np.abs(g - r).max()), "mean_gpu": float(g.mean()), "mean_ref": float(r.mean()),
                          "frame0_pixels_within_1e-4": float(ok1.mean()), "frame0_bit_exact": float((g1.astype(np.float32) == r1.astype(np.float32)).all(1).mean()),
                          "gpu_s": t_gpu, "cpu_s": t_cpu, "cpu_threads": os.cpu_count()}), flush=True)
    pupil.shutdown()


if __name__ == "__main__":
    main()
