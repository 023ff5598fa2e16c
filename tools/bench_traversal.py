#!/usr/bin/env python
"""Config C4 (BASELINE.json configs[3]): GPU BVH build time and BVH8 traversal throughput on the ~30 M-triangle
procedural terrain, with the traversal kernels' memory roofline from their own node / primitive counters.

    python tools/bench_traversal.py [--n 3873] [--rays 2097152] [--builder 0|1] [--reps 5]

Ray batches: (i) 1080p primary camera rays, (ii) incoherent cosine-bounce rays leaving the primary hit
points, (iii) shadow rays from the hit points to the area light.  Prints one JSON line per batch plus one for the build.
Rays are resident in HBM; timing = CUDA events around `reps` launches of pb2_trace_*_dev on torch's stream.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


sys.path.insert(0, str(ROOT / "tools"))
from ray_batches import bounce_rays, camera_rays, hit_points, measure_trace, shadow_rays  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3873)
    ap.add_argument("--rays", type=int, default=1 << 21)
    ap.add_argument("--builder", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--refill", type=str, default="", help="comma list of refill thresholds to sweep")
    ap.add_argument("--orders", action="store_true", help="also trace the incoherent batch in three coherent orders")
    ap.add_argument("--big", type=int, default=0, help="also trace this many incoherent / shadow rays in one launch (origins resampled with replacement): "
                    "the 2^21-ray batches of config C4 give each resident warp ~200 rays, so ramp-up and tail are a visible share of the launch")
    ap.add_argument("--l2", type=str, default="", help="comma list of persist_mb:window_mb pairs for the L2 access-policy window over the top of the node array")
    ap.add_argument("--coop", type=int, default=-1, help="warp-cooperative primitive tests: 1 on, 0 off, -1 auto")
    ap.add_argument("--ploc-radius", type=int, default=0, help="builder 2: neighbours searched on either side along the Morton curve")
    ap.add_argument("--collapse", type=int, default=-1, help="binary tree -> BVH8: 1 SAH-optimal cut (default), 0 greedy largest-area expansion")
    ap.add_argument("--prim-cost", type=int, default=0, help="collapse = 1: cost of a primitive test in per cent of a wide-node test")
    ap.add_argument("--morton-bits", type=int, default=0, help="leading bits of the 63-bit Morton key that are sorted (8 per sort pass)")
    ap.add_argument("--no-check", action="store_true", help="skip the parity check of 8192 sampled rays per batch against the oracle's CPU BVH")
    args = ap.parse_args()
    import torch
    from pupiloptixlab_b200 import pupil, scenes
    peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0

    t0 = time.perf_counter()
    desc = scenes.terrain(args.n, args.width, args.height, 8)
    t_gen = time.perf_counter() - t0
    pupil.init(0)
    if args.builder >= 0:
        pupil.set_bvh_builder(args.builder)
    t0 = time.perf_counter()
    pupil.load_scene(desc)
    scene = pupil.scene_handle()
    t_load = time.perf_counter() - t0
    bs = pupil.build_stats()
    builds = [bs.build_ms]
    if args.ploc_radius:
        pupil.scene_handle().set_option("ploc_radius", args.ploc_radius)
    if args.collapse >= 0:
        pupil.scene_handle().set_option("collapse", args.collapse)
    if args.morton_bits:
        pupil.scene_handle().set_option("morton_bits", args.morton_bits)
    if args.prim_cost:
        pupil.scene_handle().set_option("collapse_prim_cost_pct", args.prim_cost)
    for _ in range(2):  # rebuild twice more: steady-state build time (allocator warm)
        pupil.set_bvh_builder(args.builder if args.builder >= 0 else 0)
        builds.append(pupil.build_stats().build_ms)
    scene = pupil.scene_handle()
    bs = pupil.build_stats()
    tri_bytes = bs.n_triangles * 230
    print(json.dumps({"what": "bvh_build", "n_prims": bs.n_prims, "n_nodes": bs.n_nodes, "bvh_bytes": bs.bvh_bytes, "build_ms": min(builds),
                      "build_ms_all": builds, "mtris_per_s": bs.n_triangles / min(builds) / 1e3, "sah_cost": bs.sah_cost, "depth": bs.max_depth,
                      "roofline": {"bound": "hbm", "achieved": tri_bytes / (min(builds) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": tri_bytes / (min(builds) * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_triangle": 230},
                      "scene_generate_s": t_gen, "scene_load_s": t_load}), flush=True)

    stream = torch.cuda.Stream()
    scene.set_stream(stream.cuda_stream)
    scene.set_option("coop_prims", args.coop)
    s2c, c2w, _ = pupil.camera()
    prim = camera_rays(s2c, c2w, args.width, args.height)

    oracle = None
    if not args.no_check:  # the oracle is the checker here, never the thing measured (its time is not part of any number printed)
        sys.path.insert(0, str(ROOT / "tests"))
        import orc
        from gpu_util import compare_hits
        t0 = time.perf_counter()
        oracle = orc.OracleScene(orc.port(), desc)
        print(json.dumps({"what": "oracle_bvh", "seconds": time.perf_counter() - t0}), flush=True)
    check_rng = np.random.default_rng(99)

    def check(rays_np, any_hit, tuvp_t, inst_t, occ_t, label):
        """8192 sampled rays of the batch against the oracle's CPU BVH: ids exact up to ties, t within 1e-5 relative"""
        sel = check_rng.choice(len(rays_np), min(8192, len(rays_np)), replace=False)
        sub = rays_np[sel]
        if any_hit:
            ref = oracle.trace_any(sub)
            got = occ_t.cpu().numpy()[sel] != 0
            bad = int(np.count_nonzero(got != (ref != 0)))
            assert bad <= 4, f"{label}: {bad} of {len(sel)} any-hit answers differ from the oracle"
            return {"checked": len(sel), "differ": bad}
        ref, _ = oracle.trace_closest(sub, threads=os.cpu_count())
        tu, ins = tuvp_t.cpu().numpy()[sel], inst_t.cpu().numpy()[sel]
        gpu = np.zeros(len(sel), ref.dtype)
        gpu["t"], gpu["u"], gpu["v"], gpu["inst"] = tu[:, 0], tu[:, 1], tu[:, 2], ins
        gpu["prim"] = np.where(ins >= 0, np.ascontiguousarray(tu[:, 3]).view(np.uint32).astype(np.int64), -1)
        ties = compare_hits(gpu, ref, sub)
        return {"checked": len(sel), "ties_excused": ties}

    def trace(rays_np, any_hit, label):
        rec, tuvp, inst, occ = measure_trace(torch, scene, stream, rays_np, any_hit, args.reps, peak, label)
        rec["parity_vs_oracle_bvh"] = check(rays_np, any_hit, tuvp, inst, occ, label) if oracle is not None else None
        print(json.dumps(rec), flush=True)
        return tuvp.cpu().numpy(), inst.cpu().numpy()

    sweep = [tuple(int(y) for y in x.split(":")) for x in args.refill.split(",")] if args.refill else [None]

    def apply(thr):
        scene.set_option("refill_threshold", thr[0])

    tuvp, inst = trace(prim, False, "closest_primary_1080p")
    pos = hit_points(prim, tuvp[:, 0], inst)
    rng = np.random.default_rng(7)
    k = min(args.rays, len(pos))
    inco, p = bounce_rays(pos, k, rng)
    if args.orders:  # how much of the incoherent batch's cost is ordering: same rays, binned by direction octant / sorted along a Morton curve of the origins
        octant = (inco[:, 4] >= 0) * 4 + (inco[:, 5] >= 0) * 2 + (inco[:, 6] >= 0)
        trace(inco[np.argsort(octant, kind="stable")], False, "closest_incoherent_bounce binned by octant")
        lo, hi = inco[:, 0:3].min(0), inco[:, 0:3].max(0)
        q = np.clip((inco[:, 0:3] - lo) / np.maximum(hi - lo, 1e-9) * 1023, 0, 1023).astype(np.uint64)

        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            return (v | (v << 2)) & 0x09249249
        morton = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
        trace(inco[np.argsort(morton, kind="stable")], False, "closest_incoherent_bounce sorted by origin (Morton)")
        trace(inco[np.lexsort((morton, octant))], False, "closest_incoherent_bounce sorted by octant, then origin")
    for thr in sweep:
        if thr is not None:
            apply(thr)
            trace(prim, False, f"closest_primary_1080p thr={thr}")
        trace(inco, False, f"closest_incoherent_bounce thr={thr}")
    sh = shadow_rays(p, rng)
    for thr in sweep:
        if thr is not None:
            apply(thr)
        trace(sh, True, f"anyhit_shadow_to_light thr={thr}")
    if args.big:
        kb = args.big
        selb = rng.integers(0, len(pos), kb)
        pb = pos[selb]
        u1, u2 = rng.random(kb, dtype=np.float32), rng.random(kb, dtype=np.float32)
        rr, ph = np.sqrt(u1), 2 * np.pi * u2
        big = np.zeros((kb, 8), np.float32)
        big[:, 0:3], big[:, 3], big[:, 7] = pb, 1e-3, 1e16
        big[:, 4], big[:, 5], big[:, 6] = rr * np.cos(ph), np.sqrt(np.maximum(0, 1 - u1)), rr * np.sin(ph)
        trace(big, False, f"closest_incoherent_bounce {kb} rays")
        lb = np.array([0.0, 12.0, 0.0], np.float32) + rng.uniform(-3, 3, (kb, 3)).astype(np.float32) * np.array([1, 0, 1], np.float32)
        dlb = lb - pb
        db = np.linalg.norm(dlb, axis=1, keepdims=True)
        big[:, 3], big[:, 4:7], big[:, 7] = 1e-4, dlb / db, db[:, 0] - 1e-4
        trace(big, True, f"anyhit_shadow_to_light {kb} rays")
        del big
    for pair in [x for x in args.l2.split(",") if x]:
        persist, window = (int(y) for y in pair.split(":"))
        scene.set_option("l2_persist_mb", persist)
        scene.set_option("l2_window_mb", window)
        trace(prim, False, f"closest_primary_1080p l2={pair}")
        trace(inco, False, f"closest_incoherent_bounce l2={pair}")
        trace(sh, True, f"anyhit_shadow_to_light l2={pair}")
    pupil.shutdown()


if __name__ == "__main__":
    main()
