#!/usr/bin/env python
"""bench.py — Msamples/s of the pt-with-MIS hot path (BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cornell|material_grid|terrain] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one render of the named workload's full sample count through the host surface (System::Run ->
PTPass::OnRun -> pb2_render): default workload = BASELINE.json configs[1], the Cornell box at 1920x1080, 64 spp,
max depth 8.  Numbers:
  value        whole-job Msamples/s with the scene, BVH and buffers resident in HBM, CUDA events on the launching
               stream, K steps, max over ranks
  e2e          same metric through the public API from HOST data every step: scene description -> XML loader ->
               H2D of geometry/tables -> GPU BVH build -> render -> D2H of the float4 image
  roofline     the dominant wavefront kernel: algorithmic bytes (DESIGN.md "Algorithmic bytes") / its CUDA-event
               time, against the measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline the oracle (reference headers compiled as host C++ when oracle/_ref exists, else the port) on the
               box's host cores, bounded sample of the same workload
  c3 / c4 / c5 sub-records (N = 1; --no-sub skips them): the other configs of BASELINE.json measured in the same run —
               c3: material-ball grid 1080p 256 spp with k_shade's roofline; c4: 30 M-triangle terrain, cold and warm BVH build
               ms and the three traversal batches (Mrays/s, roofline fraction, nodes / primitives per ray); c5: the 4K
               1024-spp job on this many GPUs (strong scaling: the job is fixed, the ranks split its sample indices)
N > 1: one process per GPU inside the product (PTPass::SetShard -> pb2_shard_plan / pb2_comm_reduce_frames, NCCL loaded by
libpb2.so): rank r renders seeds base + r + k N into plain sums; the reduction (ncclReduce to rank 0 + finalize, or
reduce-scatter + finalize + all-gather with --reduce all) runs on its own stream and overlaps the next step's render up to its first accumulate
kernel.  --scaling weak (default): every rank renders the step's sample count; --scaling strong: the ranks split it.
torch.distributed carries the NCCL id to the ranks and the max-over-ranks of the timings; it is not on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (scene maker kwargs, spp per step, label)
    "cornell": dict(spp=64, label="Cornell box 1920x1080, 64 spp, max depth 8 (BASELINE.json configs[1])"),
    "material_grid": dict(spp=64, label="material-ball grid (7 BSDFs, const env + area light) 1920x1080, 64 spp per step, max depth 8 (configs[2] scene)"),
    "terrain": dict(spp=8, label="tessellated terrain 1920x1080, 8 spp per step, max depth 8 (configs[3] scene)"),
    "c5": dict(spp=1024, width=3840, height=2160, label="material-ball grid 3840x2160, 1024 spp per step, max depth 8 (BASELINE.json configs[4])"),
}


def make_scene(name: str, width: int, height: int, depth: int, terrain_n: int):
    from pupiloptixlab_b200 import scenes
    if name == "cornell":
        return scenes.cornell_box(width, height, depth)
    if name in ("material_grid", "c5"):
        return scenes.material_grid(width, height, depth)
    if name == "terrain":
        return scenes.terrain(terrain_n, width, height, depth)
    raise SystemExit(f"unknown workload {name}")


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        """NVML in a thread (a sample every 20 ms, no process start-up inside the timed region); the nvidia-smi loop is the fallback."""
        self.samples, self.thread, self.halt = [], None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_of = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.halt.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx, int(reasons_of(h))))
                    except Exception:
                        pass
                    self.halt.wait(0.02)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.path = tempfile.NamedTemporaryFile(prefix="clocks_", suffix=".csv", delete=False).name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if getattr(self, "thread", None) is not None:
            self.halt.set()
            self.thread.join(timeout=2)
            if self.samples:
                bits = 0
                for _, _, r in self.samples:
                    bits |= r
                names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}  # NVML clocks-event-reason bits
                out = {"sm_mhz": float(np.median([c for c, _, _ in self.samples])), "sm_max_mhz": self.samples[0][1],
                       "reasons": sorted(n for b, n in names.items() if bits & b), "samples": len(self.samples), "source": "nvml"}
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])), mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


class DevPtr:
    """__cuda_array_interface__ view of a BufferManager buffer so torch (NCCL) can use it in place."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def scene_h2d_bytes(desc) -> int:
    """bytes the loader copies host -> device for one scene (unique meshes + per-instance tables + emitters + camera)"""
    from pupiloptixlab_b200 import pupil
    total, seen = 128, set()
    for sh in desc.shapes:
        if sh.type == "obj":
            m = sh.mesh
            total += sum(np.asarray(m[k]).astype(np.float32 if k != "indices" else np.uint32).nbytes for k in ("positions", "normals", "texcoords", "indices") if m.get(k) is not None)
        elif sh.type in ("rectangle", "cube") and sh.type not in seen:
            seen.add(sh.type)
            nv, nf = (4, 2) if sh.type == "rectangle" else (24, 12)
            total += nv * 32 + nf * 12
    total += len(desc.shapes) * (144 + 288) + pupil.lib().pupil_num_area_emitters() * 192 + (192 if desc.env_radiance is not None else 0)
    return int(total)


_pinned_keepalive = []


def pin_mesh_arrays(torch, desc):
    """the scene's mesh arrays in pinned host memory (once, outside every timed region): the e2e leg uploads from there, as the
    contract's "inputs from pinned host memory" asks; the host library borrows the arrays instead of copying them"""
    for sh in desc.shapes:
        if sh.type != "obj":
            continue
        for k, dt in (("positions", np.float32), ("normals", np.float32), ("texcoords", np.float32), ("indices", np.uint32)):
            a = sh.mesh.get(k)
            if a is None:
                continue
            t = torch.from_numpy(np.ascontiguousarray(a, dt)).pin_memory()
            _pinned_keepalive.append(t)
            sh.mesh[k] = t.numpy()


def cpu_arm(args, desc, spp_sample: int):
    """The reference's CPU implementation of the path (oracle/_ref when built, else the oracle port), all host threads."""
    sys.path.insert(0, str(ROOT / "tests"))
    import orc
    lib, kind = orc.ref(), "reference"
    if lib is None:
        lib, kind = orc.port(), "port"
    sc = orc.OracleScene(lib, desc)
    cores = os.cpu_count() or 1
    sc.render(1, threads=cores)  # builds the CPU BVH, warms caches
    return sc, kind, cores


def run_reference_impl(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    desc = make_scene(args.workload, args.width, args.height, args.depth, args.terrain_n)
    # bounded sample per step: rows x 1 spp, sized for ~2 s of CPU work per step
    sc, kind, cores = cpu_arm(args, desc, 1)
    n_px = args.width * args.height
    times = []
    rays = 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = sc.render(args.cpu_spp, first_seed=i * args.cpu_spp, threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays += out["closest_rays"] + out["shadow_rays"]
    total = sum(times)
    value = n_px * args.cpu_spp * args.steps / total / 1e6
    sample = f"{args.cpu_spp} spp of the {args.width}x{args.height} depth-{args.depth} frame per step ({kind} oracle, {cores} threads)"
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["label"], "width": args.width, "height": args.height, "max_depth": args.depth, "spp_per_step": args.cpu_spp},
        "mrays_per_s": rays / total / 1e6,
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def stage_bytes(cs, build, desc, n_px, spp, n_ext, n_shadow):
    """Algorithmic HBM bytes of one step per stage (DESIGN.md "Algorithmic bytes"): 32-byte ray records, 16-byte hit, throughput and
    radiance records; 48-byte shadow-queue entries; 80 B / node, 48 B / primitive.  cs = render stats of one counting pass of the step,
    n_ext / n_shadow = trace launches per step."""
    nodes_c, prims_c = cs.nodes_visited - cs.nodes_shadow, cs.prims_tested - cs.prims_shadow
    big_mesh = desc.num_triangles() * 36 > 64e6  # vertex data does not stay in L2: count the hit triangle's attributes per vertex
    ext_rays = max(cs.closest_rays - n_px * spp, 0)  # extension rays emitted by k_shade (the rest are camera rays)
    # a BVH that fits the 126 MB L2 is read from HBM once per launch, not once per visit (the Cornell box's is 2 KB and lives in L1):
    # then the trace kernels' HBM bytes are their ray / hit records, and they are instruction-issue bound
    bvh_cached = build.bvh_bytes < 100e6
    tree_c = build.bvh_bytes * n_ext if bvh_cached else nodes_c * 80 + prims_c * 48
    tree_s = build.bvh_bytes * n_shadow if bvh_cached else cs.nodes_shadow * 80 + cs.prims_shadow * 48
    return {
        "extend": cs.closest_rays * (4 + 32 + 16) + tree_c,
        "shadow": cs.shadow_rays * 32 + cs.shadow_unoccluded * (16 + 16 + 16) + tree_s,
        # per vertex: queue entry 4 + hit 16 + ray 32 + throughput|rng 16 in; per extension ray: ray record 32 + throughput|rng 16 + queue entry 4 out;
        # per shadow ray: 48-byte queue entry out (radiance records are touched by emitter hits and misses only: not counted)
        "shade": cs.closest_rays * (4 + 64 + (24 if cs.sorted else 0) + (108 if big_mesh else 0)) + ext_rays * 52 + cs.shadow_rays * 48,
    }, bvh_cached, nodes_c, prims_c


def measured_traffic(workload: str, kernel: str):
    """DRAM bytes per launch from the committed ncu pass (profiles/roofline_traffic.json) — only while the CUDA sources it was
    measured on are the sources this run executes (tools/source_hash.py); otherwise null: ncu cannot run inside a timed bench."""
    prof = ROOT / "profiles" / "roofline_traffic.json"
    try:
        sys.path.insert(0, str(ROOT / "tools"))
        from source_hash import kernel_source_sha
        data = json.loads(prof.read_text())
        if data.get("kernel_source_sha") != kernel_source_sha():
            return None, "profiles/roofline_traffic.json was measured on other kernel sources: not reported"
        return data.get(workload, {}).get(kernel), "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/roofline_traffic.json (same kernel sources)"
    except Exception as e:  # noqa: BLE001
        return None, f"no traffic record ({e})"


def timed_render(torch, pupil, stream, n_px, spp, steps, warm=1, seed0=0):
    """`steps` progressive steps of `spp` samples through the pass on the already loaded scene; returns (ms per step, stage ms, stats)"""
    for i in range(warm):
        pupil.pass_config(frames_per_run=spp, first_seed=seed0 + i * spp)
        pupil.run(1)
    pupil.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = dict(generate=0.0, extend=0.0, shade=0.0, shadow=0.0, accumulate=0.0)
    tot = dict(launches=0, closest=0, shadow=0, n_ext=0, n_shade=0, n_shadow=0)
    e0.record(stream)
    for i in range(steps):
        pupil.pass_config(frames_per_run=spp, first_seed=seed0 + (warm + i) * spp)
        pupil.run(1)
        rs = pupil.render_stats()
        for k in stage:
            stage[k] += getattr(rs, k + "_ms")
        tot["launches"] += rs.kernel_launches
        tot["closest"] += rs.closest_rays
        tot["shadow"] += rs.shadow_rays
        tot["n_ext"] += rs.extend_launches
        tot["n_shade"] += rs.shade_launches
        tot["n_shadow"] += rs.shadow_launches
    pupil.synchronize()  # N > 1: the last step's reduction runs on the communicator's stream
    e1.record(stream)
    torch.cuda.synchronize()
    tot = {k: v / steps for k, v in tot.items()}
    return e0.elapsed_time(e1) / steps, {k: v / steps for k, v in stage.items()}, tot


def c3_record(torch, pupil, stream, args, peak):
    """BASELINE.json configs[2]: material-ball grid (7 BSDFs, analytic env light), 1080p, 256 spp — Msamples/s and k_shade's roofline"""
    from pupiloptixlab_b200 import scenes
    w, h, spp = 1920, 1080, 256
    desc = scenes.material_grid(w, h, args.depth)
    pupil.load_scene(desc)
    scene = pupil.scene_handle()
    scene.set_stream(stream.cuda_stream)
    scene.set_option("profiling", 1)
    build = pupil.build_stats()
    ms, stage, tot = timed_render(torch, pupil, stream, w * h, spp, steps=2, warm=1)
    scene.set_option("counting", 1)
    pupil.pass_config(frames_per_run=64, first_seed=0)
    pupil.run(1)
    cs = pupil.render_stats()
    scene.set_option("counting", 0)
    per64, _, nodes_c, prims_c = stage_bytes(cs, build, desc, w * h, 64, cs.extend_launches, cs.shadow_launches)
    scale = spp / 64  # the counting pass covers 64 of the step's 256 samples (same scene, same statistics)
    frac = {k: per64[k] * scale / (stage[k] * 1e-3) / 1e9 / peak for k in per64 if stage[k] > 0}
    return {"workload": f"material-ball grid {w}x{h}, {spp} spp, max depth {args.depth}", "value": w * h * spp / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms,
            "mrays_per_s": (tot["closest"] + tot["shadow"]) / (ms * 1e-3) / 1e6, "stage_ms_per_step": stage, "sorted_by_material": bool(cs.sorted),
            "k_shade": {"frac": frac.get("shade"), "achieved_gbs": frac.get("shade", 0) * peak, "algorithmic_bytes_per_step": per64["shade"] * scale,
                        "traffic": measured_traffic("material_grid", "shade")[0]},
            "all_stage_frac": frac, "nodes_per_closest_ray": nodes_c / max(cs.closest_rays, 1), "prims_per_closest_ray": prims_c / max(cs.closest_rays, 1)}


def c4_record(torch, pupil, stream, args, peak):
    """BASELINE.json configs[3]: ~30 M tessellated triangles, 1080p — cold and warm GPU BVH build ms, and the three traversal batches of
    SURVEY.md 8d (2^21 rays each for (ii), (iii)) with their roofline fraction by algorithmic bytes"""
    sys.path.insert(0, str(ROOT / "tools"))
    import ray_batches as rb
    from pupiloptixlab_b200 import scenes
    w, h = 1920, 1080
    t0 = time.perf_counter()
    desc = scenes.terrain(args.terrain_n, w, h, args.depth)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    pupil.load_scene(desc)
    t_load = time.perf_counter() - t0
    bs = pupil.build_stats()
    cold = bs.build_ms  # the first build of this scene in this process: device blocks come from cudaMalloc, kernels load lazily
    warm = []
    for _ in range(3):
        pupil.set_bvh_builder(args.builder)
        warm.append(pupil.build_stats().build_ms)
    bs = pupil.build_stats()
    scene = pupil.scene_handle()
    scene.set_stream(stream.cuda_stream)
    tri_bytes = bs.n_triangles * 230
    out = {"workload": f"tessellated terrain, {bs.n_triangles} triangles, {w}x{h}", "scene_generate_s": t_gen, "scene_load_s": t_load,
           "build": {"cold_ms": cold, "warm_ms": min(warm), "warm_ms_all": warm, "mtris_per_s": bs.n_triangles / min(warm) / 1e3, "n_nodes": bs.n_nodes, "bvh_bytes": bs.bvh_bytes,
                     "sah_cost": bs.sah_cost, "depth": bs.max_depth, "builder": args.builder,
                     "roofline": {"bound": "hbm", "achieved": tri_bytes / (min(warm) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": tri_bytes / (min(warm) * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_triangle": 230}}}
    s2c, c2w, _ = pupil.camera()
    prim = rb.camera_rays(s2c, c2w, w, h)
    rec, tuvp, inst, _ = rb.measure_trace(torch, scene, stream, prim, False, 5, peak, "closest_primary_1080p")
    out["primary"] = rec
    pos = rb.hit_points(prim, tuvp[:, 0].cpu().numpy(), inst.cpu().numpy())
    rng = np.random.default_rng(7)
    inco, p = rb.bounce_rays(pos, 1 << 21, rng)
    out["incoherent"] = rb.measure_trace(torch, scene, stream, inco, False, 5, peak, "closest_incoherent_bounce")[0]
    out["any_hit"] = rb.measure_trace(torch, scene, stream, rb.shadow_rays(p, rng), True, 5, peak, "anyhit_shadow_to_light")[0]
    # the same two distributions in renderer-sized launches (16 Mi rays, origins resampled with replacement): a 2^21-ray batch gives
    # each resident lane ~14 rays, so the drain of the persistent kernel is a visible share of a 1.3 ms launch
    big, pb = rb.bounce_rays(pos, 1 << 24, rng)
    out["incoherent_16mi"] = rb.measure_trace(torch, scene, stream, big, False, 3, peak, "closest_incoherent_bounce, 16 Mi rays in one launch")[0]
    del big
    out["any_hit_16mi"] = rb.measure_trace(torch, scene, stream, rb.shadow_rays(pb, rng), True, 3, peak, "anyhit_shadow_to_light, 16 Mi rays in one launch")[0]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cornell", choices=sorted(WORKLOADS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--spp", type=int, default=0, help="samples per pixel per step (default: the workload's)")
    ap.add_argument("--terrain-n", type=int, default=3873, help="terrain grid size n (2*n*n triangles; 3873 -> 30.0 M)")
    ap.add_argument("--cpu-spp", type=int, default=2, help="spp of the CPU sample per step / for the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: every rank renders the step's spp (weak) or the ranks split it (strong)")
    ap.add_argument("--reduce", default="root", choices=["all", "root"],
                    help="N > 1: reduce to rank 0 + finalize (default: measured faster, profiles/README.md), or reduce-scatter + finalize + all-gather")
    ap.add_argument("--builder", type=int, default=0, help="BVH builder (pb2_scene_set_builder): 0 LBVH, 2 SAH-driven clustering")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the c3 / c4 / c5 sub-records")
    ap.add_argument("--opt", action="append", default=[], help="pb2 scene option name=value (tuning experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.workload]
    args.width = args.width or wl.get("width", 1920)
    args.height = args.height or wl.get("height", 1080)
    if args.impl == "reference":
        return run_reference_impl(args)

    import torch
    from pupiloptixlab_b200 import pb2, pupil

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    pupil.init(local, log_level=1)
    if args.builder:
        pupil.set_bvh_builder(args.builder)  # sticky: every scene of this run is built with it
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        box = [pupil.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)  # plumbing: the 128-byte NCCL id of the product's own communicator
        pupil.set_shard(rank, world, box[0], strong=args.scaling == "strong", reduce_mode=pupil.REDUCE_ALL if args.reduce == "all" else pupil.REDUCE_ROOT)
    warmup = max(args.warmup, 3)
    strong = args.scaling == "strong" and world > 1
    spp = args.spp or wl["spp"]               # pass argument: per rank (weak) or for the whole step (strong)
    spp_step_total = spp if (strong or world == 1) else spp * world  # samples per pixel all ranks render in one step
    desc = make_scene(args.workload, args.width, args.height, args.depth, args.terrain_n)
    n_px = args.width * args.height

    stream = torch.cuda.Stream()
    t_load0 = time.perf_counter()
    pupil.load_scene(desc)
    load_s = time.perf_counter() - t_load0
    scene = pupil.scene_handle()
    scene.set_stream(stream.cuda_stream)
    scene.set_option("profiling", 1)  # per-stage CUDA events (a few dozen event records per batch)
    for kv in args.opt:
        k, v = kv.split("=")
        scene.set_option(k, int(v))
    if any(kv.startswith(("collapse", "ploc_radius")) for kv in args.opt):
        scene.build()  # options that change the tree invalidate it
    build = pupil.build_stats()

    def step(i: int):
        # PTPass::OnRunSharded splits the step's seeds over the ranks (pb2_shard_plan) and queues the reduction; N = 1: plain OnRun
        pupil.pass_config(frames_per_run=spp, first_seed=i * spp_step_total)
        pupil.run(1)

    def barrier():
        pupil.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage = dict(generate=0.0, extend=0.0, shade=0.0, shadow=0.0, accumulate=0.0)
    launches = closest = shadow = 0
    n_ext = n_shade = n_shadow = 0
    e0.record(stream)
    for i in range(args.steps):
        step(warmup + i)
        rs = pupil.render_stats()  # waits for the scene's stream (not for the reduction, which runs on the communicator's)
        for k in stage:
            stage[k] += getattr(rs, k + "_ms")
        launches += rs.kernel_launches + 1 if world > 1 else rs.kernel_launches  # + the finalize kernel of the reduction
        closest, shadow = closest + rs.closest_rays, shadow + rs.shadow_rays
        n_ext, n_shade, n_shadow = n_ext + rs.extend_launches, n_shade + rs.shade_launches, n_shadow + rs.shadow_launches
    pupil.synchronize()  # the last step's reduction; then the end event on the launching stream
    e1.record(stream)
    reduce_rec = None
    if world > 1:  # the last reduction of the timed region: nothing ran beside it, so this is the collective + finalize alone
        r_ms, r_bytes = pupil.last_reduction()
        reduce_rec = {"ms": r_ms, "bytes_per_rank": r_bytes, "gbs_per_rank": r_bytes / (r_ms * 1e-3) / 1e9 if r_ms > 0 else None,
                      "collective": "ncclReduce to rank 0 + finalize" if args.reduce == "root" else "ncclReduceScatter + finalize + ncclAllGather",
                      "note": "overlapped with the next step's render in steady state; exposed only after the last step"}
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        r = torch.tensor([float(closest), float(shadow)], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(r)
        closest_all, shadow_all = float(r[0].item()), float(r[1].item())
    else:
        closest_all, shadow_all = float(closest), float(shadow)
    value = n_px * spp_step_total * args.steps / (ms * 1e-3) / 1e6
    mrays = (closest_all + shadow_all) / (ms * 1e-3) / 1e6
    peak, peak_kind = load_peaks()

    # ---- roofline of the dominant kernel: counters from one untimed pass with traversal counting on -------------
    roofline = None
    if rank == 0:
        scene.set_option("counting", 1)
    step(warmup)  # on EVERY rank: a sharded step ends in a collective (pb2_comm_reduce_frames); only rank 0 counts
    cs = pupil.render_stats()
    pupil.synchronize()
    if rank == 0:
        scene.set_option("counting", 0)
        spp_rank = cs.closest_rays and (spp if not strong else max(1, -(-spp // world)))  # frames this rank rendered in the counting pass
        per_step, bvh_cached, nodes_c, prims_c = stage_bytes(cs, build, desc, n_px, spp_rank, n_ext / args.steps, n_shadow / args.steps)
        dom = max(("extend", "shade", "shadow"), key=lambda k: stage[k])
        dom_ms = stage[dom] / args.steps
        achieved = per_step[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic, traffic_note = measured_traffic(args.workload, dom)
        n_l = {"extend": n_ext, "shade": n_shade, "shadow": n_shadow}[dom] / args.steps
        roofline = {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_note,
                    "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback (B200_PROFILING.md)",
                    "algorithmic_bytes_per_launch": per_step[dom] / max(n_l, 1), "launches_per_step": n_l, "avg_launch_ms": dom_ms / max(n_l, 1),
                    "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
                    "nodes_per_closest_ray": nodes_c / max(cs.closest_rays, 1), "prims_per_closest_ray": prims_c / max(cs.closest_rays, 1),
                    "all_stage_gbs": {k: per_step[k] / (stage[k] / args.steps * 1e-3) / 1e9 if stage[k] > 0 else 0.0 for k in per_step}}
        roofline["all_stage_frac"] = {k: v / peak for k, v in roofline["all_stage_gbs"].items()}
        if bvh_cached and dom in ("extend", "shadow"):
            roofline["note"] = (f"the {build.bvh_bytes}-byte BVH is cache resident, so this traversal kernel is bound by instruction issue, not by HBM; "
                                "its HBM bytes are the ray and hit records only.  The HBM-bound kernel of the step is k_shade: see all_stage_frac; the "
                                "memory-bound traversal numbers are in the c4 record")

    # ---- e2e: scene from HOST data every step, image back to the host -----------------------------------------------
    e2e = None
    if not args.no_e2e:
        pin_mesh_arrays(torch, desc)
        t0 = 0.0
        pinned = None
        for i in range(-1, args.steps):  # iteration -1 is an untimed warm-up of the reload path
            if i == 0:
                barrier()
                t0 = time.perf_counter()
            t_a = time.perf_counter()
            pupil.load_scene(desc)  # XML text -> loader -> H2D -> BVH build
            t_b = time.perf_counter()
            scene = pupil.scene_handle()
            scene.set_stream(stream.cuda_stream)
            step(warmup + args.steps + 1 + i)
            t_c = time.perf_counter()
            if rank == 0:
                if pinned is None:
                    pinned = torch.empty(n_px * 4, dtype=torch.float32).pin_memory()
                pupil.buffer_into("final result", pinned.data_ptr(), n_px * 16)  # waits for render + reduction; D2H of the float4 frame into pinned host memory
                img = pinned.numpy().reshape(args.height, args.width, 4)
                assert np.isfinite(img[0, 0, :3]).all()
                print(f"[e2e step {i}] load {1e3 * (t_b - t_a):.1f} ms, render {1e3 * (t_c - t_b):.1f} ms, download {1e3 * (time.perf_counter() - t_c):.1f} ms", file=sys.stderr)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n_px * spp_step_total * args.steps / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": scene_h2d_bytes(desc), "d2h_bytes_per_step": n_px * 16,
               "ms_per_step": 1e3 * dt / args.steps}

    # ---- CPU baseline (rank 0, N = 1 only) --------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sc, kind, cores = cpu_arm(args, desc, args.cpu_spp)
        t0 = time.perf_counter()
        out = sc.render(args.cpu_spp, first_seed=0, threads=cores)
        dt = time.perf_counter() - t0
        cpu = {"value": n_px * args.cpu_spp / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
               "sample": f"{args.cpu_spp} spp of the same {args.width}x{args.height} depth-{args.depth} frame ({dt:.1f} s)",
               "mrays_per_s": (out["closest_rays"] + out["shadow_rays"]) / dt / 1e6}
        del sc

    # ---- the other configs of BASELINE.json, same run ---------------------------------------------------------------
    sub = {}
    if not args.no_sub:
        def guarded(name, fn):
            try:
                sub[name] = fn()
            except Exception as e:  # noqa: BLE001 — a sub-record must not cost the main line
                sub[name] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            guarded("c3", lambda: c3_record(torch, pupil, stream, args, peak))
            guarded("c4", lambda: c4_record(torch, pupil, stream, args, peak))
        # c5: the fixed 4K 1024-spp job on `world` GPUs (strong scaling); c2_strong: the headline job (1080p, 64 spp) split the same way
        def strong_job(label, maker, w, h, job_spp, steps):
            d = maker(w, h, args.depth)
            pupil.load_scene(d)
            sc = pupil.scene_handle()
            sc.set_stream(stream.cuda_stream)
            barrier()
            ms_step, _, tot = timed_render(torch, pupil, stream, w * h, job_spp, steps=steps, warm=1)
            if world > 1:
                t = torch.tensor([ms_step], device=f"cuda:{local}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_step = float(t.item())
            return {"workload": label, "n_gpus": world, "scaling": "strong", "spp_total": job_spp, "spp_per_gpu": -(-job_spp // world), "ms_per_step": ms_step,
                    "value": w * h * job_spp / (ms_step * 1e-3) / 1e6, "unit": "Msamples/s", "steps": steps,
                    "reduce": "none" if world == 1 else ("reduce-scatter + finalize + all-gather" if args.reduce == "all" else "reduce to rank 0 + finalize")}
        from pupiloptixlab_b200 import scenes as _sc
        if world > 1 and not strong:  # the sub-records split a FIXED job: switch the pass to the strong plan (same communicator)
            pupil.set_shard_plan(strong=True)
        guarded("c5", lambda: strong_job("material-ball grid 3840x2160, 1024 spp in total, max depth 8 (BASELINE.json configs[4])", _sc.material_grid, 3840, 2160, 1024, 1 if world == 1 else 2))
        if world > 1:
            guarded("c2_strong", lambda: strong_job("Cornell box 1920x1080, 64 spp in total, max depth 8 (configs[1] as a fixed job)", _sc.cornell_box, 1920, 1080, 64, 5))

    if rank == 0:
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"] if not args.spp else f"{args.workload} {args.width}x{args.height}, {spp} spp, max depth {args.depth}",
                       "width": args.width, "height": args.height, "max_depth": args.depth, "spp_per_step": spp_step_total, "spp_per_gpu_per_step": spp_step_total / world,
                       "triangles": desc.num_triangles(),
                       "sharding": (f"sample-index (seed = base + rank + k*{world}) inside the product: PTPass::SetShard -> pb2_shard_plan, plain sums, "
                                    f"pb2_comm_reduce_frames ({'ncclReduceScatter + finalize + ncclAllGather' if args.reduce == 'all' else 'ncclReduce to rank 0 + finalize'}) "
                                    "on its own stream, overlapped with the next step's render") if world > 1 else "none",
                       "l2": "path-state working set per batch (up to 128 Mi paths, ~18.8 GB; here one batch of 64 frames) and accumulation buffers exceed the 126 MB L2; no explicit flush"},
            "mrays_per_s": mrays, "rays_per_sample": (closest_all + shadow_all) / (n_px * spp_step_total * args.steps),
            "bvh": {"build_ms": build.build_ms, "n_prims": build.n_prims, "n_nodes": build.n_nodes, "bytes": build.bvh_bytes, "sah_cost": build.sah_cost},
            "scene_load_s": load_s, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        }
        if reduce_rec:
            line["reduce"] = reduce_rec
        line.update(sub)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    pupil.shutdown()


if __name__ == "__main__":
    main()
