"""The three ray batches of config C4, shared by tools/bench_traversal.py, bench.py's `c4` record and the
-m gpu parity test at the config's scale (tests/test_gpu_config_scale.py):
  (i)   primary camera rays of the frame, one per pixel, jittered
  (ii)  incoherent cosine-bounce rays leaving the primary hit points, shuffled
  (iii) shadow rays from the hit points towards the area light
Rays are (n, 8) float32: ox oy oz tmin dx dy dz tmax (the layout of pb2_trace_closest / pb2_trace_any)."""
import numpy as np


def camera_rays(s2c, c2w, w, h, seed=0):
    rng = np.random.default_rng(seed)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    u = (xs + rng.random((h, w), dtype=np.float32)) / np.float32(w)
    v = (ys + rng.random((h, w), dtype=np.float32)) / np.float32(h)
    pf = np.stack([u, v, np.zeros_like(u), np.ones_like(u)], -1).reshape(-1, 4)
    d = pf @ s2c.T
    d = d[:, :3] / d[:, 3:4]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dw = d @ c2w[:3, :3].T
    dw /= np.linalg.norm(dw, axis=1, keepdims=True)
    rays = np.zeros((w * h, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = c2w[:3, 3], 1e-3, dw, 1e16
    return rays


def hit_points(rays, t, inst):
    hit = inst >= 0
    return rays[hit, 0:3] + t[hit, None] * rays[hit, 4:7]


def bounce_rays(pos, k, rng, shuffle=True):
    """cosine-distributed directions about +Y (the terrain is a height field) from k of the points"""
    sel = rng.choice(len(pos), k, replace=len(pos) < k)
    p = pos[sel]
    u1, u2 = rng.random(k, dtype=np.float32), rng.random(k, dtype=np.float32)
    r, phi = np.sqrt(u1), 2 * np.pi * u2
    d = np.stack([r * np.cos(phi), np.sqrt(np.maximum(0, 1 - u1)), r * np.sin(phi)], -1).astype(np.float32)
    rays = np.zeros((k, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = p, 1e-3, d, 1e16
    if shuffle:
        rays = rays[rng.permutation(k)]
    return rays, p


def shadow_rays(p, rng, light_centre=(0.0, 12.0, 0.0), light_half=3.0):
    k = len(p)
    light = np.array(light_centre, np.float32) + rng.uniform(-light_half, light_half, (k, 3)).astype(np.float32) * np.array([1, 0, 1], np.float32)
    dl = light - p
    dist = np.linalg.norm(dl, axis=1, keepdims=True)
    rays = np.zeros((k, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = p, 1e-4, dl / dist, dist[:, 0] - 1e-4
    return rays


def measure_trace(torch, scene, stream, rays_np, any_hit, reps, peak_gbs, label):
    """Times `reps` launches of pb2_trace_closest_dev / pb2_trace_any_dev over rays resident in HBM (CUDA events on the
    launching stream, 3 warm-up launches), then one counting launch for nodes / primitives per ray.  Returns the JSON record
    and the device result tensors (t u v prim | instance | occluded).  Algorithmic bytes per ray: 32 in +
    20 (closest) or 4 (any) out + 80 per node visited + 48 per primitive tested."""
    n = len(rays_np)
    rays = torch.from_numpy(np.ascontiguousarray(rays_np)).cuda()
    tuvp = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    inst = torch.zeros(n, dtype=torch.int32, device="cuda")
    occ = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def launch():
        if any_hit:
            scene.trace_any_dev(rays.data_ptr(), n, occ.data_ptr())
        else:
            scene.trace_closest_dev(rays.data_ptr(), n, tuvp.data_ptr(), inst.data_ptr())
    for _ in range(3):
        launch()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            launch()
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    scene.set_option("counting", 1)
    launch()
    scene.synchronize()
    st = scene.render_stats_raw()
    scene.set_option("counting", 0)
    nodes, prims = st.nodes_visited, st.prims_tested
    bytes_ = n * (32 + (4 if any_hit else 20)) + nodes * 80 + prims * 48
    hits = int((occ != 0).sum().item()) if any_hit else int((inst >= 0).sum().item())
    rec = {"what": label, "rays": n, "ms": ms, "mrays_per_s": n / ms / 1e3, "hit_fraction": hits / n, "nodes_per_ray": nodes / n, "prims_per_ray": prims / n,
           "roofline": {"bound": "hbm", "achieved": bytes_ / (ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": bytes_ / (ms * 1e-3) / 1e9 / peak_gbs,
                        "algorithmic_bytes_per_ray": bytes_ / n}}
    return rec, tuvp, inst, occ
