"""Helpers shared by the -m gpu tests: build a pb2 scene straight from a SceneDesc *using the oracle's
resolved instance matrices* (only for traversal-level tests that bypass the host library)."""
import numpy as np

from pupiloptixlab_b200 import pb2


def random_soup(n_tris, seed, extent=10.0, size=0.6):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, size=(n_tris, 1, 3))
    v = c + rng.normal(scale=size, size=(n_tris, 3, 3))
    pos = v.reshape(-1, 3).astype(np.float32)
    idx = np.arange(n_tris * 3, dtype=np.uint32).reshape(-1, 3)
    return dict(positions=pos, indices=idx)


def random_rays(n, seed, extent=12.0, tmin=1e-3, tmax=1e16):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-extent, extent, size=(n, 3))
    tgt = rng.uniform(-extent * 0.8, extent * 0.8, size=(n, 3))
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, tmin, d, tmax
    return rays


def pb2_scene_from_oracle(desc, oscene):
    """geometry only: meshes + instances with the oracle's resolved object->world matrices"""
    from orc import OracleScene  # noqa: F401
    from pupiloptixlab_b200.host_py import builtin_mesh
    s = pb2.Scene()
    cache = {}
    for i, sh in enumerate(desc.shapes):
        xf = oscene.instance_xform(i)[:3]
        if sh.type == "sphere":
            s.add_instance(pb2.MESH_SPHERE, xf)
            continue
        if sh.type == "obj":
            mid = s.add_mesh(sh.mesh["positions"], sh.mesh["indices"], sh.mesh.get("normals"), sh.mesh.get("texcoords"))
        else:
            if sh.type not in cache:
                m = builtin_mesh(sh.type)
                cache[sh.type] = s.add_mesh(m["positions"], m["indices"], m["normals"], m["texcoords"])
            mid = cache[sh.type]
        s.add_instance(mid, xf)
    return s


def compare_hits(gpu, ref, rays, rel=1e-5, oracle=None):
    """IDs must match exactly except where the two candidates are at the same distance (exact-t ties /
    hits within `rel` of each other); t within `rel` relative.  Returns the number of tie-excused rays."""
    assert gpu.shape == ref.shape
    miss_g, miss_r = gpu["inst"] < 0, ref["inst"] < 0
    both = ~miss_g & ~miss_r
    same_id = both & (gpu["inst"] == ref["inst"]) & (gpu["prim"] == ref["prim"])
    dt = np.abs(gpu["t"] - ref["t"]) <= rel * np.maximum(np.abs(ref["t"]), 1e-3)
    assert np.all(dt[same_id]), f"t mismatch on {np.count_nonzero(~dt & same_id)} rays with equal ids"
    diff = ~same_id & ~(miss_g & miss_r)
    # a differing id is excused only when both found a hit at (nearly) the same distance
    excused = diff & both & dt
    bad = diff & ~excused
    assert not np.any(bad), (f"{np.count_nonzero(bad)} of {len(gpu)} rays disagree beyond ties; first: "
                             f"{np.flatnonzero(bad)[:5]} gpu={gpu[bad][:3]} ref={ref[bad][:3]}")
    return int(np.count_nonzero(excused))
