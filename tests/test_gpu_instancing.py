"""-m gpu: the two-level acceleration structure (bvh_build.cu / traverse.cuh) — one bottom-level tree per shared mesh, an
instance node per placement (GASManager::RefGAS + IAS of the reference, framework/world/gas_manager.cpp:10,
ias_manager.cpp:29-151) — against the oracle, which flattens every placement to world-space triangles.

A placement is intersected in OBJECT space (the ray goes through the fp32 inverse transform, as OptiX does for an IAS), the
oracle intersects world-space triangles: the two differ by rounding that grows as 1 / cos of the incidence angle.  Ids must
match exactly up to ties (north_star); t is held to 1e-5 relative on 99 % of the rays, and on EVERY ray the hit point may
be off by at most 2e-5 (relative to its distance) measured perpendicular to the surface, |dt| * cos <= 2e-5 * max(t, 1) —
the flattened build of the same scene (instancing = 0) keeps the plain 1e-5 bar of tests/test_gpu_traversal.py."""
import numpy as np
import pytest

import orc
from gpu_util import compare_hits, pb2_scene_from_oracle, random_rays, random_soup
from pupiloptixlab_b200 import pb2, pupil, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    pb2.init(0)


def _check_t(rays, a, b, meshes_of_inst, xforms):
    """a, b: hit records with equal ids; meshes_of_inst[i] = (positions, indices) or None; xforms[i] = 3x4 object -> world"""
    both = (a["inst"] >= 0) & (a["inst"] == b["inst"]) & (a["prim"] == b["prim"])
    dt = np.abs(a["t"].astype(np.float64) - b["t"])
    tt = np.maximum(np.abs(b["t"].astype(np.float64)), 1.0)
    assert np.mean(dt[both] <= 1e-5 * tt[both]) > 0.99
    cos = np.ones(len(rays))
    for i in np.flatnonzero(both & (dt > 1e-5 * tt)):  # the few that exceed the plain bar: grazing incidence?
        m = meshes_of_inst[a["inst"][i]]
        if m is None:
            continue
        P, I = m
        x = np.asarray(xforms[a["inst"][i]], np.float64)
        tri = P[I[a["prim"][i]]].astype(np.float64) @ x[:, :3].T + x[:, 3]
        n = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        cos[i] = abs(n @ rays[i, 4:7].astype(np.float64)) / max(np.linalg.norm(n), 1e-300)
    bad = both & (dt * cos > 2e-5 * tt)
    assert not bad.any(), (np.count_nonzero(bad), dt[bad][:4], cos[bad][:4], b["t"][bad][:4])


def _placements(rng, k, spread=9.0):
    out = []
    for _ in range(k):
        out.append(scenes.Xf("srt", scale=tuple(rng.uniform(0.4, 1.6, 3)), rotate_axis=tuple(rng.normal(size=3)), rotate_angle=float(rng.uniform(0, 360)),
                             translate=tuple(rng.uniform(-spread, spread, 3))))
    return out


def _instanced_desc(n_tris=4000, k=7, seed=5):
    rng = np.random.default_rng(seed)
    mesh = random_soup(n_tris, seed, extent=2.0, size=0.25)
    shapes = [scenes.Shape("obj", xf, mesh=mesh) for xf in _placements(rng, k)]        # one mesh, k placements -> one bottom-level tree
    shapes.append(scenes.Shape("obj", scenes.Xf("srt", translate=(0.0, -6.0, 0.0)), mesh=random_soup(900, seed + 1, extent=8.0, size=0.8)))  # used once: flattened
    shapes.append(scenes.Shape("rectangle", scenes.Xf("srt", scale=(12.0, 12.0, 1.0), rotate_axis=(1, 0, 0), rotate_angle=-90.0, translate=(0, -9.0, 0))))
    for j in range(3):
        shapes.append(scenes.Shape("sphere", scenes.Xf("srt", scale=(1.0, 0.7, 1.2)), center=tuple(rng.uniform(-7, 7, 3)), radius=float(rng.uniform(0.5, 1.5))))
    return scenes.SceneDesc(shapes=shapes)


def _pb2_scene_shared_meshes(desc, osc):
    """like gpu_util.pb2_scene_from_oracle, but placements that share a mesh array share ONE pb2 mesh"""
    from pupiloptixlab_b200.host_py import builtin_mesh
    s = pb2.Scene()
    cache = {}
    for i, sh in enumerate(desc.shapes):
        xf = osc.instance_xform(i)[:3]
        if sh.type == "sphere":
            s.add_instance(pb2.MESH_SPHERE, xf)
            continue
        key = id(sh.mesh) if sh.type == "obj" else sh.type
        if key not in cache:
            m = sh.mesh if sh.type == "obj" else builtin_mesh(sh.type)
            cache[key] = s.add_mesh(m["positions"], m["indices"], m.get("normals"), m.get("texcoords"))
        s.add_instance(cache[key], xf)
    return s


@pytest.mark.parametrize("coop", [0, 1])
def test_instanced_scene_matches_the_flattened_oracle(port_lib, coop):
    desc = _instanced_desc()
    osc = orc.OracleScene(port_lib, desc)
    s = _pb2_scene_shared_meshes(desc, osc)
    st = s.build()
    assert st.n_blas == 1 and st.n_instance_leaves == 7
    assert st.n_triangles == 4000 + 900 + 2 and st.n_spheres == 3  # the shared mesh counts once
    s.set_option("coop_prims", coop)
    rays = random_rays(30000, 3, extent=13.0)
    ref, _ = osc.trace_closest(rays, brute=True)
    gpu = s.trace_closest(rays)
    # ids: equal, or two different primitives at (nearly) the same distance; never a hit against a miss (t of equal ids is checked
    # geometrically below)
    same_id = (gpu["inst"] == ref["inst"]) & (gpu["prim"] == ref["prim"])
    both_hit = (gpu["inst"] >= 0) & (ref["inst"] >= 0)
    tie = ~same_id & both_hit & (np.abs(gpu["t"] - ref["t"]) <= 2e-4 * np.maximum(np.abs(ref["t"]), 1e-3))
    assert not np.any(~same_id & ~tie), np.count_nonzero(~same_id & ~tie)
    assert np.count_nonzero(tie) <= len(rays) // 200
    from pupiloptixlab_b200.host_py import builtin_mesh
    meshes = [(sh.mesh["positions"], sh.mesh["indices"]) if sh.type == "obj" else (builtin_mesh(sh.type)["positions"], builtin_mesh(sh.type)["indices"]) if sh.type != "sphere" else None
              for sh in desc.shapes]
    _check_t(rays, gpu, ref, meshes, [osc.instance_xform(i)[:3] for i in range(len(desc.shapes))])
    inst_hits = np.count_nonzero((ref["inst"] >= 0) & (ref["inst"] < 7))
    assert inst_hits > 1500, inst_hits  # the instanced placements are actually hit
    # the same loop with the seven placements intersected in OBJECT space, as the instance nodes do (the ray through the
    # fp32 inverse of the placement, the mesh's own triangles): every hit bit for bit, ids without any tie allowance
    flags = np.zeros(len(desc.shapes), np.uint8)
    flags[:7] = 1
    obj = osc.trace_closest_objspace(rays, flags)
    assert np.array_equal(gpu["inst"], obj["inst"]) and np.array_equal(gpu["prim"], obj["prim"])
    hit = obj["inst"] >= 0
    for f in ("t", "u", "v"):
        assert np.array_equal(gpu[f][hit].view(np.uint32), obj[f][hit].view(np.uint32)), f
    m = (gpu["inst"] == ref["inst"]) & (gpu["prim"] == ref["prim"]) & (ref["inst"] >= 0)
    assert np.allclose(gpu["u"][m], ref["u"][m], atol=5e-4) and np.allclose(gpu["v"][m], ref["v"][m], atol=5e-4)
    rays[:, 7] = np.random.default_rng(1).uniform(0.5, 25.0, len(rays)).astype(np.float32)
    occ, ref_occ = s.trace_any(rays), osc.trace_any(rays, brute=True)
    assert np.count_nonzero(occ != ref_occ) <= 6
    # flattening the same scene (instancing = 0) gives the same ids
    s.set_option("instancing", 0)
    st0 = s.build()
    assert st0.n_blas == 0 and st0.n_triangles == 7 * 4000 + 900 + 2
    rays[:, 7] = 1e16
    flat = s.trace_closest(rays)
    same = (flat["inst"] == gpu["inst"]) & (flat["prim"] == gpu["prim"])
    assert np.count_nonzero(~same) <= len(rays) // 100


def test_a_thousand_placements_cost_one_tree(port_lib):
    """1000 instances of one 20 000-triangle mesh: the structure holds the mesh once (RefGAS), flattening holds it 1000 times"""
    rng = np.random.default_rng(11)
    mesh = random_soup(20000, 4, extent=1.0, size=0.08)
    s = pb2.Scene()
    mid = s.add_mesh(mesh["positions"], mesh["indices"])
    xfs = []
    for k in range(1000):
        a = rng.uniform(0, 2 * np.pi)
        c, sn = np.cos(a), np.sin(a)
        sc = rng.uniform(0.5, 1.5)
        m = np.array([[c * sc, 0, sn * sc, 0], [0, sc, 0, 0], [-sn * sc, 0, c * sc, 0]], np.float32)
        m[:, 3] = rng.uniform(-40, 40, 3)
        xfs.append(m)
        s.add_instance(mid, m)
    st = s.build()
    assert st.n_blas == 1 and st.n_instance_leaves == 1000 and st.n_triangles == 20000
    one_tree_bytes = st.bvh_bytes
    assert one_tree_bytes < 3_000_000  # ~20 000 x 48 B of records + ~3 000 nodes x 80 B + 1000 instance nodes and their top level
    rays = random_rays(20000, 9, extent=45.0)
    hits = s.trace_closest(rays)
    assert np.count_nonzero(hits["inst"] >= 0) > 500
    # the same hits as the flattened scene (which needs ~1000 times the memory)
    s.set_option("instancing", 0)
    st0 = s.build()
    assert st0.n_blas == 0 and st0.n_triangles == 20_000_000 and st0.bvh_bytes > 300 * one_tree_bytes
    flat = s.trace_closest(rays)
    same = (flat["inst"] == hits["inst"]) & (flat["prim"] == hits["prim"])
    assert np.count_nonzero(~same) <= 40
    _check_t(rays, hits, flat, [(mesh["positions"], mesh["indices"])] * 1000, xfs)


def test_transform_edit_rebuilds_only_the_top_level(port_lib):
    """IAS::Update: after pb2_scene_set_instance_transform the bottom-level trees stay and the rebuild takes well under a
    millisecond of device time, for a mesh of 2 M triangles; hits are those of a scene built with the new transform"""
    mesh = scenes.heightfield_mesh(1000)  # 2.0 M triangles
    desc = scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt", translate=(0.0, 0.0, 0.0)), mesh=mesh),
                                    scenes.Shape("cube", scenes.Xf("srt", scale=(0.5, 0.5, 0.5), translate=(0.0, 4.0, 0.0)))])
    s = pb2.Scene()
    s.set_option("instancing", 2)  # every mesh gets a bottom-level tree, also the one placed once
    mid = s.add_mesh(mesh["positions"], mesh["indices"], mesh["normals"], mesh["texcoords"])
    from pupiloptixlab_b200.host_py import builtin_mesh
    cube = builtin_mesh("cube")
    cid = s.add_mesh(cube["positions"], cube["indices"], cube["normals"], cube["texcoords"])
    ident = np.eye(4, dtype=np.float32)[:3]
    s.add_instance(mid, ident)
    cm = ident.copy()
    cm[:, :3] *= 0.5
    cm[1, 3] = 4.0
    s.add_instance(cid, cm)
    st = s.build()
    assert st.n_blas == 1 and st.n_instance_leaves == 1  # the 12-triangle cube stays flattened (below the 64-triangle floor)
    full_ms = st.build_ms
    moved = np.array([[0.8, 0, 0.6, 1.5], [0, 1, 0, 0.25], [-0.6, 0, 0.8, -2.0]], np.float32)
    s.set_instance_transform(0, moved)
    st2 = s.build()
    assert st2.n_blas == 1 and st2.top_level_ms < 1.0 and st2.build_ms < 1.0, (st2.top_level_ms, st2.build_ms)
    assert st2.build_ms < full_ms / 3
    rays = random_rays(20000, 2, extent=9.0)
    rays[:, 1] = np.abs(rays[:, 1]) + 3.0  # from above
    rays[:, 5] = -np.abs(rays[:, 5]) - 0.2
    rays[:, 4:7] /= np.linalg.norm(rays[:, 4:7], axis=1, keepdims=True)
    got = s.trace_closest(rays)
    fresh = pb2.Scene()
    fresh.set_option("instancing", 0)
    m2 = fresh.add_mesh(mesh["positions"], mesh["indices"])
    c2 = fresh.add_mesh(cube["positions"], cube["indices"])
    fresh.add_instance(m2, moved)
    fresh.add_instance(c2, cm)
    fresh.build()
    want = fresh.trace_closest(rays)
    same = (got["inst"] == want["inst"]) & (got["prim"] == want["prim"])
    assert np.count_nonzero(~same) <= 40 and np.count_nonzero(want["inst"] == 0) > 5000
    _check_t(rays, got, want, [(mesh["positions"], mesh["indices"]), (cube["positions"], cube["indices"])], [moved, cm])


def test_instanced_render_matches_the_oracle(port_lib):
    """the whole path on a scene whose mesh is placed several times: System -> World (shared shape -> one pb2 mesh) -> PTPass"""
    rng = np.random.default_rng(3)
    d = scenes.material_grid(160, 90, 6, nx=2, nz=1)
    bump = scenes.heightfield_mesh(24, seed=7, size=1.2, amplitude=0.15)
    for k, xf in enumerate(_placements(rng, 5, spread=2.0)):
        xf.translate = (xf.translate[0], 0.6 + 0.3 * k, xf.translate[2])
        d.shapes.append(scenes.Shape("obj", xf, scenes.Bsdf("diffuse", params=dict(reflectance=(0.3 + 0.1 * k, 0.5, 0.6))), mesh=bump, name=f"bump{k}"))
    pupil.init(0)
    try:
        pupil.load_scene(d)
        st = pupil.build_stats()
        assert st.n_blas == 1 and st.n_instance_leaves == 5
        pupil.pass_config()
        pupil.run(1)
        ref = orc.OracleScene(port_lib, d).render(1)
        assert np.array_equal(pupil.buffer("test").reshape(-1), ref["test"])
        frame = pupil.buffer("final result")
        g, r = frame.reshape(-1, 4)[:, :3].astype(np.float64), ref["frame"][:, :3].astype(np.float64)
        ok = (np.abs(g - r) <= 1e-4 * np.maximum(1.0, np.abs(r))).all(1)
        assert ok.mean() >= 0.97, ok.mean()
        assert abs(g.mean() - r.mean()) <= 0.01 * r.mean()
        # moving one placement through the host surface: top level only, same image as a scene loaded that way
        target = len(d.shapes) - 1
        d2 = scenes.material_grid(160, 90, 6, nx=2, nz=1)
        d2.shapes = d.shapes[:]
        import copy
        moved_shape = copy.copy(d.shapes[target])
        moved_shape.to_world = scenes.Xf("srt", scale=(1.2, 0.8, 1.1), rotate_axis=(0, 1, 0), rotate_angle=50.0, translate=(0.3, 1.4, -0.2))
        d2.shapes[target] = moved_shape
        o2 = orc.OracleScene(port_lib, d2)
        pupil.set_instance_transform(target, o2.instance_xform(target))
        pupil.run(1)
        st2 = pupil.build_stats()
        assert st2.n_blas == 1 and st2.build_ms < 1.0, st2.build_ms
        got = pupil.buffer("final result").copy()
        pupil.load_scene(d2)
        pupil.pass_config()
        pupil.run(1)
        assert np.allclose(got, pupil.buffer("final result"), rtol=1e-5, atol=1e-6)
    finally:
        pupil.shutdown()


def test_instanced_emitters_and_mirrored_placements_render_like_the_oracle(port_lib):
    """an EMISSIVE mesh placed twice (one placement mirrored: negative determinant) behind instance nodes: hits inside a
    bottom-level tree must find the emitter entries of the right placement (emitter offset + primitive index), next-event
    estimation must find both, and a mirrored instance must shade like its flattened twin"""
    d = scenes.cornell_box(128, 128, 6)
    panel = scenes.heightfield_mesh(8, seed=3, size=0.5, amplitude=0.02)  # 128 triangles: above the bottom-level floor of 64
    d.shapes.append(scenes.Shape("obj", scenes.Xf("srt", scale=(1.0, 1.0, 1.0), rotate_axis=(1, 0, 0), rotate_angle=180.0, translate=(-0.4, 1.6, 0.1)),
                                 scenes.Bsdf("diffuse", params=dict(reflectance=(0.1, 0.1, 0.1))), mesh=panel, emitter=(6.0, 3.0, 1.5), name="panel_a"))
    d.shapes.append(scenes.Shape("obj", scenes.Xf("srt", scale=(-1.2, 1.0, 0.8), rotate_axis=(1, 0, 0), rotate_angle=180.0, translate=(0.45, 1.5, -0.2)),
                                 scenes.Bsdf("diffuse", params=dict(reflectance=(0.1, 0.1, 0.1))), mesh=panel, emitter=(1.0, 4.0, 6.0), name="panel_b"))
    pupil.init(0)
    try:
        pupil.load_scene(d)
        st = pupil.build_stats()
        assert st.n_blas == 1 and st.n_instance_leaves == 2
        areas, _ = pupil.emitters()
        assert len(areas) == 2 + 2 * 128
        pupil.pass_config(frames_per_run=4)
        pupil.run(1)
        got = pupil.buffer("pt accum buffer")[..., :3].reshape(-1, 3).astype(np.float64)
        ref = orc.OracleScene(port_lib, d).render(4)["accum"][:, :3].astype(np.float64)
        ok = (np.abs(got - ref) <= 1e-4 * np.maximum(1.0, np.abs(ref))).all(1)
        assert ok.mean() >= 0.985, ok.mean()
        assert abs(got.mean() - ref.mean()) <= 0.01 * ref.mean()
        # and the flattened build of the same scene gives the same picture up to the object-space rounding of the hits
        pupil.set_instancing(0)
        pupil.load_scene(d)
        assert pupil.build_stats().n_blas == 0
        pupil.pass_config(frames_per_run=4)
        pupil.run(1)
        flat = pupil.buffer("pt accum buffer")[..., :3].reshape(-1, 3).astype(np.float64)
        close = (np.abs(got - flat) <= 1e-4 * np.maximum(1.0, np.abs(flat))).all(1)
        assert close.mean() >= 0.985, close.mean()
    finally:
        pupil.set_instancing(1)
        pupil.shutdown()
