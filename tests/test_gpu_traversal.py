"""-m gpu: BVH build + closest-hit / any-hit traversal through the C ABI against the oracle's exhaustive
ray/primitive loop (north_star: primitive ids bit-exact except exact-t ties, t within 1e-5 relative)."""
import numpy as np
import pytest

import orc
from gpu_util import compare_hits, pb2_scene_from_oracle, random_rays, random_soup
from pupiloptixlab_b200 import pb2, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    pb2.init(0)


def _soup_desc(n_tris, seed, n_spheres=0):
    shapes = [scenes.Shape("obj", scenes.Xf("srt", scale=(1.0, 1.5, 0.75), rotate_axis=(0.3, 1, 0.2), rotate_angle=33.0, translate=(0.5, -1.0, 2.0)),
                           mesh=random_soup(n_tris, seed))]
    rng = np.random.default_rng(seed + 1)
    for k in range(n_spheres):
        shapes.append(scenes.Shape("sphere", scenes.Xf("srt", scale=(1.0, 0.6 + 0.1 * k, 1.3)), center=tuple(rng.uniform(-8, 8, 3)), radius=float(rng.uniform(0.3, 2.0))))
    return scenes.SceneDesc(shapes=shapes)


@pytest.mark.parametrize("n_tris,n_spheres,builder", [(1, 0, 0), (2, 0, 0), (3, 1, 0), (4, 0, 0), (37, 3, 0), (1000, 5, 0), (20000, 16, 0), (20000, 16, 1),
                                                      (1, 0, 2), (2, 0, 2), (3, 1, 2), (37, 3, 2), (1000, 5, 2), (20000, 16, 2)])
def test_closest_hit_matches_brute_force(port_lib, n_tris, n_spheres, builder):
    desc = _soup_desc(n_tris, 7 + n_tris, n_spheres)
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    s.set_builder(builder)
    st = s.build()
    assert st.n_prims == n_tris + n_spheres and st.n_nodes >= 1
    rays = random_rays(20000 if n_tris <= 1000 else 6000, 99)
    ref, _ = osc.trace_closest(rays, brute=True)
    gpu = s.trace_closest(rays)
    ties = compare_hits(gpu, ref, rays)
    assert ties <= len(rays) // 200
    if n_tris >= 1000:
        assert np.count_nonzero(ref["inst"] >= 0) > len(rays) // 20  # the batch actually hits things
    # barycentrics agree where ids agree
    m = (gpu["inst"] == ref["inst"]) & (gpu["prim"] == ref["prim"]) & (ref["inst"] >= 0)
    assert np.allclose(gpu["u"][m], ref["u"][m], atol=2e-4) and np.allclose(gpu["v"][m], ref["v"][m], atol=2e-4)


@pytest.mark.parametrize("coop", [0, 1])
def test_both_primitive_test_schedules_give_the_same_hits(port_lib, coop):
    """per-lane and warp-cooperative primitive tests (traverse.cuh) are two schedules of the same arithmetic"""
    desc = _soup_desc(20000, 11, 16)
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    s.build()
    s.set_option("coop_prims", coop)
    rays = random_rays(8000, 21)
    ref, _ = osc.trace_closest(rays, brute=True)
    gpu = s.trace_closest(rays)
    compare_hits(gpu, ref, rays)
    s.set_option("coop_prims", 1 - coop)
    other = s.trace_closest(rays)
    same = (gpu["inst"] == other["inst"]) & (gpu["prim"] == other["prim"])
    assert np.count_nonzero(~same) <= len(rays) // 200  # exact-t ties may resolve differently
    assert np.array_equal(gpu["t"][same], other["t"][same]) and np.array_equal(gpu["u"][same], other["u"][same])
    rays[:, 7] = np.random.default_rng(2).uniform(0.5, 20.0, len(rays)).astype(np.float32)
    a = s.trace_any(rays)
    s.set_option("coop_prims", coop)
    assert np.array_equal(a, s.trace_any(rays))
    # tiny scene, forced cooperative: every lane holds primitives at once
    small = scenes.cornell_box(64, 64, 8)
    o2 = orc.OracleScene(port_lib, small)
    s2 = pb2_scene_from_oracle(small, o2)
    s2.build()
    s2.set_option("coop_prims", coop)
    r2 = o2.camera_rays(seed=5)
    compare_hits(s2.trace_closest(r2), o2.trace_closest(r2, brute=True)[0], r2)


def test_any_hit_matches_brute_force(port_lib):
    desc = _soup_desc(5000, 3, 4)
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    s.build()
    rays = random_rays(20000, 5)
    rays[:, 7] = np.random.default_rng(1).uniform(0.5, 20.0, len(rays)).astype(np.float32)  # finite tmax
    ref = osc.trace_any(rays, brute=True)
    gpu = s.trace_any(rays)
    # an any-hit answer may differ only where the occluder sits at the very end of the interval
    diff = np.flatnonzero(ref != gpu)
    closest, _ = osc.trace_closest(rays[diff], brute=True)
    assert len(diff) <= 4, len(diff)
    for k, i in enumerate(diff):
        t = closest["t"][k]
        assert abs(t - rays[i, 7]) < 1e-4 * max(1.0, t) or abs(t - rays[i, 3]) < 1e-4


def test_empty_scene_and_clear():
    s = pb2.Scene()
    st = s.build()
    assert st.n_prims == 0 and st.n_nodes == 0
    rays = random_rays(64, 1)
    hits = s.trace_closest(rays)
    assert np.all(hits["inst"] == -1)
    assert not s.trace_any(rays).any()


def test_cornell_box_primary_hits(port_lib):
    desc = scenes.cornell_box(96, 96, 8)
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    st = s.build()
    assert st.n_prims == 36
    rays = osc.camera_rays(seed=3)
    ref, _ = osc.trace_closest(rays, brute=True)
    gpu = s.trace_closest(rays)
    compare_hits(gpu, ref, rays)
    assert np.all(ref["inst"] >= 0)  # closed box: every primary ray hits


def test_axis_aligned_and_degenerate_rays(port_lib):
    """rays parallel to the axes, starting on surfaces, -0.0 components: the slab test must stay conservative"""
    desc = scenes.cornell_box(32, 32, 8)
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    s.build()
    rng = np.random.default_rng(4)
    n = 6000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-0.95, 0.95, (n, 3)) + np.array([0, 1, 0])
    axis = rng.integers(0, 3, n)
    sign = rng.choice([-1.0, 1.0], n)
    rays[np.arange(n), 4 + axis] = sign
    rays[:, 4:7] = np.where(rays[:, 4:7] == 0, rng.choice([0.0, -0.0], (n, 3)), rays[:, 4:7])
    rays[:, 3], rays[:, 7] = 1e-3, 1e16
    ref, _ = osc.trace_closest(rays, brute=True)
    gpu = s.trace_closest(rays)
    compare_hits(gpu, ref, rays)


def test_error_behaviour_through_the_c_abi():
    """status codes + pb2_last_error instead of the reference's assert(false) (cuda/util.h:15-38): wrong call order and bad
    arguments fail loudly, and the scene stays usable afterwards"""
    from pupiloptixlab_b200.pb2 import LaunchParams, Pb2Error
    s = pb2.Scene()
    rays = random_rays(16, 1)
    mesh = random_soup(10, 1)
    mid = s.add_mesh(mesh["positions"], mesh["indices"])
    s.add_instance(mid)
    with pytest.raises(Pb2Error, match="pb2_bvh_build first"):
        s.trace_closest(rays)
    with pytest.raises(Pb2Error, match="pb2_bvh_build first"):
        s.render(LaunchParams(max_depth=4, width=4, height=4, n_frames=1))
    s.build()
    with pytest.raises(Pb2Error, match="accum_buffer"):
        s.render(LaunchParams(max_depth=4, width=4, height=4, n_frames=1))
    with pytest.raises(Pb2Error):
        s.add_instance(mid + 7)  # no such mesh
    with pytest.raises(Pb2Error):
        s.set_option("no_such_option", 1)
    with pytest.raises(Pb2Error):
        pb2.Bitmap(np.zeros((0, 4, 4), np.float32))
    with pytest.raises(Pb2Error):
        pb2.Bitmap(np.zeros((2, 2, 4), np.float32), address_mode=9)
    assert s.trace_closest(rays[:0]).shape == (0,)  # empty batch: fine
    hits = s.trace_closest(rays)                     # and the scene still works
    assert hits.shape == (16,)
