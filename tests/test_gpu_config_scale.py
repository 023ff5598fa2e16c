"""-m gpu: parity at the sizes BASELINE.json's configs name ("parity is unproven at the scale the configs name").

  C1  Cornell box 512 x 512, 1 spp, max depth 8 through System -> PTPass -> pb2_render against the oracle, same seed
  C4  the 30.0 M-triangle terrain: 8192 rays sampled from each of the three ray batches of tools/bench_traversal.py
      (primary, shuffled cosine bounces, shadow rays) against the oracle's CPU BVH over the same triangles
  scale-stress: degenerate and sliver triangles plus affinely transformed analytic spheres in a 2 M-primitive soup,
      against the oracle's BVH and (for a subset of the rays) its exhaustive loop
  converged: 4096 spp of the Cornell box and the material grid at reduced resolution, relMSE < 1e-3 (north_star)

Tolerances are north_star's: primitive ids exact except exact-t ties, t within 1e-5 relative (gpu_util.compare_hits).
"""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

import orc
from gpu_util import compare_hits, pb2_scene_from_oracle, random_rays
from pupiloptixlab_b200 import pb2, pupil, scenes

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import ray_batches  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _system():
    pupil.init(0)
    yield
    pupil.shutdown()


def _match(g, r, rel=1e-4):
    g, r = g.reshape(-1, g.shape[-1])[:, :3].astype(np.float64), r.reshape(-1, r.shape[-1])[:, :3].astype(np.float64)
    return (np.abs(g - r) <= rel * np.maximum(1.0, np.abs(r))).all(1)


def test_c1_cornell_512_one_sample(port_lib):
    """BASELINE.json configs[0], exactly: Cornell box 512 x 512, 1 spp, max depth 8"""
    desc = scenes.cornell_box(512, 512, 8)
    pupil.load_scene(desc)
    pupil.pass_config()
    pupil.run(1)
    ref = orc.OracleScene(port_lib, desc).render(1, threads=os.cpu_count())
    assert np.array_equal(pupil.buffer("test").reshape(-1), ref["test"])
    assert np.array_equal(pupil.buffer("albedo").reshape(-1, 3), ref["albedo"])
    assert np.abs(pupil.buffer("normal").reshape(-1, 3) - ref["normal"]).max() < 2e-3
    frame = pupil.buffer("final result")
    assert frame.shape == (512, 512, 4) and np.isfinite(frame).all() and np.all(frame[..., 3] == 1.0)
    ok = _match(frame, ref["frame"])
    assert ok.mean() >= 0.999, f"{ok.mean() * 100:.3f}% of the 512 x 512 pixels within 1e-4"
    gm, rm = frame[..., :3].mean(), ref["frame"][:, :3].mean()
    assert abs(gm - rm) <= 0.005 * rm
    rs = pupil.render_stats()
    assert abs(int(rs.closest_rays) - int(ref["closest_rays"])) <= 0.002 * ref["closest_rays"]


@pytest.mark.parametrize("name,maker,min_match", [("c2_cornell", lambda: scenes.cornell_box(1920, 1080, 8), 0.999),
                                                  ("c3_material_grid", lambda: scenes.material_grid(1920, 1080, 8), 0.97)])
def test_c2_c3_full_hd_frame_against_the_oracle_on_a_pixel_sample(port_lib, name, maker, min_match):
    """BASELINE.json configs[1] / configs[2] at their frame size: one 1920 x 1080 frame through the whole path, 6000 pixels of it
    against the oracle's per-pixel loop (same seed); the RNG-only `test` buffer of the whole frame is bit-exact"""
    desc = maker()
    pupil.load_scene(desc)
    pupil.pass_config()
    pupil.run(1)
    frame, test = pupil.buffer("final result"), pupil.buffer("test")
    assert frame.shape == (1080, 1920, 4) and np.isfinite(frame).all()
    osc = orc.OracleScene(port_lib, desc)
    rng = np.random.default_rng(8)
    xs, ys = rng.integers(0, 1920, 6000), rng.integers(0, 1080, 6000)
    ref = np.stack([osc.render_pixel(int(x), int(y), seed=0)[0] for x, y in zip(xs, ys)])
    got = frame[ys, xs, :3]
    ok = (np.abs(got - ref) <= 1e-4 * np.maximum(1.0, np.abs(ref))).all(1)
    assert ok.mean() >= min_match, f"{name}: {ok.mean() * 100:.2f}% of the sampled pixels within 1e-4"
    assert abs(got.mean() - ref.mean()) <= 0.02 * ref.mean()
    strip = orc.OracleScene(port_lib, maker()).render(1)["test"].reshape(1080, 1920) if name == "c2_cornell" else None
    if strip is not None:
        assert np.array_equal(test.reshape(1080, 1920), strip)


@pytest.fixture(scope="module")
def terrain_30m(port_lib):
    desc = scenes.terrain(3873, 1920, 1080, 8)
    assert desc.num_triangles() >= 30_000_000
    pupil.load_scene(desc)
    osc = orc.OracleScene(port_lib, desc)  # CPU BVH over the same 30 M world-space triangles (~15 s)
    return desc, osc


def _gpu_hits(scene, rays):
    return scene.trace_closest(rays)


def test_c4_sampled_rays_of_the_three_batches_match_the_oracle_bvh(terrain_30m):
    """BASELINE.json configs[3]: 8192 rays from each batch bench_traversal.py times, ids exact up to ties, t within 1e-5"""
    desc, osc = terrain_30m
    scene = pupil.scene_handle()
    st = pupil.build_stats()
    assert st.n_triangles == desc.num_triangles() and st.n_prims == st.n_triangles
    s2c, c2w, _ = pupil.camera()
    rng = np.random.default_rng(2024)
    prim = ray_batches.camera_rays(s2c, c2w, 1920, 1080)
    sel = rng.choice(len(prim), 8192, replace=False)
    gpu_all = _gpu_hits(scene, prim)  # the whole 1080p batch on the GPU, a sample of it on the CPU
    ref, _ = osc.trace_closest(prim[sel], threads=os.cpu_count())
    ties = compare_hits(gpu_all[sel], ref, prim[sel])
    assert ties <= 8 and np.count_nonzero(ref["inst"] >= 0) > 6000
    m = (gpu_all[sel]["prim"] == ref["prim"]) & (ref["inst"] >= 0)
    assert np.allclose(gpu_all[sel]["u"][m], ref["u"][m], atol=2e-4) and np.allclose(gpu_all[sel]["v"][m], ref["v"][m], atol=2e-4)

    pos = ray_batches.hit_points(prim, gpu_all["t"], gpu_all["inst"])
    inco, p = ray_batches.bounce_rays(pos, 1 << 21, rng)
    gi = _gpu_hits(scene, inco)
    sel = rng.choice(len(inco), 8192, replace=False)
    ref, _ = osc.trace_closest(inco[sel], threads=os.cpu_count())
    ties = compare_hits(gi[sel], ref, inco[sel])
    assert ties <= 8

    sh = ray_batches.shadow_rays(p, rng)
    occ = scene.trace_any(sh)
    sel = rng.choice(len(sh), 8192, replace=False)
    ref_occ = osc.trace_any(sh[sel])
    diff = np.flatnonzero(ref_occ != occ[sel])
    assert len(diff) <= 4, len(diff)  # an answer may differ only where the occluder sits at the very end of the interval
    if len(diff):
        closest, _ = osc.trace_closest(sh[sel][diff])
        for k, i in enumerate(diff):
            t = closest["t"][k]
            assert abs(t - sh[sel][i, 7]) < 1e-4 * max(1.0, t) or abs(t - sh[sel][i, 3]) < 1e-4
    assert 0.02 < occ.mean() < 0.98  # the batch has both answers


def test_c4_scene_renders_like_the_oracle_on_a_pixel_sample(terrain_30m):
    """the full path on the 30 M-triangle scene: 4096 pixels of the 1080p frame, same seed, against the oracle's per-pixel loop"""
    desc, osc = terrain_30m
    pupil.pass_config()
    pupil.run(1)
    frame = pupil.buffer("final result")
    rng = np.random.default_rng(5)
    xs, ys = rng.integers(0, 1920, 4096), rng.integers(0, 1080, 4096)
    ref = np.stack([osc.render_pixel(int(x), int(y), seed=0)[0] for x, y in zip(xs, ys)])
    got = frame[ys, xs, :3]
    ok = (np.abs(got - ref) <= 1e-4 * np.maximum(1.0, np.abs(ref))).all(1)
    assert ok.mean() >= 0.985, f"{ok.mean() * 100:.2f}% of the sampled pixels within 1e-4"
    assert abs(got.mean() - ref.mean()) <= 0.02 * ref.mean()


def _stress_soup(n_tris, seed):
    """triangle soup with the cases an intersector gets wrong first: zero-area triangles (repeated vertex, collinear
    vertices), slivers with aspect ratios up to 1e6, tiny and huge triangles next to each other"""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-30, 30, size=(n_tris, 1, 3))
    v = c + rng.normal(scale=0.25, size=(n_tris, 3, 3))
    kind = rng.integers(0, 10, n_tris)
    dup = kind == 0
    v[dup, 2] = v[dup, 1]                                                     # repeated vertex
    col = kind == 1
    v[col, 2] = v[col, 0] + (v[col, 1] - v[col, 0]) * rng.uniform(0.1, 3.0, (np.count_nonzero(col), 1))  # collinear
    sl = kind == 2
    v[sl, 2] = v[sl, 0] + (v[sl, 1] - v[sl, 0]) * 0.5 + rng.normal(scale=1e-6, size=(np.count_nonzero(sl), 3))  # slivers
    big = kind == 3
    v[big] = c[big] + rng.normal(scale=6.0, size=(np.count_nonzero(big), 3, 3))
    tiny = kind == 4
    v[tiny] = c[tiny] + rng.normal(scale=1e-3, size=(np.count_nonzero(tiny), 3, 3))
    return dict(positions=v.reshape(-1, 3).astype(np.float32), indices=np.arange(n_tris * 3, dtype=np.uint32).reshape(-1, 3))


def test_degenerate_slivers_and_transformed_spheres_at_scale(port_lib):
    n_tris = 2_000_000
    shapes = [scenes.Shape("obj", scenes.Xf("srt", scale=(1.0, 0.7, 1.3), rotate_axis=(0.2, 1, 0.1), rotate_angle=21.0, translate=(0.5, 0.25, -1.0)),
                           mesh=_stress_soup(n_tris, 17))]
    rng = np.random.default_rng(18)
    for k in range(64):  # analytic spheres under non-uniform scale + rotation
        shapes.append(scenes.Shape("sphere", scenes.Xf("srt", scale=tuple(rng.uniform(0.3, 2.5, 3)), rotate_axis=tuple(rng.normal(size=3)), rotate_angle=float(rng.uniform(0, 360)),
                                                       translate=tuple(rng.uniform(-25, 25, 3))), center=tuple(rng.uniform(-1, 1, 3)), radius=float(rng.uniform(0.5, 3.0))))
    desc = scenes.SceneDesc(shapes=shapes)
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    st = s.build()
    assert st.n_prims == n_tris + 64 and st.n_spheres == 64
    rays = random_rays(65536, 77, extent=34.0)
    # a quarter of the rays aimed at the spheres (random targets rarely find 64 small spheres among 2 M triangles)
    centres = np.stack([osc.instance_xform(1 + k)[:3, 3] for k in range(64)])
    aim = np.arange(0, len(rays), 4)
    tgt = centres[rng.integers(0, 64, len(aim))] + rng.normal(scale=0.4, size=(len(aim), 3))
    away = rng.normal(size=(len(aim), 3))
    rays[aim, 0:3] = (tgt + 4.0 * away / np.linalg.norm(away, axis=1, keepdims=True)).astype(np.float32)  # start close: the soup is dense
    d = tgt - rays[aim, 0:3]
    rays[aim, 4:7] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    gpu = s.trace_closest(rays)
    sel = np.random.default_rng(3).choice(len(rays), 8192, replace=False)
    ref, _ = osc.trace_closest(rays[sel], threads=os.cpu_count())
    ties = compare_hits(gpu[sel], ref, rays[sel])
    assert ties <= 40
    assert np.count_nonzero(ref["inst"] >= 1) > 20  # spheres are hit too (the dense soup hides most of them)
    # the oracle's BVH itself against its exhaustive loop on a subset
    sub = sel[:384]
    brute, _ = osc.trace_closest(rays[sub], brute=True, threads=os.cpu_count())
    compare_hits(gpu[sub], brute, rays[sub])
    # both schedules of the primitive tests
    for coop in (0, 1):
        s.set_option("coop_prims", coop)
        compare_hits(s.trace_closest(rays[sel]), ref, rays[sel])
    rays[:, 7] = np.random.default_rng(4).uniform(0.5, 40.0, len(rays)).astype(np.float32)
    occ = s.trace_any(rays[sel])
    ref_occ = osc.trace_any(rays[sel])
    assert np.count_nonzero(occ != ref_occ) <= 4


@pytest.mark.slow
@pytest.mark.parametrize("name,maker", [("cornell", lambda: scenes.cornell_box(96, 96, 8)), ("material_grid", lambda: scenes.material_grid(128, 72, 8))])
def test_converged_4096_spp_relmse(port_lib, name, maker):
    """north_star: converged 4096-spp images must match within relMSE < 1e-3 (same seeds 0..4095 on both sides)"""
    desc = maker()
    pupil.load_scene(desc)
    pupil.pass_config(frames_per_run=256)
    pupil.run(16)
    assert pupil.pass_state()[0] == 4096
    g = pupil.buffer("pt accum buffer")[..., :3].reshape(-1, 3).astype(np.float64)
    lib = orc.ref() or port_lib
    r = orc.OracleScene(lib, desc).render(4096, threads=os.cpu_count())["accum"][:, :3].astype(np.float64)
    relmse = float(np.mean((g - r) ** 2 / (r ** 2 + 1e-2)))
    assert relmse < 1e-3, relmse
    assert abs(g.mean() - r.mean()) <= 2e-3 * r.mean()
