"""-m gpu: the binary tree -> BVH8 step of the build (bvh_build.cu: k_refit<true> fills the cost tables bottom-up, k_collapse
follows their decisions top-down).  The reference asks OptiX for a fast-trace structure (OPTIX_BUILD_FLAG_PREFER_FAST_TRACE,
framework/world/gas_manager.cpp:191); here that is the cut of the binary tree that minimises the SAH cost of the wide tree.

Two things are held: (i) whatever the cut, hits are those of the oracle's exhaustive loop; (ii) with the primitive-test cost at
100 % of a node test the objective of the dynamic programme IS the `sah_cost` the build reports (node area + leaf area x count,
over the root's area), and the greedy largest-area expansion (collapse = 0) is one of the cuts it searches, so its cost can
never come out higher."""
import numpy as np
import pytest

import orc
from gpu_util import compare_hits, pb2_scene_from_oracle, random_rays, random_soup
from pupiloptixlab_b200 import pb2, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    pb2.init(0)


def _descs():
    yield "soup_5", scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt"), mesh=random_soup(5, 3))])
    yield "soup_37", scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt"), mesh=random_soup(37, 4))])
    yield "soup_3000+spheres", scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt", scale=(1.0, 1.4, 0.8)), mesh=random_soup(3000, 5))] + [
        scenes.Shape("sphere", scenes.Xf("srt", scale=(1.0, 0.7, 1.2)), center=(float(x), 0.5, -2.0), radius=0.9) for x in (-6, -2, 2, 6)])
    yield "heightfield_80k", scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt"), mesh=scenes.heightfield_mesh(200))])
    yield "cornell", scenes.cornell_box(64, 64, 8)


@pytest.mark.parametrize("name,desc", list(_descs()), ids=[n for n, _ in _descs()])
def test_every_cut_gives_the_oracles_hits_and_the_optimal_one_costs_no_more(port_lib, name, desc):
    osc = orc.OracleScene(port_lib, desc)
    s = pb2_scene_from_oracle(desc, osc)
    rays = random_rays(6000, 17, extent=11.0)
    ref, _ = osc.trace_closest(rays, brute=True)
    stats = {}
    for label, opts in (("greedy", dict(collapse=0)), ("optimal_100", dict(collapse=1, collapse_prim_cost_pct=100)),
                        ("optimal_30", dict(collapse=1, collapse_prim_cost_pct=30))):
        for k, v in opts.items():
            s.set_option(k, v)
        st = s.build()
        stats[label] = (st.sah_cost, st.n_nodes)
        assert st.n_prims > 0 and st.n_nodes >= 1
        gpu = s.trace_closest(rays)
        compare_hits(gpu, ref, rays)
        same = (gpu["inst"] == ref["inst"]) & (gpu["prim"] == ref["prim"]) & (ref["inst"] >= 0)
        for f in ("t", "u", "v"):  # the cut decides which boxes are tested, never the arithmetic of a hit
            assert np.array_equal(gpu[f][same].view(np.uint32), ref[f][same].view(np.uint32)), (label, f)
        occ = s.trace_any(rays)
        assert np.count_nonzero(occ != osc.trace_any(rays, brute=True)) <= 2, label
    assert stats["optimal_100"][0] <= stats["greedy"][0] * (1 + 1e-5), stats
    if name == "heightfield_80k":  # a regular mesh: the greedy cut spends a wide node on every 4 .. 8-primitive subtree it meets
        assert stats["optimal_30"][1] < stats["greedy"][1], stats


def test_option_values_are_clamped_and_invalidate_the_tree(port_lib):
    desc = scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt"), mesh=random_soup(500, 9))])
    s = pb2_scene_from_oracle(desc, orc.OracleScene(port_lib, desc))
    s.build()
    rays = random_rays(64, 1)
    s.trace_closest(rays)
    s.set_option("collapse_prim_cost_pct", 0)  # clamped to 1: a zero-cost primitive test would still give a finite objective
    with pytest.raises(pb2.Pb2Error):
        s.trace_closest(rays)  # the option invalidated the tree
    s.build()
    s.trace_closest(rays)
