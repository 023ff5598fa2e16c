"""-m gpu: the compressed 8-wide BVH as the build leaves it in device memory (pb2_bvh_download), walked on the host.

The traversal tests prove that the tree FINDS the right hits; this one checks the structure itself, for every builder and both
cuts of the binary tree (bvh_build.cu: greedy largest-area expansion, SAH-optimal cut): every primitive record sits in exactly
one leaf slot, every node is reached exactly once, the unary counts and offsets of the leaf slots tile a node's primitive range,
child indices follow the inner-slot mask, and the quantised box of every slot — decoded with the arithmetic the node test uses,
origin + byte * 2^e — contains everything below it.  The reference leaves all of this to optixAccelBuild
(framework/world/gas_manager.cpp:211-224), which does not expose its tree; the contract restated here is the one the
traversal kernels rely on (traverse.cuh)."""
import numpy as np
import pytest

import orc
from gpu_util import pb2_scene_from_oracle, random_soup
from pupiloptixlab_b200 import pb2, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    pb2.init(0)


def _byte(words, first, slot):
    return (int(words[first + slot // 4]) >> (8 * (slot % 4))) & 0xFF


def _walk(nodes, prims, root, seen_node, seen_prim, eps):
    """returns (lo, hi) of the primitives below `root` and checks the node on the way"""
    assert not seen_node[root], f"node {root} reached twice"
    seen_node[root] = True
    w = nodes[root]
    origin = w[0:3].view(np.float32)
    ebits = int(w[3])
    scale = np.array([np.ldexp(np.float32(1.0), ((ebits >> (8 * k)) & 0xFF) - 127) for k in range(3)], np.float32)
    imask = ebits >> 24
    child_base, prim_base = int(w[4]), int(w[5])
    lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
    next_off = 0
    for s in range(8):
        meta = _byte(w, 6, s)
        qlo = np.array([_byte(w, 8, s), _byte(w, 10, s), _byte(w, 12, s)], np.float32)
        qhi = np.array([_byte(w, 14, s), _byte(w, 16, s), _byte(w, 18, s)], np.float32)
        if meta == 0:
            assert not (imask >> s) & 1
            assert (qlo > qhi).all(), "an empty slot must hold an inverted box"
            continue
        box_lo, box_hi = origin + qlo * scale, origin + qhi * scale  # fp32, as the node test decodes it
        inner = (meta & 0x18) == 0x18
        assert inner == bool((imask >> s) & 1)
        if inner:
            assert meta == ((1 << 5) | (24 + s))
            child = child_base + bin(imask & ((1 << s) - 1)).count("1")
            lo, hi = _walk(nodes, prims, child, seen_node, seen_prim, eps)
        else:
            unary, off = meta >> 5, meta & 0x1F
            assert unary in (1, 3, 7) and off == next_off, "leaf slots tile the node's primitive range in slot order"
            count = bin(unary).count("1")
            next_off += count
            lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
            for k in range(count):
                slot = prim_base + off + k
                assert not seen_prim[slot], f"primitive record {slot} sits in two leaf slots"
                seen_prim[slot] = True
                rec = prims[slot]
                assert rec[11:12].view(np.uint32)[0] == 0, "triangle scenes only"
                v0, e1, e2 = rec[0:3].astype(np.float64), rec[4:7].astype(np.float64), rec[8:11].astype(np.float64)
                pts = np.stack([v0, v0 + e1, v0 + e2])
                lo, hi = np.minimum(lo, pts.min(0)), np.maximum(hi, pts.max(0))
        tol = eps * np.maximum(1.0, np.maximum(np.abs(lo), np.abs(hi)))
        assert (box_lo <= lo + tol).all() and (box_hi >= hi - tol).all(), (root, s, box_lo, lo, box_hi, hi)
        lo_all, hi_all = np.minimum(lo_all, lo), np.maximum(hi_all, hi)
    assert next_off <= 24
    return lo_all, hi_all


def _descs():
    one = scenes.Xf("srt", scale=(1.0, 1.3, 0.7), rotate_axis=(0.2, 1.0, 0.1), rotate_angle=25.0, translate=(0.5, -1.0, 2.0))
    for n in (1, 2, 3, 4, 9, 37, 3000):
        yield f"soup_{n}", scenes.SceneDesc(shapes=[scenes.Shape("obj", one, mesh=random_soup(n, 20 + n))])
    yield "heightfield_8k", scenes.SceneDesc(shapes=[scenes.Shape("obj", scenes.Xf("srt"), mesh=scenes.heightfield_mesh(64))])


@pytest.mark.parametrize("builder", [0, 1, 2])
@pytest.mark.parametrize("collapse", [0, 1])
def test_wide_tree_is_well_formed(port_lib, builder, collapse):
    for name, desc in _descs():
        s = pb2_scene_from_oracle(desc, orc.OracleScene(port_lib, desc))
        s.set_option("instancing", 0)
        s.set_option("collapse", collapse)
        s.set_builder(builder)
        st = s.build()
        nodes, prims = s.bvh_download()
        assert len(nodes) == st.n_nodes and len(prims) == st.n_prims == st.n_triangles, name
        seen_node, seen_prim = np.zeros(len(nodes), bool), np.zeros(len(prims), bool)
        _walk(nodes, prims, 0, seen_node, seen_prim, 1e-5)
        assert seen_node.all(), (name, "unreachable nodes", int((~seen_node).sum()))
        assert seen_prim.all(), (name, "primitives in no leaf", int((~seen_prim).sum()))
        assert sorted(np.ascontiguousarray(prims[:, 3]).view(np.uint32).tolist()) == list(range(st.n_prims)), name  # every triangle of the mesh, once
