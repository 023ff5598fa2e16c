"""Live diff of the two oracle builds (needs oracle/_ref/liborc_ref.so, i.e. the reference tree or a
prebuilt copy).  Wider and more random than the committed goldens."""
import ctypes as C

import numpy as np
import pytest

import kat
import orc
from pupiloptixlab_b200 import scenes


def test_kat_bit_exact(port_lib, ref_lib):
    a, b = kat.run(port_lib), kat.run(ref_lib)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_bsdf_sweep_bit_exact(port_lib, ref_lib, seed):
    rng = np.random.default_rng(100 + seed)
    mats = kat.local_bsdfs()
    for b in mats:
        for _ in range(200):
            wo = rng.normal(size=3)
            wo = (wo / np.linalg.norm(wo)).astype(np.float32)
            wi = rng.normal(size=3)
            wi = (wi / np.linalg.norm(wi)).astype(np.float32)
            st = int(rng.integers(0, 2 ** 32))
            b.alpha = float(np.float32(rng.uniform(0.01, 1.0)))
            ra, rb = orc.BsdfResult(), orc.BsdfResult()
            port_lib.orc_bsdf_sample(C.byref(b), orc.fp(wo), st, C.byref(ra))
            ref_lib.orc_bsdf_sample(C.byref(b), orc.fp(wo), st, C.byref(rb))
            assert bytes(ra) == bytes(rb), (b.type, wo, st)
            fa, fb, pa, pb = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float(), C.c_float()
            port_lib.orc_bsdf_eval(C.byref(b), orc.fp(wi), orc.fp(wo), orc.fp(fa), C.byref(pa))
            ref_lib.orc_bsdf_eval(C.byref(b), orc.fp(wi), orc.fp(wo), orc.fp(fb), C.byref(pb))
            assert np.array_equal(fa, fb, equal_nan=True) and (pa.value == pb.value or (np.isnan(pa.value) and np.isnan(pb.value)))


def test_render_bit_exact(port_lib, ref_lib):
    for desc, frames in [(scenes.cornell_box(40, 40, 8), 3), (scenes.material_grid(64, 36, 8), 3)]:
        a = orc.OracleScene(port_lib, desc).render(frames)
        b = orc.OracleScene(ref_lib, desc).render(frames)
        for k in ("accum", "albedo", "normal", "test"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k
        assert a["closest_rays"] == b["closest_rays"] and a["shadow_rays"] == b["shadow_rays"]
