"""-m gpu: every device-side restatement (RNG, sampling warps, frames, Fresnel, GGX, textures, the seven
BSDFs, emitters, emitter selection) against the oracle on the same seeded grids, through pb2_kat.

Tolerances: integer work (RNG state/draws, lobe tags, selection indices) is bit-exact.  fp32 work is
compared with rtol 2e-5 / atol 1e-6: the device build contracts a*b+c into FMAs and uses CUDA's
sinf/cosf/atan2f/acosf (<= 2 ulp), the oracle is an x86 build with -ffp-contract=off and glibc libm.
Where a formula divides by a vanishing quantity (grazing angles) the comparison is relative to the
magnitude of the oracle value.  One input class is ill-conditioned by construction and gets its own bound:
visible-normal GGX sampling with wo BELOW the surface (wo.z < 0, reachable only on back-face hits of
one-sided rough materials).  ggx::Sample then forms s = 0.5 * (1 + Vh.z) with Vh.z -> -1 as alpha -> 0
(render/material/ggx.h:44-62): at alpha = 0.01 that difference keeps ~2 significant digits, so one ulp of
contraction difference moves the sampled direction by ~1e-3.  Those rows are held to 5e-3 absolute."""
import ctypes as C

import numpy as np
import pytest

import kat
import orc
from pupiloptixlab_b200 import pb2

pytestmark = pytest.mark.gpu
F = np.float32
RTOL, ATOL = 2e-5, 1e-6


@pytest.fixture(scope="module", autouse=True)
def _init():
    pb2.init(0)


@pytest.fixture(scope="module")
def ref(port_lib):
    return kat.run(port_lib)


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    ok = np.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True) | (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)))
    assert ok.all(), f"{np.count_nonzero(~ok)} mismatches, worst: gpu={a[~ok][:4]} ref={b[~ok][:4]}"


def test_rng_bit_exact(ref):
    cases = np.array([[4, 0, 0], [4, 1, 0], [4, 12345, 7], [4, 262143, 63], [4, 2073599, 4095], [4, 0xFFFFFFFF, 0xFFFFFFFF]], np.uint32)
    out = np.zeros((len(cases), 8), F)
    pb2.kat("rng", cases, None, None, len(cases), out)
    assert np.array_equal(out[:, 0].view(np.uint32), ref["rng_state"])
    assert np.array_equal(out[:, 1:], ref["rng_stream"][:, :7])


def test_warps(ref):
    U = ref["warp_in"]
    out = np.zeros((len(U), 12), F)
    pb2.kat("warp", U, None, None, len(U), out)
    for k, name in enumerate(["tri", "sphere", "coshemi", "unihemi"]):
        close(out[:, 3 * k:3 * k + 3], ref["warp_" + name], atol=2e-6)


def test_frames(ref):
    inp = np.concatenate([ref["frame_v"], ref["frame_n"]], 1).astype(F)
    out = np.zeros((len(inp), 8), F)
    pb2.kat("frame", inp, None, None, len(inp), out)
    close(out[:, 0:3], ref["frame_local"], atol=2e-6)
    close(out[:, 3:6], ref["frame_world"], atol=2e-6)
    close(out[:, 6:8], ref["sphere_uv"], atol=2e-6)


def test_fresnel(ref):
    cos, etas = ref["fresnel_cos"], ref["fresnel_eta"]
    eta3, k3 = [0.200438, 0.924033, 1.10221], [3.91295, 2.45285, 2.14219]
    for i, e in enumerate(etas):
        inp = np.array([[e, c, *eta3, *k3] for c in cos], F)
        out = np.zeros((len(cos), 8), F)
        pb2.kat("fresnel", inp, None, None, len(cos), out)
        close(out[:, 0], ref["fresnel_dielectric"][i], atol=2e-6)
        close(out[:, 1], ref["fresnel_cos_t"][i], atol=2e-6)
        close(out[:, 2:5], ref["fresnel_conductor"], atol=2e-6)


def test_ggx(ref):
    WO, WI, WH, xi = ref["ggx_wo"], ref["ggx_wi"], ref["ggx_wh"], ref["ggx_xi"]
    for a, al in enumerate(ref["ggx_alpha"]):
        inp = np.concatenate([WI, WO, WH, np.full((len(WO), 1), al, F), xi], 1).astype(F)
        out = np.zeros((len(WO), 8), F)
        pb2.kat("ggx", inp, None, None, len(WO), out)
        close(out[:, 0:4], ref["ggx_dgp"][a], rtol=5e-5)
        up = WO[:, 2] >= 0
        close(out[up, 4:7], ref["ggx_sample"][a][up], atol=5e-6)
        close(out[~up, 4:7], ref["ggx_sample"][a][~up], atol=5e-3)  # ill-conditioned, see the module docstring


def test_textures(ref):
    from pupiloptixlab_b200.scenes import Tex
    UV = ref["tex_uv"]
    t = pb2.Texture()
    t.type = pb2.TEX_CHECKERBOARD
    t.a[:], t.b[:] = (0.8, 0.7, 0.6), (0.1, 0.2, 0.3)
    t.r0[:], t.r1[:] = (5, 0, 0, 0), (0, 3, 0, 0)
    arr = (pb2.Texture * len(UV))(*[t] * len(UV))
    out = np.zeros((len(UV), 4), F)
    pb2.kat("texture", arr, UV, None, len(UV), out)
    # a checker lookup is a discrete choice: positions within 1e-6 of a cell edge may legitimately flip
    fx, fy = (UV[:, 0] * 5) % 1.0, (UV[:, 1] * 3) % 1.0
    edge = (np.abs(fx - 0.5) < 1e-5) | (np.abs(fy - 0.5) < 1e-5) | (fx < 1e-5) | (fy < 1e-5) | (fx > 1 - 1e-5) | (fy > 1 - 1e-5)
    assert np.array_equal(out[~edge, :3], ref["tex_checker"][~edge])


def _pb2_bitmap_texture(bitmap: pb2.Bitmap, r0=(1, 0, 0, 0), r1=(0, 1, 0, 0)) -> pb2.Texture:
    t = pb2.Texture()
    t.type, t.bitmap = pb2.TEX_BITMAP, bitmap.handle
    t.r0[:], t.r1[:] = r0, r1
    return t


@pytest.mark.parametrize("wrap", ["repeat", "clamp", "mirror"])
@pytest.mark.parametrize("filt", ["nearest", "bilinear"])
def test_bitmap_textures_match_the_tex2d_rules(ref, wrap, filt):
    """cuda::Texture::Sample's bitmap branch (tex2D<float4>, framework/cuda/texture.h:52-54) on the texture unit against
    the oracle's restatement of the published fetch rules.  Stated tolerance: nearest = exact except for coordinates
    within 1e-4 texel of a texel edge; bilinear = 1/256 of the local texel range (1.8 fixed-point weights) + 1e-6."""
    img = kat.test_image(13, 7, 21)
    bm = pb2.Bitmap(img, orc.ADDR[wrap], orc.FILTER[filt])
    UV = ref["texbmp_uv"]
    t = _pb2_bitmap_texture(bm, (1.5, 0, 0, 0), (0, 0.75, 0, 0))
    arr = (pb2.Texture * len(UV))(*[t] * len(UV))
    out = np.zeros((len(UV), 4), F)
    pb2.kat("texture", arr, UV, None, len(UV), out)
    want = ref[f"texbmp_{wrap}_{filt}"]
    x, y = UV[:, 0].astype(np.float64) * 1.5 * 13, UV[:, 1].astype(np.float64) * 0.75 * 7
    if filt == "nearest":
        edge = (np.abs(x - np.round(x)) < 1e-3) | (np.abs(y - np.round(y)) < 1e-3)
        assert edge.mean() < 0.2
        assert np.array_equal(out[~edge, :3], want[~edge])
    else:
        spread = float(img[..., :3].max() - img[..., :3].min())
        edge = (np.abs(x - 0.5 - np.round(x - 0.5)) < 1e-3) | (np.abs(y - 0.5 - np.round(y - 0.5)) < 1e-3)
        err = np.abs(out[:, :3].astype(np.float64) - want)
        assert err[~edge].max() <= spread / 256 + 1e-6, err[~edge].max()
        assert np.median(err) < 1e-3
    bm.free()


def test_env_map_emitter(ref, port_lib):
    """EnvMapEmitter::SampleDirect / Eval (framework/render/emitter/env.h:24-64): the sampled cell (direction) must be the
    reference's exactly; radiance goes through the texture unit (bilinear: 1/256 of the texel range), pdf follows it."""
    img = kat.test_image(16, 8, 31, hdr=True)
    env = kat.env_map_emitter(port_lib, img)
    h, w = img.shape[:2]
    bm = pb2.Bitmap(img, 0, 1)
    tables = np.concatenate([ref["env_row_cdf"], ref["env_row_weight"], ref["env_col_cdf"]]).astype(F)
    dev = pb2.DeviceBuffer(tables.nbytes)
    dev.upload(tables)
    p = pb2.Emitter()
    p.type, p.weight, p.select_probability = pb2.EMIT_ENV_MAP, 1.0, 0.25
    p.radiance = _pb2_bitmap_texture(bm)
    p.scale, p.normalization, p.map_w, p.map_h = env.scale, float(ref["env_normalization"][0]), w, h
    p.to_world[:], p.to_local[:] = list(env.to_world), list(env.to_local)
    p.env_tables = dev.ptr.value
    XE, DIRS, HP, HN = ref["env_xi"], ref["env_dirs"], ref["emit_hit_pos"], ref["emit_hit_n"]
    n = len(XE)
    idx = np.arange(n) % 32
    in1 = np.concatenate([HP[idx], HN[idx], XE], 1).astype(F)
    in2 = np.zeros((n, 12), F)
    in2[:, 0:3], in2[:, 8:11] = (HP[idx] + DIRS).astype(F), HP[idx]
    arr = (pb2.Emitter * n)(*[p] * n)
    out = np.zeros((n, 16), F)
    pb2.kat("emitter", arr, in1, in2, n, out)
    es, ee = ref["env_sample"], ref["env_eval"]
    # xi exactly on a cdf entry may pick the neighbouring cell on the other side: none of the test points is
    close(out[:, 3:6], es[:, 3:6], atol=5e-6)       # direction = cell choice: exact up to fp32 sin/cos
    assert np.all(out[:, 6] == es[:, 6])            # MAX_DISTANCE
    spread = float(img[..., :3].max() - img[..., :3].min()) * env.scale
    assert np.abs(out[:, 0:3] - es[:, 0:3]).max() <= spread / 256 + 1e-5
    lum = lambda c: 0.2126 * c[:, 0] + 0.7152 * c[:, 1] + 0.0722 * c[:, 2]
    # pdf = lum(radiance) * weights: compare after dividing out the radiance each side saw
    ok = lum(es[:, 0:3]) > 1e-3
    close((out[:, 7] / lum(out[:, 0:3]))[ok], (es[:, 7] / lum(es[:, 0:3]))[ok], rtol=2e-4)
    assert np.abs(out[:, 8:11] - ee[:, 0:3]).max() <= spread / 256 + 1e-5
    ok = lum(ee[:, 0:3]) > 1e-3
    close((out[:, 11] / lum(out[:, 8:11]))[ok], (ee[:, 3] / lum(ee[:, 0:3]))[ok], rtol=5e-4)
    dev.free(), bm.free()


def _kat_bsdf(b: orc.LocalBsdf) -> pb2.KatBsdf:
    k = pb2.KatBsdf()
    k.type, k.alpha, k.eta, k.int_fdr, k.specular_sampling_weight, k.nonlinear = b.type, b.alpha, b.eta, b.int_fdr, b.specular_sampling_weight, b.nonlinear
    name = {v: n for n, v in orc.MAT.items()}[b.type]
    if name == "diffuse":
        k.c0[:] = b.reflectance
    elif name in ("dielectric", "roughdielectric"):
        k.c0[:], k.c1[:] = b.specular_reflectance, b.specular_transmittance
    elif name in ("conductor", "roughconductor"):
        k.c0[:], k.c1[:], k.c2[:] = b.specular_reflectance, b.eta3, b.k3
    else:
        k.c0[:], k.c1[:] = b.reflectance, b.specular_reflectance
    return k


def name_of(b):
    return {v: n for n, v in orc.MAT.items()}[b.type]


def test_bsdfs(ref):
    mats = kat.local_bsdfs()
    WO, WI, rng_in = ref["bsdf_wo"], ref["bsdf_wi"], ref["bsdf_rng_in"]
    n = len(WO)
    worst = 0
    for m, b in enumerate(mats):
        arr = (pb2.KatBsdf * n)(*[_kat_bsdf(b)] * n)
        inp = np.zeros((n, 8), F)
        inp[:, 0:3], inp[:, 3:6] = WO, WI
        inp[:, 6] = rng_in.view(F)
        out = np.zeros((n, 16), F)
        pb2.kat("bsdf", arr, inp, None, n, out)
        # integer-exact: RNG consumption and the sampled lobe tag — except where the lobe choice compares a
        # draw with a computed probability that differs in the last bit (none on this grid)
        assert np.array_equal(out[:, 8].view(np.uint32), ref["bsdf_sample_rng"][m]), f"material {m}: RNG consumption differs"
        same_lobe = out[:, 7].view(np.uint32) == ref["bsdf_sample_type"][m]
        assert same_lobe.all(), f"material {m}: lobe tags differ on {np.count_nonzero(~same_lobe)} samples"
        s_ref, e_ref = ref["bsdf_sample"][m], ref["bsdf_eval"][m]
        # f and pdf blow up at grazing angles (divisions by wi.z*wo.z): compare relative to magnitude
        rough_below = (WO[:, 2] < 0) & (name_of(b) in ("roughconductor", "roughdielectric", "roughplastic"))
        ok_rows = ~rough_below
        close(out[ok_rows, 0:3], s_ref[ok_rows, 0:3], atol=2e-5)  # sampled direction: approximate sqrt / division on the device (<= 2 ulp each)
        close(out[ok_rows, 3:7], s_ref[ok_rows, 3:7], rtol=2e-4, atol=1e-6)
        close(out[rough_below, 0:3], s_ref[rough_below, 0:3], atol=5e-3)  # ill-conditioned VNDF sampling from below
        close(out[:, 9:13], e_ref, rtol=2e-4, atol=1e-6)
        worst = max(worst, float(np.nanmax(np.abs(out[:, 3:7] - s_ref[:, 3:7]) / (np.abs(s_ref[:, 3:7]) + 1e-3))))
    print("worst relative f/pdf difference:", worst)


def _pb2_emitter(e: orc.Emitter) -> pb2.Emitter:
    p = pb2.Emitter()
    p.type, p.weight, p.select_probability, p.area = e.type, e.weight, e.select_probability, e.area
    p.radiance.type = e.radiance.type
    p.radiance.a[:], p.radiance.b[:] = e.radiance.a, e.radiance.b
    p.radiance.r0[:], p.radiance.r1[:] = e.radiance.to_uv[0:4], e.radiance.to_uv[4:8]
    for k in range(3):
        p.pos[k][:], p.nrm[k][:], p.uv[k][:] = e.pos[k], e.nrm[k], e.uv[k]
    p.center[:], p.radius = e.center, e.radius
    return p


def test_emitters(ref):
    ems = kat.emitters()
    HP, HN, XI = ref["emit_hit_pos"], ref["emit_hit_n"], ref["emit_xi"]
    n = len(HP)
    for k, e in enumerate(ems):
        arr = (pb2.Emitter * n)(*[_pb2_emitter(e)] * n)
        in1 = np.concatenate([HP, HN, XI], 1).astype(F)
        sd = ref["emit_sample"][k]  # radiance wi pos normal distance pdf is_delta
        in2 = np.zeros((n, 12), F)
        in2[:, 0:3], in2[:, 3:6], in2[:, 6:8], in2[:, 8:11] = sd[:, 6:9], sd[:, 9:12], XI, HP
        out = np.zeros((n, 16), F)
        pb2.kat("emitter", arr, in1, in2, n, out)
        if e.radiance.type == orc.TEX_RGB:
            close(out[:, 0:3], sd[:, 0:3])
        close(out[:, 3:6], sd[:, 3:6], atol=5e-6)
        close(out[:, 6], sd[:, 12], rtol=1e-4)
        close(out[:, 7], sd[:, 13], rtol=2e-4)
        if e.radiance.type == orc.TEX_RGB:
            close(out[:, 8:12], ref["emit_eval"][k], rtol=2e-4)
        else:
            close(out[:, 11], ref["emit_eval"][k][:, 3], rtol=2e-4)


def test_select_emitter(ref):
    ems = kat.emitters()[:3]
    arr = (pb2.Emitter * 3)(*[_pb2_emitter(e) for e in ems])
    ps = ref["select_p"]
    for has_env, key in ((1, "select_env"), (0, "select_noenv")):
        out = np.zeros(len(ps), np.int32)
        pb2.kat("select", arr, ps, np.array([3, has_env], np.uint32), len(ps), out)
        assert np.array_equal(out, ref[key])


def test_select_emitter_large_table_matches_the_linear_scan(port_lib):
    """a20: the device finds the entry by binary search over the same fp32 running sums — identical to the reference's
    O(n) scan (render/emitter.h:110-136), including draws that land exactly on a running sum"""
    rng = np.random.default_rng(5)
    m = 3000
    w = rng.random(m).astype(F) ** 3
    w[rng.integers(0, m, 100)] = 0.0  # zero-probability entries must never be chosen over their successor... by either side
    sp = (w / w.sum() * F(0.9)).astype(F)
    ems = []
    for i in range(m):
        e = orc.Emitter()
        e.type, e.weight, e.select_probability, e.area = orc.EMIT_TRI, 1.0, float(sp[i]), 1.0
        ems.append(e)
    oarr = (orc.Emitter * m)(*ems)
    parr = (pb2.Emitter * m)(*[_pb2_emitter(e) for e in ems])
    cum = np.zeros(m, F)
    acc = F(0)
    for i in range(m):
        acc = F(acc + sp[i])
        cum[i] = acc
    ps = np.concatenate([rng.random(4000).astype(F), cum[rng.integers(0, m, 500)], np.nextafter(cum[:200], F(2)), [F(0), F(0.95), F(0.99999994)]]).astype(F)
    for has_env in (1, 0):
        want = np.array([port_lib.orc_select_emitter(oarr, m, has_env, float(p)) for p in ps], np.int32)
        out = np.zeros(len(ps), np.int32)
        pb2.kat("select", parr, ps, np.array([m, has_env], np.uint32), len(ps), out)
        assert np.array_equal(out, want)


@pytest.mark.parametrize("n", [1, 2, 31, 4095, 4096, 4097, 100_003, 3_000_001])
def test_radix_sort_is_a_stable_sort(n):
    """csrc/radix_sort.cu (the BVH builders' Morton sort): the permutation of a stable sort by the chosen key bits, for ragged
    tile counts, heavy duplicates and partial bit ranges"""
    rng = np.random.default_rng(n)
    cases = [(rng.integers(0, 1 << 63, n, dtype=np.uint64), 0, 63),                      # full 63-bit Morton keys
             (rng.integers(0, 7, n, dtype=np.uint64) << np.uint64(20), 0, 63),            # seven distinct keys: stability decides the order
             (rng.integers(0, 1 << 40, n, dtype=np.uint64), 8, 32)]                       # three passes over a bit window: an odd pass count
    for keys, b0, b1 in cases:
        out = np.zeros(n, np.uint32)
        pb2.kat("sort", keys, None, np.array([b0, b1], np.uint32), n, out)
        window = (keys >> np.uint64(b0)) & np.uint64((1 << (b1 - b0)) - 1)
        assert np.array_equal(out, np.argsort(window, kind="stable").astype(np.uint32))
