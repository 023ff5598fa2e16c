"""Known-answer-test drivers: run one oracle library (port or reference build) over fixed, seeded
input grids and return plain numpy arrays.  Used three ways:
  * tests/golden/make_golden.py  : reference build  -> tests/golden/kat_reference.npz (committed)
  * tests/test_oracle_pinned.py  : port build vs the golden file (runs anywhere, incl. the GPU box)
  * tests/test_oracle_vs_reference.py : port vs reference build live, bit for bit (where _ref exists)
  * tests/test_gpu_kat.py        : the CUDA device functions vs the same grids
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import orc

F = np.float32


def unit_vectors(n, seed, hemisphere=None):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if hemisphere == "+":
        v[:, 2] = np.abs(v[:, 2])
    v = v.astype(F)
    # make sure the hard cases are present: grazing, normal incidence, exactly in-plane, backfacing
    special = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0.6, 0.8, 0], [0.99999, 0, 0.004472], [0.7, 0.1, 1e-4],
                        [0.3, -0.4, -0.8660254], [1e-4, 1e-4, 1.0]], F)
    special /= np.linalg.norm(special, axis=1, keepdims=True)
    v[:len(special)] = special
    return v


def local_bsdfs():
    """One orc_local_bsdf per material flavour (7 types x a few parameter sets)."""
    out = []

    def mk(type_, **kw):
        b = orc.LocalBsdf()
        b.type = orc.MAT[type_]
        b.alpha, b.eta, b.int_fdr, b.specular_sampling_weight, b.nonlinear = 0.1, 1.5, 0.0, 0.5, 0
        b.eta3[:], b.k3[:] = (0.200438, 0.924033, 1.10221), (3.91295, 2.45285, 2.14219)
        b.reflectance[:], b.specular_reflectance[:], b.specular_transmittance[:] = (0.6, 0.5, 0.3), (1, 0.9, 0.8), (0.9, 1, 0.95)
        for k, v in kw.items():
            if isinstance(v, tuple):
                getattr(b, k)[:] = v
            else:
                setattr(b, k, v)
        out.append(b)

    mk("diffuse")
    mk("dielectric", eta=1.5)
    mk("dielectric", eta=1.0 / 1.33)
    for a in (0.05, 0.35, 0.95):
        mk("roughdielectric", eta=1.5, alpha=a)
    mk("conductor")
    mk("conductor", eta3=(0.0, 0.0, 0.0), k3=(1.0, 1.0, 1.0))
    for a in (0.01, 0.25, 0.95):
        mk("roughconductor", alpha=a)
    for nl in (0, 1):
        mk("plastic", eta=1.5, int_fdr=0.5963, specular_sampling_weight=0.62, nonlinear=nl)
    for a in (0.05, 0.35, 0.95):
        mk("roughplastic", eta=1.49, int_fdr=0.59, specular_sampling_weight=0.55, alpha=a)
    return out


def emitters():
    out = []
    e = orc.Emitter()
    e.type, e.weight, e.select_probability, e.area = orc.EMIT_TRI, 1.5, 0.3, 0.0893
    e.radiance = orc.make_texture((17.0, 12.0, 4.0))
    e.pos[0][:], e.pos[1][:], e.pos[2][:] = (-0.24, 1.98, -0.22), (0.23, 1.98, -0.22), (0.23, 1.98, 0.16)
    for k in range(3):
        e.nrm[k][:] = (0.0, -1.0, 0.0)
    e.uv[0][:], e.uv[1][:], e.uv[2][:] = (0, 0), (1, 0), (1, 1)
    out.append(e)
    e = orc.Emitter()
    e.type, e.weight, e.select_probability, e.area = orc.EMIT_TRI, 0.7, 0.2, 2.0
    from pupiloptixlab_b200.scenes import Tex
    e.radiance = orc.make_texture(Tex("checkerboard", (3, 2, 1), (0.1, 0.2, 0.3), (4, 4, 1)))
    e.pos[0][:], e.pos[1][:], e.pos[2][:] = (0, 0, 3), (2, 0, 3), (0, 2, 3.5)
    e.nrm[0][:], e.nrm[1][:], e.nrm[2][:] = (0, 0.1, -1), (0.1, 0, -1), (0, 0, -1)
    e.uv[0][:], e.uv[1][:], e.uv[2][:] = (0, 0), (1, 0), (0, 1)
    out.append(e)
    e = orc.Emitter()
    e.type, e.weight, e.select_probability = orc.EMIT_SPHERE, 2.0, 0.4
    e.radiance = orc.make_texture((5.0, 5.0, 4.0))
    e.center[:], e.radius = (0.5, 3.0, -1.0), 0.75
    e.area = float(F(4 * np.pi) * F(0.75) * F(0.75))
    out.append(e)
    e = orc.Emitter()
    e.type, e.weight, e.select_probability = orc.EMIT_CONST_ENV, 1.0, 0.1
    e.radiance = orc.make_texture((1.0, 0.9, 0.8))
    out.append(e)
    return out


def test_image(w, h, seed, hdr=False):
    """deterministic float32 (h, w, 4) texels: smooth gradients + noise (+ a bright lobe when hdr)"""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(h, dtype=F), np.arange(w, dtype=F), indexing="ij")
    img = np.zeros((h, w, 4), F)
    img[..., 0] = 0.2 + 0.6 * x / max(w - 1, 1)
    img[..., 1] = 0.1 + 0.8 * y / max(h - 1, 1)
    img[..., 2] = 0.5 + 0.4 * np.sin(x * 0.7) * np.cos(y * 0.9)
    img[..., :3] += rng.random((h, w, 3), dtype=F) * F(0.15)
    if hdr:
        img[..., :3] += (40.0 * np.exp(-((x - 0.7 * w) ** 2 + (y - 0.3 * h) ** 2) / (0.02 * w * w)))[..., None].astype(F)
    img[..., 3] = 1.0
    return np.ascontiguousarray(img.astype(F))


_ENV_KEEP = []


def env_map_emitter(lib, img, scale=1.5, rotate_deg=30.0):
    """orc.Emitter of type env map over `img`, tables built by the library itself (padded like orc scenes pad them)"""
    h, w = img.shape[:2]
    rc, rw, cc = np.zeros(h + 2, F), np.zeros(h + 1, F), np.zeros((h + 1) * (w + 1), F)
    norm = lib.orc_build_env_tables(orc.fp(img), w, h, orc.fp(rc), orc.fp(rw), orc.fp(cc))
    rc[h + 1], rw[h], cc[h * (w + 1):] = 1.0, rw[h - 1], cc[(h - 1) * (w + 1):h * (w + 1)]
    _ENV_KEEP.append((img, rc, rw, cc))
    e = orc.Emitter()
    e.type, e.weight, e.select_probability = orc.EMIT_ENV_MAP, 1.0, 0.25
    from pupiloptixlab_b200.scenes import Tex
    e.radiance = orc.make_texture(Tex("bitmap", image=img, filter_type="bilinear", wrap_mode="repeat"))
    e.scale, e.normalization, e.map_w, e.map_h = scale, norm, w, h
    a = np.deg2rad(rotate_deg)
    m = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], F)
    e.to_world[:] = [float(v) for v in m.reshape(-1)]
    e.to_local[:] = [float(v) for v in m.T.reshape(-1)]
    e.row_cdf, e.row_weight, e.col_cdf = orc.fp(rc), orc.fp(rw), orc.fp(cc)
    return e


def run(lib) -> dict[str, np.ndarray]:
    P = C.POINTER
    R: dict[str, np.ndarray] = {}
    # ---- RNG (integer exact) ----
    states, streams = [], []
    for (px, seed) in [(0, 0), (1, 0), (12345, 7), (262143, 63), (2073599, 4095), (0xFFFFFFFF, 0xFFFFFFFF)]:
        out = np.zeros(64, F)
        st = C.c_uint32()
        lib.orc_rng_stream(4, px, seed, 64, C.byref(st), orc.fp(out))
        states.append(st.value)
        streams.append(out)
    R["rng_state"], R["rng_stream"] = np.array(states, np.uint32), np.stack(streams)
    # ---- sampling warps ----
    g = np.linspace(0, 1, 17, dtype=F)
    g[-1] = np.nextafter(F(1), F(0))
    U = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    R["warp_in"] = U
    for which, name in enumerate(["tri", "sphere", "coshemi", "unihemi"]):
        o = np.zeros((len(U), 3), F)
        for i, (a, b) in enumerate(U):
            lib.orc_warp(which, float(a), float(b), orc.fp(o[i]))
        R["warp_" + name] = o
    # ---- frames ----
    V, N = unit_vectors(64, 1), unit_vectors(64, 2)
    loc, wor, uv = np.zeros((64, 3), F), np.zeros((64, 3), F), np.zeros((64, 2), F)
    for i in range(64):
        lib.orc_frame(orc.fp(V[i]), orc.fp(N[i]), orc.fp(loc[i]), orc.fp(wor[i]))
        lib.orc_sphere_texcoord(orc.fp(N[i]), orc.fp(uv[i]))
    R["frame_v"], R["frame_n"], R["frame_local"], R["frame_world"], R["sphere_uv"] = V, N, loc, wor, uv
    # ---- Fresnel ----
    cos = np.concatenate([np.linspace(-1, 1, 41), [1e-4, -1e-4, 0.0]]).astype(F)
    etas = np.array([1.5, 1 / 1.5, 1.33, 1.000277, 2.419], F)
    fd, ct = np.zeros((len(etas), len(cos)), F), np.zeros((len(etas), len(cos)), F)
    for i, e in enumerate(etas):
        for j, c in enumerate(cos):
            t = C.c_float()
            fd[i, j] = lib.orc_fresnel_dielectric(float(e), float(c), C.byref(t))
            ct[i, j] = t.value
    R["fresnel_cos"], R["fresnel_eta"], R["fresnel_dielectric"], R["fresnel_cos_t"] = cos, etas, fd, ct
    fc = np.zeros((len(cos), 3), F)
    eta3, k3 = np.array([0.200438, 0.924033, 1.10221], F), np.array([3.91295, 2.45285, 2.14219], F)
    for j, c in enumerate(cos):
        lib.orc_fresnel_conductor(orc.fp(eta3), orc.fp(k3), float(c), orc.fp(fc[j]))
    R["fresnel_conductor"] = fc
    de = np.array([0.5, 1 / 1.5, 0.9, 1.0, 1.2, 1.5, 2.0, 5.0], F)
    R["fresnel_diffuse_eta"], R["fresnel_diffuse"] = de, np.array([lib.orc_fresnel_diffuse(float(e)) for e in de], F)
    # ---- GGX ----
    WO, WI = unit_vectors(48, 3, "+"), unit_vectors(48, 4, "+")
    WH = WO + WI
    WH = (WH / np.linalg.norm(WH, axis=1, keepdims=True)).astype(F)
    alphas = np.array([0.01, 0.05, 0.35, 0.95], F)
    gg, gs = np.zeros((len(alphas), 48, 4), F), np.zeros((len(alphas), 48, 3), F)
    xi = np.random.default_rng(5).random((48, 2)).astype(F)
    for a, al in enumerate(alphas):
        for i in range(48):
            lib.orc_ggx(orc.fp(WI[i]), orc.fp(WO[i]), orc.fp(WH[i]), float(al), orc.fp(gg[a, i]))
            lib.orc_ggx_sample(orc.fp(WO[i]), float(al), float(xi[i, 0]), float(xi[i, 1]), orc.fp(gs[a, i]))
    R["ggx_wo"], R["ggx_wi"], R["ggx_wh"], R["ggx_alpha"], R["ggx_xi"], R["ggx_dgp"], R["ggx_sample"] = WO, WI, WH, alphas, xi, gg, gs
    # ---- textures ----
    from pupiloptixlab_b200.scenes import Tex
    tex = orc.make_texture(Tex("checkerboard", (0.8, 0.7, 0.6), (0.1, 0.2, 0.3), (5.0, 3.0, 1.0)))
    UV = np.concatenate([U * 2 - 0.5, [[0.1, 0.1], [0.5, 0.5], [-0.3, 1.7]]]).astype(F)
    to = np.zeros((len(UV), 3), F)
    for i, (a, b) in enumerate(UV):
        lib.orc_tex_sample(C.byref(tex), float(a), float(b), orc.fp(to[i]))
    R["tex_uv"], R["tex_checker"] = UV, to
    # bitmap: every address / filter mode over the same uv set (tex2D rules of orc_tex2d.h), with a to_uv scale
    img = test_image(13, 7, 21)
    UVB = np.concatenate([UV, np.random.default_rng(22).uniform(-2.5, 3.5, (200, 2)).astype(F)]).astype(F)
    R["texbmp_uv"] = UVB
    for wrap in ("repeat", "clamp", "mirror"):
        for filt in ("nearest", "bilinear"):
            tb = orc.make_texture(Tex("bitmap", image=img, filter_type=filt, wrap_mode=wrap, uv_scale=(1.5, 0.75, 1.0)))
            tb.to_uv[0], tb.to_uv[5] = 1.5, 0.75
            o = np.zeros((len(UVB), 3), F)
            for i, (a, b) in enumerate(UVB):
                lib.orc_tex_sample(C.byref(tb), float(a), float(b), orc.fp(o[i]))
            R[f"texbmp_{wrap}_{filt}"] = o
    R["texbmp_weight"] = np.array([lib.orc_texture_weight(C.byref(orc.make_texture(Tex("bitmap", image=im)))) for im in (img, test_image(5, 9, 3), test_image(8, 8, 4))], F)
    # ---- BSDFs: Sample over (material, wo, rng state) and Eval over (material, wi, wo) ----
    mats = local_bsdfs()
    WO2 = unit_vectors(40, 6)      # both hemispheres: back-facing wo must be handled like the reference
    WI2 = unit_vectors(40, 7)
    rng_states = np.random.default_rng(8).integers(0, 2 ** 32, size=40, dtype=np.uint64).astype(np.uint32)
    S = np.zeros((len(mats), 40, 9), F)   # wi(3) f(3) pdf type rng_after(as float bits below)
    S_rng = np.zeros((len(mats), 40), np.uint32)
    S_type = np.zeros((len(mats), 40), np.uint32)
    E = np.zeros((len(mats), 40, 4), F)
    for m, b in enumerate(mats):
        for i in range(40):
            res = orc.BsdfResult()
            lib.orc_bsdf_sample(C.byref(b), orc.fp(WO2[i]), int(rng_states[i]), C.byref(res))
            S[m, i, 0:3], S[m, i, 3:6], S[m, i, 6] = list(res.wi), list(res.f), res.pdf
            S_rng[m, i], S_type[m, i] = res.rng_after, res.sampled_type
            f, pdf = np.zeros(3, F), C.c_float()
            lib.orc_bsdf_eval(C.byref(b), orc.fp(WI2[i]), orc.fp(WO2[i]), orc.fp(f), C.byref(pdf))
            E[m, i, :3], E[m, i, 3] = f, pdf.value
    R["bsdf_wo"], R["bsdf_wi"], R["bsdf_rng_in"] = WO2, WI2, rng_states
    R["bsdf_sample"], R["bsdf_sample_rng"], R["bsdf_sample_type"], R["bsdf_eval"] = S[:, :, :7], S_rng, S_type, E
    # ---- emitters ----
    ems = emitters()
    HP = (np.random.default_rng(9).random((32, 3)) * 2 - 1).astype(F)
    HN = unit_vectors(32, 10)
    XI = np.random.default_rng(11).random((32, 2)).astype(F)
    SD = np.zeros((len(ems), 32, 15), F)
    EV = np.zeros((len(ems), 32, 4), F)
    for k, e in enumerate(ems):
        for i in range(32):
            es = orc.EmitSample()
            lib.orc_emitter_sample_direct(C.byref(e), orc.fp(HP[i]), orc.fp(HN[i]), float(XI[i, 0]), float(XI[i, 1]), C.byref(es))
            SD[k, i] = list(es.radiance) + list(es.wi) + list(es.pos) + list(es.normal) + [es.distance, es.pdf, float(es.is_delta)]
            rad, pdf = np.zeros(3, F), C.c_float()
            # evaluate "an emitter hit" at the sampled point seen from the scatter position
            pos, nrm, uv = np.array(list(es.pos), F), np.array(list(es.normal), F), XI[i].copy()
            lib.orc_emitter_eval(C.byref(e), orc.fp(pos), orc.fp(nrm), orc.fp(uv), orc.fp(HP[i]), orc.fp(rad), C.byref(pdf))
            EV[k, i, :3], EV[k, i, 3] = rad, pdf.value
    R["emit_hit_pos"], R["emit_hit_n"], R["emit_xi"], R["emit_sample"], R["emit_eval"] = HP, HN, XI, SD, EV
    # ---- environment map: tables (BuildEnvMapCdfTable), SampleDirect, Eval ----
    env_img = test_image(16, 8, 31, hdr=True)
    eh, ew = env_img.shape[:2]
    rc, rw, cc = np.zeros(eh + 1, F), np.zeros(eh, F), np.zeros(eh * (ew + 1), F)
    R["env_normalization"] = np.array([lib.orc_build_env_tables(orc.fp(env_img), ew, eh, orc.fp(rc), orc.fp(rw), orc.fp(cc))], F)
    R["env_row_cdf"], R["env_row_weight"], R["env_col_cdf"] = rc, rw, cc
    env = env_map_emitter(lib, env_img)
    XE = np.concatenate([np.random.default_rng(12).random((96, 2)), [[0.0, 0.0], [0.999999, 0.999999], [1e-6, 0.5], [0.5, 1e-6]]]).astype(F)
    DIRS = unit_vectors(len(XE), 13)
    ES, EE = np.zeros((len(XE), 8), F), np.zeros((len(XE), 4), F)
    zero3 = np.zeros(3, F)
    for i in range(len(XE)):
        es = orc.EmitSample()
        lib.orc_emitter_sample_direct(C.byref(env), orc.fp(HP[i % 32]), orc.fp(HN[i % 32]), float(XE[i, 0]), float(XE[i, 1]), C.byref(es))
        ES[i] = list(es.radiance) + list(es.wi) + [es.distance, es.pdf]
        rad, pdf = np.zeros(3, F), C.c_float()
        pos = (HP[i % 32] + DIRS[i]).astype(F)
        lib.orc_emitter_eval(C.byref(env), orc.fp(pos), orc.fp(zero3), orc.fp(np.zeros(2, F)), orc.fp(HP[i % 32]), orc.fp(rad), C.byref(pdf))
        EE[i, :3], EE[i, 3] = rad, pdf.value
    R["env_xi"], R["env_dirs"], R["env_sample"], R["env_eval"] = XE, DIRS, ES, EE
    arr = (orc.Emitter * 3)(*ems[:3])
    ps = np.concatenate([np.linspace(0, 1, 33), [0.3, 0.5, 0.9, 0.90000004]]).astype(F)
    R["select_p"] = ps
    R["select_env"] = np.array([lib.orc_select_emitter(arr, 3, 1, float(p)) for p in ps], np.int32)
    R["select_noenv"] = np.array([lib.orc_select_emitter(arr, 3, 0, float(p)) for p in ps], np.int32)
    return R


def render_cases():
    """(name, SceneDesc, frames) rendered for the image-level goldens; small enough to commit."""
    from pupiloptixlab_b200 import scenes
    return [
        ("cornell64_d8", scenes.cornell_box(64, 64, 8), 1),
        ("cornell64_d8_4spp", scenes.cornell_box(64, 64, 8), 4),
        ("cornell48x32_d4", scenes.cornell_box(48, 32, 4), 2),
        ("grid96x54_d8", scenes.material_grid(96, 54, 8), 2),
        ("grid64x36_noarea_d6", scenes.material_grid(64, 36, 6, nx=4, nz=2, with_area_light=False), 2),
        ("envmap64x36_d6", scenes.envmap_scene(64, 36, 6), 2),
    ]


def run_renders(lib) -> dict[str, np.ndarray]:
    out = {}
    for name, desc, frames in render_cases():
        s = orc.OracleScene(lib, desc)
        r = s.render(frames)
        out[name + "/accum"] = r["accum"]
        out[name + "/albedo"], out[name + "/normal"], out[name + "/test"] = r["albedo"], r["normal"], r["test"]
        out[name + "/rays"] = np.array([r["closest_rays"], r["shadow_rays"]], np.uint64)
        s2c, c2w, fov = s.camera()
        out[name + "/s2c"], out[name + "/c2w"] = s2c, c2w
    return out
