"""ctypes binding of the ORACLE libraries (test infrastructure only).

  port()  -> oracle/_build/liborc_port.so  (restated arithmetic; built on demand with `make -C oracle port`)
  ref()   -> oracle/_ref/liborc_ref.so     (reference headers compiled on the host; None when absent)

Nothing in pupiloptixlab_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE = ROOT / "oracle"

MAT = dict(unknown=0, diffuse=1, dielectric=2, roughdielectric=3, conductor=4, roughconductor=5, plastic=6, roughplastic=7)
TEX_RGB, TEX_BITMAP, TEX_CHECKER = 0, 1, 2
EMIT_TRI, EMIT_SPHERE, EMIT_CONST_ENV, EMIT_ENV_MAP = 1, 2, 3, 4
ADDR = dict(repeat=0, clamp=1, mirror=2)
FILTER = dict(nearest=0, bilinear=1)
SHAPE = dict(obj=1, sphere=2, cube=3, rectangle=4)
XF_IDENTITY, XF_MATRIX16, XF_MATRIX9, XF_LOOKAT, XF_SRT = range(5)

f32, i32, u32 = C.c_float, C.c_int32, C.c_uint32


class Texture(C.Structure):
    _fields_ = [("type", i32), ("a", f32 * 3), ("b", f32 * 3), ("to_uv", f32 * 16),
                ("bitmap_w", i32), ("bitmap_h", i32), ("address_mode", i32), ("filter_mode", i32), ("bitmap", C.POINTER(f32))]


class Material(C.Structure):
    _fields_ = [("type", i32), ("twosided", i32), ("int_ior", f32), ("ext_ior", f32), ("nonlinear", i32),
                ("alpha", Texture), ("eta", Texture), ("k", Texture), ("reflectance", Texture),
                ("specular_reflectance", Texture), ("specular_transmittance", Texture)]


class Transform(C.Structure):
    _fields_ = [("kind", i32), ("m", f32 * 16), ("origin", f32 * 3), ("target", f32 * 3), ("up", f32 * 3),
                ("has_scale", i32), ("has_rotate", i32), ("has_translate", i32),
                ("scale", f32 * 3), ("axis", f32 * 3), ("angle", f32), ("translate", f32 * 3)]


class LocalBsdf(C.Structure):
    _fields_ = [("type", i32), ("alpha", f32), ("eta", f32), ("int_fdr", f32), ("specular_sampling_weight", f32),
                ("nonlinear", i32), ("eta3", f32 * 3), ("k3", f32 * 3), ("reflectance", f32 * 3),
                ("specular_reflectance", f32 * 3), ("specular_transmittance", f32 * 3)]


class BsdfResult(C.Structure):
    _fields_ = [("wi", f32 * 3), ("f", f32 * 3), ("pdf", f32), ("sampled_type", u32), ("rng_after", u32)]


class Emitter(C.Structure):
    _fields_ = [("type", i32), ("weight", f32), ("select_probability", f32), ("radiance", Texture), ("area", f32),
                ("pos", (f32 * 3) * 3), ("nrm", (f32 * 3) * 3), ("uv", (f32 * 2) * 3), ("center", f32 * 3), ("radius", f32),
                ("scale", f32), ("normalization", f32), ("map_w", u32), ("map_h", u32), ("to_world", f32 * 9), ("to_local", f32 * 9),
                ("row_cdf", C.POINTER(f32)), ("row_weight", C.POINTER(f32)), ("col_cdf", C.POINTER(f32))]


class EmitSample(C.Structure):
    _fields_ = [("radiance", f32 * 3), ("wi", f32 * 3), ("pos", f32 * 3), ("normal", f32 * 3), ("distance", f32),
                ("pdf", f32), ("is_delta", i32)]


class Hit(C.Structure):
    _fields_ = [("t", f32), ("u", f32), ("v", f32), ("inst", i32), ("prim", i32)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("inst", "<i4"), ("prim", "<i4")])

_IDENTITY = (1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1)


def fp(a):
    return a.ctypes.data_as(C.POINTER(f32))


def vec3(v):
    return (f32 * 3)(*[float(x) for x in v])


def _declare(lib):
    P = C.POINTER
    lib.orc_backend_name.restype = C.c_char_p
    lib.orc_rng_stream.argtypes = [u32, u32, u32, u32, P(u32), P(f32)]
    lib.orc_warp.argtypes = [C.c_int, f32, f32, P(f32)]
    lib.orc_frame.argtypes = [P(f32)] * 4
    lib.orc_sphere_texcoord.argtypes = [P(f32), P(f32)]
    lib.orc_fresnel_dielectric.argtypes = [f32, f32, P(f32)]
    lib.orc_fresnel_dielectric.restype = f32
    lib.orc_fresnel_conductor.argtypes = [P(f32), P(f32), f32, P(f32)]
    lib.orc_fresnel_diffuse.argtypes = [f32]
    lib.orc_fresnel_diffuse.restype = f32
    lib.orc_ggx.argtypes = [P(f32), P(f32), P(f32), f32, P(f32)]
    lib.orc_ggx_sample.argtypes = [P(f32), f32, f32, f32, P(f32)]
    lib.orc_tex_sample.argtypes = [P(Texture), f32, f32, P(f32)]
    lib.orc_bsdf_sample.argtypes = [P(LocalBsdf), P(f32), u32, P(BsdfResult)]
    lib.orc_bsdf_eval.argtypes = [P(LocalBsdf), P(f32), P(f32), P(f32), P(f32)]
    lib.orc_emitter_sample_direct.argtypes = [P(Emitter), P(f32), P(f32), f32, f32, P(EmitSample)]
    lib.orc_emitter_eval.argtypes = [P(Emitter), P(f32), P(f32), P(f32), P(f32), P(f32), P(f32)]
    lib.orc_select_emitter.argtypes = [P(Emitter), C.c_int, C.c_int, f32]
    lib.orc_select_emitter.restype = C.c_int
    lib.orc_resolve_transform.argtypes = [P(Transform), P(f32)]
    lib.orc_load_material.argtypes = [P(Material), P(f32), P(f32), P(f32)]
    lib.orc_scene_new.restype = C.c_void_p
    lib.orc_scene_free.argtypes = [C.c_void_p]
    lib.orc_set_integrator.argtypes = [C.c_void_p, C.c_int]
    lib.orc_set_sensor.argtypes = [C.c_void_p, f32, C.c_int, f32, f32, P(Transform), C.c_int, C.c_int]
    lib.orc_add_shape.argtypes = [C.c_void_p, C.c_int, P(Transform), P(Material), C.c_int, P(Texture), C.c_int, P(f32), f32,
                                  C.c_int, u32, u32, P(f32), P(f32), P(f32), P(u32)]
    lib.orc_add_shape.restype = C.c_int
    lib.orc_set_env_const.argtypes = [C.c_void_p, f32, f32, f32]
    lib.orc_set_env_map.argtypes = [C.c_void_p, P(f32), u32, u32, f32, P(Transform)]
    lib.orc_build_env_tables.argtypes = [P(f32), u32, u32, P(f32), P(f32), P(f32)]
    lib.orc_build_env_tables.restype = f32
    lib.orc_texture_weight.argtypes = [P(Texture)]
    lib.orc_texture_weight.restype = f32
    lib.orc_finalize.argtypes = [C.c_void_p]
    lib.orc_get_camera.argtypes = [C.c_void_p, P(f32), P(f32), P(f32)]
    lib.orc_num_area_emitters.argtypes = [C.c_void_p]
    lib.orc_get_area_emitters.argtypes = [C.c_void_p, P(Emitter)]
    lib.orc_get_env_emitter.argtypes = [C.c_void_p, P(Emitter)]
    lib.orc_num_instances.argtypes = [C.c_void_p]
    lib.orc_get_instance_xform.argtypes = [C.c_void_p, C.c_int, P(f32)]
    lib.orc_num_triangles.argtypes = [C.c_void_p]
    lib.orc_num_triangles.restype = C.c_uint64
    lib.orc_trace_closest.argtypes = [C.c_void_p, P(f32), C.c_uint64, C.c_void_p, C.c_int, C.c_int, P(C.c_uint64)]
    if hasattr(lib, "orc_trace_closest_objspace"):
        lib.orc_trace_closest_objspace.argtypes = [C.c_void_p, P(f32), C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    lib.orc_trace_any.argtypes = [C.c_void_p, P(f32), C.c_uint64, P(C.c_uint8), C.c_int]
    lib.orc_hit_geometry.argtypes = [C.c_void_p, C.c_void_p, P(f32), P(f32), P(f32), P(f32), P(C.c_int)]
    lib.orc_camera_rays.argtypes = [C.c_void_p, u32, P(f32)]
    lib.orc_render.argtypes = [C.c_void_p, u32, u32, u32, C.c_int, C.c_int, C.c_int, P(f32), P(f32), P(f32), P(f32), P(f32),
                               P(C.c_uint64)]
    lib.orc_render_pixel.argtypes = [C.c_void_p, u32, u32, u32, C.c_int, P(f32), P(u32)]
    return lib


_cache: dict[str, object] = {}


def port():
    if "port" not in _cache:
        so = ORACLE / "_build" / "liborc_port.so"
        srcs = [p for p in ORACLE.glob("orc_*") if p.is_file()]
        if not so.exists() or any(p.stat().st_mtime > so.stat().st_mtime for p in srcs):
            subprocess.run(["make", "-C", str(ORACLE), "port"], check=True, capture_output=True)
        _cache["port"] = _declare(C.CDLL(str(so)))
    return _cache["port"]


def ref():
    """The reference-header build, or None when neither the reference tree nor a prebuilt .so exists."""
    if "ref" not in _cache:
        so = ORACLE / "_ref" / "liborc_ref.so"
        ref_root = Path(os.environ.get("PUPIL_REF", "/root/reference"))
        if (ref_root / "framework").is_dir():
            subprocess.run(["make", "-C", str(ORACLE), "ref", f"PUPIL_REF={ref_root}"], check=True, capture_output=True)
        _cache["ref"] = _declare(C.CDLL(str(so))) if so.exists() else None
    return _cache["ref"]


# ---------------------------------------------------------------------------------------------
# scene description (pupiloptixlab_b200.scenes.SceneDesc) -> oracle scene
# ---------------------------------------------------------------------------------------------
def make_texture(t) -> Texture:
    """t: None | float | (r,g,b) | scenes.Tex"""
    out = Texture()
    out.to_uv[:] = _IDENTITY
    if t is None:
        return out
    if isinstance(t, (int, float)):
        out.type = TEX_RGB
        out.a[:] = [float(t)] * 3
        return out
    if isinstance(t, (tuple, list, np.ndarray)):
        out.type = TEX_RGB
        out.a[:] = [float(x) for x in t]
        return out
    if t.kind == "rgb":
        out.type = TEX_RGB
        out.a[:] = [float(x) for x in t.color0]
    elif t.kind == "bitmap":  # texels are borrowed: t.image must outlive the oracle scene (it lives in the SceneDesc)
        img = t.image
        assert img.dtype == np.float32 and img.flags["C_CONTIGUOUS"] and img.ndim == 3 and img.shape[2] == 4
        out.type = TEX_BITMAP
        out.bitmap, out.bitmap_w, out.bitmap_h = fp(img), img.shape[1], img.shape[0]
        out.address_mode, out.filter_mode = ADDR[t.wrap_mode], FILTER[t.filter_type]
    else:
        out.type = TEX_CHECKER
        out.a[:] = [float(x) for x in t.color0]
        out.b[:] = [float(x) for x in t.color1]
    if t.uv_scale is not None:  # <transform name="to_uv"><scale .../>: Transform::Scale on identity
        m = list(_IDENTITY)
        m[0], m[5], m[10] = [float(x) for x in t.uv_scale]
        out.to_uv[:] = m
    return out


def make_transform(x) -> Transform:
    out = Transform()
    if x is None:
        out.kind = XF_IDENTITY
        return out
    if x.kind == "matrix":
        vals = [float(v) for v in x.matrix]
        out.kind = XF_MATRIX16 if len(vals) == 16 else XF_MATRIX9
        for i, v in enumerate(vals):
            out.m[i] = v
    elif x.kind == "lookat":
        out.kind = XF_LOOKAT
        out.origin[:], out.target[:], out.up[:] = x.origin, x.target, x.up
    elif x.kind == "srt":
        out.kind = XF_SRT
        if x.scale is not None:
            out.has_scale, out.scale[:] = 1, [float(v) for v in x.scale]
        if x.rotate_axis is not None:
            out.has_rotate, out.axis[:], out.angle = 1, [float(v) for v in x.rotate_axis], float(x.rotate_angle)
        if x.translate is not None:
            out.has_translate, out.translate[:] = 1, [float(v) for v in x.translate]
    else:
        out.kind = XF_IDENTITY
    return out


# defaults of resource/material.cpp:26-147
_DEFAULTS = dict(int_ior_glass=1.5046, int_ior_plastic=1.49, ext_ior=1.000277)


def make_material(b) -> Material:
    m = Material()
    for name in ("alpha", "eta", "k", "reflectance", "specular_reflectance", "specular_transmittance"):
        getattr(m, name).to_uv[:] = _IDENTITY
    if b is None:
        return m
    m.type = MAT[b.type]
    m.twosided = int(b.twosided)
    p = b.params
    plastic = b.type in ("plastic", "roughplastic")
    m.int_ior = float(p.get("int_ior", _DEFAULTS["int_ior_plastic"] if plastic else _DEFAULTS["int_ior_glass"]))
    m.ext_ior = float(p.get("ext_ior", _DEFAULTS["ext_ior"]))
    m.nonlinear = int(bool(p.get("nonlinear", False)))
    m.alpha = make_texture(p.get("alpha", 0.1))
    m.eta = make_texture(p.get("eta", 0.0))
    m.k = make_texture(p.get("k", 1.0))
    if b.type == "diffuse":
        m.reflectance = make_texture(p.get("reflectance", 0.5))
    else:
        m.reflectance = make_texture(p.get("diffuse_reflectance", 0.5))
    m.specular_reflectance = make_texture(p.get("specular_reflectance", 1.0))
    m.specular_transmittance = make_texture(p.get("specular_transmittance", 1.0))
    return m


class OracleScene:
    def __init__(self, lib, desc):
        self.lib, self.desc = lib, desc
        self.h = C.c_void_p(lib.orc_scene_new())
        s = desc.sensor
        lib.orc_set_integrator(self.h, int(desc.max_depth))
        xf = make_transform(s.to_world)
        lib.orc_set_sensor(self.h, float(s.fov), int(s.fov_axis == "x"), float(s.near_clip), float(s.far_clip), C.byref(xf),
                           int(s.width), int(s.height))
        self._keep = []
        for sh in desc.shapes:
            xf = make_transform(sh.to_world)
            mat = make_material(sh.bsdf)
            rad = make_texture(sh.emitter) if sh.emitter is not None else None
            center = vec3(sh.center)
            null_f, null_u = C.POINTER(f32)(), C.POINTER(u32)()
            nv = nf = 0
            pos = nrm = uv = null_f
            idx = null_u
            if sh.type == "obj":
                P = np.ascontiguousarray(sh.mesh["positions"], np.float32)
                I = np.ascontiguousarray(sh.mesh["indices"], np.uint32)
                N = sh.mesh.get("normals")
                T = sh.mesh.get("texcoords")
                N = None if N is None else np.ascontiguousarray(N, np.float32)
                T = None if T is None else np.ascontiguousarray(T, np.float32)
                self._keep += [P, I, N, T]
                nv, nf = P.shape[0], I.shape[0]
                pos, idx = fp(P), I.ctypes.data_as(C.POINTER(u32))
                nrm = fp(N) if N is not None else null_f
                uv = fp(T) if T is not None else null_f
            lib.orc_add_shape(self.h, SHAPE[sh.type], C.byref(xf), C.byref(mat), int(sh.emitter is not None),
                              C.byref(rad) if rad is not None else None, int(sh.flip_normals), center, float(sh.radius),
                              int(sh.flip_tex_coords), nv, nf, pos, nrm, uv, idx)
        if desc.env_radiance is not None:
            lib.orc_set_env_const(self.h, *[float(x) for x in desc.env_radiance])
        if getattr(desc, "env_map", None) is not None:
            em = desc.env_map
            xf = make_transform(em.to_world)
            lib.orc_set_env_map(self.h, fp(em.image), em.image.shape[1], em.image.shape[0], float(em.scale), C.byref(xf))
        lib.orc_finalize(self.h)
        self.w, self.h_px = int(s.width), int(s.height)

    def __del__(self):
        try:
            self.lib.orc_scene_free(self.h)
        except Exception:
            pass

    def camera(self):
        s2c, c2w = np.zeros(16, np.float32), np.zeros(16, np.float32)
        fov = f32()
        self.lib.orc_get_camera(self.h, fp(s2c), fp(c2w), C.byref(fov))
        return s2c.reshape(4, 4), c2w.reshape(4, 4), fov.value

    def instance_xform(self, i):
        m = np.zeros(16, np.float32)
        self.lib.orc_get_instance_xform(self.h, i, fp(m))
        return m.reshape(4, 4)

    def area_emitters(self):
        n = self.lib.orc_num_area_emitters(self.h)
        arr = (Emitter * max(n, 1))()
        if n:
            self.lib.orc_get_area_emitters(self.h, arr)
        return list(arr)[:n]

    def env_emitter(self):
        e = Emitter()
        return e if self.lib.orc_get_env_emitter(self.h, C.byref(e)) else None

    def camera_rays(self, seed=0):
        rays = np.zeros((self.w * self.h_px, 8), np.float32)
        self.lib.orc_camera_rays(self.h, seed, fp(rays))
        return rays

    def trace_closest(self, rays, brute=False, threads=0):
        rays = np.ascontiguousarray(rays, np.float32)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        tests = C.c_uint64()
        self.lib.orc_trace_closest(self.h, fp(rays), rays.shape[0], hits.ctypes.data, int(brute), threads, C.byref(tests))
        return hits, tests.value

    def trace_closest_objspace(self, rays, objspace, threads=0):
        """exhaustive loop; instances with objspace[i] != 0 are intersected in object space (a mesh behind an instance node)"""
        rays = np.ascontiguousarray(rays, np.float32)
        flags = np.ascontiguousarray(objspace, np.uint8)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        self.lib.orc_trace_closest_objspace(self.h, fp(rays), rays.shape[0], hits.ctypes.data, flags.ctypes.data, threads)
        return hits

    def trace_any(self, rays, brute=False):
        rays = np.ascontiguousarray(rays, np.float32)
        occ = np.zeros(rays.shape[0], np.uint8)
        self.lib.orc_trace_any(self.h, fp(rays), rays.shape[0], occ.ctypes.data_as(C.POINTER(C.c_uint8)), int(brute))
        return occ

    def hit_geometry(self, hit, ray):
        h = Hit(float(hit["t"]), float(hit["u"]), float(hit["v"]), int(hit["inst"]), int(hit["prim"]))
        ray = np.ascontiguousarray(ray, np.float32)
        pos, nrm, uv = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(2, np.float32)
        ei = C.c_int()
        self.lib.orc_hit_geometry(self.h, C.byref(h), fp(ray), fp(pos), fp(nrm), fp(uv), C.byref(ei))
        return pos, nrm, uv, ei.value

    def render(self, n_frames=1, first_seed=0, sample_cnt0=0, max_depth=0, accumulate=True, threads=0, accum=None):
        n = self.w * self.h_px
        out = dict(accum=np.zeros((n, 4), np.float32) if accum is None else accum, frame=np.zeros((n, 4), np.float32),
                   albedo=np.zeros((n, 3), np.float32), normal=np.zeros((n, 3), np.float32), test=np.zeros(n, np.float32))
        counts = (C.c_uint64 * 2)()
        self.lib.orc_render(self.h, first_seed, n_frames, sample_cnt0, max_depth, int(accumulate), threads, fp(out["accum"]),
                            fp(out["frame"]), fp(out["albedo"]), fp(out["normal"]), fp(out["test"]), counts)
        out["closest_rays"], out["shadow_rays"] = counts[0], counts[1]
        return out

    def render_pixel(self, x, y, seed=0, max_depth=0):
        rad = np.zeros(3, np.float32)
        rays = (u32 * 2)()
        self.lib.orc_render_pixel(self.h, x, y, seed, max_depth, fp(rad), rays)
        return rad, (rays[0], rays[1])
