"""Scene descriptions at the level of the reference's mitsuba3-style XML, procedural generators for
the benchmark configurations, and an XML writer.

A `SceneDesc` says exactly what an XML file would say (resource::Scene level, before any world
precompute), so one description can be
  * written out with `to_xml()` and loaded by the host library's XML loader (the drop-in route), or
  * handed to the host library field by field (`pupil_scene_*`, for multi-million-triangle meshes
    that should not round-trip through text), and
  * handed, by the tests, to the CPU oracle.

Generators (SURVEY.md §8d "Synthetic inputs"):
  cornell_box()    C1/C2  — the numbers of data/static/cornellbox.xml (reference), 36 triangles
  material_grid()  C3/C5  — spheres cycling the seven BSDFs on a checkerboard floor under a constant env
  terrain()        C4     — displaced height field, ~2*n*n triangles, one area light + constant env
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

MAT_NAMES = ("diffuse", "dielectric", "roughdielectric", "conductor", "roughconductor", "plastic", "roughplastic")


@dataclass
class Tex:
    """<rgb> (kind='rgb', color0), <texture type="checkerboard"> (color0/color1 + to_uv scale) or
    <texture type="bitmap"> (image: float32 (H, W, 4) linear texels, row 0 = first row of the file; or filename)."""
    kind: str = "rgb"
    color0: Sequence[float] = (0.5, 0.5, 0.5)
    color1: Sequence[float] = (0.2, 0.2, 0.2)
    uv_scale: Optional[Sequence[float]] = None
    image: Optional[np.ndarray] = None
    filename: Optional[str] = None      # written to the XML when set; otherwise the image is registered in memory
    filter_type: str = "bilinear"       # bilinear | nearest   (resource/scene.cpp:151-155)
    wrap_mode: str = "repeat"           # repeat | mirror | clamp (scene.cpp:157-165)


@dataclass
class EnvMap:
    """<emitter type="envmap">: lat-long radiance image + scale + to_world (resource/scene.cpp:207-219)."""
    image: Optional[np.ndarray] = None  # float32 (H, W, 4)
    filename: Optional[str] = None
    scale: float = 1.0
    to_world: Optional["Xf"] = None


@dataclass
class Xf:
    """<transform name="to_world">: matrix (16 or 9 values) | lookat | scale/rotate/translate."""
    kind: str = "identity"
    matrix: Optional[Sequence[float]] = None
    origin: Sequence[float] = (1, 0, 0)
    target: Sequence[float] = (0, 0, 0)
    up: Sequence[float] = (0, 1, 0)
    scale: Optional[Sequence[float]] = None
    rotate_axis: Optional[Sequence[float]] = None
    rotate_angle: float = 0.0
    translate: Optional[Sequence[float]] = None


@dataclass
class Bsdf:
    type: str = "diffuse"
    twosided: bool = False
    params: dict = field(default_factory=dict)  # values: float | (r,g,b) | Tex | bool


@dataclass
class Shape:
    type: str = "rectangle"  # rectangle | cube | sphere | obj
    to_world: Optional[Xf] = None
    bsdf: Optional[Bsdf] = None
    emitter: Optional[object] = None  # area-light radiance: (r,g,b) | Tex
    flip_normals: bool = False
    center: Sequence[float] = (0.0, 0.0, 0.0)
    radius: float = 1.0
    flip_tex_coords: bool = True  # obj default, resource/shape.cpp:146
    mesh: Optional[dict] = None   # obj: positions (nv,3), indices (nf,3), normals?, texcoords?
    name: str = ""


@dataclass
class Sensor:
    fov: float = 90.0
    fov_axis: str = "x"
    near_clip: float = 0.01
    far_clip: float = 10000.0
    to_world: Optional[Xf] = None
    width: int = 768
    height: int = 576


@dataclass
class SceneDesc:
    max_depth: int = 1
    sensor: Sensor = field(default_factory=Sensor)
    shapes: list = field(default_factory=list)
    env_radiance: Optional[Sequence[float]] = None
    env_map: Optional[EnvMap] = None
    name: str = "scene"

    def num_triangles(self) -> int:
        n = 0
        for s in self.shapes:
            n += {"rectangle": 2, "cube": 12, "sphere": 0}.get(s.type, 0)
            if s.type == "obj":
                n += int(np.asarray(s.mesh["indices"]).shape[0])
        return n


# ------------------------------------------------------------------------------------------------
# generators
# ------------------------------------------------------------------------------------------------
def cornell_box(width: int = 512, height: int = 512, max_depth: int = 8) -> SceneDesc:
    """The Cornell box of the reference's data/static/cornellbox.xml (same matrices, colours and light);
    only film size and max_depth are parameters (BASELINE.json configs C1/C2 override them to depth 8)."""
    white = (0.725, 0.71, 0.68)

    def diffuse(c):
        return Bsdf("diffuse", twosided=True, params=dict(reflectance=c))

    def M(s):
        return Xf("matrix", matrix=[float(v) for v in s.split()])

    shapes = [
        Shape("rectangle", M("-4.37114e-008 1 4.37114e-008 0 0 -8.74228e-008 2 0 1 4.37114e-008 1.91069e-015 0 0 0 0 1"), diffuse(white), name="Floor"),
        Shape("rectangle", M("-1 7.64274e-015 -1.74846e-007 0 8.74228e-008 8.74228e-008 -2 2 0 -1 -4.37114e-008 0 0 0 0 1"), diffuse(white), name="Ceiling"),
        Shape("rectangle", M("1.91069e-015 1 1.31134e-007 0 1 3.82137e-015 -8.74228e-008 1 -4.37114e-008 1.31134e-007 -2 -1 0 0 0 1"), diffuse(white), name="BackWall"),
        Shape("rectangle", M("4.37114e-008 -1.74846e-007 2 1 1 3.82137e-015 -8.74228e-008 1 3.82137e-015 1 2.18557e-007 0 0 0 0 1"), diffuse((0.14, 0.45, 0.091)), name="RightWall"),
        Shape("rectangle", M("-4.37114e-008 8.74228e-008 -2 -1 1 3.82137e-015 -8.74228e-008 1 0 -1 -4.37114e-008 0 0 0 0 1"), diffuse((0.63, 0.065, 0.05)), name="LeftWall"),
        Shape("cube", M("0.0851643 0.289542 1.31134e-008 0.328631 3.72265e-009 1.26563e-008 -0.3 0.3 -0.284951 0.0865363 5.73206e-016 0.374592 0 0 0 1"), diffuse(white), name="ShortBox"),
        Shape("cube", M("0.286776 0.098229 -2.29282e-015 -0.335439 -4.36233e-009 1.23382e-008 -0.6 0.6 -0.0997984 0.282266 2.62268e-008 -0.291415 0 0 0 1"), diffuse(white), name="TallBox"),
        Shape("rectangle", M("0.235 -1.66103e-008 -7.80685e-009 -0.005 -2.05444e-008 3.90343e-009 -0.0893 1.98 2.05444e-008 0.19 8.30516e-009 -0.03 0 0 0 1"),
              diffuse((0.0, 0.0, 0.0)), emitter=(17.0, 12.0, 4.0), name="Light"),
    ]
    sensor = Sensor(fov=19.5, fov_axis="x", to_world=M("-1 0 0 0 0 1 0 1 0 0 -1 6.8 0 0 0 1"), width=width, height=height)
    return SceneDesc(max_depth=max_depth, sensor=sensor, shapes=shapes, name="cornell_box")


def _bsdf_cycle(i: int) -> Bsdf:
    """The seven BSDFs, rough ones at alpha in {0.05, 0.35, 0.95}; conductor eta/k and plastic
    reflectance are the values of the reference's data/static/material_test.xml."""
    cu_eta, cu_k = (0.200438, 0.924033, 1.10221), (3.91295, 2.45285, 2.14219)
    alpha = (0.05, 0.35, 0.95)[(i // 7) % 3]
    kind = MAT_NAMES[i % 7]
    if kind == "diffuse":
        cols = [(0.8, 0.25, 0.2), (0.2, 0.6, 0.8), (0.7, 0.7, 0.3)]
        return Bsdf("diffuse", params=dict(reflectance=cols[(i // 7) % 3]))
    if kind == "dielectric":
        return Bsdf("dielectric", params=dict(int_ior=1.5, ext_ior=1.0))
    if kind == "roughdielectric":
        return Bsdf("roughdielectric", params=dict(int_ior=1.5, ext_ior=1.0, alpha=alpha))
    if kind == "conductor":
        return Bsdf("conductor", params=dict(eta=cu_eta, k=cu_k, specular_reflectance=(0.9, 0.9, 0.9)))
    if kind == "roughconductor":
        return Bsdf("roughconductor", params=dict(eta=cu_eta, k=cu_k, alpha=alpha, specular_reflectance=(0.9, 0.9, 0.9)))
    if kind == "plastic":
        return Bsdf("plastic", params=dict(int_ior=1.5, ext_ior=1.0, nonlinear=bool((i // 7) % 2),
                                           diffuse_reflectance=(0.647814, 0.3, 0.2)))
    return Bsdf("roughplastic", params=dict(int_ior=1.5, ext_ior=1.0, alpha=alpha, nonlinear=False,
                                            diffuse_reflectance=(0.2, 0.647814, 0.3)))


def material_grid(width: int = 1920, height: int = 1080, max_depth: int = 8, nx: int = 7, nz: int = 5,
                  env=(1.0, 1.0, 1.0), with_area_light: bool = True) -> SceneDesc:
    """C3/C5: nx*nz analytic unit spheres cycling all seven BSDFs over a checkerboard floor, lit by a
    constant environment (and one small rectangular area light so both NEE paths are exercised)."""
    shapes = []
    spacing = 2.6
    x0, z0 = -(nx - 1) * spacing / 2, -(nz - 1) * spacing / 2
    for iz in range(nz):
        for ix in range(nx):
            i = iz * nx + ix
            shapes.append(Shape("sphere", center=(x0 + ix * spacing, 1.0, z0 + iz * spacing), radius=1.0, bsdf=_bsdf_cycle(i),
                                name=f"ball{i}"))
    floor_tex = Tex("checkerboard", color0=(0.8, 0.8, 0.8), color1=(0.15, 0.15, 0.15), uv_scale=(16.0, 16.0, 1.0))
    half = max(nx, nz) * spacing
    shapes.append(Shape("rectangle", Xf("srt", scale=(half, half, 1.0), rotate_axis=(1, 0, 0), rotate_angle=-90.0),
                        Bsdf("diffuse", params=dict(reflectance=floor_tex)), name="floor"))
    if with_area_light:
        shapes.append(Shape("rectangle", Xf("srt", scale=(2.0, 2.0, 1.0), rotate_axis=(1, 0, 0), rotate_angle=90.0, translate=(0.0, 9.0, 0.0)),
                            Bsdf("diffuse", params=dict(reflectance=(0.0, 0.0, 0.0))), emitter=(30.0, 28.0, 25.0), name="light"))
    cam = Xf("lookat", origin=(0.0, 9.0, 14.0 + nz), target=(0.0, 0.5, 0.0), up=(0, 1, 0))
    sensor = Sensor(fov=45.0, fov_axis="x", to_world=cam, width=width, height=height)
    return SceneDesc(max_depth=max_depth, sensor=sensor, shapes=shapes, env_radiance=env, name="material_grid")


def procedural_image(w: int, h: int, seed: int, hdr: bool = False) -> np.ndarray:
    """deterministic float32 (h, w, 4) texels: gradients + noise, plus a bright sun lobe when hdr (env maps)"""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    img = np.zeros((h, w, 4), np.float32)
    img[..., 0] = 0.25 + 0.5 * x / max(w - 1, 1)
    img[..., 1] = 0.2 + 0.6 * y / max(h - 1, 1)
    img[..., 2] = 0.5 + 0.35 * np.sin(x * 0.5) * np.cos(y * 0.8)
    img[..., :3] += rng.random((h, w, 3), dtype=np.float32) * np.float32(0.1)
    if hdr:
        img[..., :3] += (30.0 * np.exp(-((x - 0.7 * w) ** 2 + (y - 0.3 * h) ** 2) / (0.01 * w * w)))[..., None].astype(np.float32)
    img[..., 3] = 1.0
    return np.ascontiguousarray(img.astype(np.float32))


def envmap_scene(width: int = 1920, height: int = 1080, max_depth: int = 8, env_w: int = 64, env_h: int = 32) -> SceneDesc:
    """Bitmap-textured floor and spheres under a lat-long environment map (rows a18 / a19 of SURVEY.md 8: bitmap
    texture sampling and the EnvMapEmitter), with one small area light so both emitter kinds are selected."""
    floor_tex = Tex("bitmap", image=procedural_image(32, 32, 5), filter_type="bilinear", wrap_mode="repeat", uv_scale=(4.0, 4.0, 1.0))
    ball_tex = Tex("bitmap", image=procedural_image(16, 8, 6), filter_type="nearest", wrap_mode="mirror")
    shapes = [
        Shape("rectangle", Xf("srt", scale=(6.0, 6.0, 1.0), rotate_axis=(1, 0, 0), rotate_angle=-90.0), Bsdf("diffuse", True, dict(reflectance=floor_tex)), name="floor"),
        Shape("sphere", None, Bsdf("plastic", False, dict(diffuse_reflectance=ball_tex, int_ior=1.5, ext_ior=1.0)), center=(-1.6, 0.8, 0.0), radius=0.8, name="ball_tex"),
        Shape("sphere", None, Bsdf("roughconductor", False, dict(alpha=0.2, eta=(0.200438, 0.924033, 1.10221), k=(3.91295, 2.45285, 2.14219))),
              center=(0.2, 0.8, 0.3), radius=0.8, name="ball_metal"),
        Shape("sphere", None, Bsdf("dielectric", False, dict(int_ior=1.5, ext_ior=1.0)), center=(2.0, 0.8, -0.2), radius=0.8, name="ball_glass"),
        Shape("rectangle", Xf("srt", scale=(0.4, 0.4, 1.0), rotate_axis=(1, 0, 0), rotate_angle=90.0, translate=(0.0, 3.5, 0.0)),
              Bsdf("diffuse", True, dict(reflectance=(0.0, 0.0, 0.0))), emitter=(6.0, 5.0, 4.0), name="lamp"),
    ]
    sensor = Sensor(fov=45.0, fov_axis="x", to_world=Xf("lookat", origin=(0.0, 2.2, 6.5), target=(0.0, 0.7, 0.0), up=(0, 1, 0)), width=width, height=height)
    env = EnvMap(image=procedural_image(env_w, env_h, 7, hdr=True), scale=1.25, to_world=Xf("srt", rotate_axis=(0, 1, 0), rotate_angle=40.0))
    return SceneDesc(max_depth=max_depth, sensor=sensor, shapes=shapes, env_map=env, name="envmap_scene")


def heightfield_mesh(n: int, seed: int = 42, size: float = 20.0, amplitude: float = 1.2) -> dict:
    """(n+1)^2 vertices, 2*n*n triangles: a sum of seeded sinusoids (deterministic, no file IO)."""
    rng = np.random.default_rng(seed)
    k = 12
    freq = rng.uniform(0.3, 6.0, size=(k, 2)).astype(np.float32)
    phase = rng.uniform(0, 2 * math.pi, size=k).astype(np.float32)
    amp = (amplitude / (1.0 + np.arange(k, dtype=np.float32))).astype(np.float32)
    lin = np.linspace(-size / 2, size / 2, n + 1, dtype=np.float32)
    X, Z = np.meshgrid(lin, lin, indexing="xy")
    Y = np.zeros_like(X)
    dYdx = np.zeros_like(X)
    dYdz = np.zeros_like(X)
    for j in range(k):
        arg = freq[j, 0] * X + freq[j, 1] * Z + phase[j]
        Y += amp[j] * np.sin(arg)
        c = amp[j] * np.cos(arg)
        dYdx += c * freq[j, 0]
        dYdz += c * freq[j, 1]
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    nrm = np.stack([-dYdx, np.ones_like(X), -dYdz], -1).reshape(-1, 3)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    uv = np.stack([(X / size + 0.5), (Z / size + 0.5)], -1).reshape(-1, 2).astype(np.float32)
    i = np.arange(n, dtype=np.uint32)
    a = (i[:, None] * (n + 1) + i[None, :]).reshape(-1)  # (row z, col x) -> vertex z*(n+1)+x
    idx = np.empty((2 * n * n, 3), np.uint32)
    idx[0::2] = np.stack([a, a + (n + 1), a + 1], -1)          # counter-clockwise seen from +Y
    idx[1::2] = np.stack([a + 1, a + (n + 1), a + (n + 2)], -1)
    return dict(positions=pos, normals=nrm, texcoords=uv, indices=idx)


def terrain(n: int = 3873, width: int = 1920, height: int = 1080, max_depth: int = 8, seed: int = 42) -> SceneDesc:
    """C4: ~2*n*n tessellated triangles (n=3873 -> 30.0 M), diffuse, one area light + constant env."""
    mesh = heightfield_mesh(n, seed)
    shapes = [
        Shape("obj", None, Bsdf("diffuse", twosided=True, params=dict(reflectance=(0.6, 0.55, 0.45))), mesh=mesh,
              flip_tex_coords=False, name="terrain"),
        Shape("rectangle", Xf("srt", scale=(3.0, 3.0, 1.0), rotate_axis=(1, 0, 0), rotate_angle=90.0, translate=(0.0, 12.0, 0.0)),
              Bsdf("diffuse", params=dict(reflectance=(0.0, 0.0, 0.0))), emitter=(40.0, 38.0, 35.0), name="light"),
    ]
    cam = Xf("lookat", origin=(0.0, 7.0, 16.0), target=(0.0, 0.0, 0.0), up=(0, 1, 0))
    sensor = Sensor(fov=50.0, fov_axis="x", to_world=cam, width=width, height=height)
    return SceneDesc(max_depth=max_depth, sensor=sensor, shapes=shapes, env_radiance=(0.4, 0.5, 0.7), name=f"terrain_{2 * n * n}")


# ------------------------------------------------------------------------------------------------
# XML writer (the dialect of framework/resource/xml/* in the reference)
# ------------------------------------------------------------------------------------------------
def _f(v) -> str:
    return repr(float(v))


def _csv(v) -> str:
    return ", ".join(_f(x) for x in v)


def _tex_xml(name: str, t, ind: str) -> str:
    if isinstance(t, bool):
        return f'{ind}<boolean name="{name}" value="{"true" if t else "false"}" />\n'
    if isinstance(t, (int, float)):
        return f'{ind}<float name="{name}" value="{_f(t)}" />\n'
    if isinstance(t, (tuple, list, np.ndarray)):
        return f'{ind}<rgb name="{name}" value="{_csv(t)}" />\n'
    if t.kind == "rgb":
        return f'{ind}<rgb name="{name}" value="{_csv(t.color0)}" />\n'
    if t.kind == "bitmap":
        s = f'{ind}<texture name="{name}" type="bitmap">\n{ind}\t<string name="filename" value="{image_name(t)}" />\n'
        s += f'{ind}\t<string name="filter_type" value="{t.filter_type}" />\n{ind}\t<string name="wrap_mode" value="{t.wrap_mode}" />\n'
    else:
        s = f'{ind}<texture name="{name}" type="checkerboard">\n'
        s += f'{ind}\t<rgb name="color0" value="{_csv(t.color0)}" />\n{ind}\t<rgb name="color1" value="{_csv(t.color1)}" />\n'
    if t.uv_scale is not None:
        s += f'{ind}\t<transform name="to_uv">\n{ind}\t\t<scale x="{_f(t.uv_scale[0])}" y="{_f(t.uv_scale[1])}" z="{_f(t.uv_scale[2])}" />\n{ind}\t</transform>\n'
    return s + f"{ind}</texture>\n"


def image_name(t) -> str:
    """file name of a bitmap texture / env map: its own, or the key the in-memory image is registered under"""
    return t.filename if t.filename else f"mem:image_{id(t.image):x}"


def images_of(scene: "SceneDesc") -> list:
    """every Tex(kind='bitmap') / EnvMap of the scene that carries its texels in memory"""
    out = []
    for sh in scene.shapes:
        cands = [sh.emitter] + (list(sh.bsdf.params.values()) if sh.bsdf is not None else [])
        out += [c for c in cands if isinstance(c, Tex) and c.kind == "bitmap" and c.filename is None]
    if scene.env_map is not None and scene.env_map.filename is None:
        out.append(scene.env_map)
    return out


def _xf_xml(x: Optional[Xf], ind: str) -> str:
    if x is None or x.kind == "identity":
        return ""
    s = f'{ind}<transform name="to_world">\n'
    if x.kind == "matrix":
        s += f'{ind}\t<matrix value="{" ".join(_f(v) for v in x.matrix)}" />\n'
    elif x.kind == "lookat":
        s += f'{ind}\t<lookat origin="{_csv(x.origin)}" target="{_csv(x.target)}" up="{_csv(x.up)}" />\n'
    else:
        if x.scale is not None:
            s += f'{ind}\t<scale x="{_f(x.scale[0])}" y="{_f(x.scale[1])}" z="{_f(x.scale[2])}" />\n'
        if x.rotate_axis is not None:
            s += f'{ind}\t<rotate value="{_csv(x.rotate_axis)}" angle="{_f(x.rotate_angle)}" />\n'
        if x.translate is not None:
            s += f'{ind}\t<translate x="{_f(x.translate[0])}" y="{_f(x.translate[1])}" z="{_f(x.translate[2])}" />\n'
    return s + f"{ind}</transform>\n"


def _bsdf_xml(b: Optional[Bsdf], ind: str) -> str:
    if b is None:
        return ""
    inner_ind = ind + "\t" if b.twosided else ind
    s = f'{inner_ind}<bsdf type="{b.type}">\n'
    for k, v in b.params.items():
        s += _tex_xml(k, v, inner_ind + "\t")
    s += f"{inner_ind}</bsdf>\n"
    if b.twosided:
        s = f'{ind}<bsdf type="twosided">\n{s}{ind}</bsdf>\n'
    return s


def write_obj(path: Path, mesh: dict) -> None:
    P, I = np.asarray(mesh["positions"]), np.asarray(mesh["indices"])
    N, T = mesh.get("normals"), mesh.get("texcoords")
    with open(path, "w") as f:
        for p in P:
            f.write(f"v {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
        if T is not None:
            for t in np.asarray(T):
                f.write(f"vt {float(t[0])!r} {float(t[1])!r}\n")
        if N is not None:
            for n in np.asarray(N):
                f.write(f"vn {float(n[0])!r} {float(n[1])!r} {float(n[2])!r}\n")
        for a, b, c in I + 1:
            if T is not None and N is not None:
                f.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
            elif N is not None:
                f.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")
            elif T is not None:
                f.write(f"f {a}/{a} {b}/{b} {c}/{c}\n")
            else:
                f.write(f"f {a} {b} {c}\n")


def to_xml_string(scene: SceneDesc, obj_names: dict) -> str:
    """`scene` in the mitsuba3-style dialect the reference's loader accepts.  obj_names maps shape index ->
    the file name (or "mem:KEY") written for that obj shape."""
    s = '<scene version="3.0.0">\n'
    s += f'\t<default name="max_depth" value="{int(scene.max_depth)}" />\n'
    s += f'\t<default name="resx" value="{int(scene.sensor.width)}" />\n\t<default name="resy" value="{int(scene.sensor.height)}" />\n'
    s += '\t<integrator type="path">\n\t\t<integer name="max_depth" value="$max_depth" />\n\t</integrator>\n'
    se = scene.sensor
    s += '\t<sensor type="perspective">\n'
    s += f'\t\t<float name="fov" value="{_f(se.fov)}" />\n\t\t<string name="fov_axis" value="{se.fov_axis}" />\n'
    s += f'\t\t<float name="near_clip" value="{_f(se.near_clip)}" />\n\t\t<float name="far_clip" value="{_f(se.far_clip)}" />\n'
    s += _xf_xml(se.to_world, "\t\t")
    s += '\t\t<film type="hdrfilm">\n\t\t\t<integer name="width" value="$resx" />\n\t\t\t<integer name="height" value="$resy" />\n\t\t</film>\n'
    s += "\t</sensor>\n"
    for i, sh in enumerate(scene.shapes):
        s += f'\t<shape type="{sh.type}" id="{sh.name or f"shape{i}"}">\n'
        if sh.type == "obj":
            s += f'\t\t<string name="filename" value="{obj_names[i]}" />\n'
            s += f'\t\t<boolean name="flip_tex_coords" value="{"true" if sh.flip_tex_coords else "false"}" />\n'
        if sh.type == "sphere":
            s += f'\t\t<point name="center" x="{_f(sh.center[0])}" y="{_f(sh.center[1])}" z="{_f(sh.center[2])}" />\n'
            s += f'\t\t<float name="radius" value="{_f(sh.radius)}" />\n'
        if sh.flip_normals:
            s += '\t\t<boolean name="flip_normals" value="true" />\n'
        s += _xf_xml(sh.to_world, "\t\t")
        s += _bsdf_xml(sh.bsdf, "\t\t")
        if sh.emitter is not None:
            s += '\t\t<emitter type="area">\n' + _tex_xml("radiance", sh.emitter, "\t\t\t") + "\t\t</emitter>\n"
        s += "\t</shape>\n"
    if scene.env_radiance is not None:
        s += f'\t<emitter type="constant">\n\t\t<rgb name="radiance" value="{_csv(scene.env_radiance)}" />\n\t</emitter>\n'
    if scene.env_map is not None:
        em = scene.env_map
        s += f'\t<emitter type="envmap">\n\t\t<string name="filename" value="{image_name(em)}" />\n\t\t<float name="scale" value="{_f(em.scale)}" />\n'
        s += _xf_xml(em.to_world, "\t\t") + "\t</emitter>\n"
    s += "</scene>\n"
    return s


def to_xml(scene: SceneDesc, path) -> Path:
    """Write `scene` as an XML file; obj meshes are written next to it as Wavefront .obj files."""
    path = Path(path)
    names = {}
    for i, sh in enumerate(scene.shapes):
        if sh.type == "obj":
            names[i] = f"{path.stem}_{i}.obj"
            write_obj(path.parent / names[i], sh.mesh)
    path.write_text(to_xml_string(scene, names))
    return path
