"""CPU-only: the C++ host library (pupiloptixlab_b200/host, include/pupil_host.h) — XML dialect, transforms,
camera matrices, material precompute, emitter table — against the oracle's independent restatement of the
same reference code (framework/resource, framework/world, framework/util).  Bit-exact: both are host fp32."""
import os
from pathlib import Path

import numpy as np
import pytest

import orc
from pupiloptixlab_b200 import pb2, pupil, scenes

REF_DATA = Path(os.environ.get("PUPIL_REF", "/root/reference")) / "data" / "static"


def _tex_equal(p: pb2.Texture, o: orc.Texture):
    return (p.type == o.type and list(p.a) == list(o.a) and (p.type != pb2.TEX_CHECKERBOARD or list(p.b) == list(o.b))
            and list(p.r0) == list(o.to_uv[0:4]) and list(p.r1) == list(o.to_uv[4:8]))


def _check_against_oracle(desc, port_lib):
    pupil.load_scene(desc, host_only=True)
    o = orc.OracleScene(port_lib, desc)
    s2c, c2w, fov = pupil.camera()
    os2c, oc2w, ofov = o.camera()
    assert np.array_equal(s2c, os2c) and np.array_equal(c2w, oc2w) and fov == ofov
    assert pupil.film() == (desc.sensor.width, desc.sensor.height, desc.max_depth)
    ins = pupil.instances()
    assert len(ins) == len(desc.shapes)
    for i, (inst, sh) in enumerate(zip(ins, desc.shapes)):
        assert np.array_equal(inst["xform"], o.instance_xform(i)), f"instance {i} transform"
        assert inst["is_sphere"] == (sh.type == "sphere")
        # material precompute (optix_material.cpp:41-130)
        m = inst["material"]
        om = orc.make_material(sh.bsdf)
        eta, fdr, w = np.zeros(1, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float32)
        port_lib.orc_load_material(om, orc.fp(eta), orc.fp(fdr), orc.fp(w))
        assert m.type == om.type and bool(m.twosided) == bool(om.twosided)
        name = {v: k for k, v in orc.MAT.items()}[om.type]
        if name not in ("diffuse", "conductor", "roughconductor", "unknown"):
            assert m.eta == eta[0]
        if name in ("plastic", "roughplastic"):
            assert m.int_fdr == fdr[0] and m.specular_sampling_weight == w[0] and bool(m.nonlinear) == bool(om.nonlinear)
        slots = dict(diffuse=["reflectance"], dielectric=["specular_reflectance", "specular_transmittance"],
                     roughdielectric=["specular_reflectance", "specular_transmittance", "alpha"],
                     conductor=["specular_reflectance", "eta", "k"], roughconductor=["specular_reflectance", "eta", "k", "alpha"],
                     plastic=["reflectance", "specular_reflectance"], roughplastic=["reflectance", "specular_reflectance", "alpha"]).get(name, [])
        for k, slot in enumerate(slots):
            assert _tex_equal(m.tex[k], getattr(om, slot)), f"instance {i} texture slot {slot}"
    areas, env = pupil.emitters()
    oareas = o.area_emitters()
    assert len(areas) == len(oareas)
    for a, b in zip(areas, oareas):
        assert a.type == b.type and a.weight == b.weight and a.select_probability == b.select_probability and a.area == b.area
        assert _tex_equal(a.radiance, b.radiance)
        for k in range(3):
            assert list(a.pos[k]) == list(b.pos[k]) and list(a.nrm[k]) == list(b.nrm[k]) and list(a.uv[k]) == list(b.uv[k])
        assert list(a.center) == list(b.center) and a.radius == b.radius
    oenv = orc.Emitter()
    has = port_lib.orc_get_env_emitter(o.h, oenv)
    assert (env is not None) == bool(has)
    if env is not None:
        assert env.type == oenv.type and list(env.radiance.a) == list(oenv.radiance.a) and env.select_probability == oenv.select_probability
    if env is not None and env.type == pb2.EMIT_ENV_MAP:  # EmitterHelper::AddEmitter + BuildEnvMapCdfTable, world/emitter.cpp:107-149,293-312
        assert env.radiance.type == pb2.TEX_BITMAP and oenv.radiance.type == orc.TEX_BITMAP
        assert (env.map_w, env.map_h) == (oenv.map_w, oenv.map_h) and env.scale == oenv.scale and env.normalization == oenv.normalization
        assert list(env.to_world) == list(oenv.to_world) and list(env.to_local) == list(oenv.to_local)
        rc, rw, cc = pupil.env_tables()
        h, w = env.map_h, env.map_w
        assert np.array_equal(rc, np.ctypeslib.as_array(oenv.row_cdf, (h + 1,)))
        assert np.array_equal(rw, np.ctypeslib.as_array(oenv.row_weight, (h,)))
        assert np.array_equal(cc.reshape(-1), np.ctypeslib.as_array(oenv.col_cdf, (h * (w + 1),)))
    # emitter offsets: running sum over emitting instances (pt_pass.cpp:178-193)
    off = 0
    for inst, sh in zip(ins, desc.shapes):
        if sh.emitter is not None:
            assert inst["emitter_offset"] == off
            off += inst["n_prims"]
        else:
            assert inst["emitter_offset"] == -1
    return o


@pytest.mark.parametrize("maker", [lambda: scenes.cornell_box(64, 48, 8), lambda: scenes.material_grid(96, 54, 6),
                                   lambda: scenes.terrain(24, 80, 45, 5), lambda: scenes.envmap_scene(64, 36, 6)],
                         ids=["cornell", "material_grid", "terrain", "envmap"])
def test_world_precompute_matches_oracle(port_lib, maker):
    port_lib.orc_get_env_emitter.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").POINTER(orc.Emitter)]
    _check_against_oracle(maker(), port_lib)


def test_sphere_emitter_and_flip_flags(port_lib):
    port_lib.orc_get_env_emitter.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").POINTER(orc.Emitter)]
    d = scenes.material_grid(32, 18, 4, nx=2, nz=1)
    d.shapes.append(scenes.Shape("sphere", scenes.Xf("srt", scale=(1.0, 2.0, 1.0), translate=(0, 5, 0)), scenes.Bsdf("diffuse"), emitter=(5.0, 4.0, 3.0),
                                 center=(0.5, 0.0, 0.0), radius=0.25, name="bulb"))
    d.shapes.append(scenes.Shape("rectangle", scenes.Xf("srt", translate=(0, 1, 0)), scenes.Bsdf("diffuse"), flip_normals=True, name="flipped"))
    _check_against_oracle(d, port_lib)
    ins = pupil.instances()
    # rectangle is a shared singleton: the LAST <shape type="rectangle"> decides flip_normals for all of them (shape.cpp:105)
    rect_flags = [i["flags"] & pb2.INST_FLIP_NORMALS for i, s in zip(ins, d.shapes) if s.type == "rectangle"]
    assert rect_flags and all(rect_flags)


XML = """<?xml version="1.0" encoding="utf-8"?>
<!-- dialect probe -->
<scene version="3.0.0">
  <default name="res" value="40"/>
  <default name="resx" value="64"/>
  <default name="depth" value="7"/>
  <integrator type="path"><integer name="max_depth" value="$depth"/></integrator>
  <bsdf type="twosided" id="red"><bsdf type="diffuse"><rgb name="reflectance" value="0.5, 0.1, 0.1"/></bsdf></bsdf>
  <bsdf type="roughconductor" id="silver"><string name="material" value="Ag"/><float name="alpha" value="0.2"/></bsdf>
  <sensor type="perspective">
    <float name="fov" value="39.3077"/><string name="fov_axis" value="y"/>
    <transform name="to_world"><lookat origin="0, 1, 5" target="0, 1, 0" up="0, 1, 0"/></transform>
    <sampler type="independent"><integer name="sample_count" value="64"/></sampler>
    <film type="hdrfilm"><integer name="width" value="$resx"/><integer name="height" value="$res"/>
      <rfilter type="gaussian"/></film>
  </sensor>
  <shape type="rectangle" id="a"><ref id="red"/>
    <transform name="to_world"><translate x="1" y="2" z="3"/><rotate y="1" angle="90"/><scale value="2"/></transform></shape>
  <shape type="cube" id="b"><ref id="silver"/>
    <transform name="to_world"><matrix value="1 0 0 0 1 0 0 0 2"/></transform></shape>
  <shape type="sphere" id="c"><point name="center" x="1" y="0.5" z="0"/><float name="radius" value="0.5"/>
    <bsdf type="dielectric"><string name="int_ior" value="water"/><float name="ext_ior" value="1"/></bsdf>
    <emitter type="area"><rgb name="radiance" value="3"/></emitter></shape>
  <shape type="hair" id="skipped"><string name="filename" value="x.hair"/></shape>
  <emitter type="constant"><rgb name="radiance" value="0.25, 0.5, 1"/></emitter>
  <emitter type="point"><point name="position" x="0" y="0" z="0"/></emitter>
</scene>
"""


def test_xml_dialect():
    check = pupil.check
    check(pupil.lib().pupil_parse_scene_xml_string(XML.encode(), None))
    assert pupil.film() == (64, 40, 7)          # $resx is not eaten by $res; $depth substituted
    s2c, c2w, fov = pupil.camera()
    assert abs(fov - 39.3077) < 1e-5            # fov_axis = y: unchanged
    # lookat from (0,1,5) toward -z: camera looks down its own -Z; after the double handedness flip (loader +
    # sensor) the matrix is the plain look-at frame
    assert np.allclose(c2w[:3, 3], [0, 1, 5]) and np.allclose(c2w[:3, :3], np.eye(3), atol=1e-6)
    ins = pupil.instances()
    assert len(ins) == 3                        # hair skipped, unknown <sampler>/<rfilter> ignored
    # fixed S -> R -> T order regardless of element order: p' = T R S p
    a = ins[0]["xform"]
    c, s = np.cos(np.pi / 2), np.sin(np.pi / 2)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)
    assert np.allclose(a[:3, :3], R * 2, atol=1e-6) and np.allclose(a[:3, 3], [1, 2, 3])
    m = ins[0]["material"]
    assert m.type == pb2.MAT["diffuse"] and m.twosided == 1 and np.allclose(list(m.tex[0].a), [0.5, 0.1, 0.1])
    b = ins[1]
    assert np.allclose(b["xform"][:3, :3], np.diag([1, 1, 2])) and b["xform"][3, 3] == 1   # 9-value matrix -> upper-left 3x3
    assert b["material"].type == pb2.MAT["roughconductor"]
    assert np.allclose(list(b["material"].tex[1].a), [0.15494, 0.11648, 0.13809]) and np.allclose(list(b["material"].tex[3].a), [0.2] * 3)
    sph = ins[2]
    assert sph["is_sphere"] and np.allclose(sph["xform"][:3, :3], np.eye(3) * 0.5) and np.allclose(sph["xform"][:3, 3], [1, 0.5, 0])
    assert abs(sph["material"].eta - 1.333) < 1e-6
    areas, env = pupil.emitters()
    assert len(areas) == 1 and areas[0].type == pb2.EMIT_SPHERE and abs(areas[0].radius - 0.5) < 1e-6 and list(areas[0].radiance.a) == [3, 3, 3]
    assert env is not None and env.type == pb2.EMIT_CONST_ENV and list(env.radiance.a) == [0.25, 0.5, 1.0]
    assert abs(areas[0].select_probability - 0.5) < 1e-7 and abs(env.select_probability - 0.5) < 1e-7


def test_xml_errors_are_reported():
    L = pupil.lib()
    assert L.pupil_parse_scene_xml_string(b"<scene><shape type='cube'></scene>", None) != 0
    assert L.pupil_parse_scene_xml_string(b"<notascene/>", None) != 0
    assert L.pupil_parse_scene_xml(b"/nonexistent/scene.xml") != 0
    assert b"exist" in L.pupil_last_error()


def test_obj_round_trip(tmp_path, port_lib):
    """to_xml writes Wavefront .obj files; the host's own reader must return the same triangles in the same order"""
    d = scenes.terrain(6, 32, 18, 3)
    path = scenes.to_xml(d, tmp_path / "t.xml")
    pupil.parse_scene_xml(path)
    ins = pupil.instances()
    assert ins[0]["n_prims"] == 72 and ins[1]["n_prims"] == 2
    areas, env = pupil.emitters()
    assert len(areas) == 2 and env is not None
    o = orc.OracleScene(port_lib, d)
    assert np.array_equal(ins[1]["xform"], o.instance_xform(1))


@pytest.mark.skipif(not REF_DATA.is_dir(), reason="reference data directory absent")
@pytest.mark.parametrize("name,n_inst,n_area,has_env", [("cornellbox.xml", 8, 2, False), ("mis.xml", None, None, None),
                                                        ("material_test.xml", None, None, None), ("restir_test.xml", None, 5, None),
                                                        ("default.xml", None, None, None)])
def test_reference_scene_files_load_unchanged(name, n_inst, n_area, has_env):
    pupil.lib().pupil_set_log_level(0)
    try:
        pupil.parse_scene_xml(REF_DATA / name)
    finally:
        pupil.lib().pupil_set_log_level(1)
    ins = pupil.instances()
    areas, env = pupil.emitters()
    assert len(ins) > 0 and len(areas) > 0
    if n_inst is not None:
        assert len(ins) == n_inst
    if n_area is not None:
        assert len(areas) == n_area
    if has_env is not None:
        assert (env is not None) == has_env
    total_p = sum(a.select_probability for a in areas) + (env.select_probability if env is not None else 0.0)
    assert abs(total_p - 1.0) < 1e-5
    if name == "cornellbox.xml":
        w, h, depth = pupil.film()
        assert (w, h) == (512, 512)
        assert all(i["material"].type == pb2.MAT["diffuse"] and i["material"].twosided for i in ins)


def test_instance_transform_update_matches_a_fresh_scene(port_lib):
    """RenderObject::UpdateTransform -> World's RenderInstanceTransform handler (world.cpp:15-43): the object's area emitters
    are rebuilt in place and the selection probabilities recomputed, exactly as if the scene had been loaded that way"""
    desc = scenes.cornell_box(48, 48, 6)
    pupil.load_scene(desc, host_only=True)
    lamp = [i for i, sh in enumerate(desc.shapes) if sh.emitter is not None][0]
    moved = scenes.Xf("srt", scale=(0.3, 0.2, 1.0), rotate_axis=(1, 0, 0), rotate_angle=80.0, translate=(0.1, 1.7, 0.2))
    fresh = scenes.cornell_box(48, 48, 6)
    fresh.shapes[lamp].to_world = moved
    o = orc.OracleScene(port_lib, fresh)
    pupil.set_instance_transform(lamp, o.instance_xform(lamp))
    ins = pupil.instances()
    assert np.array_equal(ins[lamp]["xform"], o.instance_xform(lamp))
    areas, _ = pupil.emitters()
    for a, b in zip(areas, o.area_emitters()):
        assert a.weight == b.weight and a.select_probability == b.select_probability and a.area == b.area
        for k in range(3):
            assert list(a.pos[k]) == list(b.pos[k]) and list(a.nrm[k]) == list(b.nrm[k])


def _two_light_scene():
    d = scenes.cornell_box(48, 48, 6)
    # a second emitting object in FRONT of the ceiling lamp in the object list, plus one behind it
    d.shapes.insert(0, scenes.Shape("cube", scenes.Xf("srt", scale=(0.1, 0.1, 0.1), translate=(-0.5, 0.3, 0.2)), scenes.Bsdf("diffuse"), emitter=(9.0, 2.0, 1.0), name="glow_cube"))
    d.shapes.append(scenes.Shape("sphere", scenes.Xf("srt", translate=(0.4, 0.3, 0.3)), scenes.Bsdf("diffuse"), emitter=(1.0, 5.0, 2.0), center=(0, 0, 0), radius=0.1, name="bulb"))
    return d


def test_removing_an_emitting_object_takes_its_emitters_out_of_the_table(port_lib):
    """World::RemoveRenderObject: the removed object's entries leave EmitterHelper's table, later objects' offsets move down
    and the selection probabilities are those of a scene that never held the object (world.cpp)"""
    d = _two_light_scene()
    pupil.load_scene(d, host_only=True)
    before = pupil.instances()
    assert [i["emitter_offset"] for i in before if i["emitter_offset"] >= 0] == [0, 12, 14]
    pupil.remove_instance(0)
    fresh = _two_light_scene()
    del fresh.shapes[0]
    o = orc.OracleScene(port_lib, fresh)
    ins = pupil.instances()
    assert len(ins) == len(fresh.shapes)
    assert [i["emitter_offset"] for i in ins if i["emitter_offset"] >= 0] == [0, 2]
    areas, _ = pupil.emitters()
    oareas = o.area_emitters()
    assert len(areas) == len(oareas) == 3
    for a, b in zip(areas, oareas):
        assert a.type == b.type and a.weight == b.weight and a.select_probability == b.select_probability and a.area == b.area
        for k in range(3):
            assert list(a.pos[k]) == list(b.pos[k])
    # removing the LAST emitter shifts nothing
    pupil.remove_instance(len(ins) - 1)
    assert [i["emitter_offset"] for i in pupil.instances() if i["emitter_offset"] >= 0] == [0]
    assert len(pupil.emitters()[0]) == 2
    with pytest.raises(pupil.PupilError):
        pupil.remove_instance(99)


def test_hostile_xml_fails_instead_of_crashing():
    """element nesting is capped (the reader recurses per level), numeric character references decode, a film without
    pixels falls back to the default size (xml.cpp / Scene::LoadXmlObj)"""
    L = pupil.lib()
    deep = b"<scene>" + b"<a>" * 200000 + b"</a>" * 200000 + b"</scene>"
    assert L.pupil_parse_scene_xml_string(deep, None) != 0
    ok_depth = b"<scene version='3.0.0'>" + b"<x>" * 100 + b"</x>" * 100 + b"<shape type='cube' id='c&#49;&#x32;'/></scene>"
    assert L.pupil_parse_scene_xml_string(ok_depth, None) == 0
    assert len(pupil.instances()) == 1
    bad_film = XML.replace('value="$resx"', 'value="0"').replace('value="$res"', 'value="-5"')
    pupil.lib().pupil_set_log_level(0)
    try:
        assert L.pupil_parse_scene_xml_string(bad_film.encode(), None) == 0
    finally:
        pupil.lib().pupil_set_log_level(1)
    w, h, _ = pupil.film()
    assert (w, h) == (768, 576)
    s2c, _, _ = pupil.camera()
    assert np.isfinite(s2c).all()
