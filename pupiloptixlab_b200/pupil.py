"""ctypes binding of libpupil_host.so (include/pupil_host.h): the reference's host surface —
System::Init / AddPass(PTPass) / SetScene / Run, World, BufferManager — re-implemented in C++ under
pupiloptixlab_b200/host/ on top of the pb2 CUDA back end.

    from pupiloptixlab_b200 import pupil, scenes
    pupil.init(0)
    pupil.load_scene(scenes.cornell_box(512, 512, 8))       # or pupil.load_scene_xml("cornellbox.xml")
    pupil.run(64)                                            # 64 x PTPass::OnRun = 64 spp
    img = pupil.buffer("final result")                       # (H, W, 4) float32, row 0 = bottom

No CPU path exists: without the built libraries or without a CUDA device, init() raises.
"""
from __future__ import annotations

import os
import ctypes as C
from pathlib import Path

import numpy as np

from . import pb2
from .scenes import SceneDesc, to_xml_string

LIB_PATH = pb2.PKG / os.environ.get("PB2_BUILD_DIR", "_build") / "libpupil_host.so"
u32, i32, f32, u64 = C.c_uint32, C.c_int32, C.c_float, C.c_uint64
_lib = None
_mesh_serial = 0
_keepalive = []


class PupilError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        pb2.lib()  # libpb2.so first (also resolved through the rpath)
        if not LIB_PATH.exists():
            raise PupilError(f"{LIB_PATH} not found: run `python -m pupiloptixlab_b200.build`")
        L = C.CDLL(str(LIB_PATH))
        P, vp = C.POINTER, C.c_void_p
        L.pupil_last_error.restype = C.c_char_p
        sigs = {
            "pupil_init": [C.c_int], "pupil_shutdown": [], "pupil_set_log_level": [C.c_int], "pupil_load_scene_xml": [C.c_char_p],
            "pupil_load_scene_xml_string": [C.c_char_p, C.c_char_p],
            "pupil_parse_scene_xml": [C.c_char_p], "pupil_parse_scene_xml_string": [C.c_char_p, C.c_char_p], "pupil_register_mesh": [C.c_char_p, vp, vp, vp, vp, u32, u32],
            "pupil_clear_shapes": [], "pupil_pass_config": [C.c_int, C.c_int, u32, u32, u32, C.c_int], "pupil_run": [u64],
            "pupil_pass_state": [P(u32), P(u32)], "pupil_buffer_info": [C.c_char_p, P(vp), P(u32), P(u32), P(u32)],
            "pupil_buffer_download": [C.c_char_p, vp, u64], "pupil_buffer_upload": [C.c_char_p, vp, u64],
            "pupil_get_film": [P(u32), P(u32), P(u32)], "pupil_get_camera": [P(f32), P(f32), P(f32)], "pupil_num_instances": [],
            "pupil_get_instance": [u32, P(f32), P(pb2.Material), P(i32), P(u32), P(u32), P(i32)], "pupil_num_area_emitters": [],
            "pupil_get_emitters": [P(pb2.Emitter), P(pb2.Emitter), P(i32)], "pupil_scene_handle": [P(vp)], "pupil_set_bvh_builder": [C.c_int], "pupil_set_instancing": [C.c_int],
            "pupil_build_stats": [P(pb2.BuildStats)], "pupil_render_stats": [P(pb2.RenderStats)], "pupil_camera_move": [f32, f32, f32],
            "pupil_camera_rotate": [f32, f32], "pupil_camera_set_fov": [f32],
            "pupil_register_image": [C.c_char_p, vp, u32, u32], "pupil_image_load": [C.c_char_p, P(u32), P(u32), vp, u64],
            "pupil_image_save": [C.c_char_p, vp, u32, u32, C.c_int], "pupil_save_buffer": [C.c_char_p, C.c_char_p, C.c_int],
            "pupil_get_env_tables": [P(u32), P(u32), vp, vp, vp], "pupil_set_instance_transform": [u32, P(f32)],
            "pupil_remove_instance": [u32], "pupil_comm_unique_id": [vp], "pupil_set_shard": [C.c_int, C.c_int, vp, C.c_int, C.c_int],
            "pupil_synchronize": [], "pupil_set_shard_plan": [C.c_int], "pupil_last_reduction": [P(f32), P(u64)],
            "pupil_register_mesh_borrowed": [C.c_char_p, vp, vp, vp, vp, u32, u32], "pupil_unregister_mesh": [C.c_char_p],
            "pupil_checkpoint_save": [C.c_char_p], "pupil_checkpoint_load": [C.c_char_p],
        }
        for name, args in sigs.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, C.c_int
        _lib = L
    return _lib


def check(code: int):
    if code != 0:
        raise PupilError(lib().pupil_last_error().decode())


def init(device: int = 0, log_level: int = 1):
    L = lib()
    if pb2.lib().pb2_device_count() <= 0:
        raise PupilError("no CUDA device visible: there is no CPU fallback")
    L.pupil_set_log_level(log_level)
    check(L.pupil_init(device))


def shutdown():
    lib().pupil_shutdown()
    lib().pupil_clear_shapes()
    _keepalive.clear()


def load_scene_xml(path):
    check(lib().pupil_load_scene_xml(str(path).encode()))


def parse_scene_xml(path):
    """Host-only: parse + World precompute without a device (inspection through camera() / instances() / emitters())."""
    check(lib().pupil_parse_scene_xml(str(path).encode()))


_borrowed = {}  # key -> the arrays a borrowed registration points at (kept alive here), of the scene loaded last


def load_scene(desc: SceneDesc, host_only: bool = False, borrow: bool = True):
    """SceneDesc -> XML text (in memory) -> the host library's loader; triangle meshes go across as arrays.
    borrow: the library points at the arrays instead of copying them (pupil_register_mesh_borrowed); this module keeps them
    alive until the next scene is loaded.  Arrays that live in pinned memory (e.g. numpy views of pinned torch tensors) make
    the upload a straight DMA."""
    global _mesh_serial, _borrowed
    names, keep, by_mesh = {}, {}, {}
    for i, sh in enumerate(desc.shapes):
        if sh.type != "obj":
            continue
        if id(sh.mesh) in by_mesh:  # shapes that share a mesh share ONE registered shape: the device holds it once (bottom-level tree + instances)
            names[i] = by_mesh[id(sh.mesh)]
            continue
        _mesh_serial += 1
        key = f"mem:{desc.name}_{i}_{_mesh_serial}"
        m = sh.mesh
        P = np.ascontiguousarray(m["positions"], np.float32).reshape(-1, 3)
        I = np.ascontiguousarray(m["indices"], np.uint32).reshape(-1, 3)
        N = None if m.get("normals") is None else np.ascontiguousarray(m["normals"], np.float32)
        T = None if m.get("texcoords") is None else np.ascontiguousarray(m["texcoords"], np.float32)
        fn = lib().pupil_register_mesh_borrowed if borrow else lib().pupil_register_mesh
        check(fn(key.encode(), pb2._ptr(P), pb2._ptr(N), pb2._ptr(T), pb2._ptr(I), P.shape[0], I.shape[0]))
        names[i] = by_mesh[id(sh.mesh)] = key
        if borrow:
            keep[key] = (P, I, N, T)
    from . import scenes as _scenes
    for t in _scenes.images_of(desc):  # bitmap textures / env maps that carry their texels in memory
        img = np.ascontiguousarray(t.image, np.float32)
        check(lib().pupil_register_image(_scenes.image_name(t).encode(), pb2._ptr(img), img.shape[1], img.shape[0]))
    fn = lib().pupil_parse_scene_xml_string if host_only else lib().pupil_load_scene_xml_string
    try:
        check(fn(to_xml_string(desc, names).encode(), None))
    finally:
        for key in _borrowed:  # the previous scene's render objects are gone now (or the load failed and left no scene)
            lib().pupil_unregister_mesh(key.encode())
        _borrowed = keep


IMAGE_FORMATS = dict(hdr=0, exr=1, pfm=2, png=3, png_aces=4)


def set_instance_transform(index: int, xform):
    """RenderObject::UpdateTransform: row-major 4x4 object-to-world matrix of the index-th render object"""
    m = np.ascontiguousarray(xform, np.float32).reshape(16)
    check(lib().pupil_set_instance_transform(index, m.ctypes.data_as(C.POINTER(f32))))


def remove_instance(index: int):
    """World::RemoveRenderObject: the index-th render object leaves the scene, with its area emitters"""
    check(lib().pupil_remove_instance(index))


def image_load(path) -> np.ndarray:
    """util::BitmapTexture::Load: float32 (H, W, 4), row 0 = first row of the file; 8-bit sources linearised (gamma 2.2)"""
    w, h = u32(), u32()
    check(lib().pupil_image_load(str(path).encode(), C.byref(w), C.byref(h), None, 0))
    out = np.zeros((h.value, w.value, 4), np.float32)
    check(lib().pupil_image_load(str(path).encode(), C.byref(w), C.byref(h), pb2._ptr(out), out.size))
    return out


def image_save(path, rgba: np.ndarray, fmt: str | None = None):
    """util::BitmapTexture::Save: rgba (H, W, 4) with row 0 = BOTTOM of the picture (frame-buffer order)"""
    img = np.ascontiguousarray(rgba, np.float32)
    fmt = fmt or str(path).rsplit(".", 1)[-1].lower()
    check(lib().pupil_image_save(str(path).encode(), pb2._ptr(img), img.shape[1], img.shape[0], IMAGE_FORMATS[fmt]))


def save_buffer(name: str, path, fmt: str | None = None):
    fmt = fmt or str(path).rsplit(".", 1)[-1].lower()
    check(lib().pupil_save_buffer(name.encode(), str(path).encode(), IMAGE_FORMATS[fmt]))


def env_tables():
    """(row_cdf[h+1], row_weight[h], col_cdf[h, w+1]) of the scene's env-map emitter"""
    w, h = u32(), u32()
    check(lib().pupil_get_env_tables(C.byref(w), C.byref(h), None, None, None))
    rc, rw, cc = np.zeros(h.value + 1, np.float32), np.zeros(h.value, np.float32), np.zeros((h.value, w.value + 1), np.float32)
    check(lib().pupil_get_env_tables(C.byref(w), C.byref(h), pb2._ptr(rc), pb2._ptr(rw), pb2._ptr(cc)))
    return rc, rw, cc


def pass_config(max_depth: int = 0, accumulate: bool = True, frames_per_run: int = 1, first_seed: int = 0, seed_stride: int = 1,
                sum_mode: bool = False):
    check(lib().pupil_pass_config(max_depth, int(accumulate), frames_per_run, first_seed, seed_stride, int(sum_mode)))


REDUCE_ROOT, REDUCE_ALL = 0, 1


def comm_unique_id() -> bytes:
    """rank 0: the 128-byte NCCL id the other ranks need for set_shard (hand it over with any out-of-band channel)"""
    buf = (C.c_uint8 * 128)()
    check(lib().pupil_comm_unique_id(buf))
    return bytes(buf)


def set_shard(rank: int, world: int, comm_id: bytes | None, strong: bool = False, reduce_mode: int = REDUCE_ROOT):
    """collective: this process becomes rank `rank` of `world` (one process per GPU); world <= 0 switches sharding off"""
    buf = (C.c_uint8 * 128)(*comm_id) if comm_id else None
    check(lib().pupil_set_shard(rank, world, buf, int(strong), reduce_mode))


def set_shard_plan(strong: bool):
    check(lib().pupil_set_shard_plan(int(strong)))


def last_reduction():
    """(device ms, bytes of one rank's sum buffer) of the last multi-GPU reduction; waits for it"""
    ms, nbytes = f32(), u64()
    check(lib().pupil_last_reduction(C.byref(ms), C.byref(nbytes)))
    return ms.value, nbytes.value


def synchronize():
    check(lib().pupil_synchronize())


def run(n_pass_runs: int = 1):
    check(lib().pupil_run(n_pass_runs))


def checkpoint_save(path):
    """PTPass::SaveCheckpoint: accum + frame buffers, sample count, next seed, pass settings"""
    check(lib().pupil_checkpoint_save(str(path).encode()))


def checkpoint_load(path):
    """PTPass::LoadCheckpoint (the checkpoint's scene must be loaded): the next run() continues where the checkpoint stopped"""
    check(lib().pupil_checkpoint_load(str(path).encode()))


def pass_state():
    a, b = u32(), u32()
    check(lib().pupil_pass_state(C.byref(a), C.byref(b)))
    return a.value, b.value


def buffer_info(name: str):
    p, w, h, s = C.c_void_p(), u32(), u32(), u32()
    check(lib().pupil_buffer_info(name.encode(), C.byref(p), C.byref(w), C.byref(h), C.byref(s)))
    return p.value, w.value, h.value, s.value


def buffer(name: str) -> np.ndarray:
    """Download a named buffer as (H, W, C) float32 (row 0 = bottom of the image)."""
    _, w, h, stride = buffer_info(name)
    out = np.empty((h, w, stride // 4), np.float32)
    check(lib().pupil_buffer_download(name.encode(), pb2._ptr(out), out.nbytes))
    return out


def buffer_into(name: str, host_ptr: int, nbytes: int):
    """Download a named buffer into caller-owned host memory (pinned memory makes the copy a straight DMA)."""
    check(lib().pupil_buffer_download(name.encode(), C.c_void_p(host_ptr), nbytes))


def film():
    w, h, d = u32(), u32(), u32()
    check(lib().pupil_get_film(C.byref(w), C.byref(h), C.byref(d)))
    return w.value, h.value, d.value


def camera():
    a, b, fov = np.zeros(16, np.float32), np.zeros(16, np.float32), f32()
    check(lib().pupil_get_camera(a.ctypes.data_as(C.POINTER(f32)), b.ctypes.data_as(C.POINTER(f32)), C.byref(fov)))
    return a.reshape(4, 4), b.reshape(4, 4), fov.value


def instances():
    out = []
    for i in range(lib().pupil_num_instances()):
        xf, mat = np.zeros(16, np.float32), pb2.Material()
        eo, fl, npr, sph = i32(), u32(), u32(), i32()
        check(lib().pupil_get_instance(i, xf.ctypes.data_as(C.POINTER(f32)), C.byref(mat), C.byref(eo), C.byref(fl), C.byref(npr), C.byref(sph)))
        out.append(dict(xform=xf.reshape(4, 4), material=mat, emitter_offset=eo.value, flags=fl.value, n_prims=npr.value, is_sphere=bool(sph.value)))
    return out


def emitters():
    n = lib().pupil_num_area_emitters()
    arr = (pb2.Emitter * max(1, n))()
    env, has = pb2.Emitter(), i32()
    check(lib().pupil_get_emitters(arr, C.byref(env), C.byref(has)))
    return list(arr)[:n], (env if has.value else None)


class _BorrowedScene(pb2.Scene):
    """pb2.Scene view of the World's device scene (not owned: never destroyed from Python)."""

    def __init__(self, handle):  # noqa: super().__init__ would create a new scene
        self.h = C.c_void_p(handle)

    def close(self):
        self.h = C.c_void_p()

    def __del__(self):
        pass


def scene_handle() -> pb2.Scene:
    h = C.c_void_p()
    check(lib().pupil_scene_handle(C.byref(h)))
    return _BorrowedScene(h.value)


def set_bvh_builder(builder: int):
    check(lib().pupil_set_bvh_builder(builder))


def set_instancing(mode: int):
    check(lib().pupil_set_instancing(mode))


def build_stats() -> pb2.BuildStats:
    st = pb2.BuildStats()
    check(lib().pupil_build_stats(C.byref(st)))
    return st


def render_stats() -> pb2.RenderStats:
    st = pb2.RenderStats()
    check(lib().pupil_render_stats(C.byref(st)))
    return st


def camera_move(dx, dy, dz):
    check(lib().pupil_camera_move(dx, dy, dz))


def camera_rotate(dx, dy):
    check(lib().pupil_camera_rotate(dx, dy))


def camera_set_fov(fov_y):
    check(lib().pupil_camera_set_fov(fov_y))
