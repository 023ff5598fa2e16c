"""ctypes binding of libpb2.so — the C ABI declared in include/pb2.h (CUDA back end, sm_100a).

There is no fallback: if the shared library is missing or no CUDA device is present, loading /
initialising raises.  torch is not needed here; device buffers are plain pointers (a torch tensor's
`data_ptr()` can be passed wherever a device pointer is expected).
"""
from __future__ import annotations

import os
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / os.environ.get("PB2_BUILD_DIR", "_build") / "libpb2.so"

f32, i32, u32, u64 = C.c_float, C.c_int32, C.c_uint32, C.c_uint64

MAT = dict(unknown=0, diffuse=1, dielectric=2, roughdielectric=3, conductor=4, roughconductor=5, plastic=6, roughplastic=7)
TEX_RGB, TEX_BITMAP, TEX_CHECKERBOARD = 0, 1, 2
EMIT_NONE, EMIT_TRI, EMIT_SPHERE, EMIT_CONST_ENV, EMIT_ENV_MAP = range(5)
INST_FLIP_NORMALS, INST_FLIP_TEX = 1, 2
MESH_SPHERE = 0xFFFFFFFF


class Pb2Error(RuntimeError):
    pass


class Texture(C.Structure):
    _fields_ = [("type", i32), ("a", f32 * 3), ("b", f32 * 3), ("r0", f32 * 4), ("r1", f32 * 4), ("pad0", u32), ("bitmap", u64)]


class Material(C.Structure):
    _fields_ = [("type", i32), ("twosided", i32), ("eta", f32), ("nonlinear", i32), ("int_fdr", f32),
                ("specular_sampling_weight", f32), ("tex", Texture * 4)]


class Emitter(C.Structure):
    _fields_ = [("type", i32), ("weight", f32), ("select_probability", f32), ("radiance", Texture), ("area", f32),
                ("pos", (f32 * 3) * 3), ("nrm", (f32 * 3) * 3), ("uv", (f32 * 2) * 3), ("center", f32 * 3), ("radius", f32),
                ("scale", f32), ("normalization", f32), ("map_w", u32), ("map_h", u32), ("to_world", f32 * 9), ("to_local", f32 * 9),
                ("env_tables", C.c_void_p)]


class Hit(C.Structure):
    _fields_ = [("t", f32), ("u", f32), ("v", f32), ("inst", i32), ("prim", i32)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("inst", "<i4"), ("prim", "<i4")])


class LaunchParams(C.Structure):
    _fields_ = [("max_depth", u32), ("accumulate", u32), ("width", u32), ("height", u32), ("random_seed", u32), ("seed_stride", u32),
                ("sample_cnt", u32), ("n_frames", u32), ("accum_buffer", C.c_void_p), ("frame_buffer", C.c_void_p),
                ("normal_buffer", C.c_void_p), ("albedo_buffer", C.c_void_p), ("test_buffer", C.c_void_p)]


class BuildStats(C.Structure):
    _fields_ = [("n_prims", u64), ("n_triangles", u64), ("n_spheres", u64), ("n_nodes", u64), ("bvh_bytes", u64),
                ("build_ms", f32), ("sah_cost", f32), ("max_depth", u32), ("n_blas", u32), ("n_instance_leaves", u64), ("top_level_ms", f32), ("pad0", u32)]


class RenderStats(C.Structure):
    _fields_ = [("closest_rays", u64), ("shadow_rays", u64), ("kernel_launches", u64), ("total_ms", f32), ("generate_ms", f32),
                ("extend_ms", f32), ("shade_ms", f32), ("shadow_ms", f32), ("accumulate_ms", f32), ("nodes_visited", u64),
                ("prims_tested", u64), ("nodes_shadow", u64), ("prims_shadow", u64), ("batches", u32), ("rounds", u32),
                ("extend_launches", u32), ("shade_launches", u32), ("shadow_launches", u32), ("other_launches", u32), ("shaded_paths", u64), ("shadow_unoccluded", u64), ("sorted", u32), ("pad0", u32)]


class KatBsdf(C.Structure):  # csrc/kat.cu
    _fields_ = [("type", i32), ("alpha", f32), ("eta", f32), ("int_fdr", f32), ("specular_sampling_weight", f32), ("nonlinear", i32),
                ("c0", f32 * 3), ("c1", f32 * 3), ("c2", f32 * 3)]


_lib = None


def lib():
    """Load libpb2.so (raises if it has not been built — there is no CPU path)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise Pb2Error(f"{LIB_PATH} not found: run `python -m pupiloptixlab_b200.build` (needs nvcc); there is no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        P, vp = C.POINTER, C.c_void_p
        L.pb2_last_error.restype = C.c_char_p
        sigs = {
            "pb2_init": [C.c_int], "pb2_device_count": [], "pb2_malloc": [P(vp), u64], "pb2_free": [vp], "pb2_trim": [], "pb2_bitmap_create": [vp, u32, u32, C.c_int, C.c_int, P(u64)],
            "pb2_bitmap_destroy": [u64], "pb2_upload": [vp, vp, u64],
            "pb2_download": [vp, vp, u64], "pb2_memset": [vp, C.c_int, u64], "pb2_scene_create": [P(vp)], "pb2_scene_destroy": [vp],
            "pb2_scene_clear": [vp], "pb2_scene_set_stream": [vp, vp],
            "pb2_scene_add_mesh": [vp, vp, vp, vp, vp, u32, u32, P(u32)],
            "pb2_scene_add_instance": [vp, u32, P(f32), u32, P(Material), i32, P(u32)],
            "pb2_scene_set_emitters": [vp, P(Emitter), u32, P(Emitter)], "pb2_scene_set_instance_transform": [vp, u32, P(f32)], "pb2_scene_set_camera": [vp, P(f32), P(f32)],
            "pb2_bvh_build": [vp, P(BuildStats)], "pb2_scene_set_builder": [vp, C.c_int],
            "pb2_trace_closest": [vp, vp, u64, vp], "pb2_trace_any": [vp, vp, u64, vp],
            "pb2_trace_closest_dev": [vp, vp, u64, vp, vp], "pb2_trace_any_dev": [vp, vp, u64, vp],
            "pb2_bvh_download": [vp, vp, P(u64), vp, P(u64)], "pb2_render": [vp, P(LaunchParams)], "pb2_synchronize": [vp],
            "pb2_render_stats_get": [vp, P(RenderStats)], "pb2_scene_set_option": [vp, C.c_char_p, C.c_int64],
            "pb2_finalize_sum": [vp, vp, vp, u64, u32],
            "pb2_comm_unique_id": [vp], "pb2_comm_create": [P(vp), C.c_int, C.c_int, vp], "pb2_comm_destroy": [vp],
            "pb2_comm_reduce_frames": [vp, vp, vp, vp, u64, u32, C.c_int, C.c_int], "pb2_comm_synchronize": [vp], "pb2_comm_nccl_version": [P(C.c_int)], "pb2_comm_last_reduction": [vp, P(f32), P(u64)],
            "pb2_shard_plan": [C.c_int, C.c_int, u32, u32, C.c_int, P(u32), P(u32), P(u32), P(u32)],
        }
        for name, args in sigs.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, C.c_int
        _lib = L
    return _lib


def check(code: int):
    if code != 0:
        raise Pb2Error(f"pb2 error {code}: {lib().pb2_last_error().decode()}")


def init(device: int = 0):
    L = lib()
    if L.pb2_device_count() <= 0:
        raise Pb2Error("no CUDA device visible: the pb2 back end has no CPU fallback")
    check(L.pb2_init(device))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceBuffer:
    """cudaMalloc'ed, zero-initialised buffer (BufferManager::AllocBuffer)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = C.c_void_p()
        check(lib().pb2_malloc(C.byref(self.ptr), self.nbytes))

    def upload(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        check(lib().pb2_upload(self.ptr, _ptr(arr), arr.nbytes))

    def download(self, dtype=np.float32, shape=None) -> np.ndarray:
        out = np.empty(self.nbytes // np.dtype(dtype).itemsize, dtype)
        check(lib().pb2_download(_ptr(out), self.ptr, out.nbytes))
        return out.reshape(shape) if shape is not None else out

    def zero(self):
        check(lib().pb2_memset(self.ptr, 0, self.nbytes))

    def free(self):
        if self.ptr:
            lib().pb2_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Bitmap:
    """pb2_bitmap_create: a float4 texture object (normalised coordinates); `handle` goes into Texture.bitmap"""

    def __init__(self, rgba: np.ndarray, address_mode: int = 0, filter_mode: int = 1):
        img = np.ascontiguousarray(rgba, np.float32)
        assert img.ndim == 3 and img.shape[2] == 4
        h = u64()
        check(lib().pb2_bitmap_create(_ptr(img), img.shape[1], img.shape[0], address_mode, filter_mode, C.byref(h)))
        self.handle = h.value

    def free(self):
        if self.handle:
            lib().pb2_bitmap_destroy(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Scene:
    """Thin object wrapper over a pb2_scene handle."""

    def __init__(self):
        self.h = C.c_void_p()
        check(lib().pb2_scene_create(C.byref(self.h)))

    def close(self):
        if self.h:
            lib().pb2_scene_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_mesh(self, positions, indices, normals=None, texcoords=None) -> int:
        P = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        I = np.ascontiguousarray(indices, np.uint32).reshape(-1, 3)
        N = None if normals is None else np.ascontiguousarray(normals, np.float32).reshape(-1, 3)
        T = None if texcoords is None else np.ascontiguousarray(texcoords, np.float32).reshape(-1, 2)
        mid = u32()
        check(lib().pb2_scene_add_mesh(self.h, _ptr(P), _ptr(N), _ptr(T), _ptr(I), P.shape[0], I.shape[0], C.byref(mid)))
        return mid.value

    def add_instance(self, mesh_id: int, xform=None, flags: int = 0, material: Material | None = None, emitter_offset: int = -1) -> int:
        x = np.eye(4, dtype=np.float32)[:3] if xform is None else np.ascontiguousarray(xform, np.float32).reshape(-1)[:12]
        x = np.ascontiguousarray(x, np.float32).reshape(-1)
        iid = u32()
        check(lib().pb2_scene_add_instance(self.h, mesh_id, x.ctypes.data_as(C.POINTER(f32)), flags,
                                            C.byref(material) if material is not None else None, emitter_offset, C.byref(iid)))
        return iid.value

    def set_instance_transform(self, instance_id: int, xform):
        """new 3x4 (or 4x4) object->world transform; the next build() keeps the bottom-level trees"""
        m = np.ascontiguousarray(np.asarray(xform, np.float32).reshape(-1)[:12])
        check(lib().pb2_scene_set_instance_transform(self.h, instance_id, m.ctypes.data_as(C.POINTER(f32))))

    def set_emitters(self, areas: list, env: Emitter | None = None):
        arr = (Emitter * max(1, len(areas)))(*areas)
        check(lib().pb2_scene_set_emitters(self.h, arr, len(areas), C.byref(env) if env is not None else None))

    def set_camera(self, s2c, c2w):
        a = np.ascontiguousarray(s2c, np.float32).reshape(-1)
        b = np.ascontiguousarray(c2w, np.float32).reshape(-1)
        check(lib().pb2_scene_set_camera(self.h, a.ctypes.data_as(C.POINTER(f32)), b.ctypes.data_as(C.POINTER(f32))))

    def set_builder(self, builder: int):
        check(lib().pb2_scene_set_builder(self.h, builder))

    def set_option(self, name: str, value: int):
        check(lib().pb2_scene_set_option(self.h, name.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr: int | None):
        check(lib().pb2_scene_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def build(self) -> BuildStats:
        st = BuildStats()
        check(lib().pb2_bvh_build(self.h, C.byref(st)))
        return st

    def trace_closest(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], HIT_DTYPE)
        check(lib().pb2_trace_closest(self.h, _ptr(rays), rays.shape[0], _ptr(hits)))
        return hits

    def trace_any(self, rays) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], np.uint8)
        check(lib().pb2_trace_any(self.h, _ptr(rays), rays.shape[0], _ptr(occ)))
        return occ

    def trace_closest_dev(self, rays_ptr: int, n: int, tuvp_ptr: int, inst_ptr: int):
        check(lib().pb2_trace_closest_dev(self.h, C.c_void_p(rays_ptr), n, C.c_void_p(tuvp_ptr), C.c_void_p(inst_ptr)))

    def trace_any_dev(self, rays_ptr: int, n: int, occ_ptr: int):
        check(lib().pb2_trace_any_dev(self.h, C.c_void_p(rays_ptr), n, C.c_void_p(occ_ptr)))

    def bvh_download(self):
        nn, npr = u64(), u64()
        check(lib().pb2_bvh_download(self.h, None, C.byref(nn), None, C.byref(npr)))
        nodes = np.zeros((nn.value, 20), np.uint32)
        prims = np.zeros((npr.value, 12), np.float32)
        check(lib().pb2_bvh_download(self.h, _ptr(nodes), C.byref(nn), _ptr(prims), C.byref(npr)))
        return nodes, prims

    def render(self, params: LaunchParams):
        check(lib().pb2_render(self.h, C.byref(params)))

    def synchronize(self):
        check(lib().pb2_synchronize(self.h))

    def render_stats(self) -> RenderStats:
        st = RenderStats()
        check(lib().pb2_render_stats_get(self.h, C.byref(st)))
        return st

    def render_stats_raw(self) -> RenderStats:
        """counters of the last pb2_trace_* call (no render needed)"""
        return self.render_stats()

    def finalize_sum(self, sum_ptr: int, frame_ptr: int, n_pixels: int, total_spp: int):
        check(lib().pb2_finalize_sum(self.h, C.c_void_p(sum_ptr), C.c_void_p(frame_ptr), n_pixels, total_spp))


def kat(what: str, in0, in1, in2, n: int, out: np.ndarray):
    def p(x):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            return _ptr(x)
        return C.cast(x, C.c_void_p)
    code = kat_lib().pb2_kat(what.encode(), p(in0), p(in1), p(in2), n, _ptr(out))
    if code != 0:
        raise Pb2Error(f"pb2_kat error {code}: {kat_lib().pb2_kat_last_error().decode()}")
    return out


_kat_lib = None


def kat_lib():
    """libpb2_kat.so: the known-answer test hooks (include/pb2_kat.h) — test infrastructure, not in libpb2.so"""
    global _kat_lib
    if _kat_lib is None:
        lib()  # libpb2.so first: the hooks link against it
        path = LIB_PATH.parent / "libpb2_kat.so"
        if not path.exists():
            raise Pb2Error(f"{path} not found: run `python -m pupiloptixlab_b200.build`")
        L = C.CDLL(str(path))
        L.pb2_kat.argtypes, L.pb2_kat.restype = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, u64, C.c_void_p], C.c_int
        L.pb2_kat_last_error.restype = C.c_char_p
        _kat_lib = L
    return _kat_lib
