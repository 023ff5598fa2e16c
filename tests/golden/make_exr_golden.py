#!/usr/bin/env python
"""Writes the small PIZ-compressed OpenEXR fixtures under tests/golden/exr/ and what the reference's reader makes of them
(tests/golden/exr_reference.npz), both through the reference's own tinyexr compiled by oracle/Makefile into
oracle/_ref/libtinyexr_ref.so (so this script runs only where /root/reference exists).  tests/test_image_exr.py holds the
host library's reader to these on any machine.

    python tests/golden/make_exr_golden.py
"""
import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
T = C.CDLL(str(HERE.parent.parent / "oracle" / "_ref" / "libtinyexr_ref.so"))
T.exr_ref_save.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
T.exr_ref_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
T.exr_ref_free.argtypes = [C.c_void_p]


def ref_load(path):
    ptr, w, h = C.POINTER(C.c_float)(), C.c_int(), C.c_int()
    assert T.exr_ref_load(str(path).encode(), C.byref(ptr), C.byref(w), C.byref(h)) == 0
    a = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 4)).copy()
    T.exr_ref_free(ptr)
    return a


def picture(h, w, ch, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([0.5 + 0.5 * np.sin(x / 5 + y / 9), x / max(1, w - 1) * 3.0, np.exp(-((x - w / 2) ** 2 + (y - h / 2) ** 2) / 50.0) * 40.0, y / max(1, h - 1)], -1)
    # little noise: a block that PIZ cannot shrink is stored raw, and the fixtures are there to exercise the PIZ decoder
    q = 256 if seed == 1 else 4  # the FLOAT and the luminance fixture need coarser values before PIZ pays
    return (np.round(base[..., :ch] * q) / q + (rng.random((h, w, ch)) < 0.02) * 0.5).astype(np.float32)


def main():
    out = {}
    for name, (h, w, ch, half, seed) in {"piz_rgb_half": (45, 37, 3, 1, 1), "piz_rgba_float": (33, 20, 4, 0, 2), "piz_y_half": (9, 70, 1, 1, 3)}.items():
        path = HERE / "exr" / f"{name}.exr"
        img = picture(h, w, ch, seed)
        assert T.exr_ref_save(str(path).encode(), img.ctypes.data, w, h, ch, 4, half) == 0
        out[name] = ref_load(path)
        print(name, path.stat().st_size, "bytes")
    # a single-level tiled file (16 x 8 tiles, PIZ, ragged right and bottom edges)
    T.exr_ref_save_tiled.argtypes = [C.c_char_p, C.c_void_p] + [C.c_int] * 7
    img = picture(21, 43, 3, 4)
    path = HERE / "exr" / "tiled_piz_rgb_half.exr"
    assert T.exr_ref_save_tiled(str(path).encode(), img.ctypes.data, 43, 21, 3, 4, 1, 16, 8) == 0
    out["tiled_piz_rgb_half"] = ref_load(path)
    print("tiled_piz_rgb_half", path.stat().st_size, "bytes")
    np.savez_compressed(HERE / "exr_reference.npz", **out)


if __name__ == "__main__":
    main()
