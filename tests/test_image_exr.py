"""CPU-only: the host library's OpenEXR reader (pupiloptixlab_b200/host/image.cpp, image_piz.cpp) against the reference's own
reader — tinyexr's LoadEXR (framework/util/texture.cpp:131-149).

  * committed fixtures: tests/golden/exr/*.exr are PIZ files written by the reference's tinyexr and tests/golden/exr_reference.npz
    is what its LoadEXR returns for them (tests/golden/make_exr_golden.py); the reader must return the same floats, bit for bit;
  * where the reference tree was present at build time, oracle/_ref/libtinyexr_ref.so (tinyexr compiled from it, oracle/tinyexr_ref.cc)
    writes fresh files in all five compressions the reader knows, HALF and FLOAT, 1 / 3 / 4 channels, ragged sizes, and its
    LoadEXR is compared live."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from pupiloptixlab_b200 import pupil

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def test_piz_fixtures_match_the_reference_reader():
    want = np.load(GOLD / "exr_reference.npz")
    assert sorted(want.files) == ["piz_rgb_half", "piz_rgba_float", "piz_y_half", "tiled_piz_rgb_half"]
    for name in want.files:
        got = pupil.image_load(GOLD / "exr" / f"{name}.exr")
        assert got.shape == want[name].shape and np.array_equal(got.view(np.uint32), want[name].view(np.uint32)), name
    # a lone channel lands in all four slots, alpha included: LoadEXR's rule, which the reference passes on unchanged
    y = pupil.image_load(GOLD / "exr" / "piz_y_half.exr")
    assert np.array_equal(y[..., 0], y[..., 3])


def _tinyexr():
    p = ROOT / "oracle" / "_ref" / "libtinyexr_ref.so"
    if not p.exists():
        return None
    lib = C.CDLL(str(p))
    lib.exr_ref_save.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.exr_ref_load.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.exr_ref_free.argtypes = [C.c_void_p]
    return lib


T = _tinyexr()
needs_tinyexr = pytest.mark.skipif(T is None, reason="oracle/_ref/libtinyexr_ref.so is built only where the reference tree exists")


def _ref_load(path):
    ptr, w, h = C.POINTER(C.c_float)(), C.c_int(), C.c_int()
    assert T.exr_ref_load(str(path).encode(), C.byref(ptr), C.byref(w), C.byref(h)) == 0
    a = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, 4)).copy()
    T.exr_ref_free(ptr)
    return a


@needs_tinyexr
@pytest.mark.parametrize("compression", [0, 1, 2, 3, 4], ids=["none", "rle", "zips", "zip", "piz"])
@pytest.mark.parametrize("half", [0, 1], ids=["float", "half"])
def test_reader_equals_tinyexr(tmp_path, compression, half):
    rng = np.random.default_rng(compression * 2 + half)
    written = 0
    for ch, (h, w), kind in ((3, (37, 19), "noise"), (4, (64, 100), "smooth"), (1, (5, 7), "noise"), (3, (33, 130), "smooth"), (3, (1, 1), "noise"),
                             (4, (2, 257), "noise"), (3, (100, 3), "smooth"), (3, (31, 31), "flat"), (3, (40, 40), "hdr")):
        if kind == "noise":
            img = rng.random((h, w, ch), dtype=np.float32) * 4
        elif kind == "flat":
            img = np.full((h, w, ch), 0.25, np.float32)
        elif kind == "hdr":  # wide dynamic range incl. zeros, denormal halves and large values: every 16-bit pattern class
            img = np.exp(rng.uniform(-20, 10, (h, w, ch))).astype(np.float32) * (rng.random((h, w, ch)) > 0.1)
        else:
            y, x = np.mgrid[0:h, 0:w].astype(np.float32)
            img = np.stack([np.sin(x / 7 + y / 3) + 1, x / w, y / h, 0.5 + 0 * x], -1)[..., :ch] + rng.normal(0, 0.002, (h, w, ch))
        img = np.ascontiguousarray(img, np.float32)
        path = tmp_path / f"t{ch}_{h}x{w}.exr"
        if compression == 4:
            # tinyexr's own PIZ WRITER corrupts its heap on pictures that do not compress (its output buffer is sized by the input)
            # and takes the process down: it runs in a child, and a case it cannot write is skipped
            np.save(tmp_path / "img.npy", img)
            child = ("import ctypes as C, numpy as np, sys; a = np.load(sys.argv[2]); T = C.CDLL(sys.argv[1]); "
                     "T.exr_ref_save.argtypes = [C.c_char_p, C.c_void_p] + [C.c_int] * 5; "
                     "sys.exit(T.exr_ref_save(sys.argv[3].encode(), a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], 4, int(sys.argv[4])))")
            r = subprocess.run([sys.executable, "-c", child, str(ROOT / "oracle" / "_ref" / "libtinyexr_ref.so"), str(tmp_path / "img.npy"), str(path), str(half)],
                               capture_output=True)
            if r.returncode != 0:
                continue
            written += 1
        else:
            assert T.exr_ref_save(str(path).encode(), img.ctypes.data, w, h, ch, compression, half) == 0
        want, got = _ref_load(path), pupil.image_load(path)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (ch, h, w, kind)
    assert compression != 4 or written >= 5


@needs_tinyexr
@pytest.mark.parametrize("compression", [0, 1, 3, 4], ids=["none", "rle", "zip", "piz"])
def test_tiled_files_equal_tinyexr(tmp_path, compression):
    """single-level tiled files (the other layout LoadEXR reads): tiles that do not divide the picture, tiles larger than it"""
    child = ("import ctypes as C, numpy as np, sys; a = np.load(sys.argv[2]); T = C.CDLL(sys.argv[1]); "
             "T.exr_ref_save_tiled.argtypes = [C.c_char_p, C.c_void_p] + [C.c_int] * 7; "
             "sys.exit(T.exr_ref_save_tiled(sys.argv[3].encode(), a.ctypes.data, a.shape[1], a.shape[0], a.shape[2], int(sys.argv[4]), int(sys.argv[5]), "
             "int(sys.argv[6]), int(sys.argv[7])))")
    written = 0
    for ch, (h, w), (tw, th), half in ((3, (40, 50), (16, 16), 1), (4, (33, 70), (32, 8), 0), (1, (20, 20), (8, 8), 1), (3, (64, 64), (64, 64), 0), (3, (17, 5), (4, 16), 1)):
        y, x = np.mgrid[0:h, 0:w].astype(np.float32)
        img = np.ascontiguousarray(np.stack([np.round((np.sin(x / 7 + y / 3) + 1) * 16) / 16, x / w, y / h, 0 * x + 0.5], -1)[..., :ch], np.float32)
        np.save(tmp_path / "img.npy", img)
        path = tmp_path / f"t{ch}_{h}x{w}.exr"
        r = subprocess.run([sys.executable, "-c", child, str(ROOT / "oracle" / "_ref" / "libtinyexr_ref.so"), str(tmp_path / "img.npy"), str(path), str(compression),
                            str(half), str(tw), str(th)], capture_output=True)
        if r.returncode != 0:
            continue  # tinyexr's writer refuses tiles larger than the picture and is fragile on incompressible data
        written += 1
        want, got = _ref_load(path), pupil.image_load(path)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (ch, h, w, tw, th)
    assert written >= 3


@needs_tinyexr
def test_corrupted_piz_files_are_decoded_or_refused(tmp_path):
    rng = np.random.default_rng(11)
    data = (GOLD / "exr" / "piz_rgb_half.exr").read_bytes()
    pupil.lib().pupil_set_log_level(0)
    try:
        for it in range(300):
            d = bytearray(data)
            for _ in range(int(rng.integers(1, 6))):
                d[int(rng.integers(300, len(d)))] = int(rng.integers(0, 256))  # past the header: the compressed blocks
            if it % 5 == 0:
                d = d[:int(rng.integers(300, len(d)))]
            (tmp_path / "c.exr").write_bytes(bytes(d))
            try:
                got = pupil.image_load(tmp_path / "c.exr")
                assert got.shape == (45, 37, 4)
            except pupil.PupilError:
                pass
    finally:
        pupil.lib().pupil_set_log_level(1)
