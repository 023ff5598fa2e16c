"""CPU-only: the host library's JPEG / BMP / TGA decoders and Adam7 PNG (pupiloptixlab_b200/host/image_ldr.cpp, image.cpp) —
the formats stb_image adds on the reference's 8-bit texture path (framework/util/texture.cpp:106-117).

Two checkers.  (1) The reference's own decoder: oracle/_ref/libstb_ref.so is stb_image compiled from the reference tree
(oracle/stb_ref.c); where it exists the texels must be EQUAL to what stbi_load returns, after the reference's
pow(x / 255, 2.2) — bit-exact 8-bit values.  (2) Pillow (libjpeg-turbo, an independent implementation): JPEG decoders may
differ by rounding in the inverse DCT and in chroma upsampling, so that comparison carries a stated tolerance; BMP / TGA / PNG
are lossless and must match exactly.  Files are written by Pillow or by the small encoders below."""
import ctypes as C
import io
import struct
import zlib
from pathlib import Path

import numpy as np
import pytest

from pupiloptixlab_b200 import pupil

PIL = pytest.importorskip("PIL.Image")
F = np.float32
ROOT = Path(__file__).resolve().parent.parent


def _lin(u8):
    return np.power(np.asarray(u8, F) * F(1.0) / F(255.0), F(2.2)).astype(F)


def _stb():
    p = ROOT / "oracle" / "_ref" / "libstb_ref.so"
    if not p.exists():
        return None
    lib = C.CDLL(str(p))
    lib.stb_ref_load.restype = C.POINTER(C.c_ubyte)
    lib.stb_ref_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.stb_ref_free.argtypes = [C.c_void_p]
    return lib


STB = _stb()


def stb_load(data: bytes):
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    p = STB.stb_ref_load(data, len(data), C.byref(w), C.byref(h), C.byref(c))
    if not p:
        return None
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, c.value)).copy()
    STB.stb_ref_free(p)
    return a


def expected_rgba(u8):
    """(h, w, c) 8-bit pixels -> what BitmapTexture::Load makes of them (grey spread over RGB; alpha / 255 or 1)."""
    h, w, c = u8.shape
    out = np.ones((h, w, 4), F)
    if c >= 3:
        out[..., :3] = _lin(u8[..., :3])
    else:
        out[..., :3] = _lin(u8[..., :1])
    if c in (2, 4):
        out[..., 3] = u8[..., -1].astype(F) * F(1.0) / F(255.0)
    return out


def check(path: Path, data: bytes, pil_tol=None, pil=True):
    """decode `data` through the host library and hold it to stb_image (exact) and to Pillow (exact or within pil_tol levels)"""
    path.write_bytes(data)
    got = pupil.image_load(path)
    if STB is not None:
        ref = stb_load(data)
        assert ref is not None, "the reference's decoder rejects this file: not a fair test"
        want = expected_rgba(ref)
        assert got.shape == want.shape
        assert np.allclose(got, want, rtol=3e-6, atol=1e-7), f"{path.name}: differs from stb_image in {int((~np.isclose(got, want, rtol=3e-6, atol=1e-7)).sum())} values"
    if not pil or (pil_tol is not None and STB is not None and min(got.shape[:2]) < 16):
        return got  # Pillow expands 5-bit channels differently; thin subsampled JPEGs are dominated by the libraries' edge rules
    try:
        im = PIL.open(io.BytesIO(data))
        im.load()
    except OSError:
        return got  # Pillow's TGA reader rejects some legal files (top-down + RLE combinations)
    mode = {"L": "L", "LA": "LA", "RGBA": "RGBA", "1": "L", "P": "RGBA" if "transparency" in im.info else "RGB"}.get(im.mode, "RGB")
    ref = np.asarray(im.convert(mode))
    if ref.ndim == 2:
        ref = ref[..., None]
    assert got.shape[:2] == ref.shape[:2]
    if pil_tol is None:
        assert np.allclose(got[..., :3], expected_rgba(ref)[..., :3], rtol=3e-6, atol=1e-7), f"{path.name}: differs from Pillow"
    else:
        # back to 8-bit levels for a tolerance in levels
        lv = np.rint(np.power(got[..., :3].astype(np.float64), 1 / 2.2) * 255)
        want = ref[..., :3].astype(np.float64) if ref.shape[2] >= 3 else np.repeat(ref[..., :1], 3, 2).astype(np.float64)
        d = np.abs(lv - want)
        assert d.max() <= pil_tol[0] and d.mean() <= pil_tol[1], f"{path.name}: max {d.max()} mean {d.mean():.3f} levels from Pillow"
    return got


def _picture(h, w, seed=0, smooth=True):
    """a colourful test card: gradients + a few sharp edges + mild noise (smooth) or pure noise"""
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(x / 7.0 + y / 13.0), 255 * x / max(1, w - 1), 255 * y / max(1, h - 1)], -1)
    img[h // 3:h // 2, w // 4:w // 2] = (250, 20, 30)
    img[h // 2:, : w // 5] = (10, 200, 240)
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def _jpeg(arr, **kw):
    b = io.BytesIO()
    PIL.fromarray(arr if arr.shape[-1] != 1 else arr[..., 0]).save(b, "JPEG", **kw)
    return b.getvalue()


# ---- JPEG -----------------------------------------------------------------------------------------------------------
# tolerance against Pillow: (max, mean) in 8-bit levels.  4:4:4 differs only by inverse-DCT rounding; subsampled chroma is
# upsampled with different rules at sharp colour edges in the two libraries (a few pixels, tens of levels; the mean stays < 1).
@pytest.mark.parametrize("subsampling,tol", [(0, (3, 0.6)), (1, (48, 0.8)), (2, (48, 0.8))])
@pytest.mark.parametrize("size", [(64, 48), (37, 53), (8, 8), (1, 1), (17, 1), (1, 19), (100, 3)])
def test_jpeg_baseline(tmp_path, subsampling, tol, size):
    h, w = size
    data = _jpeg(_picture(h, w, seed=h * 131 + w), quality=90, subsampling=subsampling)
    got = check(tmp_path / "a.jpg", data, pil_tol=tol)
    assert got.shape == (h, w, 4) and np.all(got[..., 3] == 1.0)


@pytest.mark.parametrize("quality", [30, 75, 100])
def test_jpeg_quality_and_noise(tmp_path, quality):
    check(tmp_path / "n.jpg", _jpeg(_picture(40, 56, 5, smooth=False), quality=quality, subsampling=0), pil_tol=(4, 0.8))
    check(tmp_path / "s.jpg", _jpeg(_picture(40, 56, 6), quality=quality, subsampling=2), pil_tol=(48, 1.0))


@pytest.mark.parametrize("subsampling", [0, 1, 2])
@pytest.mark.parametrize("size", [(64, 48), (37, 53), (9, 70), (1, 1)])
def test_jpeg_progressive(tmp_path, subsampling, size):
    """spectral selection + successive approximation (DC first / refine, AC first / refine with EOB runs)"""
    h, w = size
    data = _jpeg(_picture(h, w, seed=77 + h), quality=85, subsampling=subsampling, progressive=True)
    assert b"\xff\xc2" in data
    check(tmp_path / "p.jpg", data, pil_tol=(48, 0.8))


def test_jpeg_grey_and_optimised_tables(tmp_path):
    g = _picture(33, 47, 9)[..., :1]
    got = check(tmp_path / "g.jpg", _jpeg(g, quality=88, optimize=True), pil_tol=(3, 0.6))
    assert np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 0], got[..., 2])
    check(tmp_path / "gp.jpg", _jpeg(g, quality=60, progressive=True), pil_tol=(3, 0.6))
    check(tmp_path / "o.jpg", _jpeg(_picture(50, 50, 10), quality=95, optimize=True, subsampling=2), pil_tol=(48, 0.8))


def test_jpeg_restart_intervals(tmp_path):
    img = _picture(48, 80, 12)
    for kw in (dict(restart_marker_blocks=3), dict(restart_marker_rows=1)):
        try:
            data = _jpeg(img, quality=85, subsampling=2, **kw)
        except TypeError:
            pytest.skip("this Pillow cannot write restart markers")
        if b"\xff\xdd" not in data:
            pytest.skip("this Pillow ignores the restart options")
        check(tmp_path / "r.jpg", data, pil_tol=(48, 0.8))
        data = _jpeg(img, quality=85, subsampling=0, progressive=True, **kw)
        check(tmp_path / "rp.jpg", data, pil_tol=(48, 0.8))


def test_jpeg_malformed_files_fail_loudly(tmp_path):
    good = _jpeg(_picture(24, 24, 3), quality=80)
    for name, data in {
        "cut_header.jpg": good[:30],
        "no_tables.jpg": good[:2] + good[good.index(b"\xff\xc0"):],
        "twelve_bit.jpg": good.replace(b"\xff\xc0\x00\x11\x08", b"\xff\xc0\x00\x11\x0c"),
        "arithmetic.jpg": good.replace(b"\xff\xc0", b"\xff\xc9"),
    }.items():
        (tmp_path / name).write_bytes(data)
        with pytest.raises(pupil.PupilError):
            pupil.image_load(tmp_path / name)
    # a file cut inside the entropy-coded data still decodes (missing bits read as zeros), like stb_image and libjpeg
    (tmp_path / "cut_scan.jpg").write_bytes(good[:(good.index(b"\xff\xda") + len(good)) // 2])
    assert pupil.image_load(tmp_path / "cut_scan.jpg").shape == (24, 24, 4)


# ---- BMP ------------------------------------------------------------------------------------------------------------
def _bmp(w, h, bpp, rows, palette=None, compression=0, masks=None, top_down=False, header=40):
    """rows: list of h byte strings (unpadded), first = top of the picture"""
    stride = (w * bpp + 31) // 32 * 4
    body = b"".join(r + bytes(stride - len(r)) for r in (rows if top_down else rows[::-1]))
    pal = b"".join(bytes([b, g, r, 0]) for r, g, b in (palette or []))
    extra = b""
    if masks is not None:
        extra = b"".join(struct.pack("<I", m) for m in masks)
    if header == 40:
        info = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bpp, compression, len(body), 2835, 2835, len(palette or []), 0) + extra
    else:  # BITMAPV4HEADER: masks (4) are part of it
        m = list(masks or (0, 0, 0)) + [0] * (4 - len(masks or (0, 0, 0)))
        info = struct.pack("<IiiHHIIiiII", 108, w, -h if top_down else h, 1, bpp, compression, len(body), 2835, 2835, 0, 0)
        info += struct.pack("<IIII", *m) + b"BGRs" + bytes(36) + bytes(12)
    off = 14 + len(info) + len(pal)
    return b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) + info + pal + body


def test_bmp_true_colour_and_row_order(tmp_path):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)
    rows = [img[y, :, ::-1].tobytes() for y in range(7)]
    a = check(tmp_path / "a.bmp", _bmp(5, 7, 24, rows))
    b = check(tmp_path / "b.bmp", _bmp(5, 7, 24, rows, top_down=True))
    assert np.array_equal(a, b) and np.allclose(a[..., :3], _lin(img), rtol=3e-6, atol=1e-7) and np.all(a[..., 3] == 1.0)
    rgba = rng.integers(0, 256, (4, 6, 4), dtype=np.uint8)
    rows = [rgba[y][:, [2, 1, 0, 3]].tobytes() for y in range(4)]
    got = check(tmp_path / "c.bmp", _bmp(6, 4, 32, rows))
    assert np.allclose(got[..., 3], rgba[..., 3] / 255.0, rtol=1e-6)
    rgba[..., 3] = 0  # the usual "XRGB" file: alpha byte left zero -> opaque
    rows = [rgba[y][:, [2, 1, 0, 3]].tobytes() for y in range(4)]
    got = check(tmp_path / "d.bmp", _bmp(6, 4, 32, rows))
    assert np.all(got[..., 3] == 1.0)


@pytest.mark.parametrize("bpp", [1, 4, 8])
def test_bmp_palette(tmp_path, bpp):
    rng = np.random.default_rng(bpp)
    n = 1 << bpp
    palette = [tuple(int(v) for v in rng.integers(0, 256, 3)) for _ in range(n)]
    idx = rng.integers(0, n, (6, 11))
    rows = []
    for y in range(6):
        bits = "".join(format(int(v), f"0{bpp}b") for v in idx[y])
        bits += "0" * (-len(bits) % 8)
        rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    got = check(tmp_path / "p.bmp", _bmp(11, 6, bpp, rows, palette=palette))
    assert np.allclose(got[..., :3], _lin(np.array(palette, np.uint8)[idx]), rtol=3e-6, atol=1e-7)


def test_bmp_bit_fields(tmp_path):
    rng = np.random.default_rng(4)
    v = rng.integers(0, 1 << 16, (5, 9), dtype=np.uint16)
    rows = [v[y].astype("<u2").tobytes() for y in range(5)]
    spread = lambda x, bits: (x << (8 - bits) | x >> (2 * bits - 8)).astype(np.uint8)  # noqa: E731  bit replication to 8 bits
    got = check(tmp_path / "555.bmp", _bmp(9, 5, 16, rows), pil=False)  # BI_RGB 16 bit = 5-5-5
    want = np.stack([spread((v >> 10) & 31, 5), spread((v >> 5) & 31, 5), spread(v & 31, 5)], -1)
    assert np.allclose(got[..., :3], _lin(want), rtol=3e-6, atol=1e-7)
    got = check(tmp_path / "565.bmp", _bmp(9, 5, 16, rows, compression=3, masks=(0xF800, 0x07E0, 0x001F)), pil=False)
    want = np.stack([spread((v >> 11) & 31, 5), spread((v >> 5) & 63, 6), spread(v & 31, 5)], -1)
    assert np.allclose(got[..., :3], _lin(want), rtol=3e-6, atol=1e-7)
    v32 = rng.integers(0, 1 << 32, (3, 4), dtype=np.uint64).astype(np.uint32)
    rows = [v32[y].astype("<u4").tobytes() for y in range(3)]
    got = check(tmp_path / "v4.bmp", _bmp(4, 3, 32, rows, compression=3, masks=(0x000000FF, 0x0000FF00, 0x00FF0000, 0xFF000000), header=108))
    want = np.stack([v32 & 255, (v32 >> 8) & 255, (v32 >> 16) & 255], -1).astype(np.uint8)
    assert np.allclose(got[..., :3], _lin(want), rtol=3e-6, atol=1e-7) and np.allclose(got[..., 3], ((v32 >> 24) & 255) / 255.0, rtol=1e-6)


def test_bmp_rle_and_truncated_fail_loudly(tmp_path):
    rows = [bytes(12)] * 4
    for name, data in {"rle.bmp": _bmp(4, 4, 8, [bytes(4)] * 4, palette=[(0, 0, 0)] * 256, compression=1), "cut.bmp": _bmp(4, 4, 24, rows)[:-20]}.items():
        (tmp_path / name).write_bytes(data)
        with pytest.raises(pupil.PupilError):
            pupil.image_load(tmp_path / name)


# ---- TGA ------------------------------------------------------------------------------------------------------------
def _tga(arr, rle=False, top_down=False, grey=False, palette=None, bits16=False):
    """arr: (h, w, c) uint8 (c = 1, 3, 4), or (h, w) palette indices, or (h, w) uint16 5-5-5 words when bits16"""
    h, w = arr.shape[:2]
    if palette is not None:
        elems = [bytes([int(v)]) for v in arr.reshape(-1)]
        bpp, typ = 8, 1
    elif bits16:
        elems = [struct.pack("<H", int(v)) for v in arr.reshape(-1)]
        bpp, typ = 16, 2
    else:
        c = arr.shape[2]
        order = {1: [0], 3: [2, 1, 0], 4: [2, 1, 0, 3]}[c]
        elems = [bytes(px[order]) for px in arr.reshape(-1, c)]
        bpp, typ = 8 * c, 3 if c == 1 else 2
    rows = [elems[y * w:(y + 1) * w] for y in range(h)]
    if not top_down:
        rows = rows[::-1]
    flat = [e for r in rows for e in r]
    if rle:
        typ += 8
        body, i = b"", 0
        while i < len(flat):
            run = 1
            while i + run < len(flat) and run < 128 and flat[i + run] == flat[i]:
                run += 1
            if run > 1:
                body += bytes([0x80 | (run - 1)]) + flat[i]
                i += run
            else:
                lit = 1
                while i + lit < len(flat) and lit < 128 and (i + lit + 1 >= len(flat) or flat[i + lit] != flat[i + lit + 1]):
                    lit += 1
                body += bytes([lit - 1]) + b"".join(flat[i:i + lit])
                i += lit
    else:
        body = b"".join(flat)
    cmap = b"".join(bytes([b, g, r]) for r, g, b in palette) if palette is not None else b""
    hdr = struct.pack("<BBBHHBHHHHBB", 0, 1 if palette is not None else 0, typ, 0, len(palette or []), 24 if palette is not None else 0, 0, 0, w, h, bpp,
                      (0x20 if top_down else 0) | (8 if (not bits16 and palette is None and arr.shape[2] == 4) else 0))
    return hdr + cmap + body


@pytest.mark.parametrize("rle", [False, True])
@pytest.mark.parametrize("top_down", [False, True])
def test_tga_true_colour_grey_and_alpha(tmp_path, rle, top_down):
    rng = np.random.default_rng(8)
    for c in (1, 3, 4):
        img = rng.integers(0, 4 if rle else 256, (9, 13, c), dtype=np.uint8) * (60 if rle else 1)  # few colours: real runs
        got = check(tmp_path / f"t{c}.tga", _tga(img, rle=rle, top_down=top_down))
        assert np.allclose(got, expected_rgba(img), rtol=3e-6, atol=1e-7)


def test_tga_palette_and_16_bit(tmp_path):
    rng = np.random.default_rng(9)
    palette = [tuple(int(v) for v in rng.integers(0, 256, 3)) for _ in range(40)]
    idx = rng.integers(0, 40, (6, 10))
    for rle in (False, True):
        got = check(tmp_path / "p.tga", _tga(idx, palette=palette, rle=rle))
        assert np.allclose(got[..., :3], _lin(np.array(palette, np.uint8)[idx]), rtol=3e-6, atol=1e-7)
    v = rng.integers(0, 1 << 15, (5, 7), dtype=np.uint16)
    got = check(tmp_path / "h.tga", _tga(v, bits16=True))
    want = np.stack([((v >> 10) & 31) * 255 // 31, ((v >> 5) & 31) * 255 // 31, (v & 31) * 255 // 31], -1).astype(np.uint8)
    assert np.allclose(got[..., :3], _lin(want), rtol=3e-6, atol=1e-7)


def test_tga_needs_its_extension_and_a_sane_header(tmp_path):
    img = np.zeros((4, 4, 3), np.uint8)
    (tmp_path / "x.dat").write_bytes(_tga(img))  # TGA has no signature: only files named .tga are tried
    with pytest.raises(pupil.PupilError):
        pupil.image_load(tmp_path / "x.dat")
    (tmp_path / "cut.tga").write_bytes(_tga(img)[:-10])
    with pytest.raises(pupil.PupilError):
        pupil.image_load(tmp_path / "cut.tga")


# ---- PNG, Adam7 -----------------------------------------------------------------------------------------------------
def _png_adam7(arr, color_type, depth=8):
    h, w, ch = arr.shape
    raw = b""
    for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
        sub = arr[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        for row in sub:
            if depth == 16:
                line = row.astype(">u2").tobytes()
            elif depth == 8:
                line = row.astype(np.uint8).tobytes()
            else:
                bits = "".join(format(int(v), f"0{depth}b") for v in row[:, 0])
                bits += "0" * (-len(bits) % 8)
                line = bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))
            raw += b"\0" + line  # filter type 0; the filters themselves are covered by test_image_io.py

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 1)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")


@pytest.mark.parametrize("size", [(1, 1), (3, 2), (8, 8), (9, 17), (33, 5)])
def test_png_adam7(tmp_path, size):
    h, w = size
    rng = np.random.default_rng(h * 100 + w)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = check(tmp_path / "i.png", _png_adam7(rgb, 2))
    assert np.allclose(got[..., :3], _lin(rgb), rtol=3e-6, atol=1e-7)
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    got = check(tmp_path / "ia.png", _png_adam7(rgba, 6))
    assert np.allclose(got, expected_rgba(rgba), rtol=3e-6, atol=1e-7)
    grey = rng.integers(0, 4, (h, w, 1), dtype=np.uint8)
    got = check(tmp_path / "g2.png", _png_adam7(grey, 0, depth=2))
    assert np.allclose(got[..., 0], _lin(grey[..., 0] * 85), rtol=3e-6, atol=1e-7)


def test_png_adam7_written_by_pillow_matches_the_flat_file(tmp_path):
    """the same picture stored flat and interlaced decodes to the same texels"""
    img = _picture(29, 43, 21)
    b = io.BytesIO()
    PIL.fromarray(img).save(b, "PNG")
    flat = check(tmp_path / "f.png", b.getvalue())
    inter = check(tmp_path / "i.png", _png_adam7(img, 2))
    assert np.array_equal(flat, inter)


def test_corrupted_files_are_decoded_or_refused_never_worse(tmp_path):
    """byte flips and truncations of valid files: every outcome is an image of the header's size or a PupilError (the decoders were
    also run over 10 800 such files under AddressSanitizer + UBSan while they were written)"""
    rng = np.random.default_rng(3)
    img = _picture(24, 40, 2)
    seeds = {"a.jpg": _jpeg(img, quality=85, subsampling=2), "p.jpg": _jpeg(img, quality=85, subsampling=1, progressive=True),
             "b.bmp": _bmp(5, 7, 24, [bytes(15)] * 7), "t.tga": _tga(img[:9, :13], rle=True), "i.png": _png_adam7(img[:17, :9], 2)}
    pupil.lib().pupil_set_log_level(0)
    try:
        for name, data in seeds.items():
            for it in range(120):
                d = bytearray(data)
                for _ in range(int(rng.integers(1, 6))):
                    d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
                if it % 4 == 0:
                    d = d[:int(rng.integers(1, len(d)))]
                (tmp_path / name).write_bytes(bytes(d))
                try:
                    got = pupil.image_load(tmp_path / name)
                    assert got.ndim == 3 and got.shape[2] == 4 and np.isfinite(got).all()
                except pupil.PupilError:
                    pass
    finally:
        pupil.lib().pupil_set_log_level(1)


# ---- PGM / PPM ------------------------------------------------------------------------------------------------------
def test_binary_pnm(tmp_path):
    rng = np.random.default_rng(14)
    rgb = rng.integers(0, 256, (7, 9, 3), dtype=np.uint8)
    got = check(tmp_path / "a.ppm", b"P6\n# a comment\n9 7\n255\n" + rgb.tobytes())
    assert np.allclose(got, expected_rgba(rgb), rtol=3e-6, atol=1e-7)
    grey = rng.integers(0, 256, (5, 4, 1), dtype=np.uint8)
    got = check(tmp_path / "g.pgm", b"P5 4 5 255\n" + grey.tobytes())
    assert np.allclose(got, expected_rgba(grey), rtol=3e-6, atol=1e-7)
    deep = rng.integers(0, 65536, (3, 5, 3), dtype=np.uint16)
    got = check(tmp_path / "d.ppm", b"P6\n5 3\n65535\n" + deep.astype(">u2").tobytes(), pil=False)  # Pillow rescales 16-bit PPM differently
    assert np.allclose(got, expected_rgba((deep & 255).astype(np.uint8)), rtol=3e-6, atol=1e-7)  # the byte stb_image keeps on little-endian hosts
    for name, data in {"ascii.ppm": b"P3\n1 1\n255\n1 2 3\n", "cut.ppm": b"P6\n9 7\n255\n" + rgb.tobytes()[:-1], "zero.pgm": b"P5 0 5 255\n"}.items():
        (tmp_path / name).write_bytes(data)
        with pytest.raises(pupil.PupilError):
            pupil.image_load(tmp_path / name)


def test_jpeg_four_components(tmp_path):
    """Adobe CMYK (APP14 transform 0, what Pillow writes), and the same file relabelled YCCK (2) and "YCbCr + K" (1): stb_image returns
    RGB for all three, and so does the host library — compared with stb only (Pillow keeps CMYK)"""
    if STB is None:
        pytest.skip("needs the reference's stb_image as checker")
    cmyk = PIL.fromarray(_picture(40, 56, 3)).convert("CMYK")
    for kw in (dict(quality=90), dict(quality=85, progressive=True), dict(quality=90, subsampling=0)):
        b = io.BytesIO()
        cmyk.save(b, "JPEG", **kw)
        data = bytearray(b.getvalue())
        flag = data.index(b"Adobe") + 11
        for transform in (0, 1, 2):
            data[flag] = transform
            got = check(tmp_path / f"c{transform}.jpg", bytes(data), pil=False)
            assert got.shape == (40, 56, 4) and np.all(got[..., 3] == 1.0)
