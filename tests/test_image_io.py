"""CPU-only: the host library's image reader / writer (pupiloptixlab_b200/host/image.cpp) — the role of
util::BitmapTexture::Load / Save (framework/util/texture.cpp:13-174; stb_image, stb_image_write and tinyexr there).
Files are produced here by independent Python encoders (zlib, struct), so the decoders are checked against the
formats' definitions, and the writers against the decoders and the reference's conventions (vertical flip on save,
B,G,R float channels in EXR, pow(x/255, 2.2) on 8-bit sources)."""
import struct
import zlib

import numpy as np
import pytest

from pupiloptixlab_b200 import pupil

F = np.float32


def _png(path, arr, color_type, depth=8, filters=(0, 1, 2, 3, 4), palette=None, trns=None):
    """arr: (h, w, channels) uint8/uint16 samples (palette: indices).  Rows cycle through the five PNG filter types."""
    h, w = arr.shape[:2]
    ch = arr.shape[2]
    if depth == 16:
        raw_rows = [arr[y].astype(">u2").tobytes() for y in range(h)]
    elif depth == 8:
        raw_rows = [arr[y].astype(np.uint8).tobytes() for y in range(h)]
    else:  # packed sub-byte samples, one channel
        raw_rows = []
        for y in range(h):
            bits = "".join(format(int(v), f"0{depth}b") for v in arr[y, :, 0])
            bits += "0" * (-len(bits) % 8)
            raw_rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
    bpp = max(1, ch * depth // 8)
    out, prev = b"", bytes(len(raw_rows[0]))
    for y, row in enumerate(raw_rows):
        f = filters[y % len(filters)]
        enc = bytearray(len(row))
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            if f == 0:
                p = 0
            elif f == 1:
                p = a
            elif f == 2:
                p = b
            elif f == 3:
                p = (a + b) >> 1
            else:
                pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                p = a if pa <= pb and pa <= pc else (b if pb <= pc else c)
            enc[i] = (v - p) & 255
        out += bytes([f]) + bytes(enc)
        prev = row

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color_type, 0, 0, 0))
    if palette is not None:
        data += chunk(b"PLTE", bytes(palette))
    if trns is not None:
        data += chunk(b"tRNS", bytes(trns))
    comp = zlib.compress(out, 6)
    data += chunk(b"IDAT", comp[:len(comp) // 2]) + chunk(b"IDAT", comp[len(comp) // 2:]) + chunk(b"IEND", b"")
    path.write_bytes(data)


def _lin(u8):
    return np.power(np.asarray(u8, F) * F(1.0) / F(255.0), F(2.2)).astype(F)


def test_png_rgb_rgba_grey_16bit(tmp_path):
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (9, 7, 3), dtype=np.uint16)
    _png(tmp_path / "rgb.png", rgb, 2)
    got = pupil.image_load(tmp_path / "rgb.png")
    assert got.shape == (9, 7, 4)
    assert np.allclose(got[..., :3], _lin(rgb), rtol=2e-6, atol=1e-7) and np.all(got[..., 3] == 1.0)
    rgba = rng.integers(0, 256, (5, 11, 4), dtype=np.uint16)
    _png(tmp_path / "rgba.png", rgba, 6)
    got = pupil.image_load(tmp_path / "rgba.png")
    assert np.allclose(got[..., :3], _lin(rgba[..., :3]), rtol=2e-6, atol=1e-7)
    assert np.array_equal(got[..., 3], (rgba[..., 3].astype(F) * F(1.0) / F(255.0)))
    grey = rng.integers(0, 256, (4, 6, 1), dtype=np.uint16)
    _png(tmp_path / "g.png", grey, 0)
    got = pupil.image_load(tmp_path / "g.png")
    assert np.allclose(got[..., 0], _lin(grey[..., 0]), rtol=2e-6) and np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 0], got[..., 2])
    ga = rng.integers(0, 256, (4, 6, 2), dtype=np.uint16)
    _png(tmp_path / "ga.png", ga, 4)
    got = pupil.image_load(tmp_path / "ga.png")
    assert np.allclose(got[..., 2], _lin(ga[..., 0]), rtol=2e-6) and np.allclose(got[..., 3], ga[..., 1] / 255.0, rtol=1e-6)
    deep = rng.integers(0, 65536, (6, 5, 3), dtype=np.uint16)
    _png(tmp_path / "deep.png", deep, 2, depth=16)
    got = pupil.image_load(tmp_path / "deep.png")
    assert np.allclose(got[..., :3], _lin(deep >> 8), rtol=2e-6, atol=1e-7)  # 16 -> 8 bit keeps the high byte (stb_image's 8-bit API)


@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_png_palette_and_packed_grey(tmp_path, depth):
    rng = np.random.default_rng(depth)
    n = 1 << depth
    idx = rng.integers(0, n, (7, 13, 1), dtype=np.uint16)
    pal = rng.integers(0, 256, (n, 3), dtype=np.uint16)
    trns = rng.integers(0, 256, n // 2 or 1, dtype=np.uint16)
    _png(tmp_path / "p.png", idx, 3, depth=depth, palette=pal.reshape(-1).tolist(), trns=trns.tolist())
    got = pupil.image_load(tmp_path / "p.png")
    assert np.allclose(got[..., :3], _lin(pal[idx[..., 0]]), rtol=2e-6, atol=1e-7)
    alpha = np.where(idx[..., 0] < len(trns), trns[np.minimum(idx[..., 0], len(trns) - 1)], 255)
    assert np.allclose(got[..., 3], alpha / 255.0, rtol=1e-6)
    if depth < 8:
        _png(tmp_path / "g.png", idx, 0, depth=depth)
        got = pupil.image_load(tmp_path / "g.png")
        assert np.allclose(got[..., 0], _lin(idx[..., 0] * (255 // (n - 1))), rtol=2e-6, atol=1e-7)


def _rgbe(img):
    m = img[..., :3].max(-1)
    e = np.where(m < 1e-32, 0, np.frexp(m)[1])
    scale = np.where(m < 1e-32, 0.0, np.frexp(m)[0] * 256.0 / np.maximum(m, 1e-38))
    out = np.zeros(img.shape[:2] + (4,), np.uint8)
    out[..., :3] = (img[..., :3] * scale[..., None]).astype(np.uint8)
    out[..., 3] = np.where(m < 1e-32, 0, e + 128).astype(np.uint8)
    return out


def _decode_rgbe(q):
    f = np.ldexp(F(1.0), q[..., 3].astype(np.int32) - 136).astype(F)
    return np.where(q[..., 3:4] == 0, F(0), q[..., :3].astype(F) * f[..., None])


def test_hdr_flat_and_rle(tmp_path):
    rng = np.random.default_rng(2)
    img = (rng.random((6, 40, 3)) * np.array([0.01, 1.0, 300.0])).astype(F)
    img[2, 3] = 0.0
    q = _rgbe(img)
    head = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 6 +X 40\n"
    (tmp_path / "flat.hdr").write_bytes(head + q.tobytes())
    got = pupil.image_load(tmp_path / "flat.hdr")
    assert got.shape == (6, 40, 4) and np.array_equal(got[..., :3], _decode_rgbe(q)) and np.all(got[..., 3] == 1.0)
    assert np.all(np.abs(got[..., :3] - img) <= img.max(-1, keepdims=True) / 128 + 1e-6)  # shared exponent: 8 bits of the largest channel
    body = b""
    for y in range(6):  # new-style RLE: per channel, runs (>128) and literals mixed
        body += bytes([2, 2, 0, 40])
        for k in range(4):
            row, x = q[y, :, k].tolist(), 0
            while x < 40:
                run = 1
                while x + run < 40 and row[x + run] == row[x] and run < 127:
                    run += 1
                if run >= 3:
                    body += bytes([128 + run, row[x]])
                    x += run
                else:
                    n = min(5, 40 - x)
                    body += bytes([n] + row[x:x + n])
                    x += n
    (tmp_path / "rle.hdr").write_bytes(head + body)
    assert np.array_equal(pupil.image_load(tmp_path / "rle.hdr"), got)


def test_pfm_and_hdr_and_exr_writers_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    buf = np.ones((21, 33, 4), F)  # frame-buffer order: row 0 = bottom
    buf[..., :3] = (rng.random((21, 33, 3)) * 4.0).astype(F)
    pupil.image_save(tmp_path / "a.pfm", buf)
    got = pupil.image_load(tmp_path / "a.pfm")   # load returns file order: row 0 = top
    assert np.array_equal(got[..., :3], buf[::-1, :, :3]) and np.all(got[..., 3] == 1.0)
    raw = (tmp_path / "a.pfm").read_bytes()
    assert raw.startswith(b"PF\n33 21\n-1.0\n")
    assert np.array_equal(np.frombuffer(raw[-21 * 33 * 12:], "<f4").reshape(21, 33, 3), buf[..., :3])  # PFM itself is bottom-up
    pupil.image_save(tmp_path / "a.exr", buf)
    got = pupil.image_load(tmp_path / "a.exr")
    assert np.array_equal(got[..., :3], buf[::-1, :, :3]) and np.all(got[..., 3] == 1.0)  # flipped on save (texture.cpp:37-44)
    exr = (tmp_path / "a.exr").read_bytes()
    assert exr[:4] == bytes([0x76, 0x2f, 0x31, 0x01]) and exr.index(b"B\0") < exr.index(b"G\0") < exr.index(b"R\0")
    pupil.image_save(tmp_path / "a.hdr", buf)
    got = pupil.image_load(tmp_path / "a.hdr")
    assert np.all(np.abs(got[..., :3] - buf[::-1, :, :3]) <= buf[::-1, :, :3].max(-1, keepdims=True) / 128 + 1e-6)


def test_png_display_output(tmp_path):
    """what the reference's canvas shows (system/gui/output.hlsl:30-72): gamma 2.2 by default, ACES tone mapping on request"""
    rng = np.random.default_rng(6)
    buf = np.ones((17, 23, 4), F)
    buf[..., :3] = (rng.random((17, 23, 3)) ** 2 * 3.0).astype(F)  # some values above 1
    pupil.image_save(tmp_path / "g.png", buf, "png")
    got = pupil.image_load(tmp_path / "g.png")  # the reader linearises 8-bit sources again: pow(x / 255, 2.2)
    want = np.clip(buf[::-1, :, :3], 0, 1)
    u8 = np.floor(np.power(want, F(1 / 2.2)) * 255 + 0.5)
    assert np.allclose(got[..., :3], np.power(u8 / 255, 2.2), atol=1e-6) and np.abs(got[..., :3] - want).max() < 0.02
    pupil.image_save(tmp_path / "a.png", buf, "png_aces")
    got = pupil.image_load(tmp_path / "a.png")
    c = buf[::-1, :, :3].astype(np.float64)
    aces = (c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14)
    u8 = np.floor(np.clip(np.power(aces, 1 / 2.2), 0, 1) * 255 + 0.5)
    back = np.round(np.power(got[..., :3].astype(np.float64), 1 / 2.2) * 255)
    assert np.abs(back - u8).max() <= 1  # fp32 pow on either side of a rounding boundary


def _exr(path, img, compression, half):
    """minimal scan-line EXR written by hand: channels A? no — B, G, R (+ optional HALF), compression 0 / 2 / 3"""
    h, w = img.shape[:2]
    def attr(name, typ, body):
        return name + b"\0" + typ + b"\0" + struct.pack("<I", len(body)) + body
    ch = b""
    for n in (b"B", b"G", b"R"):
        ch += n + b"\0" + struct.pack("<iBBBBii", 1 if half else 2, 0, 0, 0, 0, 1, 1)
    ch += b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    hdr = struct.pack("<II", 20000630, 2) + attr(b"channels", b"chlist", ch) + attr(b"compression", b"compression", bytes([compression]))
    hdr += attr(b"dataWindow", b"box2i", box) + attr(b"displayWindow", b"box2i", box) + attr(b"lineOrder", b"lineOrder", b"\0")
    hdr += attr(b"pixelAspectRatio", b"float", struct.pack("<f", 1)) + attr(b"screenWindowCenter", b"v2f", struct.pack("<ff", 0, 0))
    hdr += attr(b"screenWindowWidth", b"float", struct.pack("<f", 1)) + b"\0"
    lines = {0: 1, 2: 1, 3: 16}[compression]
    blocks = []
    for y0 in range(0, h, lines):
        raw = b""
        for y in range(y0, min(h, y0 + lines)):
            for c in (2, 1, 0):
                raw += img[y, :, c].astype("<f2" if half else "<f4").tobytes()
        if compression:
            a = np.frombuffer(raw, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
            t[1:] = (t[1:] - t[:-1] + 128) & 255
            comp = zlib.compress(t.astype(np.uint8).tobytes())
            raw = comp if len(comp) < len(raw) else raw
        blocks.append(struct.pack("<iI", y0, len(raw)) + raw)
    off = len(hdr) + 8 * len(blocks)
    table = b""
    for b in blocks:
        table += struct.pack("<Q", off)
        off += len(b)
    path.write_bytes(hdr + table + b"".join(blocks))


@pytest.mark.parametrize("compression,half", [(0, False), (2, False), (3, False), (3, True), (0, True)])
def test_exr_reader(tmp_path, compression, half):
    rng = np.random.default_rng(4)
    img = (rng.random((37, 19, 3)) * 8.0).astype(F)
    img[5:9] = 0.25  # compressible rows
    _exr(tmp_path / "t.exr", img, compression, half)
    got = pupil.image_load(tmp_path / "t.exr")
    want = img.astype(np.float16).astype(F) if half else img
    assert got.shape == (37, 19, 4) and np.array_equal(got[..., :3], want) and np.all(got[..., 3] == 1.0)


def test_unreadable_images_fail_loudly(tmp_path):
    (tmp_path / "x.jpg").write_bytes(b"\xff\xd8\xff\xe0 not really")
    with pytest.raises(pupil.PupilError):
        pupil.image_load(tmp_path / "x.jpg")
    with pytest.raises(pupil.PupilError):
        pupil.image_load(tmp_path / "missing.png")
    (tmp_path / "trunc.png").write_bytes(b"\x89PNG\r\n\x1a\n" + b"\0" * 5)
    with pytest.raises(pupil.PupilError):
        pupil.image_load(tmp_path / "trunc.png")


def test_scene_with_image_files_loads(tmp_path, port_lib):
    """<texture type="bitmap"> and <emitter type="envmap"> with real files below the scene directory (scene.cpp:144-166,207-219)"""
    import orc
    from pupiloptixlab_b200 import scenes
    desc = scenes.envmap_scene(48, 27, 4)
    env_img = desc.env_map.image
    pupil.image_save(tmp_path / "sky.pfm", env_img[::-1])  # saved bottom-up so the file's first row is env_img[0]
    desc.env_map.filename = "sky.pfm"
    xml = scenes.to_xml(desc, tmp_path / "scene.xml")
    for t in scenes.images_of(desc):  # the bitmap textures stay in memory
        pupil.lib().pupil_register_image(scenes.image_name(t).encode(), t.image.ctypes.data, t.image.shape[1], t.image.shape[0])
    pupil.parse_scene_xml(xml)
    areas, env = pupil.emitters()
    assert env is not None and env.type == 4 and (env.map_w, env.map_h) == (env_img.shape[1], env_img.shape[0])
    o = orc.OracleScene(port_lib, desc)
    rc, rw, cc = pupil.env_tables()
    assert np.array_equal(rc, np.ctypeslib.as_array(o.env_emitter().row_cdf, (len(rc),)))
