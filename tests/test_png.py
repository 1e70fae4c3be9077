"""PNG on both sides of the path (the reference's default `image` feature): Bitmap::read_ldr_image = image::open(..).to_rgb8() / 255
(structure.rs:649-668) for textures and environment maps, Bitmap::save_ldr_image + Color::to_rgba (structure.rs:160-167, 471-485)
for `-o out.png`.  The files are made / decoded here with zlib and struct only, independently of the C++ reader / writer."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from rustlight_b200 import SceneLoaderManager
from rustlight_b200.host import SceneError, read_image, save_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def make_png(rows, width, depth, ctype, filters, palette=None, interlace=0, extra=b""):
    """rows: list of bytes objects (packed samples of one scanline each); filters: filter type per row."""
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = max(1, channels * depth // 8)
    out, prev = bytearray(), bytes(len(rows[0]))
    for row, ft in zip(rows, filters):
        enc = bytearray()
        for i, x in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) // 2, _paeth(a, b, c))[ft]
            enc.append((x - pred) & 0xFF)
        out += bytes([ft]) + enc
        prev = row
    ihdr = struct.pack(">IIBBBBB", width, len(rows), depth, ctype, 0, 0, interlace)
    z = zlib.compress(bytes(out))
    half = len(z) // 2  # two IDAT chunks: the reader must concatenate them
    body = _chunk(b"IHDR", ihdr) + extra
    if palette is not None:
        body += _chunk(b"PLTE", bytes(palette))
    return b"\x89PNG\r\n\x1a\n" + body + _chunk(b"IDAT", z[:half]) + _chunk(b"IDAT", z[half:]) + _chunk(b"IEND", b"")


def decode_rgb8_png(data):
    """Minimal decoder for what Bitmap::save_png writes: 8-bit RGB, any row filter, one or more IDAT chunks."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(kind + body) & 0xFFFFFFFF
        if kind == b"IHDR":
            w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", body)
            assert (depth, ctype, comp, flt, inter) == (8, 2, 0, 0, 0)
        elif kind == b"IDAT":
            idat += body
        pos += 12 + n
    raw = zlib.decompress(idat)
    stride = 3 * w
    img = np.zeros((h, w, 3), np.uint8)
    prev = bytes(stride)
    for y in range(h):
        ft = raw[(stride + 1) * y]
        row = bytearray(raw[(stride + 1) * y + 1:(stride + 1) * (y + 1)])
        for i in range(stride):
            a = row[i - 3] if i >= 3 else 0
            b = prev[i]
            c = prev[i - 3] if i >= 3 else 0
            row[i] = (row[i] + (0, a, b, (a + b) // 2, _paeth(a, b, c))[ft]) & 0xFF
        img[y] = np.frombuffer(bytes(row), np.uint8).reshape(w, 3)
        prev = bytes(row)
    return img


@pytest.mark.parametrize("ctype,depth", [(2, 8), (6, 8), (0, 8), (4, 8), (2, 16), (6, 16), (0, 16), (4, 16)])
def test_png_reader_colour_types_and_filters(tmp_path, ctype, depth):
    rng = np.random.default_rng(ctype * 100 + depth)
    w, h = 13, 10
    channels = {0: 1, 2: 3, 4: 2, 6: 4}[ctype]
    maxv = (1 << depth) - 1
    px = rng.integers(0, maxv + 1, size=(h, w, channels), dtype=np.uint32)
    px[0, 0], px[0, 1] = 0, maxv
    rows = [px[y].astype(">u2" if depth == 16 else np.uint8).tobytes() for y in range(h)]
    filters = [y % 5 for y in range(h)]  # none, sub, up, average, paeth -- twice
    path = tmp_path / "t.png"
    path.write_bytes(make_png(rows, w, depth, ctype, filters, extra=_chunk(b"gAMA", struct.pack(">I", 45455))))  # gamma is ignored
    got = read_image(str(path))
    v8 = (px + 128) // 257 if depth == 16 else px  # image 0.24: u16 -> u8
    rgb = v8[..., :3] if channels >= 3 else np.repeat(v8[..., :1], 3, axis=2)  # alpha dropped, luma replicated
    assert got.dtype == np.float32 and got.shape == (h, w, 3)
    assert np.array_equal(got, rgb.astype(np.float32) / np.float32(255.0))


@pytest.mark.parametrize("depth", [1, 2, 4])
def test_png_reader_low_bit_depths_and_palette(tmp_path, depth):
    rng = np.random.default_rng(depth)
    w, h = 11, 6  # not a multiple of the samples per byte
    idx = rng.integers(0, 1 << depth, size=(h, w), dtype=np.uint32)

    def pack(row):
        bits = "".join(format(int(v), f"0{depth}b") for v in row)
        bits += "0" * (-len(bits) % 8)
        return bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))

    rows = [pack(idx[y]) for y in range(h)]
    grey = tmp_path / "g.png"
    grey.write_bytes(make_png(rows, w, depth, 0, [0, 1, 2, 3, 4, 1]))
    scale = 255 // ((1 << depth) - 1)  # sub-byte grey samples are expanded to the full 8-bit range
    assert np.array_equal(read_image(str(grey)), np.repeat((idx * scale)[..., None], 3, axis=2).astype(np.float32) / np.float32(255.0))
    pal = rng.integers(0, 256, size=((1 << depth), 3), dtype=np.uint8)
    palf = tmp_path / "p.png"
    palf.write_bytes(make_png(rows, w, depth, 3, [0] * h, palette=pal.tobytes()))
    assert np.array_equal(read_image(str(palf)), pal[idx].astype(np.float32) / np.float32(255.0))


def test_png_reader_refuses_what_it_cannot_read(tmp_path):
    rows = [bytes(6)] * 2
    ok = make_png(rows, 2, 8, 2, [0, 0])
    p = tmp_path / "x.png"
    p.write_bytes(make_png(rows, 2, 8, 2, [0, 0], interlace=1))
    with pytest.raises(SceneError):
        read_image(str(p))  # Adam7
    bad = bytearray(ok)
    bad[40] ^= 0xFF  # inside the first IDAT: CRC mismatch
    p.write_bytes(bytes(bad))
    with pytest.raises(SceneError):
        read_image(str(p))
    p.write_bytes(ok[:30])
    with pytest.raises(SceneError):
        read_image(str(p))
    p.write_bytes(b"P6 not a png")
    with pytest.raises(SceneError):
        read_image(str(p))
    with pytest.raises(SceneError):
        read_image(str(tmp_path / "missing.png"))
    with pytest.raises(SceneError):
        read_image(str(tmp_path / "noextension"))


def test_png_writer_is_to_rgba(tmp_path):
    """(min(c, 1)^(1/2.2) * 255) as u8: truncation, > 1 and NaN clamp to 1 (f32::min returns the other operand), negatives -> NaN -> 0
    (Rust's saturating cast)."""
    rng = np.random.default_rng(7)
    img = rng.random((9, 14, 3), dtype=np.float32) * np.float32(1.3) - np.float32(0.1)
    img[0, 0] = (0.0, 1.0, 2.5)
    img[0, 1] = (-0.5, np.nan, np.inf)
    img[0, 2] = (1e-30, 0.5, 0.999999)
    path = tmp_path / "o.png"
    save_image(str(path), img)
    got = decode_rgb8_png(path.read_bytes())
    with np.errstate(invalid="ignore"):
        v = np.power(np.fmin(img, np.float32(1.0)).astype(np.float64), np.float64(np.float32(1.0) / np.float32(2.2))) * 255.0
    want = np.where(v > 0, np.minimum(np.floor(v), 255), 0)  # NaN compares false
    assert got.shape == (9, 14, 3)
    # powf in f32 (glibc) against pow in f64: a value within 1e-4 of an integer may truncate to the neighbour
    near = np.abs(v - np.round(v)) < 1e-4
    assert np.array_equal(got[~near], want[~near].astype(np.uint8))
    assert np.all(np.abs(got[near].astype(int) - want[near].astype(int)) <= 1)
    assert tuple(got[0, 0]) == (0, 255, 255) and tuple(got[0, 1]) == (0, 255, 255)  # f32::min(NaN, 1) = 1 and got[0, 2, 0] == 0
    # what was written reads back as value / 255
    assert np.array_equal(read_image(str(path)), got.astype(np.float32) / np.float32(255.0))


def test_png_texture_and_cli_output(tmp_path):
    """A PBRT imagemap texture from a .png equals the same texels given as a .ppm; the extension of -o picks the writer."""
    rng = np.random.default_rng(3)
    tex = rng.integers(0, 256, size=(4, 5, 3), dtype=np.uint8)
    (tmp_path / "t.png").write_bytes(make_png([tex[y].tobytes() for y in range(4)], 5, 8, 2, [4, 3, 2, 1]))
    (tmp_path / "t.ppm").write_bytes(b"P6\n5 4\n255\n" + tex.tobytes())
    scenes = {}
    for ext in ("png", "ppm"):
        src = open(os.path.join(ROOT, "data", "cbox.pbrt")).read()
        src = src.replace("WorldBegin", f'WorldBegin\nTexture "tx" "spectrum" "imagemap" "string filename" "t.{ext}"', 1)
        f = tmp_path / f"s_{ext}.pbrt"
        f.write_text(src)
        scenes[ext] = SceneLoaderManager().load(str(f)).to_json()
    assert scenes["png"] == scenes["ppm"] and '"bitmap"' in scenes["png"]
    cli = os.path.join(ROOT, "rustlight_b200", "rustlight-b200")
    r = subprocess.run([cli, "-o", str(tmp_path / "out.exr"), os.path.join(ROOT, "data", "cbox.pbrt"), "path"], capture_output=True, text=True)
    assert r.returncode != 0 and ".png" in r.stderr


def test_buffer_collection_saves_by_extension(tmp_path):
    """BufferCollection::save -> Bitmap::save (structure.rs:528-545): the extension picks the writer; anything else is refused."""
    from rustlight_b200.device import BufferCollection
    from rustlight_b200.host import read_pfm
    img = np.random.default_rng(1).random((6, 7, 3), dtype=np.float32)
    bc = BufferCollection(img)
    bc.save("primal", str(tmp_path / "a.pfm"))
    bc.save("primal", str(tmp_path / "a.png"))
    assert np.array_equal(read_pfm(str(tmp_path / "a.pfm")), img)
    assert np.abs(read_image(str(tmp_path / "a.png")) - np.minimum(img, 1.0) ** (1 / 2.2)).max() < 1.0 / 255.0 + 1e-6
    with pytest.raises(SceneError):
        bc.save("primal", str(tmp_path / "a.exr"))
    with pytest.raises(SceneError):
        bc.save("primal", str(tmp_path / "a.bmp"))
