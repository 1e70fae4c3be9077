"""JPEG reader (rustlight_b200/host/jpeg.cpp): Bitmap::read_ldr_image for .jpg -- the format of the reference's texture-light picture
(butterfly.jpg, examples/cli.rs:424: progressive, 4:2:0).  The files are written by PIL (libjpeg-turbo) and the decoded pixels must equal
PIL's own decoding bit for bit: baseline and progressive, 4:4:4 / 4:2:2 / 4:2:0, grey, odd sizes, restart intervals, optimised tables."""
import io
import os

import numpy as np
import pytest

from rustlight_b200 import SceneLoaderManager
from rustlight_b200.host import SceneError, read_image

PIL = pytest.importorskip("PIL.Image")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def picture(w, h, seed):
    """Smooth gradients + edges + noise: every coefficient band and both chroma planes carry signal."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 120 * np.sin(x / 7.0 + seed) * np.cos(y / 5.0), 255.0 * x / max(w - 1, 1), 255.0 * ((x // 6 + y // 4) % 2)], axis=2)
    img += rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [dict(quality=90, subsampling=0), dict(quality=75, subsampling=1), dict(quality=75, subsampling=2), dict(quality=30, subsampling=2),
         dict(quality=85, subsampling=2, progressive=True), dict(quality=60, subsampling=0, progressive=True), dict(quality=95, subsampling=1, progressive=True),
         dict(quality=80, subsampling=2, optimize=True), dict(quality=80, subsampling=2, restart_marker_blocks=3),
         dict(quality=80, subsampling=2, progressive=True, restart_marker_rows=1)]


@pytest.mark.parametrize("size", [(64, 48), (37, 29), (8, 8), (1, 1), (17, 50)])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_jpeg_equals_libjpeg(tmp_path, size, case):
    w, h = size
    kw = CASES[case]
    src = picture(w, h, case)
    path = tmp_path / "t.jpg"
    try:
        PIL.fromarray(src).save(path, "JPEG", **kw)
    except TypeError:
        pytest.skip("this PIL does not take these encoder options")
    want = np.asarray(PIL.open(path).convert("RGB"), np.uint8)
    got = read_image(str(path))
    assert got.shape == (h, w, 3)
    assert np.array_equal(got, want.astype(np.float32) / np.float32(255.0))


@pytest.mark.parametrize("progressive", [False, True])
def test_grey_jpeg(tmp_path, progressive):
    src = picture(45, 31, 3)[..., 0]
    path = tmp_path / "g.jpeg"
    PIL.fromarray(src, "L").save(path, "JPEG", quality=80, progressive=progressive)
    want = np.asarray(PIL.open(path).convert("RGB"), np.uint8)
    assert np.array_equal(read_image(str(path)), want.astype(np.float32) / np.float32(255.0))


def test_jpeg_texture_and_refusals(tmp_path):
    src = picture(16, 16, 1)
    PIL.fromarray(src).save(tmp_path / "t.jpg", "JPEG", quality=90)
    want = np.asarray(PIL.open(tmp_path / "t.jpg").convert("RGB"), np.float32) / np.float32(255.0)
    pbrt = open(os.path.join(ROOT, "data", "cbox.pbrt")).read().replace("WorldBegin", 'WorldBegin\nTexture "tx" "spectrum" "imagemap" "string filename" "t.jpg"', 1)
    (tmp_path / "s.pbrt").write_text(pbrt)
    sc = SceneLoaderManager().load(str(tmp_path / "s.pbrt"))
    t = sc.desc.contents.textures[0]
    assert (t.width, t.height) == (16, 16) and np.array_equal(np.ctypeslib.as_array(t.pixels, (16, 16, 3)), want)
    data = (tmp_path / "t.jpg").read_bytes()
    (tmp_path / "cut.jpg").write_bytes(data[:60])
    with pytest.raises(SceneError):
        read_image(str(tmp_path / "cut.jpg"))
    (tmp_path / "no.jpg").write_bytes(b"\x89PNG\r\n\x1a\n")
    with pytest.raises(SceneError):
        read_image(str(tmp_path / "no.jpg"))
    cmyk = io.BytesIO()
    PIL.fromarray(np.zeros((8, 8, 4), np.uint8), "CMYK").save(cmyk, "JPEG")
    (tmp_path / "cmyk.jpg").write_bytes(cmyk.getvalue())
    with pytest.raises(SceneError):
        read_image(str(tmp_path / "cmyk.jpg"))  # 4 components: refused


@pytest.mark.parametrize("mode,rle", [("RGB", False), ("RGB", True), ("RGBA", False), ("RGBA", True), ("L", False), ("L", True), ("P", False)])
def test_tga_equals_pil(tmp_path, mode, rle):
    """Truevision TGA (textures of the pbrt-v3 scenes): true colour / grey / colour-mapped, raw and RLE, bottom-up and top-down rows."""
    src = picture(23, 17, 4)
    src[5:9, 3:20] = (200, 10, 30)  # runs for the RLE packets
    im = PIL.fromarray(src)
    im = im.convert("RGBA") if mode == "RGBA" else im.convert("L") if mode == "L" else im.quantize(64) if mode == "P" else im
    for orientation in (1, -1):
        path = tmp_path / "t.tga"
        im.save(path, "TGA", compression="tga_rle" if rle else None, orientation=orientation)
        want = np.asarray(PIL.open(path).convert("RGB"), np.uint8)
        got = read_image(str(path))
        assert np.array_equal(got, want.astype(np.float32) / np.float32(255.0))
