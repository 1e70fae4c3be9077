"""`rustlight-b200` command line (examples/cli.rs mirror) and the -a / -e wrappers (avg.rs, equal_time.rs)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, ROOT, has_gpu, load_cbox
from rustlight_b200 import _abi
from rustlight_b200.host import read_pfm

CLI = os.path.join(ROOT, "rustlight_b200", "rustlight-b200")
CBOX = os.path.join(DATA, "cbox.pbrt")


def run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=600)


def test_cli_argument_errors():
    assert run("-o", "x.pfm", CBOX).returncode == 2                       # no subcommand
    assert run("-o", "x.exr", CBOX, "path").returncode == 2               # unsupported output type (.pfm and .png are written)
    assert run(CBOX, "path").returncode == 2                              # missing -o
    assert run("-o", "x.pfm", "-m", "0.5", CBOX, "path").returncode == 2  # media are out of scope
    assert run("-o", "x.pfm", "-x", "vpl", CBOX, "path").returncode == 2  # an extra option outside the GPU path (`-x ats` and `-x no-shading` are accepted)
    assert run("-o", "x.pfm", CBOX, "path", "-s", "nope").returncode == 2  # "invalid strategy", cli.rs:536-541
    assert run("-o", "x.pfm", CBOX, "ao", "-q").returncode == 2
    r = run("-o", "x.pfm", "nothing.obj", "path")
    assert r.returncode == 1 and "scene loader" in r.stderr               # scene_loader.rs:40-43


def test_cli_without_gpu_fails_loudly():
    if has_gpu():
        pytest.skip("GPU present")
    r = run("-n", "1", "-o", "/tmp/never.pfm", CBOX, "path")
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cli_matches_library(tmp_path, gpu_ctx):
    from rustlight_b200.device import DeviceScene
    out = str(tmp_path / "a.pfm")
    r = run("-n", "4", "-r", "independent:5", "-s", "0.25", "-o", out, CBOX, "path", "-m", "6", "-s", "all")
    assert r.returncode == 0, r.stderr
    assert "Elapsed Integrator:" in r.stderr                               # integrators/mod.rs:334
    sc = load_cbox().scale_image(0.25)
    img, _ = DeviceScene(gpu_ctx, sc).render(_abi.path_desc(max_depth=6), 4, seed=5)
    assert np.array_equal(read_pfm(out), np.abs(img))
    out3 = str(tmp_path / "ao.pfm")
    assert run("-n", "2", "-s", "0.25", "-o", out3, CBOX, "ao", "-d", "0.5", "-n").returncode == 0
    img, _ = DeviceScene(gpu_ctx, sc).render(_abi.ao_desc(0.5, True), 2, seed=0)
    assert np.array_equal(read_pfm(out3), img)
    out2 = str(tmp_path / "d.pfm")
    assert run("-n", "2", "-s", "0.25", "-o", out2, os.path.join(DATA, "cbox.json"), "direct", "-b", "2", "-l", "1").returncode == 0
    img, _ = DeviceScene(gpu_ctx, sc).render(_abi.direct_desc(2, 1), 2, seed=0)
    assert np.array_equal(read_pfm(out2), np.abs(img))
    # -o x.png: Bitmap::save_ldr_image (gamma 2.2, 8 bit); reading it back gives value / 255 of what the library's own writer produces
    from rustlight_b200.host import read_image, save_image
    outp, refp = str(tmp_path / "d.png"), str(tmp_path / "ref.png")
    assert run("-n", "2", "-s", "0.25", "-o", outp, os.path.join(DATA, "cbox.json"), "direct", "-b", "2", "-l", "1").returncode == 0
    save_image(refp, img)
    got = read_image(outp)
    assert got.shape == img.shape and np.array_equal(got, read_image(refp)) and 0.05 < got.mean() < 0.9


@pytest.mark.gpu
def test_average_wrapper(tmp_path, gpu_ctx):
    """`-a 3`: three passes of 2 spp over sample indices [0,2), [2,4), [4,6), combined exactly like avg.rs:51-62 --
    `iteration` is already 2 on the second pass, so the reference computes (2a+b)/3, then (2a+b+c)/4: the first
    pass keeps a double weight (a quirk of the reference, mirrored) -- one dump per iteration + _time.csv (avg.rs:69-106)."""
    from rustlight_b200.device import DeviceScene
    out = str(tmp_path / "avg.pfm")
    assert run("-n", "2", "-a", "3", "-s", "0.125", "-o", out, CBOX, "path").returncode == 0
    dev = DeviceScene(gpu_ctx, load_cbox().scale_image(0.125))
    passes = [dev.render(_abi.path_desc(), 2, seed=0, sample_offset=2 * p)[0] for p in range(3)]
    acc = passes[0].copy()
    for it, nb in enumerate(passes[1:], start=2):
        acc = ((acc * np.float32(it)) + nb) * np.float32(np.float32(1.0) / np.float32(it + 1))
    assert np.array_equal(read_pfm(out), np.abs(acc))
    for it in (1, 2, 3):
        assert os.path.exists(str(tmp_path / f"avg_{it}.pfm"))
    assert len(open(str(tmp_path / "avg_time.csv")).read().strip().splitlines()) == 3
    # a convex combination (weights 2:1:1) of disjoint sample sets of one 6-spp render
    want = (2.0 * passes[0].astype(np.float64) + passes[1] + passes[2]) / 4.0
    assert np.allclose(acc, want, rtol=2e-5, atol=1e-6)


@pytest.mark.gpu
def test_equal_time_wrapper(tmp_path):
    out = str(tmp_path / "eq.pfm")
    r = run("-n", "1", "-e", "0.2", "-s", "0.25", "-o", out, CBOX, "path")
    assert r.returncode == 0, r.stderr
    img = read_pfm(out)
    assert img.shape == (128, 128, 3) and np.isfinite(img).all() and 0.05 < img.mean() < 0.3
