"""A plain C program (tests/c_abi_smoke.c) against include/rl_b200.h + librl_b200.so: compiled with gcc, no Python or C++ in
between.  Without a GPU it must fail loudly in rl_create (exit code 3: the product has no CPU path); on the GPU box it renders."""
import os
import subprocess

import pytest

from conftest import ROOT, has_gpu


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    pkg = os.path.join(ROOT, "rustlight_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"),
                           "-L" + pkg, "-lrl_b200", "-lrl_host", "-Wl,-rpath," + pkg, "-lm", "-o", exe])
    return exe


def test_c_program_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if has_gpu():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr and "no CPU path" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_program_renders_through_the_c_abi(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout
