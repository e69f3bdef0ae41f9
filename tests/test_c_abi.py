"""The C ABI from plain C (tests/c_smoke.c): compile with gcc against include/voxcuda.h, link
libvoxcuda.so, run.  The CPU half needs no GPU; the GPU half runs a tiny pipeline."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "voxlogica_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_smoke")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-O1", os.path.join(HERE, "c_smoke.c"), "-o", exe,
                           "-L" + LIBDIR, "-lvoxcuda", "-Wl,-rpath," + LIBDIR, "-lm"])
    return exe


def test_c_caller_without_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path), "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.gpu
def test_c_caller_on_gpu(tmp_path):
    out = subprocess.run([_build(tmp_path), "gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_smoke gpu ok" in out.stdout
