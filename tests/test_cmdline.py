"""The command line front end keeps the reference's flags (mathmap_cmdline.c:491-521)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_png_rgb

TWIRL = os.path.join(GOLDEN, "filters", "examples", "Distorts", "Twirl.mm")


def run_cli(*args):
    return subprocess.run([sys.executable, "-m", "mathmap_b200.cmdline"] + list(args), cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_compile_only_and_errors_need_no_gpu(tmp_path):
    r = run_cli("--bench-only-compile", "-i", "-f", TWIRL, str(tmp_path / "o.png"))
    assert r.returncode == 0 and "compiled twirl" in r.stderr
    r = run_cli("--bench-no-backend", "filter f (image in) in(xy + q) end", str(tmp_path / "o.png"))
    assert r.returncode == 1 and "Undefined variable q" in r.stderr
    r = run_cli("--bench-only-compile", "-Dnope=1", "-f", TWIRL, str(tmp_path / "o.png"))
    assert r.returncode == 0  # defines are only checked when rendering, like the reference (after compilation)


@pytest.mark.gpu
def test_reference_command_line_reproduces_golden(tmp_path):
    """`mathmap -i -f Twirl.mm -Din=marlene.png out.png` (tests/run_tests.sh:125) -> distorts_twirl.png, bit-exact."""
    out = str(tmp_path / "twirl.png")
    r = run_cli("-i", "-f", TWIRL, "-Din=" + os.path.join(GOLDEN, "png", "marlene.png"), out)
    assert r.returncode == 0, r.stderr
    from PIL import Image
    got = np.array(Image.open(out).convert("RGB"))
    assert np.array_equal(got, load_png_rgb("distorts_twirl.png"))


@pytest.mark.gpu
def test_frames_and_size_flags(tmp_path):
    out = str(tmp_path / "m_%03d.png")
    r = run_cli("-s", "64x48", "-F", "3", "-Dnum_iterations=16", "-f", os.path.join(GOLDEN, "filters", "examples", "Render", "Mandelbrot.mm"), out)
    assert r.returncode == 0, r.stderr
    from PIL import Image
    for i in range(3):
        assert Image.open(str(tmp_path / ("m_%03d.png" % i))).size == (64, 48)
    r = run_cli("-f", TWIRL, str(tmp_path / "x.png"))
    assert r.returncode == 1 and "image size not set" in r.stderr
