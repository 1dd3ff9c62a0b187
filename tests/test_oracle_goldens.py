"""The oracle (oracle/) pinned against every golden vector the reference's own
test suite holds for the render path (tests/run_tests.sh: 80 command lines at
256x256, `-i`, t = 0).  This also pins the product's front end, whose IR the
oracle consumes.  CPU only."""
import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import compare_u8, filter_source, load_manifest, load_png_rgb, marlene_rgba
from oracle.oracle import OracleFilter

MANIFEST = load_manifest()
# FFTW-based native filters are a "next" row: no oracle yet
UNSUPPORTED = set()
# goldens whose residual is libm-version noise in glibc's float/complex functions (<= 2 LSB on < 0.05 % of pixels)
BIT_EXACT_EXPECTED_MIN = 99.9


@pytest.mark.parametrize("entry", MANIFEST, ids=[e["golden"] for e in MANIFEST])
def test_oracle_matches_reference_golden(entry):
    if entry["golden"] in UNSUPPORTED:
        pytest.skip("native filter visualize_fft (FFTW) is not restated yet")
    ir = mb.Module(source=filter_source(entry["script"])).ir
    f = OracleFilter(ir)
    uservals = dict(entry["uservals"])
    if entry["kind"] == "modify":
        img = marlene_rgba()
        for typ, name, _ in f.main.uservals:
            if typ == "image":
                uservals.setdefault(name, img)
        h, w = img.shape[:2]
    else:
        w = h = 256
    out = f.render(w, h, uservals, t=0.0, antialiasing=True)
    golden = load_png_rgb(entry["golden"])
    exact, le1, mx = compare_u8(out[:, :, :3], golden)
    assert exact >= BIT_EXACT_EXPECTED_MIN, "only %.4f %% of pixels exact (max diff %d)" % (exact, mx)
    assert le1 >= 99.99 and mx <= 2, "%.4f %% within 1 LSB, max diff %d" % (le1, mx)


def test_bit_exact_goldens_count():
    """The five BASELINE configs' goldens that depend only on exactly-specified arithmetic are bit-exact."""
    for golden, script, uv in [("utilities_ident.png", "examples/Utilities/Ident.mm", {}),
                               ("distorts_twirl.png", "examples/Distorts/Twirl.mm", {}),
                               ("distorts_sea.png", "examples/Distorts/Sea.mm", {}),
                               ("render_mandelbrot.png", "examples/Render/Mandelbrot.mm", {}),
                               ("blur_gaussian_blur.png", "examples/Blur/Gaussian Blur.mm", {"dev": 0.1}),
                               ("render_perlin_noise.png", "examples/Render/Perlin Noise.mm", {})]:
        f = OracleFilter(mb.Module(source=filter_source(script)).ir)
        uservals = dict(uv)
        for typ, name, _ in f.main.uservals:
            if typ == "image":
                uservals[name] = marlene_rgba()
        out = f.render(256, 256, uservals, antialiasing=True)
        assert np.array_equal(out[:, :, :3], load_png_rgb(golden)), golden
