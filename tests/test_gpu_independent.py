"""The filters tests/test_independent_numpy.py evaluates by hand (which pins the oracle and the front end away from the square
goldens) rendered on the device against the oracle, at the same non-square sizes and t != 0: closures and filter calls,
render() of a closure (whose t is 0 inside the rendering), stretched / pixel coordinate systems, curves and gradients, HSV
conversions, vector builtins, guards."""
import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import synthetic_rgba
from oracle.oracle import OracleFilter

SOURCES = {
    "closure_and_call": ("filter inner (image in, float k: 0-2 (1))\n  in(xy * k)\nend\n\n"
                         "filter outer (image in)\n  half = inner(in, 0.5);\n  half(xy + xy:[0.1, 0]) * 0.5 + inner(in, 1.5, xy * 0.5) * 0.25\nend\n"),
    "render_of_closure": ("filter inner (image in, float k: 0-2 (1))\n  in(xy * k + xy:[t, 0])\nend\n\n"
                          "filter outer (image in)\n  rr = render(inner(in, 0.8));\n  rr(xy * 0.9 + xy:[t * 0.1, 0])\nend\n"),
    "stretched": "stretched filter s (stretched image in)\n  in(xy * 0.8 + xy:[0.1, -0.05] + xy:[W, H] * 0.01)\nend\n",
    "pixel": "pixel filter p (pixel image in)\n  in(xy + xy:[3.5, -2.25])\nend\n",
    "default_on_pixel_image": "filter m (pixel image in)\n  in(xy * 20)\nend\n",
    "hsv": "filter c (image in)\n  h = toHSVA(in(xy));\n  toRGBA(hsva:[h[0] + t, h[1] * 0.75, h[2], h[3]])\nend\n",
    "guards": ("filter g ()\n  k = floor(y * 3);\n"
               "  rgba:[x / k / 8 + 0.5, asin(x * 2) / 3 + 0.5, log(x) / 6 + 0.5, (x * 4) % k / 8 + 0.5]\nend\n"),
    "curve_gradient": "filter cg (curve c, gradient g)\n  v = c(x * 0.6 + 0.5);\n  g(v * 0.8 + y * 0.3)\nend\n",
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SOURCES))
def test_hand_evaluated_filters_on_the_device(name):
    W, H, t = 93, 58, 0.3
    m = mb.Module(source=SOURCES[name])
    oracle = OracleFilter(m.ir)
    vals = {}
    for uname, kind, _lo, _hi, _default in m.uservals():
        if kind == mb.USERVAL_IMAGE:
            vals[uname] = synthetic_rgba(70, 81, seed=23)
        elif kind == mb.USERVAL_CURVE:
            vals[uname] = np.sqrt(np.arange(1024, dtype=np.float64) / 1023.0).astype(np.float32)
        elif kind == mb.USERVAL_GRADIENT:
            vals[uname] = np.random.RandomState(5).randint(0, 1 << 32, size=1024, dtype=np.uint64).astype(np.uint32)
    for aa in (False, True):
        inv = mb.Invocation(m, W, H, antialiasing=aa)
        for k, v in vals.items():
            inv.set(k, v)
        got = inv.render(0, t)
        want = oracle.render(W, H, vals, t=t, antialiasing=aa)
        assert np.array_equal(got, want), "%s aa=%s: %d pixels differ" % (name, aa, int((np.abs(got.astype(int) - want.astype(int)).max(axis=2) > 0).sum()))
