"""Seeded random cases from tools/fuzz_compositions.py and tools/fuzz_uservals.py (the tools take any seed and count; these
are the fixed ones of the suite): compositions of example filters that exist nowhere in the reference's tree, and example
filters away from their default arguments -- booleans pick the specialised kernel variants -- against the oracle."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.gpu
def test_random_compositions_match_oracle():
    import fuzz_compositions
    cases, failures = fuzz_compositions.run(2024, 24)
    assert cases == 24 and not failures, failures


@pytest.mark.gpu
@pytest.mark.parametrize("seed,count,only,settings", [(7, 40, "", False), (8, 12, "Droste", False), (31, 48, "", True)])
def test_random_arguments_match_oracle(seed, count, only, settings):
    """settings: also random edge behaviour and colours, supersampling, bytes per pixel, one-pixel-wide to several-tiles-wide
    frames (NaN coordinates, x86 conversions, the samplers' general paths)."""
    import fuzz_uservals
    cases, failures = fuzz_uservals.run(seed, count, only, settings)
    assert cases == count and not failures, failures


WILD_COORDINATES = {
    "nan_x": "filter z (image in)\n  in(xy:[log(abs(x)-0.5), y])\nend\n",
    "nan_y": "filter z (image in)\n  in(xy:[x, sqrt(y)])\nend\n",
    # (the language has no exponent notation, scanner.c:286-336)
    "huge": "filter z (image in)\n  in(xy:[x*1000000000000000000000000000000.0*(y+0.3), y*3000000000.0])\nend\n",
    "around_2_31": "filter z (image in)\n  in(xy:[x*44000000.0, y*2200000000.0+x*100000.0])\nend\n",
    "around_2_32": "filter z (image in)\n  c = in(xy:[x*4294967296.0/(W/2), y*8589934592.0/(H/2)]);\n  rgba:[1-c[0], c[1], c[2]*0.5, c[3]]\nend\n",
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(WILD_COORDINATES))
def test_nan_and_huge_sample_coordinates_under_every_edge_mode(name):
    """A NaN or out-of-range coordinate becomes INT_MIN in the samplers (x86 conversion).  The reference then negates it in the
    reflect / rotate modes -- signed overflow: gcc -O2 drops the range check after it and the reference reads in front of the
    image.  Oracle and CUDA path let the negation wrap instead (oracle/runtime/images.c, mm_runtime.cuh); before that fix NVVM
    dropped the same check on the device.  Found by tools/fuzz_uservals.py --settings."""
    import numpy as np
    import mathmap_b200 as mb
    from conftest import synthetic_rgba
    from oracle.oracle import OracleFilter
    m = mb.Module(source=WILD_COORDINATES[name])
    oracle = OracleFilter(m.ir)
    for (w, h, iw, ih) in [(9, 13, 96, 96), (64, 48, 64, 48)]:
        img = synthetic_rgba(iw, ih, seed=5)
        for aa in (False, True):
            for edge in [(0, 0), (1, 1), (2, 2), (3, 3), (3, 0), (0, 3), (1, 3), (2, 1)]:
                inv = mb.Invocation(m, w, h, antialiasing=aa)
                inv.set_edge_behaviour(edge[0], edge[1], 0x11223344, 0x55667788)
                inv.set("in", img)
                got = inv.render(0, 0.0)
                want = oracle.render(w, h, {"in": img}, t=0.0, antialiasing=aa, edge_behaviour=edge, edge_colors=(0x11223344, 0x55667788))
                assert np.array_equal(got, want), (name, w, h, aa, edge)


@pytest.mark.gpu
@pytest.mark.parametrize("rel", ["Utilities/Visualize FFT.mm", "Combine/Convolve.mm", "Blur/Gaussian Blur.mm", "Utilities/Ident.mm", "Distorts/Twirl.mm"])
def test_one_pixel_wide_frames(rel):
    """W = 1 or H = 1: (size - 1) / 2 is 0 and every virtual coordinate is NaN or infinite (opmacros.h:156-157).  What comes out
    is decided by conversions: the drawable samplers' (int) gives INT_MIN, get_floatmap_pixel's (int)lrintf gives 0 on x86-64
    (builtins.c:257-258) -- texel (0, 0) instead of black."""
    import os
    import numpy as np
    import mathmap_b200 as mb
    from conftest import synthetic_rgba
    from oracle.oracle import OracleFilter
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filters", "examples")
    m = mb.Module.from_file(os.path.join(root, rel))
    oracle = OracleFilter(m.ir)
    for (w, h) in [(1, 1), (1, 7), (7, 1), (3, 1), (2, 2)]:
        for floatmap in (False, True):
            for aa in (False, True):
                inv = mb.Invocation(m, w, h, antialiasing=aa)
                vals, k = {}, 0
                for name, kind, _lo, _hi, _default in m.uservals():
                    if kind == mb.USERVAL_IMAGE:
                        vals[name] = synthetic_rgba(w, h, seed=5 + k)
                        k += 1
                    elif kind == mb.USERVAL_FLOAT:
                        vals[name] = 0.3
                for name, v in vals.items():
                    inv.set(name, v)
                got = inv.render(0, 0.3, floatmap=floatmap)
                want = oracle.render(w, h, vals, t=0.3, antialiasing=aa, floatmap=floatmap)
                if floatmap:
                    assert np.allclose(got, want, rtol=1e-5, atol=1e-6, equal_nan=True), (rel, w, h, aa)
                else:
                    assert np.array_equal(got, want), (rel, w, h, aa)


@pytest.mark.gpu
def test_random_slices_equal_the_whole_frame():
    """mmb_calc_lines_slice (the reference's calc_lines parameters: region, row range, row stride) with random values against
    the same pixels of the whole frame; bytes between rows and behind the last one stay untouched (tools/fuzz_slices.py)."""
    import fuzz_slices
    cases, failures = fuzz_slices.run(5, 150)
    assert not failures, failures
