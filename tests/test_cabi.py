"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares; render entry points
fail loudly (no CPU fallback).  CPU only."""
import ctypes
import os
import re

import pytest

import mathmap_b200 as mb
from conftest import ROOT, filter_source


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mathmap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "mathmap_b200", "libmathmap_b200.so"))
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/mathmap_b200.h but not exported: %s" % missing


def test_python_binding_covers_header():
    L = mb.lib()
    for n in declared_symbols():
        assert getattr(L, n).argtypes is not None, n


def test_compile_error_reporting():
    with pytest.raises(mb.MathMapError) as e:
        mb.Module(source="filter f (image in)\n  in(xy + q)\nend\n")
    assert "Undefined variable q" in str(e.value) and str(e.value).startswith("2:")
    with pytest.raises(mb.MathMapError):
        mb.Module(source="filter f () 1 end")  # result must be rgba:4
    with pytest.raises(mb.MathMapError):
        mb.Module(ir="(not-mmir)")


def test_userval_metadata():
    m = mb.Module(source=filter_source("examples/Render/Mandelbrot.mm"))
    uv = m.uservals()
    assert [u[0] for u in uv] == ["pj", "pk", "c1", "ci", "cj", "ck", "num_iterations"]
    assert uv[-1][1] == mb.USERVAL_INT and uv[-1][2:] == (2.0, 256.0, 32.0)
    assert uv[0][1] == mb.USERVAL_FLOAT and uv[0][2:] == (-2.0, 2.0, 0.0)
    assert m.name == "render_mandelbrot"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = mb.Module(source=filter_source("examples/Render/Mandelbrot.mm"))
    with pytest.raises(mb.MathMapError) as e:
        mb.Invocation(m, 16, 16)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_nvrtc_compiles_for_sm100a_without_gpu():
    for script in ["examples/Distorts/Twirl.mm", "examples/Map/Droste.mm", "examples/Map/IFS Functional.mm"]:
        m = mb.Module(source=filter_source(script))
        for precise in (False, True):
            assert m.compile_check(antialiasing=True, precise=precise) > 10000
    src = mb.Module(source=filter_source("examples/Distorts/Twirl.mm")).cuda_source
    assert "mm_kernel_twirl" in src and "mm_orig_val" in src and "__grid_constant__" in src


def test_byte_unit_byte_round_trip_is_identity():
    """The direct-output fast path (mm_orig_val_out) stores a sample's rounded bytes without converting them to floats
    and back.  That is exact because the reference's own chain -- (float)(k / 255.0) (opmacros.h:147-150), clamp,
    times 255.0 in double, truncate (new_template.c.in:281-292) -- returns k for every byte k."""
    import numpy as np
    k = np.arange(256)
    unit = (k / 255.0).astype(np.float32)
    clamped = np.maximum(np.float32(0), np.minimum(np.float32(1), unit))
    back = (clamped.astype(np.float64) * 255.0).astype(np.uint8)
    assert (back == k).all()


def test_bilinear_blend_of_equal_bytes_returns_the_byte():
    """The exterior fast path of the bilinear sampler (mm_bilinear_exterior) returns the edge colour without blending.
    In the reference the four texels are then the same byte c and the blend ((c*p1 + c*p2) + c*p3) + c*p4 in float
    arithmetic (builtins.c:226-240), rounded with rintf, must give c: checked over random sub-pixel positions."""
    import numpy as np
    rng = np.random.default_rng(7)
    f32 = np.float32
    px = (rng.random(200000) * 4096).astype(f32)
    py = (rng.random(200000) * 4096).astype(f32)
    x2f = px - np.floor(px).astype(f32)
    y2f = py - np.floor(py).astype(f32)
    x1f = (f32(1.0) - x2f).astype(f32)
    y1f = (f32(1.0) - y2f).astype(f32)
    p1, p2, p3, p4 = x1f * y1f, x1f * y2f, x2f * y1f, x2f * y2f
    for c in (0, 1, 127, 128, 254, 255):
        cf = f32(c)
        s = ((cf * p1 + cf * p2).astype(f32) + cf * p3).astype(f32) + cf * p4
        assert (np.rint(s.astype(f32)) == c).all()


def test_persistent_cubin_cache(tmp_path):
    """mmb_set_cubin_cache_dir: the first compile of a configuration writes DIR/<key>.cubin, later ones (a new module,
    as in a new process) load it instead of running NVRTC; another configuration gets another file."""
    import time
    src = open(os.path.join(ROOT, "tests", "golden", "filters", "examples", "Distorts", "Twirl.mm")).read()
    plain = mb.Module(source=src).compile_check(antialiasing=True, precise=True)
    mb.set_cubin_cache_dir(str(tmp_path))
    try:
        t0 = time.perf_counter()
        first = mb.Module(source=src).compile_check(antialiasing=True, precise=True)
        t_first = time.perf_counter() - t0
        files = sorted(os.listdir(tmp_path))
        assert len(files) == 1 and files[0].endswith(".cubin") and len(files[0]) == 32 + 6
        assert os.path.getsize(os.path.join(tmp_path, files[0])) == first + 24
        t0 = time.perf_counter()
        second = mb.Module(source=src).compile_check(antialiasing=True, precise=True)
        t_second = time.perf_counter() - t0
        assert first == second == plain
        assert t_second < t_first / 3, (t_first, t_second)
        mb.Module(source=src).compile_check(antialiasing=False, precise=True)
        assert len(os.listdir(tmp_path)) == 2
        # a damaged file is ignored and replaced
        path = os.path.join(tmp_path, files[0])
        open(path, "wb").write(b"garbage")
        assert mb.Module(source=src).compile_check(antialiasing=True, precise=True) == plain
        assert os.path.getsize(path) == plain + 24
    finally:
        mb.set_cubin_cache_dir(None)


def test_reference_side_backend_type_checks_against_the_reference_headers():
    """integration/backends/cuda.c (the file a maintainer adds to the reference: gen_and_load_cuda_code, the IR printer, the
    three mathfuncs_t functions) is compiled -fsyntax-only against the reference's own compiler-internals.h, mathmap.h,
    drawable.h and userval.h and against include/mathmap_b200.h.  Needs the reference tree: skipped where it is absent."""
    import subprocess
    if not os.path.isdir("/root/reference"):
        pytest.skip("the reference tree is not on this machine")
    r = subprocess.run(["sh", os.path.join(ROOT, "integration", "check.sh")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_null_invocation_is_an_error_not_a_crash():
    L = mb.lib()
    assert L.mmb_set_antialiasing(None, 1) != 0
    assert L.mmb_set_output_bpp(None, 4) != 0
    assert L.mmb_set_edge_behaviour(None, 0, 0, 0, 0) != 0
    assert L.mmb_set_render_size(None, 4, 4) != 0
    assert L.mmb_synchronize(None) != 0
    assert L.mmb_calc_lines_slice(None, None, 0, 1, None, 0) != 0
    assert b"NULL" in L.mmb_last_error() or b"bad arguments" in L.mmb_last_error()
