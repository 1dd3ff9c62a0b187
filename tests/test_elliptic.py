"""Elliptic integrals / Jacobi elliptic functions (reference opmacros.h:101-125, evaluated there by GSL, which is not
in the reference tree).  CPU-only: the oracle's independent evaluation is checked against scipy.special, and the product's
restatement of GSL's algorithms is exercised through constant folding (host) and the NVRTC build check (device)."""
import ctypes
import re

import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import filter_source
from oracle.oracle import OracleFilter

sp = pytest.importorskip("scipy.special")


@pytest.fixture(scope="module")
def olib():
    lib = OracleFilter(mb.Module(source="filter f () rgba:[ell_int_Kcomp(x), 0, 0, 1] end").ir).lib
    D = ctypes.c_double
    for name, n in [("mmo_ellint_Kcomp", 1), ("mmo_ellint_Ecomp", 1), ("mmo_ellint_F", 2), ("mmo_ellint_E", 2), ("mmo_ellint_RC", 2),
                    ("mmo_ellint_RD", 3), ("mmo_ellint_RF", 3), ("mmo_ellint_RJ", 4)]:
        getattr(lib, name).restype = D
        getattr(lib, name).argtypes = [D] * n
    lib.mmo_elljac.argtypes = [D, D] + [ctypes.POINTER(D)] * 3
    return lib


def test_oracle_elliptic_against_scipy(olib):
    rng = np.random.default_rng(7)
    for _ in range(500):
        k, phi = rng.uniform(-0.999, 0.999), rng.uniform(-7, 7)
        x, y, z, p = rng.uniform(0.01, 20, 4)
        assert olib.mmo_ellint_Kcomp(k) == pytest.approx(sp.ellipk(k * k), rel=1e-9)
        assert olib.mmo_ellint_Ecomp(k) == pytest.approx(sp.ellipe(k * k), rel=1e-9)
        assert olib.mmo_ellint_F(phi, k) == pytest.approx(sp.ellipkinc(phi, k * k), rel=1e-8, abs=1e-9)
        assert olib.mmo_ellint_E(phi, k) == pytest.approx(sp.ellipeinc(phi, k * k), rel=1e-8, abs=1e-9)
        assert olib.mmo_ellint_RC(x, y) == pytest.approx(sp.elliprc(x, y), rel=1e-8)
        assert olib.mmo_ellint_RD(x, y, z) == pytest.approx(sp.elliprd(x, y, z), rel=1e-8)
        assert olib.mmo_ellint_RF(x, y, z) == pytest.approx(sp.elliprf(x, y, z), rel=1e-8)
        assert olib.mmo_ellint_RJ(x, y, z, p) == pytest.approx(sp.elliprj(x, y, z, p), rel=1e-8)
        u, m = rng.uniform(-10, 10), rng.uniform(0.001, 0.999)
        s, c, d = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        olib.mmo_elljac(u, m, s, c, d)
        sn, cn, dn, _ = sp.ellipj(u, m)
        assert (s.value, c.value, d.value) == pytest.approx((sn, cn, dn), abs=1e-12)


@pytest.mark.parametrize("expr,want", [("ell_int_Kcomp(0.5)", lambda: sp.ellipk(0.25)), ("ell_int_Ecomp(0.5)", lambda: sp.ellipe(0.25)),
                                        ("ell_int_F(1.0, 0.5)", lambda: sp.ellipkinc(1.0, 0.25)), ("ell_int_E(4.0, 0.5)", lambda: sp.ellipeinc(4.0, 0.25)),
                                        ("ell_int_RC(1.0, 2.0)", lambda: sp.elliprc(1.0, 2.0)), ("ell_int_RD(1.0, 2.0, 3.0)", lambda: sp.elliprd(1.0, 2.0, 3.0)),
                                        ("ell_int_RF(1.0, 2.0, 3.0)", lambda: sp.elliprf(1.0, 2.0, 3.0)),
                                        ("ell_int_RJ(1.0, 2.0, 3.0, 4.0)", lambda: sp.elliprj(1.0, 2.0, 3.0, 4.0))])
def test_elliptic_constant_folding(expr, want):
    """Constant arguments are folded by the host evaluator (ops.lisp:216-227 marks the integrals foldable)."""
    ir = mb.Module(source="filter f () grayColor(%s) end" % expr).ir
    assert "ELL_" not in ir
    consts = [float(v) for v in re.findall(r"f:(-?[0-9.]+(?:e-?[0-9]+)?)", ir)]
    assert any(abs(c - want()) <= 2e-7 * abs(want()) for c in consts), (consts, want())


def test_quincuncial_builds_for_sm100a():
    """The one reference example that calls ell_jac per pixel; NVRTC cross-compiles without a GPU."""
    m = mb.Module(source=filter_source("examples/Map/Quincuncial.mm"))
    assert "ELL_JAC" in m.ir
    assert m.compile_check(antialiasing=True) > 0
