"""IR ops and builtins that no example filter of the reference uses (found by listing the ops in the IR of all 189 examples
against the op table, ir/ir.cpp): hyperbolic functions, ceil, gamma, beta, the complex elementary functions, 3x3 linear
solves, matrix and vector builtins, colour conversions.  Device against the oracle (glibc / the oracle's own restatements);
both evaluate in double where the reference does and narrow to float, so a rare last-bit difference may move a channel by one."""
import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import compare_u8, synthetic_rgba
from oracle.oracle import OracleFilter

OP_FILTERS = {
    "hyperbolic": """filter f ()
        la = x * 3; lb = y * 2;
        lp = rgba:[sinh(la) / 20 + 0.5, cosh(lb) / 8, tanh(la * lb) / 2 + 0.5, asinh(la * 5) / 6 + 0.5];
        lq = rgba:[acosh(abs(la) + 1) / 4, atanh(x * 0.99) / 6 + 0.5, ceil(la * 3) / 20 + 0.5, floor(lb * 3) / 14 + 0.5];
        if y > 0 then lp else lq end
    end""",
    "hyperbolic_out_of_domain": """filter f ()
        rgba:[acosh(x * 2) / 3, atanh(x * 2) / 4 + 0.5, asinh(x * 100000000 * y) / 40 + 0.5, cosh(y * 100) / 1000000000]
    end""",
    "gamma_beta": """filter f ()
        rgba:[gamma(x * 4 + 4.5) / 120, gamma(x * 3) / 20 + 0.5, beta(x + 1.5, y + 1.5), beta(x * 3, y * 3) / 20 + 0.5]
    end""",
    "complex_sqrt_exp_log": """filter f ()
        lz = ri:[x * 3, y * 3];
        la = sqrt(lz); lb = exp(lz); lc = log(lz);
        rgba:[la[0] / 4 + 0.5, la[1] / 4 + 0.5, lb[0] / 40 + 0.5, lc[1] / 7 + 0.5] * 0.5 + rgba:[lb[1] / 40 + 0.5, lc[0] / 4 + 0.5, arg(lz) / 7 + 0.5, abs(lz) / 5] * 0.5
    end""",
    "complex_trig": """filter f ()
        lz = ri:[x * 2.5, y * 2];
        la = sin(lz); lb = cos(lz); lc = tan(lz);
        rgba:[la[0] / 8 + 0.5, la[1] / 8 + 0.5, lb[0] / 8 + 0.5, lb[1] / 8 + 0.5] * 0.5 + rgba:[lc[0] / 4 + 0.5, lc[1] / 4 + 0.5, 0.5, 1] * 0.5
    end""",
    "complex_inverse_trig": """filter f ()
        lz = ri:[x * 2.5, y * 2];
        la = asin(lz); lb = acos(lz); lc = atan(lz);
        rgba:[la[0] / 4 + 0.5, la[1] / 4 + 0.5, lb[0] / 4, lb[1] / 4 + 0.5] * 0.5 + rgba:[lc[0] / 4 + 0.5, lc[1] / 4 + 0.5, 0.5, 1] * 0.5
    end""",
    "complex_hyperbolic": """filter f ()
        lz = ri:[x * 2, y * 2.5];
        la = sinh(lz); lb = cosh(lz); lc = tanh(lz);
        rgba:[la[0] / 8 + 0.5, la[1] / 8 + 0.5, lb[0] / 8 + 0.5, lb[1] / 8 + 0.5] * 0.5 + rgba:[lc[0] / 4 + 0.5, lc[1] / 4 + 0.5, 0.5, 1] * 0.5
    end""",
    "complex_inverse_hyperbolic": """filter f ()
        lz = ri:[x * 2, y * 2.5];
        la = asinh(lz); lb = acosh(lz); lc = atanh(lz);
        rgba:[la[0] / 4 + 0.5, la[1] / 4 + 0.5, lb[0] / 4, lb[1] / 7 + 0.5] * 0.5 + rgba:[lc[0] / 4 + 0.5, lc[1] / 4 + 0.5, 0.5, 1] * 0.5
    end""",
    "complex_gamma_pow": """filter f ()
        lz = ri:[x * 3 + 0.3, y * 3];
        lg = gamma(lz); lp = lz ^ ri:[1.5, 0.25]; lq = lz ^ 3;
        rgba:[lg[0] / 6 + 0.5, lg[1] / 6 + 0.5, lp[0] / 12 + 0.5, lp[1] / 12 + 0.5] * 0.5 + rgba:[lq[0] / 60 + 0.5, lq[1] / 60 + 0.5, 0.5, 1] * 0.5
    end""",
    "linear_algebra": """filter f ()
        lm = m3x3:[2 + x, 0.3, y, 0.1, 1.5 - y, 0.2, x * y, 0.4, 1 + x * x];
        lv = v3:[x, y, 1] / lm;
        ln = m2x2:[1 + x, y, 0.5 - y, 2] * m2x2:[x, 1, 1, y];
        lw = v2:[x, y] / m2x2:[1.2 + x, y, 0.3, 1 + y * y];
        lc = crossp(v3:[x, y, 0.5], v3:[0.2, x * y, 1]);
        rgba:[lv[0] / 2 + 0.5, lv[1] / 2 + 0.5, lv[2] / 2 + 0.5, det(lm) / 8 + 0.5] * 0.5 +
            rgba:[det(ln) / 8 + 0.5, lw[0] / 2 + 0.5, lc[2] / 2 + 0.5, dotp(normalize(v3:[x, y, 0.3]), v3:[0.5, 0.5, 0.7]) / 2 + 0.5] * 0.5
    end""",
    "singular_solves": """filter f ()
        lv = v3:[x, y, 1] / m3x3:[1, 2, 3, 2, 4, 6, x, y, 1];
        lw = v2:[x, y] / m2x2:[x, x, y, y];
        rgba:[lv[0] + 0.5 + x * 0.2, lv[1] + 0.5 + y * 0.2, lw[0] + 0.5, lw[1] + 0.5]
    end""",
    "colour_conversions": """filter f (image in)
        lp = in(xy);
        lh = toHSVA(lp);
        lq = toRGBA(hsva:[lh[0] + t * 0.3, lh[1] * 0.9, lh[2], lh[3]]);
        lg = gray(lp);
        rgba:[lq[0], lerp(0.3, lq[1], lg), clamp(lq[2] * 1.5 - 0.2, 0.1, 0.8), inintv(lg, 0.2, 0.7) * 0.5 + 0.25]
    end""",
    "scalar_helpers": """filter f ()
        la = pmod(x * 7, 1.3); lb = sign(x * y); lc = scale(x, -1, 1, 0, 0.8); ld = rad2deg(deg2rad(x * 90)) / 180 + 0.5;
        le = x * 5 % 1.7; li = floor(x * 4) % 3;
        rgba:[la / 1.3, lb / 2 + 0.5, lc + li / 30, ld] * 0.5 + rgba:[le / 4 + 0.5, max(x, y) / 2 + 0.5, min(min(x, y), 0.3) / 2 + 0.5, 1] * 0.5
    end""",
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(OP_FILTERS))
def test_ops_no_example_uses_match_oracle(name):
    m = mb.Module(source=OP_FILTERS[name])
    vals = {}
    inv = mb.Invocation(m, 200, 150, antialiasing=True)
    if "image in" in OP_FILTERS[name]:
        vals["in"] = synthetic_rgba(200, 150)
        inv.set("in", vals["in"])
    got = inv.render(0, 0.4)
    want = OracleFilter(m.ir).render(200, 150, vals, t=0.4, antialiasing=True)
    exact, le1, mx = compare_u8(got, want)
    assert exact >= 99.5 and le1 >= 99.9, "%s: %.4f %% exact, %.4f %% within 1 LSB, max %d" % (name, exact, le1, mx)
    assert got[..., :3].std() > 1.0, "flat picture"
