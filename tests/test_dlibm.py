"""mathmap_b200/csrc/runtime/mm_dlibm.h evaluates sin, cos, asin and acos of float arguments in double with its own
polynomials (coefficients in constant memory) and narrows to float.  The reference computes RN_float(libm(x)) with the
host's double libm (ops.lisp:126-147); this compiles the header for the host and compares the two on a prime-stride
sample of all float bit patterns (the full sweep, stride 1, has no mismatch either: 4.3e9 arguments)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_double_libm_restatement_matches_host_libm(tmp_path):
    exe = str(tmp_path / "dlibm_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "tools", "dlibm_check.cpp"), "-lm"])
    r = subprocess.run([exe, "1021"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    fields = r.stdout.split()
    mism = {name: int(fields[fields.index(name) + 1]) for name in ("sin", "cos", "acos", "asin")}
    total = int(fields[fields.index("trig") + 1])
    assert total > 500000
    # a few double ulps of error can flip the float rounding for about one argument in 10^7
    assert all(v <= 2 for v in mism.values()), r.stdout
    worst = [int(v) for v in fields[fields.index("ulps") + 1:]]
    assert max(worst) <= 1, r.stdout
