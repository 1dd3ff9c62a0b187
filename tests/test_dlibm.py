"""mathmap_b200/csrc/runtime/mm_dlibm.h evaluates sin, cos, asin, acos, atan, atan2, exp and log of float arguments in double with its own
polynomials (coefficients in constant memory) and narrows to float.  The reference computes RN_float(libm(x)) with the
host's double libm (ops.lisp:126-147); this compiles the header for the host and compares the two on a prime-stride
sample of all float bit patterns (the full sweep, stride 1: no mismatch for sin, cos, acos, atan, exp and log; asin differs by one
float ulp for 2 of 2.1e9 arguments, atan2 for 2 of 8.6e9 pairs)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_double_libm_restatement_matches_host_libm(tmp_path):
    exe = str(tmp_path / "dlibm_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "tools", "dlibm_check.cpp"), "-lm"])
    r = subprocess.run([exe, "1021"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    counts, rest = r.stdout.split("mismatches")
    mism_text, worst_text = rest.split("worst ulps")
    counts, mism = counts.split(), mism_text.split()
    totals = {counts[i]: int(counts[i + 1]) for i in range(1, len(counts), 2)}
    mismatches = {mism[i]: int(mism[i + 1]) for i in range(0, len(mism), 2)}
    assert set(mismatches) == {"sin", "cos", "acos", "asin", "exp", "log", "atan", "atan2"}
    assert all(v > 500000 for v in totals.values()), r.stdout
    # a few double ulps of error can flip the float rounding for about one argument in 10^8
    assert all(v <= 2 for v in mismatches.values()), r.stdout
    assert max(int(v) for v in worst_text.split()) <= 1, r.stdout
