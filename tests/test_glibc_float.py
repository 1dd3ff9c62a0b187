"""The float routines inside glibc's float complex functions (sinhf, coshf via expm1f, atan2f via atanf, log1pf) are
restated in mathmap_b200/csrc/runtime/mm_glibc_float.h for the device.  This compiles that header for the host and
compares it with the host's libm bit for bit on a prime-stride sample of all float bit patterns (the full sweep,
stride 1, takes a few minutes and also has no mismatch)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restated_float_routines_match_host_libm(tmp_path):
    exe = str(tmp_path / "glibc_float_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "tools", "glibc_float_check.cpp"), "-lm"])
    r = subprocess.run([exe, "4099"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert "mismatches atanf 0 expm1f 0 sinhf 0 coshf 0 log1pf 0 atan2f 0" in r.stdout
