"""Renders every example filter under a directory on the GPU and with the oracle (default uservals, synthetic image inputs)
and reports the agreement.  Usage: python tests/tools/sweep_examples.py DIR [SIZE] > report.txt"""
import glob, os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mathmap_b200 as mb
from oracle.oracle import OracleFilter
from conftest import synthetic_rgba, compare_u8

d = sys.argv[1]
size = int(sys.argv[2]) if len(sys.argv) > 2 else 96
files = sorted(glob.glob(os.path.join(d, "**", "*.mm"), recursive=True))
img = synthetic_rgba(size, size)
bad = 0
for f in files:
    rel = os.path.relpath(f, d)
    t0 = time.time()
    try:
        m = mb.Module.from_file(f)
        inv = mb.Invocation(m, size, size, antialiasing=True)
        vals = {}
        for name, kind in [(u[0], u[1]) for u in m.uservals()]:
            if kind == mb.USERVAL_IMAGE:
                inv.set(name, img); vals[name] = img
        got = inv.render(0, 0.25)
        want = OracleFilter(m.ir).render(size, size, vals, t=0.25, antialiasing=True)
        exact, le1, mx = compare_u8(got, want)
        flag = "" if exact >= 99.9 else ("  <-- CHECK" if exact >= 99.0 else "  <-- MISMATCH")
        bad += exact < 99.9
        print("%-50s %8.4f %8.4f %4d  %.1fs%s" % (rel, exact, le1, mx, time.time() - t0, flag), flush=True)
    except Exception as e:
        bad += 1
        print("%-50s ERROR %s" % (rel, str(e).replace("\n", " ")[:160]), flush=True)
print("files", len(files), "not exact", bad)
