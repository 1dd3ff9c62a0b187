"""Debug aid: renders a filter truncated after each top-level statement on the GPU and with the oracle and reports where they
first diverge.  Usage: python tests/tools/debug_stages.py FILTER.mm VAR [WIDTH HEIGHT]  (VAR: a 2-tuple variable to visualise)"""
import os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mathmap_b200 as mb
from oracle.oracle import OracleFilter
from conftest import synthetic_rgba, compare_u8

src = open(sys.argv[1]).read()
var = sys.argv[2]
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (96, 96)
extra = dict(kv.split("=") for kv in sys.argv[5:])
head_end = src.index(")", src.index("filter ")) + 1
while src.count("(", 0, head_end) != src.count(")", 0, head_end):
    head_end = src.index(")", head_end) + 1
head, body = src[:head_end], src[head_end:]
lines = body.split("\n")
img = synthetic_rgba(256, 128)
depth = 0
for i, line in enumerate(lines):
    code = line.split("#")[0]
    depth += len(re.findall(r"\b(then|do)\b", code)) - len(re.findall(r"\bend\b", code))
    if depth != 0 or not code.strip().endswith(";"):
        continue
    a, b = var.split(",") if "," in var else (var + "[0]", var + "[1]")
    if any(re.search(r"\b%s\b" % re.escape(v.split("[")[0]), "\n".join(lines[:i + 1])) is None for v in (a, b)):
        continue
    text = head + "\n".join(lines[:i + 1]) + "\nrgba:[(%s)/8+0.5, (%s)/8+0.5, 0, 1]\nend\n" % (a, b)
    try:
        m = mb.Module(source=text)
    except Exception as e:
        print(i, "compile error", str(e)[:100]); continue
    inv = mb.Invocation(m, W, H, antialiasing=True)
    vals = {"in": img}
    inv.set("in", img)
    for k, v in extra.items():
        v = float(v) if "." in v else int(v)
        inv.set(k, v); vals[k] = v
    got = inv.render(0, 0.0)
    want = OracleFilter(m.ir).render(W, H, vals, antialiasing=True)
    print(i, repr(line.strip()[:60]), compare_u8(got, want))
