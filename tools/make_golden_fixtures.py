"""Copies the reference's own golden vectors for the render path into tests/golden/.

The reference's test suite (tests/run_tests.sh) renders a list of example filters
with `mathmap -i -f SCRIPT ARGS OUT.png` at 256x256 and compares against golden
PNGs.  /root/reference does not exist on the GPU box, so the golden PNGs, the
input image and the filter scripts those command lines name (inputs of the
golden vectors, in the MathMap language) are copied here as fixtures, together
with a manifest of the command lines.  Nothing else is taken from the reference.

Run in the build container:  python tools/make_golden_fixtures.py
"""
import json
import os
import re
import shlex
import shutil

REF = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

os.makedirs(os.path.join(OUT, "png"), exist_ok=True)
os.makedirs(os.path.join(OUT, "filters"), exist_ok=True)
manifest = []
for line in open(os.path.join(REF, "run_tests.sh")):
    m = re.match(r"run_(render|modify)_test\s+(\S.*)", line.strip())
    if not m or m.group(2).startswith("()"):
        continue
    parts = shlex.split(m.group(2))
    script, golden = parts[0], parts[1]
    args = parts[2] if len(parts) > 2 else ""
    src = os.path.normpath(os.path.join(REF, script))
    rel = os.path.relpath(src, "/root/reference")  # e.g. examples/Distorts/Twirl.mm or tests/Apply.mm
    dst = os.path.join(OUT, "filters", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    shutil.copyfile(os.path.join(REF, golden), os.path.join(OUT, "png", golden))
    uservals = {}
    for a in args.split():
        k, v = a[2:].split("=")
        uservals[k] = float(v) if "." in v else int(v)
    manifest.append({"kind": m.group(1), "script": rel, "golden": golden, "uservals": uservals,
                     "cmdline": "mathmap -i -f %s %s%s OUT.png" % (script, "-s 256x256 " if m.group(1) == "render" else "-Din=marlene.png ", args)})
# every example filter: the reference's test list has golden pictures for 80 of them, the parity tests render the
# rest too (CUDA path against the oracle)
import glob
for src in sorted(glob.glob("/root/reference/examples/**/*.mm", recursive=True) + glob.glob("/root/reference/examples/**/*.mmc", recursive=True)):
    rel = os.path.relpath(src, "/root/reference")
    dst = os.path.join(OUT, "filters", rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
shutil.copyfile(os.path.join(REF, "marlene.png"), os.path.join(OUT, "png", "marlene.png"))
json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
print(len(manifest), "golden vectors")
