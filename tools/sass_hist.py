"""Histogram of executed SASS opcodes (and stall samples) from an .ncu-rep source page.
Usage: python tools/sass_hist.py gpurun_out/prof.ncu-rep [top] [pixels]
With `pixels` (how many pixels the captured launch rendered) the counts are also given as thread instructions per pixel
(warp instructions x 32 / pixels: an upper bound, partial warps count as full)."""
import collections
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops = collections.Counter()
samples = collections.Counter()
total = 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    ins = r[isrc].strip()
    if ins.startswith("@"):
        ins = ins.split(None, 1)[1]
    op = ins.split()[0].rstrip(";")
    base = op.split(".")[0]
    n = int(r[iex] or 0)
    ops[op if base in ("MUFU", "F2I", "I2F", "F2F", "I2FP", "F2FP", "FRND", "DADD", "DMUL", "DFMA", "LDG", "STG", "LDGSTS", "LDS", "STS") else base] += n
    samples[base] += int(r[isamp] or 0)
    total += n
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pixels = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
print("total warp instructions executed:", total, " static SASS instructions:", len(rows) - 2)
if pixels:
    print("thread instructions per pixel (warp instructions x 32 / %d pixels): %.1f" % (pixels, total * 32.0 / pixels))
for op, n in ops.most_common(top):
    print("  %-28s %14d  %5.1f %%%s" % (op, n, 100.0 * n / total, ("  %7.2f /pixel" % (n * 32.0 / pixels)) if pixels else ""))
nsamp = sum(samples.values())
if nsamp:
    print("stall samples by opcode (where warps were when the sampler looked; %d samples):" % nsamp)
    for op, n in samples.most_common(12):
        print("  %-28s %14d  %5.1f %%" % (op, n, 100.0 * n / nsamp))
