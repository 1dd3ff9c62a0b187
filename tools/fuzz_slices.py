"""Developer tool: mmb_calc_lines_slice with random regions, row ranges and row strides against the same pixels of the whole
frame rendered by mmb_calc_lines (needs a GPU; no oracle involved -- the whole frame is what the parity tests pin).
Usage: python tools/fuzz_slices.py SEED COUNT"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mathmap_b200 as mb  # noqa: E402
from conftest import synthetic_rgba  # noqa: E402

EX = os.path.join(ROOT, "tests", "golden", "filters", "examples")
FILTERS = ["Utilities/Ident.mm", "Colors/Invert.mm", "Distorts/Sea.mm", "Distorts/Twirl.mm", "Render/Mandelbrot.mm", "Blur/Gaussian Blur.mm",
           "Geometry/Zoom.mm", "Map/Droste.mm", "Render/Perlin Noise.mm"]


def run(seed, count):
    """Returns (cases, failure lines)."""
    rng = random.Random(seed)
    failures = []
    cache = {}
    for _ in range(count):
        rel = rng.choice(FILTERS)
        W, H = rng.choice([(64, 48), (77, 53), (131, 40), (260, 33)])
        aa = bool(rng.getrandbits(1))
        bpp = rng.choice([4, 4, 3, 1])
        floatmap = rng.random() < 0.25
        key = (rel, W, H, aa, bpp)
        if key not in cache:
            m = mb.Module.from_file(os.path.join(EX, rel))
            inv = mb.Invocation(m, W, H, antialiasing=aa)
            for name, kind, _lo, _hi, _default in m.uservals():
                if kind == mb.USERVAL_IMAGE:
                    inv.set(name, synthetic_rgba(W, H, seed=3))
                elif kind == mb.USERVAL_FLOAT and name == "dev":
                    inv.set(name, 0.03)
            inv.set_output_bpp(bpp)
            inv.init_frame(0, 0.4)
            cache[key] = (inv, inv.calc_lines(0, H), inv.calc_lines(0, H, floatmap=True))
        inv, whole, whole_f = cache[key]
        rw, rh = rng.randint(1, W), rng.randint(1, H)
        rx, ry = rng.randint(0, W - rw), rng.randint(0, H - rh)
        first = rng.randint(max(0, ry - 3), ry + rh - 1)
        last = rng.randint(first + 1, min(H, ry + rh + 3))
        fr, lr = max(first, 0), min(last, ry + rh)  # like the reference: rows above the region are rendered (new_template.c.in:238-243)
        px = 16 if floatmap else bpp
        stride = W * 16 if floatmap else rw * bpp + rng.choice([0, 0, 1, 5, 16])
        desc = "%s %dx%d aa=%s bpp=%d floatmap=%s region=(%d,%d,%d,%d) rows=[%d,%d) stride=%d" % (rel, W, H, aa, bpp, floatmap, rx, ry, rw, rh, first, last, stride)
        try:
            nbytes = max(0, lr - fr - 1) * stride + rw * px if lr > fr else 0
            buf = np.full(nbytes + 32, 0xA5, dtype=np.uint8)
            out = buf[:nbytes] if nbytes else buf[:0]
            if nbytes == 0:
                continue
            inv.calc_lines_slice(first, last, out, region=(rx, ry, rw, rh), row_stride=None if floatmap else stride, floatmap=floatmap)
            if not np.all(buf[nbytes:] == 0xA5):
                failures.append("WROTE PAST THE BUFFER: " + desc)
                continue
            ref = (whole_f if floatmap else whole)[fr:lr, rx:rx + rw]
            ok = True
            for i in range(lr - fr):
                row = out[i * stride:i * stride + rw * px]
                want = ref[i].view(np.uint8).reshape(-1)
                if not np.array_equal(row, want):
                    ok = False
                    break
                gap = out[i * stride + rw * px:(i + 1) * stride] if i + 1 < lr - fr and not floatmap else out[:0]
                if gap.size and not np.all(gap == 0xA5):
                    ok = False
                    break
            if not ok:
                failures.append("MISMATCH at slice row %d: %s" % (i, desc))
        except Exception as e:  # noqa: BLE001
            failures.append("ERROR %s: %s" % ((str(e).splitlines() or [type(e).__name__])[0][:200], desc))
    return count, failures


def main():
    t0 = time.time()
    done, failures = run(int(sys.argv[1]), int(sys.argv[2]))
    for line in failures:
        print(line)
    print("slices %d, failures %d, %.0f s" % (done, len(failures), time.time() - t0))


if __name__ == "__main__":
    main()
