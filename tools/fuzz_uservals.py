"""Developer tool: example filters with random argument values inside their declared ranges, random sizes, t and sampler,
rendered on the GPU and compared with the oracle (needs a GPU).  Booleans flip the specialised kernel variants.
Usage: python tools/fuzz_uservals.py SEED COUNT [NAME_SUBSTRING] [--settings]"""
import glob
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
from oracle.oracle import OracleFilter  # noqa: E402

EX = os.path.join(ROOT, "tests", "golden", "filters", "examples")


def run(seed, count, only="", settings=False):
    """Returns (cases, failure lines).  settings: also randomise edge behaviour and colours, supersampling, output bytes per
    pixel, and draw sizes from a list with one-pixel-wide, tiny and several-tiles-wide frames."""
    failures = []
    rng = random.Random(seed)
    paths = [p for p in sorted(glob.glob(EX + "/*/*.mm")) if only in p]
    done = 0
    oracles = {}
    while done < count:
        p = rng.choice(paths)
        done += 1
        W, H = rng.choice([(64, 48), (77, 53), (50, 90), (129, 65)])
        aa = bool(rng.getrandbits(1))
        t = rng.choice([0.0, 0.2, 0.5, 0.9])
        ss, bpp, edge, edge_colors = False, 4, (0, 0), (0, 0)
        if settings:
            W, H = rng.choice([(1, 1), (1, 7), (7, 1), (2, 2), (3, 5), (33, 1), (5, 130), (517, 67), (1030, 19), (131, 260)])
            ss = rng.random() < 0.3
            bpp = rng.choice([4, 4, 3, 2, 1])
            edge = (rng.randint(0, 3), rng.randint(0, 3))
            edge_colors = (rng.getrandbits(32), rng.getrandbits(32))
        desc = ""
        try:
            m = mb.Module.from_file(p)
            inv = mb.Invocation(m, W, H, antialiasing=aa, supersampling=ss)
            if settings:
                inv.set_edge_behaviour(edge[0], edge[1], edge_colors[0], edge_colors[1])
                inv.set_output_bpp(bpp)
            vals, k = {}, 0
            for name, kind, lo, hi, _default in m.uservals():
                if kind == mb.USERVAL_IMAGE:
                    iw, ih = rng.choice([(W, H), (61, 47), (96, 96)])
                    vals[name] = synthetic_rgba(iw, ih, seed=5 + k)
                    k += 1
                elif kind == mb.USERVAL_FLOAT:
                    vals[name] = rng.choice([lo, hi, lo + (hi - lo) * rng.random(), lo + (hi - lo) * rng.random()])
                elif kind == mb.USERVAL_INT:
                    vals[name] = rng.randint(int(lo), int(hi))
                elif kind == mb.USERVAL_BOOL:
                    vals[name] = rng.getrandbits(1)
                elif kind == mb.USERVAL_COLOR:
                    vals[name] = tuple(rng.choice([0.0, 1.0, rng.random()]) for _ in range(4))
                elif kind == mb.USERVAL_CURVE:
                    xs = np.arange(1024, dtype=np.float32) / np.float32(1023)
                    vals[name] = (xs ** np.float32(rng.choice([0.5, 2.0, 3.0]))).astype(np.float32) if rng.getrandbits(1) else (np.float32(1) - xs)
                elif kind == mb.USERVAL_GRADIENT:
                    vals[name] = np.random.RandomState(rng.randint(0, 1 << 30)).randint(0, 1 << 32, size=1024, dtype=np.uint64).astype(np.uint32)
                else:
                    continue
                inv.set(name, vals[name])
            desc = "%s %dx%d aa=%s t=%s ss=%s bpp=%d edge=%r/%r %r" % (os.path.relpath(p, EX), W, H, aa, t, ss, bpp, edge, edge_colors,
                                                                      {k2: v for k2, v in vals.items() if not hasattr(v, "shape")})
            got = inv.render(0, t)
            if p not in oracles:
                oracles[p] = OracleFilter(m.ir)
            want = oracles[p].render(W, H, vals, t=t, antialiasing=aa, supersampling=ss, edge_behaviour=edge, edge_colors=edge_colors, bpp=bpp)
            d = np.abs(got.astype(np.int32) - want.astype(np.int32)).max(axis=2)
            exact = float((d == 0).mean()) * 100.0
            # frames of a few hundred pixels: a filter with discontinuities (Droste's levels) may put a handful of pixels on the
            # other side of one for a last-bit difference in a float libm function; a bug moves many
            if exact < 99.9 and int((d != 0).sum()) > 12:
                failures.append("MISMATCH %.3f %% exact, max %d: %s" % (exact, int(d.max()), desc))
        except Exception as e:  # noqa: BLE001
            failures.append("ERROR %s: %s" % ((str(e).splitlines() or [type(e).__name__])[0][:200], desc or p))
    return done, failures


def main():
    t0 = time.time()
    args = [a for a in sys.argv[1:] if a != "--settings"]
    done, failures = run(int(args[0]), int(args[1]), args[2] if len(args) > 2 else "", settings="--settings" in sys.argv)
    for line in failures:
        print(line)
    print("cases %d, failures %d, %.0f s" % (done, len(failures), time.time() - t0))


if __name__ == "__main__":
    main()
