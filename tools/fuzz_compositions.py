"""Developer tool: random two- and three-node compositions of the example filters, rendered on the GPU and compared with the
oracle (needs a GPU).  Prints one line per mismatch or error and a summary.
Usage: python tools/fuzz_compositions.py SEED COUNT"""
import collections
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


def catalogue():
    filters = []
    for p in sorted(glob.glob(EX + "/*/*.mm")):
        m = mb.Module.from_file(p)
        imgs = [u[0] for u in m.uservals() if u[1] == mb.USERVAL_IMAGE]
        filters.append((m.name, imgs))
    cnt = collections.Counter(f[0] for f in filters)
    return [f for f in filters if cnt[f[0]] == 1]


def random_design(rng, filters, with_img):
    a, b = rng.choice(filters), rng.choice(with_img)
    nodes = '(node :name "a" :type "%s" :input-slots ()) (node :name "b" :type "%s" :input-slots (("%s" "a" "out")))' % (a[0], b[0], rng.choice(b[1]))
    names, root = {a[0], b[0]}, "b"
    if rng.random() < 0.4:
        c = rng.choice(with_img)
        slots = '("%s" "b" "out")' % rng.choice(c[1])
        nodes += ' (node :name "c" :type "%s" :input-slots (%s))' % (c[0], slots)
        names.add(c[0])
        root = "c"
        if len(names) < 3:
            return None
    elif len(names) < 2:
        return None
    return '(design %s :name "comp" :root "%s")' % (nodes, root)


def run(seed, count):
    """Returns (cases, failure lines)."""
    failures = []
    rng = random.Random(seed)
    filters = catalogue()
    with_img = [f for f in filters if f[1]]
    W, H = 80, 56
    imgs = [synthetic_rgba(W, H, seed=s) for s in (1, 2, 3)]
    done = 0
    while done < count:
        design = random_design(rng, filters, with_img)
        if design is None:
            continue
        done += 1
        aa = bool(rng.getrandbits(1))
        t = rng.choice([0.0, 0.3, 0.75])
        try:
            m = mb.Module(source=mb.design_to_source(design, EX))
            inv = mb.Invocation(m, W, H, antialiasing=aa)
            vals, k = {}, 0
            for name, kind, _lo, _hi, _default in m.uservals():
                if kind == mb.USERVAL_IMAGE:
                    vals[name] = imgs[k % 3]
                    inv.set(name, vals[name])
                    k += 1
            got = inv.render(0, t)
            want = OracleFilter(m.ir).render(W, H, vals, t=t, antialiasing=aa)
            d = np.abs(got.astype(np.int32) - want.astype(np.int32)).max(axis=2)
            exact = float((d == 0).mean()) * 100.0
            # frames of a few hundred pixels: a filter with discontinuities (Droste's levels) may put a handful of pixels on the
            # other side of one for a last-bit difference in a float libm function; a bug moves many
            if exact < 99.9 and int((d != 0).sum()) > 12:
                failures.append("MISMATCH %.3f %% exact, max %d, aa=%s t=%s: %s" % (exact, int(d.max()), aa, t, design))
        except Exception as e:  # noqa: BLE001
            msg = str(e).splitlines()[0][:200] if str(e) else type(e).__name__
            if "defined more than once" in msg:
                continue
            failures.append("ERROR %s: %s" % (msg, design))
    return done, failures


def main():
    t0 = time.time()
    done, failures = run(int(sys.argv[1]), int(sys.argv[2]))
    for line in failures:
        print(line)
    print("compositions %d, failures %d, %.0f s" % (done, len(failures), time.time() - t0))


if __name__ == "__main__":
    main()
