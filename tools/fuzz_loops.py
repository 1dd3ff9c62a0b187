"""Developer tool (no GPU needed): random filters with `while` loops whose conditions share subexpressions with their bodies,
rendered by the oracle from the IR with and without the loop-carried value pass (MMB_LOOP_CARRY, csrc/ir/passes.cpp).  The two
renders must be the same bytes.  Prints one line per mismatch and a summary.
Usage: python tools/fuzz_loops.py SEED COUNT"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mathmap_b200 as mb  # noqa: E402
from oracle.oracle import OracleFilter  # noqa: E402

VARS = ["p", "q", "u"]


def atom(rng, scope):
    k = rng.random()
    if k < 0.55:
        return rng.choice(scope)
    if k < 0.7:
        return rng.choice(["x", "y", "t"])
    if k < 0.8:
        return rng.choice(["k1", "k2"])  # float uservals: frame constants
    return rng.choice(["0.3", "1.7", "2", "0.5", "-0.8"])


def expr(rng, scope, depth):
    if depth == 0 or rng.random() < 0.25:
        return atom(rng, scope)
    k = rng.random()
    a, b = expr(rng, scope, depth - 1), expr(rng, scope, depth - 1)
    if k < 0.45:
        return "(%s*%s)" % (a, b)
    if k < 0.7:
        return "(%s+%s)" % (a, b)
    if k < 0.85:
        return "(%s-%s)" % (a, b)
    if k < 0.92:
        return "abs(%s)" % a
    return "min(%s,%s)" % (a, b)


def random_filter(rng):
    nv = rng.randint(1, 3)
    scope = VARS[:nv]
    shared = [expr(rng, scope, 2) for _ in range(rng.randint(1, 3))]  # appear in the condition AND in the body
    lines = ["filter lf (float k1: -2-2 (0.4), float k2: -2-2 (-0.7))"]
    for v in scope:
        lines.append("  %s = %s;" % (v, rng.choice(["x", "y", "x*0.7", "y+0.2", "k1", "0", "x*y"])))
    lines.append("  n = 0;")
    cond = "+".join(rng.sample(shared, rng.randint(1, len(shared))))
    lines.append("  while %s < %s && n < %d do" % (cond, rng.choice(["2", "3.5", "1.2"]), rng.randint(2, 14)))
    body = []
    for v in rng.sample(scope, len(scope)):
        e = rng.choice(shared) if rng.random() < 0.7 else expr(rng, scope, 2)
        tail = rng.choice(["", " + %s" % atom(rng, scope), " * 0.9", " - %s" % rng.choice(shared)])
        body.append("    %s = %s%s;" % (v, e, tail))
    if rng.random() < 0.3:  # a nested loop reading the outer variables
        body.append("    m = 0; w = %s;" % scope[0])
        body.append("    while w*w < 1.5 && m < 3 do w = w*w + %s; m = m + 1 end;" % atom(rng, scope))
        body.append("    %s = %s + w * 0.1;" % (scope[0], scope[0]))
    body.append("    n = n + 1")
    lines += body
    lines.append("  end;")
    lines.append("  rgba:[%s * 0.2 + 0.5, %s * 0.1 + 0.5, n / 14, 1]" % (scope[0], scope[-1]))
    lines.append("end")
    return "\n".join(lines) + "\n"


def ir_of(src, carry):
    os.environ["MMB_LOOP_CARRY"] = "1" if carry else "0"
    try:
        return mb.Module(source=src).ir
    finally:
        del os.environ["MMB_LOOP_CARRY"]


def run(seed, count):
    """Returns (cases, cases the pass changed, failure lines)."""
    rng = random.Random(seed)
    done = changed = 0
    failures = []
    while done < count:
        src = random_filter(rng)
        try:
            plain, carried = ir_of(src, False), ir_of(src, True)
        except mb.MathMapError as e:
            failures.append("COMPILE %s: %r" % (str(e)[:120], src))
            done += 1
            continue
        done += 1
        if plain == carried:
            continue
        changed += 1
        uv = {"k1": rng.choice([0.4, 0.0, -1.3]), "k2": rng.choice([-0.7, 1.1])}
        t = rng.choice([0.0, 0.6])
        a = OracleFilter(plain).render(41, 29, dict(uv), t=t, antialiasing=False)
        b = OracleFilter(carried).render(41, 29, dict(uv), t=t, antialiasing=False)
        if not np.array_equal(a, b):
            failures.append("MISMATCH %d pixels, uv=%s t=%s: %r" % (int((a != b).any(axis=2).sum()), uv, t, src))
    return done, changed, failures


def main():
    t0 = time.time()
    done, changed, failures = run(int(sys.argv[1]), int(sys.argv[2]))
    for line in failures:
        print(line)
    print("filters %d, changed by the pass %d, failures %d, %.0f s" % (done, changed, len(failures), time.time() - t0))


if __name__ == "__main__":
    main()
