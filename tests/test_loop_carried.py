"""The loop-carried value pass (csrc/ir/passes.cpp: loop_carried_values): a rotated `while` computes the parts of its
condition at the end of iteration k-1 and again, through the loop's phis, at the top of iteration k; the pass carries them in
a phi instead.  Pinned here without the device: the oracle renders the IR with the pass and the IR without it
(MMB_LOOP_CARRY=0) to the same bytes, with arguments away from their defaults; what the device runs is the same IR."""
import os

import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import filter_source
from oracle.oracle import OracleFilter


def ir_of(source, carry):
    old = os.environ.get("MMB_LOOP_CARRY")
    os.environ["MMB_LOOP_CARRY"] = "1" if carry else "0"
    try:
        return mb.Module(source=source).ir
    finally:
        if old is None:
            del os.environ["MMB_LOOP_CARRY"]
        else:
            os.environ["MMB_LOOP_CARRY"] = old


def loop_of(ir):
    i = ir.index("(while")
    return ir[i:]


def test_mandelbrot_carries_its_four_squares():
    src = filter_source("examples/Render/Mandelbrot.mm")
    plain, carried = loop_of(ir_of(src, False)), loop_of(ir_of(src, True))
    # quaternion square: 16 products + 4 squares for |c| per iteration; the 4 squares of the product are last iteration's
    assert plain.count("(op MUL") == 20 and carried.count("(op MUL") == 16
    assert carried.count("(phi ") == plain.count("(phi ") + 4
    assert plain.count("(op ADD") == carried.count("(op ADD")


CASES = [
    ("examples/Render/Mandelbrot.mm", {}),
    ("examples/Render/Mandelbrot.mm", {"num_iterations": 256}),
    ("examples/Render/Mandelbrot.mm", {"pj": 0.3, "pk": -0.2, "c1": 0.1, "ci": -0.05, "cj": 0.2, "ck": 0.15, "num_iterations": 97}),
    ("examples/Render/Mandelbrot.mm", {"c1": 2.0, "num_iterations": 5}),  # |c| >= 2 on entry: the loop never runs
    ("examples/Render/Mandelbrot.mm", {"num_iterations": 2}),
    ("examples/Render/Fancy Mandelbrot.mm", {}),
]


@pytest.mark.parametrize("rel,uv", CASES)
def test_oracle_renders_the_same_bytes_with_and_without_the_pass(rel, uv):
    src = filter_source(rel)
    a, b = OracleFilter(ir_of(src, False)), OracleFilter(ir_of(src, True))
    for (w, h) in ((301, 203), (64, 1)):
        for aa in (False, True):
            want = a.render(w, h, dict(uv), t=0.4, antialiasing=aa)
            got = b.render(w, h, dict(uv), t=0.4, antialiasing=aa)
            assert np.array_equal(got, want), (rel, uv, w, h, aa)


SHAPES = {
    # complex escape loop: the squares of |z|^2 and of z*z coincide
    "julia": ("filter j (float cr: -2-2 (-0.8), float ci: -2-2 (0.156))\n  zr = x * 1.5; zi = y * 1.5; n = 0;\n"
              "  while zr*zr + zi*zi < 4 && n < 40 do\n    nr = zr*zr - zi*zi + cr;\n    zi = 2*zr*zi + ci;\n    zr = nr;\n    n = n + 1\n  end;\n"
              "  grayColor(n / 40)\nend\n"),
    # the carried expression mixes a phi with a loop invariant and a constant
    "invariant": ("filter k (float s: 0-2 (1.25))\n  v = x; n = 0;\n  while v * s + 0.5 < 3 && n < 30 do\n    v = (v * s + 0.5) * 1.1 + y;\n    n = n + 1\n  end;\n"
                  "  grayColor(n / 30)\nend\n"),
    # a loop inside a loop, both with carried parts
    "nested": ("filter m ()\n  p = x; i = 0; acc = 0;\n  while p*p < 2 && i < 6 do\n    b = y; j = 0;\n    while b*b + p*p < 3 && j < 5 do\n      b = b*b + p*p - 0.3;\n      j = j + 1\n    end;\n"
               "    acc = acc + j;\n    p = p*p + 0.2;\n    i = i + 1\n  end;\n  grayColor(acc / 30)\nend\n"),
    # the carried product is also a variable that is read behind the loop (its own phi keeps the last executed iteration's)
    "read_after": ("filter ra ()\n  v = x * 1.3; q = 0; n = 0;\n  while v*v < 2.5 && n < 12 do\n    q = v*v;\n    v = q + y * 0.5 + 0.1;\n    n = n + 1\n  end;\n"
                   "  rgba:[q * 0.3, v * 0.2, n / 12, 1]\nend\n"),
    # a swap in the loop: phis whose back-edge operands are other phis' values
    "swap": ("filter sw ()\n  p = x; q = y; n = 0;\n  while p*p + q*q < 3 && n < 9 do\n    s = p*p - q*q + 0.3;\n    p = q;\n    q = s;\n    n = n + 1\n  end;\n"
             "  rgba:[p * 0.3 + 0.5, q * 0.3 + 0.5, n / 9, 1]\nend\n"),
    # a frame-constant loop (evaluated by the host replay / once per frame)
    "frame_constant": ("filter c (float k: 0-2 (0.7))\n  v = k; n = 0;\n  while v*v < 50 && n < 20 do\n    v = v*v + k;\n    n = n + 1\n  end;\n  grayColor(n / 20 + x * 0.1)\nend\n"),
}


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_loop_shapes_with_and_without_the_pass(name):
    src = SHAPES[name]
    plain, carried = ir_of(src, False), ir_of(src, True)
    # fewer products inside the (outermost) loop: the carried ones are phis now
    assert loop_of(carried).count("(op MUL") < loop_of(plain).count("(op MUL"), "the pass found nothing to carry in " + name
    a, b = OracleFilter(plain), OracleFilter(carried)
    for (w, h) in ((157, 90), (33, 47)):
        assert np.array_equal(a.render(w, h, {}, t=0.25), b.render(w, h, {}, t=0.25)), name
    assert mb.Module(source=src).compile_check(antialiasing=False, precise=True) >= 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["frame_constant", "invariant", "julia", "nested"])
def test_loop_shapes_on_the_device(name):
    src = SHAPES[name]
    m = mb.Module(source=src)
    want = OracleFilter(ir_of(src, False)).render(157, 90, {}, t=0.25)  # the oracle on the IR WITHOUT the pass
    got = mb.Invocation(m, 157, 90).render(0, 0.25)
    assert np.array_equal(got, want), name


@pytest.mark.gpu
@pytest.mark.parametrize("rel,uv", CASES)
def test_mandelbrot_on_the_device_against_the_plain_ir(rel, uv):
    src = filter_source(rel)
    m = mb.Module(source=src)
    want = OracleFilter(ir_of(src, False)).render(301, 203, dict(uv), t=0.4)
    inv = mb.Invocation(m, 301, 203)
    for k, v in uv.items():
        inv.set(k, v)
    assert np.array_equal(inv.render(0, 0.4), want), (rel, uv)


def test_ir_loaded_over_the_boundary_gets_the_same_pass():
    """mmb_load_ir is what the reference-side binding calls with the reference compiler's IR (no such pass there): loading the
    plain IR must end in the kernels the front end builds, and loading carried IR must change nothing."""
    sources = [filter_source("examples/Render/Mandelbrot.mm"), filter_source("examples/Render/Fancy Mandelbrot.mm")] + [SHAPES[k] for k in sorted(SHAPES)]
    for src in sources:
        plain, carried = ir_of(src, False), ir_of(src, True)
        assert plain != carried
        loaded = mb.Module(ir=plain)
        assert loaded.cuda_source == mb.Module(source=src).cuda_source
        assert mb.Module(ir=carried).ir == carried
        # and with the pass switched off the loader leaves the text alone
        os.environ["MMB_LOOP_CARRY"] = "0"
        try:
            assert mb.Module(ir=plain).ir == plain
        finally:
            del os.environ["MMB_LOOP_CARRY"]


def test_the_compiled_mandelbrot_loop_is_33_instructions(tmp_path):
    """What the pass buys, read from the cubin NVRTC builds for the benchmark's kernel (no GPU needed): the hot loop executes
    10 FMUL + 19 FADD + counter, two compares and the branch -- 33 instructions an iteration instead of 37."""
    import collections
    import glob
    import re
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    mb.set_cubin_cache_dir(str(tmp_path))
    try:
        m = mb.Module(source=filter_source("examples/Render/Mandelbrot.mm"))
        assert m.compile_check(antialiasing=False, precise=True) > 0
    finally:
        mb.set_cubin_cache_dir(None)
    (path,) = glob.glob(os.path.join(str(tmp_path), "*.cubin"))
    blob = open(path, "rb").read()
    elf = os.path.join(str(tmp_path), "k.elf")
    with open(elf, "wb") as f:
        f.write(blob[blob.index(b"\x7fELF"):])
    sass = subprocess.run(["cuobjdump", "-sass", elf], stdout=subprocess.PIPE, text=True, check=True).stdout
    lines = [(int(mm.group(1), 16), mm.group(3)) for mm in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(@!?U?P\d )?([^;]+);", sass)]
    loops = []
    for addr, text in lines:
        mm = re.match(r"BRA (P\d, )?0x([0-9a-f]+)", text)
        if mm and int(mm.group(2), 16) < addr:
            loops.append(collections.Counter(t.split()[0].split(".")[0] for a, t in lines if int(mm.group(2), 16) <= a <= addr))
    hot = [c for c in loops if c["FMUL"] >= 8]
    assert len(hot) == 1, loops
    assert hot[0]["FMUL"] == 10 and hot[0]["FADD"] == 19 and sum(hot[0].values()) == 33, hot[0]


def test_compositions_with_an_escape_loop_inside():
    """Seeded two- and three-node compositions whose first node is one of the Mandelbrot filters: the loop ends up inlined into
    other filters' sampling code (closures, filter calls, other loops).  The pass changes every one of these IRs; the oracle
    renders the same bytes with and without it."""
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_compositions as fc
    from conftest import synthetic_rgba
    rng = random.Random(5)
    filters = fc.catalogue()
    with_img = [f for f in filters if f[1]]
    loops = [f for f in filters if "andelbrot" in f[0]]
    assert len(loops) == 2
    W, H = 80, 56
    imgs = [synthetic_rgba(W, H, seed=s) for s in (1, 2, 3)]
    done = 0
    while done < 8:
        a, b = rng.choice(loops), rng.choice(with_img)
        nodes = '(node :name "a" :type "%s" :input-slots ()) (node :name "b" :type "%s" :input-slots (("%s" "a" "out")))' % (a[0], b[0], rng.choice(b[1]))
        root = "b"
        if rng.random() < 0.4:
            c = rng.choice(with_img)
            if c[0] in (a[0], b[0]):
                continue
            nodes += ' (node :name "c" :type "%s" :input-slots (("%s" "b" "out")))' % (c[0], rng.choice(c[1]))
            root = "c"
        design = '(design %s :name "comp" :root "%s")' % (nodes, root)
        try:
            src = mb.design_to_source(design, fc.EX)
            plain, carried = ir_of(src, False), ir_of(src, True)
        except mb.MathMapError:
            continue  # e.g. an argument name defined twice in the generated source
        done += 1
        assert plain != carried, design
        vals, k = {}, 0
        for name, kind, _lo, _hi, _default in mb.Module(ir=carried).uservals():
            if kind == mb.USERVAL_IMAGE:
                vals[name] = imgs[k % 3]
                k += 1
        t, aa = rng.choice([0.0, 0.3, 0.75]), bool(rng.getrandbits(1))
        assert np.array_equal(OracleFilter(plain).render(W, H, vals, t=t, antialiasing=aa),
                              OracleFilter(carried).render(W, H, vals, t=t, antialiasing=aa)), design


def test_random_loop_filters_with_and_without_the_pass():
    """tools/fuzz_loops.py (any seed and count; profiles/r02_fuzz_loops.log keeps a run of 600): random loops whose conditions
    share subexpressions with their bodies, chains of carried values, nested loops."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_loops
    done, changed, failures = fuzz_loops.run(3, 30)
    assert done == 30 and changed >= 10 and not failures, failures
