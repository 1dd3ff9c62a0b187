"""Front end and IR: language rules that decide what a filter means, IR text round trip, constness levels.  CPU only."""
import os
import re

import pytest

import mathmap_b200 as mb
from conftest import filter_source, load_manifest


def ir_of(src):
    return mb.Module(source=src).ir


def test_all_fixture_filters_compile_and_round_trip():
    """mmb_load_ir(mmb_module_ir(m)) reproduces the same IR text: the boundary format is self-consistent."""
    for e in load_manifest():
        ir = ir_of(filter_source(e["script"]))
        again = mb.Module(ir=ir).ir
        assert again == ir, e["script"]


def test_integer_and_float_typing():
    ir = ir_of("filter f () p = 3; q = p / 2; c = floor(q * 3); grayColor(c % 2) end")
    # 3/2 is folded as a float division (DIV always yields float), floor gives an int, % is fmod (float)
    assert "f:1.5" not in ir  # folded further
    assert re.search(r"\(tuple f:0\.0 f:0\.0 f:0\.0 i:1\)", ir), ir


def test_division_by_zero_guard_and_short_circuit_shape():
    ir = ir_of("filter f (float p: 0-1) grayColor(1 / p) end")
    assert "(op EQ" in ir and "(op DIV i:1" in ir and "(phi" in ir


def test_operator_precedence_and_casts():
    # unary minus binds tighter than ^ ; cast binds tightest: ri:x*2 is (ri:x)*2
    a = ir_of("filter f () v = -2 ^ 2; grayColor(v) end")
    assert "f:4.0" in a
    with pytest.raises(mb.MathMapError):
        ir_of("filter f () v = xy:[1,2] + ri:[1,2,3]; grayColor(1) end")


def test_variable_typing_errors():
    with pytest.raises(mb.MathMapError) as e:
        ir_of("filter f () v = 1; v = [1,2]; grayColor(v) end")
    assert "two different types" in str(e.value)
    with pytest.raises(mb.MathMapError) as e:
        ir_of("filter f () x = 1; grayColor(x) end")
    assert "internal variable" in str(e.value)


def test_for_loop_and_do_while():
    ir = ir_of("filter f () s = 0; for i = 1 .. 4 do s = s + i end; grayColor(s / 10) end")
    assert "(while" in ir  # loops are never unrolled; the whole loop is frame-constant (level 0)
    assert re.search(r"\(while \(phis[^)]*\n\s*\(phi %\d+\.\d+ \d+ 0 ", ir), ir
    or_ir = ir_of("filter f (int n: 1-10 (3)) s = 0; i = 0; do s = s + i; i = i + 1 while i < n end; grayColor(s) end")
    assert "(while" in or_ir


def test_levels_hoist_frame_constants_and_rows():
    ir = ir_of(filter_source("examples/Distorts/Sea.mm"))
    sin_line = [l for l in ir.splitlines() if "(op sin" in l][0]
    assert re.search(r"\(assign %\d+\.\d+ \d+ 1 \(op sin", sin_line), "sin(t*2*pi + f(y)) is constant along a row"
    uv_line = [l for l in ir.splitlines() if "USERVAL_FLOAT_ACCESS" in l][0]
    assert " 7 0 " in uv_line, "uservals are frame constants"
    orig = [l for l in ir.splitlines() if "ORIG_VAL" in l][0]
    assert " 0 3 " in orig


def test_closure_inlining_and_recursion():
    src = filter_source("tests/Twice.mm")
    ir = ir_of(src)
    assert "(filter " in ir.split("(main")[0]
    rec = ir_of(filter_source("examples/Map/IFS Functional.mm"))
    assert re.search(r"\(filter \w+ ", rec.split("(code", 1)[1]), "recursive calls stay calls"


def test_native_filter_closure_is_frame_constant():
    ir = ir_of(filter_source("examples/Blur/Gaussian Blur.mm"))
    line = [l for l in ir.splitlines() if "(closure gaussian_blur" in l][0]
    assert re.search(r" 7 0 \(closure gaussian_blur", line)


TREE_VECTOR_SRC = """filter tv (int k: 0-7 (2))
    v = rgba:[x * 0.5 + 0.5, y * 0.5 + 0.5, 0.25, 1];
    i = floor((x + 1) * 2);
    q = v[i];
    w = v;
    w[i] = 1 - q;
    w[k] = w[k] * 0.5;
    rgba:[w[0], w[1], w[2], v[floor(y * 3)]]
end"""


def test_computed_subscripts_become_tree_vectors():
    """compiler.c:2521-2570, 1838-1873: a variable subscripted by a computed index is held as one tree-vector compvar,
    reads are TREE_VECTOR_NTH, element stores SET_TREE_VECTOR_NTH (also for literal subscripts of such a variable)."""
    ir = ir_of(TREE_VECTOR_SRC)
    assert re.search(r"\(\d+ tree_vector 4\)", ir)
    assert "(tree-vector " in ir
    assert ir.count("(op SET_TREE_VECTOR_NTH") == 2
    assert re.search(r"\(op TREE_VECTOR_NTH i:0 %\d+\.\d+\)", ir)
    # a select with only literal subscripts on an ordinary variable stays element-wise
    plain = ir_of("filter p () v = rgba:[x, y, 0, 1]; rgba:[v[1], v[0], v[2], v[3]] end")
    assert "TREE_VECTOR" not in plain and "tree_vector" not in plain
    # the IR text round-trips through the loader
    m2 = mb.Module(ir=ir)
    assert m2.ir == ir


def test_committed_workload_ir_is_what_the_front_end_produces():
    """tests/golden/ir/*.mmir (tools/make_golden_ir.py) feed the oracle in bench.py's CPU legs without loading the CUDA
    library; they must not drift from the front end."""
    import bench
    for name, wl in bench.WORKLOADS.items():
        m = mb.Module.from_file(os.path.join(bench.FILTERS, wl["script"]))
        assert m.ir == bench.workload_ir(name), "%s: run tools/make_golden_ir.py" % name


def test_every_example_survives_the_boundary_format():
    """The reference-side binding hands the optimised IR over as "mmir 1" text (integration/backends/cuda.c -> mmb_load_ir).  For
    all 189 example filters: the text loads, prints back identically, and the CUDA generated from the loaded IR is the CUDA
    generated from the front end's own -- so a filter compiled by the reference and one compiled here run the same kernels."""
    import glob
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filters", "examples")
    files = sorted(glob.glob(os.path.join(root, "*", "*.mm")))
    assert len(files) >= 180
    for path in files:
        m = mb.Module.from_file(path)
        again = mb.Module(ir=m.ir)
        assert again.ir == m.ir, path
        assert again.cuda_source == m.cuda_source, path
        assert again.uservals() == m.uservals() and again.name == m.name, path
