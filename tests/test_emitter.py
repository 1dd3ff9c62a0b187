"""Host-side checks of the CUDA emitter's structural decisions (no GPU needed: the generated source is inspected)."""
import os

import mathmap_b200 as mb

FILTERS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filters", "examples")


def source_of(rel=None, text=None):
    m = mb.Module.from_file(os.path.join(FILTERS, rel)) if rel else mb.Module(source=text)
    return m.cuda_source


def test_direct_output_only_when_the_pixel_is_the_sample():
    # every distortion filter ends in in(...): the sampler may hand its bytes to the store (mm_orig_val_out)
    src = source_of("Distorts/Twirl.mm")
    assert "mm_orig_val_out(" in src and "mm_store_word(" in src
    # ... four pixels at a time where the sample's row does not depend on the column (quad kernels)
    for rel in ("Distorts/Sea.mm", "Utilities/Ident.mm", "Geometry/Zoom.mm"):
        src = source_of(rel)
        assert "mm_orig_val_out_quad(" in src and "mm_store_quad(" in src, rel
    # a sample that is post-processed, used twice, or produced inside a conditional keeps the general path
    post = "filter f (image in)\n  c = in(xy);\n  rgba:[1-c[0], c[1], c[2], c[3]]\nend\n"
    twice = "filter g (image in)\n  c = in(xy);\n  d = c[0];\n  if d > 0.5 then c else rgba:[d, d, d, 1] end\nend\n"
    for text in (post, twice):
        src = source_of(text=text)
        assert "mm_orig_val_out(" not in src and "mm_orig_val_out_quad(" not in src and "mm_store_word(" not in src
    assert "mm_orig_val_out(" not in source_of("Render/Mandelbrot.mm")


def test_quad_kernels_for_samples_whose_row_is_shared():
    """the local-access kernel variant (four pixels of a row per thread) is chosen for straight-line pixel
    code with a sample whose y and frame do not depend on the column; its per-pixel values are arrays of four, row-level
    values stay scalars, a per-pixel `if` is printed once per pixel."""
    for rel in ("Utilities/Ident.mm", "Colors/Invert.mm", "Geometry/Translate.mm", "Geometry/Scale.mm", "Distorts/Sea.mm"):
        src = source_of(rel)
        assert "mm_pixel_coords_quad(" in src and ("mm_orig_val_quad(" in src or "mm_orig_val_out_quad(" in src), rel
    # a sample whose row depends on the column (rotation, twirl), per-pixel loops, or no sample at all: one pixel per thread
    for rel in ("Distorts/Twirl.mm", "Geometry/Rotate.mm", "Render/Mandelbrot.mm", "Map/Droste.mm"):
        assert "mm_pixel_coords_quad(" not in source_of(rel), rel
    cond = "filter q (image in)\n  c = in(xy);\n  if c[0] > 0.5 then c else in(xy*0.5) end\nend\n"
    src = source_of(text=cond)
    assert "mm_pixel_coords_quad(" in src and src.count("mm_orig_val_quad(") == 1
    assert src.count("mm_orig_val(P,") == 4  # the sample inside the per-pixel branch: once per pixel, through the one-pixel sampler


def test_tiles_per_block_follow_the_kernel_shape():
    # straight-line pixel code takes the tile count from the launch parameters; per-pixel loops are compiled for one tile
    assert "const int mm_rows = P.rows;" in source_of("Distorts/Twirl.mm")
    assert "const int mm_rows = P.rows;" in source_of("Colors/Invert.mm")
    for rel in ("Render/Mandelbrot.mm", "Map/Droste.mm"):
        src = source_of(rel)
        assert "constexpr int mm_rows = 1;" in src and "P.rows" not in src, rel


def test_pow_2_is_an_exact_product():
    src = source_of(text="filter h ()\n  v = x^2 + y^2;\n  grayColor(v)\nend\n")
    assert "mm_sqr(" in src and "mm_pow(" not in src
    src = source_of(text="filter k ()\n  grayColor(abs(x)^2.5)\nend\n")
    assert "mm_pow(" in src


def test_every_example_filter_compiles_for_sm100a():
    """All .mm files of the reference's examples tree go through the front end, the IR passes, the CUDA emitter and NVRTC
    (sm_100a, bilinear sampler, precise math) on the CPU: about a minute and a half, and the one check of the emitter's
    less common shapes (closures, row pre-kernels, calls, tree vectors) that needs no GPU."""
    import glob
    files = sorted(glob.glob(os.path.join(FILTERS, "*", "*.mm")))
    assert len(files) >= 180
    failed = []
    for path in files:
        try:
            if mb.Module.from_file(path).compile_check(antialiasing=True, precise=True) <= 0:
                failed.append((os.path.relpath(path, FILTERS), "empty cubin"))
        except mb.MathMapError as e:
            failed.append((os.path.relpath(path, FILTERS), str(e).splitlines()[0] if str(e) else "error"))
    assert not failed, failed


def test_fast_compile_is_a_different_smaller_kernel():
    """mmb_set_fast_compile: the complex elementary functions and everything of the samplers but a drawable's interior path as
    real calls -- one body per module instead of one per call site.  Droste's cubin shrinks by a third (and its NVRTC time by
    about 45 %: DESIGN.md, filter compile time); a filter without complex functions whose sample is its pixel already calls the
    general sampler and compiles to the same size."""
    droste = mb.Module.from_file(os.path.join(FILTERS, "Map/Droste.mm"))
    assert droste.compile_check(antialiasing=True, precise=True, fast_compile=True) < 0.8 * droste.compile_check(antialiasing=True, precise=True)
    twirl = mb.Module.from_file(os.path.join(FILTERS, "Distorts/Twirl.mm"))
    assert twirl.compile_check(antialiasing=True, precise=True, fast_compile=True) == twirl.compile_check(antialiasing=True, precise=True)
