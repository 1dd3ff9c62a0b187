"""The drop-in boundary as the reference calls it (VERDICT W2): mathfuncs_t.calc_lines with the slice / frame /
invocation parameters of new_template.c.in:208-312, driven by the reference's own call patterns.  GPU only."""
import numpy as np
import pytest

import mathmap_b200 as mb
from conftest import compare_u8, filter_source, synthetic_rgba
from oracle.oracle import OracleFilter

pytestmark = pytest.mark.gpu


def zoom_invocation(w, h, aa, supersampling=False, bpp=4):
    img = synthetic_rgba(w, h)
    m = mb.Module(source=filter_source("examples/Geometry/Zoom.mm"))
    inv = mb.Invocation(m, w, h, antialiasing=aa, supersampling=supersampling, precise=True)
    inv.set("in", img)
    inv.set("factor", 0.83)
    inv.set_output_bpp(bpp)
    return m, inv, img


@pytest.mark.parametrize("aa", [False, True], ids=["nearest", "bilinear"])
@pytest.mark.parametrize("bpp", [4, 3])
def test_reference_supersampling_call_pattern_through_calc_lines_slice(aa, bpp):
    """call_invocation with invocation->supersampling (mathmap_common.c:880-927) replayed call by call: a short slice at
    offset 0 and a slice one column wider at (-0.5, -0.5), three calc_lines calls per row, combined on bytes by the
    CALLER.  Must equal the oracle's -o render (which restates the same loop on the host) in every byte."""
    W, H = 131, 97
    m, inv, img = zoom_invocation(W, H, aa, supersampling=True, bpp=bpp)
    inv.init_frame(0, 0.0)
    region_x, region_y, region_w, region_h = 0, 0, W, H
    line1 = np.zeros(((region_w + 1) * bpp,), np.uint8)
    line2 = np.zeros((region_w * bpp,), np.uint8)
    line3 = np.zeros(((region_w + 1) * bpp,), np.uint8)
    out = np.zeros((H, W, bpp), np.uint8)
    short = dict(region=(region_x, region_y, region_w, region_h), offset=(0.0, 0.0), frame_size=(W, H), row_stride=W * bpp)
    long_ = dict(region=(region_x, region_y, region_w + 1, region_h), offset=(-0.5, -0.5), frame_size=(W, H), row_stride=W * bpp)
    inv.calc_lines_slice(region_y, region_y + 1, line1, **long_)
    for row in range(region_y, region_y + region_h):
        inv.calc_lines_slice(row, row + 1, line2, **short)
        inv.calc_lines_slice(row + 1, row + 2, line3, **long_)  # clamped away on the last row: line3 keeps its bytes
        l1 = line1.reshape(region_w + 1, bpp).astype(np.int32)
        l2 = line2.reshape(region_w, bpp).astype(np.int32)
        l3 = line3.reshape(region_w + 1, bpp).astype(np.int32)
        out[row - region_y] = ((l1[:-1] + l1[1:] + 2 * l2 + l3[:-1] + l3[1:]) // 6).astype(np.uint8)
        line1[:] = line3
    want = OracleFilter(m.ir).render(W, H, {"in": img, "factor": 0.83}, antialiasing=aa, supersampling=True, bpp=bpp)
    assert np.array_equal(out, want), compare_u8(out, want)
    # and the library's own whole-band entry (mmb_calc_lines = call_invocation) gives the same frame
    assert np.array_equal(inv.render(0, 0.0), want)


def test_tile_regions_with_row_stride_like_the_gimp_render_loop():
    """mathmap.c:1160-1175: GIMP hands over tile-sized regions (region_x != 0) of a destination whose rowstride is not
    region_width * bpp.  Every tile must carry the same bytes as that rectangle of a whole-frame render, and bytes of the
    destination outside the tile must stay untouched."""
    W, H, bpp = 150, 101, 4
    m, inv, img = zoom_invocation(W, H, True)
    whole = inv.render(0, 0.0)
    inv.init_frame(0, 0.0)
    stride = 64 * bpp + 12  # a tile buffer wider than any region
    for rx, ry, rw, rh in [(0, 0, 64, 64), (64, 0, 64, 64), (128, 0, 22, 64), (0, 64, 64, 37), (128, 64, 22, 37), (37, 11, 1, 1)]:
        buf = np.full((rh * stride,), 0xAB, np.uint8)
        inv.calc_lines_slice(ry, ry + rh, buf, region=(rx, ry, rw, rh), frame_size=(W, H), row_stride=stride)
        rows = buf.reshape(rh, stride)
        assert np.array_equal(rows[:, :rw * bpp].reshape(rh, rw, bpp), whole[ry:ry + rh, rx:rx + rw]), (rx, ry, rw, rh)
        assert (rows[:-1, rw * bpp:] == 0xAB).all() if rh > 1 else True
    # first_row / last_row are clamped to the region like new_template.c.in:238-239; q is the first RENDERED row
    buf = np.zeros((10 * W * bpp,), np.uint8)
    inv.calc_lines_slice(20, 200, buf, region=(0, 20, W, 10), frame_size=(W, H), row_stride=W * bpp)
    assert np.array_equal(buf.reshape(10, W, bpp), whole[20:30])


def test_floatmap_rows_advance_by_the_frame_width():
    """new_template.c.in:299-302: floatmap output advances q by frame_render_width pixels per row whatever the region."""
    W, H = 96, 40
    m, inv, img = zoom_invocation(W, H, True)
    whole = inv.render(0, 0.0, floatmap=True)
    inv.init_frame(0, 0.0)
    rx, ry, rw, rh = 10, 5, 30, 7
    buf = np.full((rh, W, 4), -7.0, np.float32)
    inv.calc_lines_slice(ry, ry + rh, buf, region=(rx, ry, rw, rh), frame_size=(W, H), floatmap=True)
    assert np.array_equal(buf[:, :rw], whole[ry:ry + rh, rx:rx + rw])
    assert (buf[:-1, rw:] == -7.0).all()


def test_scaled_preview_render_size():
    """The GIMP preview renders a smaller frame of the same image (mathmap.c:2191-2223: render_width/height = preview size):
    coordinates come from the frame size, __renderPixelW/H from the invocation's render size."""
    W, H = 200, 120
    src = "filter f (image in) in(xy) * (__renderPixelW / 100) end"
    img = synthetic_rgba(W, H)
    m = mb.Module(source=src)
    inv = mb.Invocation(m, W, H, antialiasing=True)
    inv.set("in", img)
    inv.set_render_size(100, 60)
    inv.init_frame(0, 0.0)
    got = inv.calc_lines_slice(0, 60, np.zeros((60, 100, 4), np.uint8), region=(0, 0, 100, 60), frame_size=(100, 60))
    # the oracle renders a 100x60 frame of the same input: same virtual coordinates, __renderPixelW = 100
    want = OracleFilter(m.ir).render(100, 60, {"in": img}, antialiasing=True)
    assert np.array_equal(got, want), compare_u8(got, want)


def test_user_stream_is_ordered_after_init_frame_work():
    """a caller's non-blocking stream does not synchronise with the legacy default stream the library queues
    init_frame's blur on; the *_device entry points order the two with events.  The blur makes the race window wide."""
    import torch
    W, H = 2048, 1536
    img = synthetic_rgba(W, H)
    m = mb.Module(source=filter_source("examples/Blur/Gaussian Blur.mm"))
    inv = mb.Invocation(m, W, H, antialiasing=True)
    inv.set("in", img)
    inv.set("dev", 0.02)
    want = inv.render(0, 0.0)
    stream = torch.cuda.Stream()  # non-blocking with respect to the legacy default stream
    out = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    for _ in range(5):
        out.zero_()
        torch.cuda.synchronize()
        inv.init_frame(0, 0.0)
        inv.calc_lines_device(out.data_ptr(), 0, H, stream=stream.cuda_stream)
        # the next frame's init_frame recycles the pool blocks this frame's kernel still reads
        inv.init_frame(0, 0.0)
        stream.synchronize()
        assert np.array_equal(out.cpu().numpy(), want)
    frames = torch.zeros((3, H, W, 4), dtype=torch.uint8, device="cuda")
    inv.render_frames_device(frames.data_ptr(), [0.0, 0.0, 0.0], [0, 1, 2], stream=stream.cuda_stream)
    inv.synchronize()
    stream.synchronize()
    for i in range(3):
        assert np.array_equal(frames[i].cpu().numpy(), want), i


RESIZED_FLOATMAP = """
filter inner (pixel image im, float s: 0-1 (0.02))
  b = gaussian_blur(im, s, s);
  b(xy)
end
filter outer (image in)
  rendered = render(in);
  inner(rendered, 0.02, xy)
end
"""


def test_native_filter_on_a_resize_wrapped_floatmap():
    """a RESIZE wrapper around a floatmap (an image argument whose coordinate flags differ from the filter's,
    compiler.c:1710-1773) is not a floatmap to render_image / gaussian_blur (builtins.c:275, gauss.c:655-657): it is resampled
    with the factors first and the blurred result carries none."""
    W, H = 160, 96
    img = synthetic_rgba(W, H)
    m = mb.Module(source=RESIZED_FLOATMAP)
    inv = mb.Invocation(m, W, H, antialiasing=True)
    inv.set("in", img)
    got = inv.render(0, 0.0)
    want = OracleFilter(m.ir).render(W, H, {"in": img}, antialiasing=True)
    exact, le1, mx = compare_u8(got, want)
    assert exact >= 99.99, (exact, le1, mx)
