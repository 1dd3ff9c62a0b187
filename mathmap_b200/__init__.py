"""mathmap_b200 — B200-native evaluation backend for MathMap's per-pixel render path.

Python is plumbing here: a ctypes binding of the C ABI in include/mathmap_b200.h
(the product is libmathmap_b200.so: front end + IR passes + CUDA emitter + NVRTC
driver + device runtime + kernels) and a small host-side mirror of the
reference's command-line flow (mathmap_cmdline.c:464-871): compile a .mm script,
bind uservals (-D), choose -i / -o, render frames.

There is no CPU rendering path: every render call needs the CUDA library and a
GPU and raises otherwise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmathmap_b200.so")
_lib = None

USERVAL_INT, USERVAL_FLOAT, USERVAL_BOOL, USERVAL_COLOR, USERVAL_CURVE, USERVAL_GRADIENT, USERVAL_IMAGE = range(7)
EDGE_COLOR, EDGE_WRAP, EDGE_REFLECT, EDGE_ROTATE = range(4)
CURVE_POINTS = 1024


class MathMapError(RuntimeError):
    pass


class Slice(ctypes.Structure):
    """mmb_slice: what the reference's calc_lines reads from mathmap_slice_t / mathmap_frame_t / invocation->row_stride."""
    _fields_ = [("frame_render_width", ctypes.c_int), ("frame_render_height", ctypes.c_int),
                ("region_x", ctypes.c_int), ("region_y", ctypes.c_int), ("region_width", ctypes.c_int), ("region_height", ctypes.c_int),
                ("sampling_offset_x", ctypes.c_float), ("sampling_offset_y", ctypes.c_float), ("row_stride", ctypes.c_int)]


def lib():
    """Loads libmathmap_b200.so (built in-tree by mathmap_b200.build); fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise MathMapError("%s is missing: run `python -m mathmap_b200.build` (there is no fallback path)" % _LIB_PATH)
    L = ctypes.CDLL(_LIB_PATH)
    vp, ci, cf, cc = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_char_p
    sigs = {
        "mmb_compile": (vp, [cc]), "mmb_load_ir": (vp, [cc]), "mmb_module_free": (None, [vp]),
        "mmb_design_to_source": (vp, [cc, cc]), "mmb_free_string": (None, [vp]), "mmb_compile_design": (vp, [cc, cc]),
        "mmb_module_ir": (cc, [vp]), "mmb_module_cuda_source": (cc, [vp]), "mmb_module_main_filter_name": (cc, [vp]),
        "mmb_module_num_uservals": (ci, [vp]),
        "mmb_module_userval_info": (ci, [vp, ci, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ci), ctypes.POINTER(cf), ctypes.POINTER(cf), ctypes.POINTER(cf)]),
        "mmb_module_userval_index": (ci, [vp, cc]),
        "mmb_module_compile_check": (ctypes.c_long, [vp, ci, ci]), "mmb_module_compile_check_fast": (ctypes.c_long, [vp, ci, ci]), "mmb_set_cubin_cache_dir": (ci, [ctypes.c_char_p]),
        "mmb_invoke": (vp, [vp, ci, ci, ci]), "mmb_invocation_free": (None, [vp]),
        "mmb_set_antialiasing": (ci, [vp, ci]), "mmb_set_supersampling": (ci, [vp, ci]),
        "mmb_set_edge_behaviour": (ci, [vp, ci, ci, ctypes.c_uint32, ctypes.c_uint32]),
        "mmb_set_output_bpp": (ci, [vp, ci]), "mmb_set_precise_math": (ci, [vp, ci]), "mmb_set_warp_shape": (ci, [vp, ci]), "mmb_set_specialize": (ci, [vp, ci]), "mmb_set_fast_compile": (ci, [vp, ci]), "mmb_set_rows_per_thread": (ci, [vp, ci]),
        "mmb_set_userval_int": (ci, [vp, ci, ci]), "mmb_set_userval_float": (ci, [vp, ci, cf]), "mmb_set_userval_bool": (ci, [vp, ci, ci]),
        "mmb_set_userval_color": (ci, [vp, ci, cf, cf, cf, cf]),
        "mmb_set_userval_color_packed": (ci, [vp, ci, ctypes.c_uint32]),
        "mmb_set_userval_curve": (ci, [vp, ci, vp]), "mmb_set_userval_gradient": (ci, [vp, ci, vp]),
        "mmb_set_userval_image_host": (ci, [vp, ci, vp, ci, ci]), "mmb_set_userval_image_device": (ci, [vp, ci, vp, ci, ci]),
        "mmb_init_frame": (ci, [vp, ci, cf]),
        "mmb_calc_lines": (ci, [vp, ci, ci, vp, ci]), "mmb_calc_lines_device": (ci, [vp, ci, ci, vp, ci, vp]),
        "mmb_calc_lines_slice": (ci, [vp, vp, ci, ci, vp, ci]), "mmb_calc_lines_slice_device": (ci, [vp, vp, ci, ci, vp, ci, vp]),
        "mmb_set_render_size": (ci, [vp, ci, ci]),
        "mmb_render_frames_device": (ci, [vp, ci, vp, vp, vp, vp]),
        "mmb_calc_lines_interleaved_device": (ci, [vp, ci, ci, vp, vp]),
        "mmb_synchronize": (ci, [vp]), "mmb_launch_count": (ctypes.c_long, [vp]), "mmb_kernel_name": (cc, [vp]),
        "mmb_gaussian_blur_device": (ci, [ci, vp, vp, ci, ci, cf, cf, vp]),
        "mmb_gauss_iir_constants": (None, [cf, vp]),
        "mmb_last_error": (cc, []), "mmb_version": (cc, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _err():
    return lib().mmb_last_error().decode("utf-8", "replace")


def set_cubin_cache_dir(path):
    """Opt-in persistent cache of compiled kernels (None turns it off); see mmb_set_cubin_cache_dir."""
    lib().mmb_set_cubin_cache_dir(None if not path else os.fsencode(path))


def design_to_source(design_text, filter_path):
    """The MathMap source of a composition (reference: make_filter_source_from_design, designer_filter.c:278)."""
    L = lib()
    p = L.mmb_design_to_source(design_text.encode(), os.fsencode(filter_path))
    if not p:
        raise MathMapError(_err())
    try:
        return ctypes.string_at(p).decode()
    finally:
        L.mmb_free_string(p)


def _check_out(out, need_bytes, dtype=None):
    """The C ABI writes need_bytes at out's address: refuse anything but a large enough, writable, C-contiguous array."""
    if not isinstance(out, np.ndarray) or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
        raise ValueError("out must be a writable C-contiguous numpy array")
    if dtype is not None and out.dtype != dtype:
        raise ValueError("out must be of dtype %s" % np.dtype(dtype).name)
    if out.nbytes < need_bytes:
        raise ValueError("out holds %d bytes, the call writes %d" % (out.nbytes, need_bytes))


class Module:
    """A compiled filter module (reference: mathmap_t after compile_mathmap, mathmap_common.c:504)."""

    def __init__(self, source=None, ir=None):
        L = lib()
        if (source is None) == (ir is None):
            raise ValueError("give exactly one of source= or ir=")
        self._h = L.mmb_compile(source.encode()) if source is not None else L.mmb_load_ir(ir.encode())
        if not self._h:
            raise MathMapError(_err())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mmb_module_free(h)

    @classmethod
    def from_file(cls, path, filter_path=None):
        """A `.mm` filter, or a `.mmc` composition whose node types are looked up under `filter_path`
        (default: the directory tree the file lives in, two levels up like the reference's examples/ layout)."""
        with open(path) as f:
            text = f.read()
        if path.endswith(".mmc"):
            if filter_path is None:
                filter_path = os.path.dirname(os.path.dirname(os.path.abspath(path)))
            return cls(source=design_to_source(text, filter_path))
        return cls(source=text)

    @property
    def ir(self):
        return lib().mmb_module_ir(self._h).decode()

    @property
    def name(self):
        return lib().mmb_module_main_filter_name(self._h).decode()

    @property
    def cuda_source(self):
        s = lib().mmb_module_cuda_source(self._h)
        if s is None:
            raise MathMapError(_err())
        return s.decode()

    def uservals(self):
        """[(name, type, min, max, default)] of the main filter's arguments."""
        L = lib()
        out = []
        for i in range(L.mmb_module_num_uservals(self._h)):
            name = ctypes.create_string_buffer(256)
            t, lo, hi, d = ctypes.c_int(), ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
            L.mmb_module_userval_info(self._h, i, name, 256, ctypes.byref(t), ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(d))
            out.append((name.value.decode(), t.value, lo.value, hi.value, d.value))
        return out

    def compile_check(self, antialiasing=True, precise=False, fast_compile=False):
        """NVRTC-compiles for sm_100a (no GPU needed); returns the cubin size."""
        fn = lib().mmb_module_compile_check_fast if fast_compile else lib().mmb_module_compile_check
        n = fn(self._h, int(antialiasing), int(precise))
        if n < 0:
            raise MathMapError(_err())
        return n


class Invocation:
    """reference: mathmap_invocation_t from invoke_mathmap (mathmap_common.c:747)."""

    def __init__(self, module, width, height, device=0, antialiasing=False, supersampling=False, precise=True, warp_width=None, rows_per_thread=None,
                 specialize=None, fast_compile=None):
        self.module = module
        self.width, self.height = width, height
        self.bpp = 4
        self._h = lib().mmb_invoke(module._h, width, height, device)
        if not self._h:
            raise MathMapError(_err())
        self._uv = {u[0]: (i, u[1]) for i, u in enumerate(module.uservals())}
        self._keep = {}
        self._ck(lib().mmb_set_antialiasing(self._h, int(antialiasing)))
        self._ck(lib().mmb_set_supersampling(self._h, int(supersampling)))
        self._ck(lib().mmb_set_precise_math(self._h, int(precise)))
        if warp_width is not None:
            self._ck(lib().mmb_set_warp_shape(self._h, warp_width))
        if rows_per_thread is not None:
            self._ck(lib().mmb_set_rows_per_thread(self._h, rows_per_thread))
        if specialize is not None:
            self._ck(lib().mmb_set_specialize(self._h, int(specialize)))
        if fast_compile is not None:
            self._ck(lib().mmb_set_fast_compile(self._h, int(fast_compile)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.mmb_invocation_free(h)

    def _ck(self, rc):
        if rc != 0:
            raise MathMapError(_err())

    def set_edge_behaviour(self, mode_x, mode_y, color_x=0, color_y=0):
        self._ck(lib().mmb_set_edge_behaviour(self._h, mode_x, mode_y, color_x, color_y))

    def set_output_bpp(self, bpp):
        self._ck(lib().mmb_set_output_bpp(self._h, bpp))
        self.bpp = bpp

    def set(self, name, value):
        """Binds one argument: int/float/bool, (r,g,b,a), a 1024-sample curve/gradient, an RGBA8 [H,W,4]
        numpy array (copied to the device) or a CUDA torch tensor of that shape (adopted in place)."""
        if name not in self._uv:
            raise KeyError("filter %s has no argument %r" % (self.module.name, name))
        idx, typ = self._uv[name]
        L = lib()
        if typ == USERVAL_INT:
            self._ck(L.mmb_set_userval_int(self._h, idx, int(value)))
        elif typ == USERVAL_FLOAT:
            self._ck(L.mmb_set_userval_float(self._h, idx, float(value)))
        elif typ == USERVAL_BOOL:
            self._ck(L.mmb_set_userval_bool(self._h, idx, int(bool(value))))
        elif typ == USERVAL_COLOR:
            r, g, b, a = value
            self._ck(L.mmb_set_userval_color(self._h, idx, r, g, b, a))
        elif typ == USERVAL_CURVE:
            arr = np.ascontiguousarray(value, dtype=np.float32)
            assert arr.shape == (CURVE_POINTS,)
            self._ck(L.mmb_set_userval_curve(self._h, idx, arr.ctypes.data))
        elif typ == USERVAL_GRADIENT:
            arr = np.ascontiguousarray(value, dtype=np.uint32)
            assert arr.shape == (CURVE_POINTS,)
            self._ck(L.mmb_set_userval_gradient(self._h, idx, arr.ctypes.data))
        elif typ == USERVAL_IMAGE:
            if hasattr(value, "data_ptr"):  # torch tensor on the device
                t = value
                assert t.is_cuda and t.is_contiguous() and t.dim() == 3 and t.shape[2] == 4 and t.element_size() == 1
                self._keep[name] = t
                self._ck(L.mmb_set_userval_image_device(self._h, idx, t.data_ptr(), t.shape[1], t.shape[0]))
            else:
                arr = np.ascontiguousarray(value, dtype=np.uint8)
                assert arr.ndim == 3 and arr.shape[2] == 4, "images are RGBA8 [H, W, 4]"
                self._ck(L.mmb_set_userval_image_host(self._h, idx, arr.ctypes.data, arr.shape[1], arr.shape[0]))

    def init_frame(self, frame=0, t=0.0):
        self._ck(lib().mmb_init_frame(self._h, int(frame), float(t)))

    def calc_lines(self, first_row=0, last_row=None, floatmap=False, out=None):
        """Renders rows [first_row, last_row) into a host array (copied back from the device)."""
        last_row = self.height if last_row is None else last_row
        rows = max(0, min(last_row, self.height) - max(0, first_row))
        if out is None:
            out = np.empty((rows, self.width, 4), dtype=np.float32) if floatmap else np.empty((rows, self.width, self.bpp), dtype=np.uint8)
        else:
            _check_out(out, rows * self.width * (16 if floatmap else self.bpp), np.float32 if floatmap else np.uint8)
        self._ck(lib().mmb_calc_lines(self._h, first_row, last_row, out.ctypes.data, int(floatmap)))
        return out

    def set_render_size(self, render_width, render_height):
        self._ck(lib().mmb_set_render_size(self._h, render_width, render_height))

    def calc_lines_slice(self, first_row, last_row, out, region=None, offset=(0.0, 0.0), frame_size=None, row_stride=None, floatmap=False):
        """mathfuncs_t.calc_lines with the reference's slice parameters (mmb_calc_lines_slice): region = (x, y, w, h) of the
        frame_size = (w, h) frame; `out` is a C-contiguous numpy buffer whose first byte is the first rendered row."""
        fw, fh = frame_size or (self.width, self.height)
        rx, ry, rw, rh = region or (0, 0, fw, fh)
        if row_stride is None:
            row_stride = rw * self.bpp
        fr, lr = max(0, first_row), min(last_row, ry + rh)
        need = 0
        if lr > fr and rw > 0:
            pitch = fw * 16 if floatmap else row_stride
            need = (lr - fr - 1) * pitch + rw * (16 if floatmap else self.bpp)
        _check_out(out, need)
        sl = Slice(fw, fh, rx, ry, rw, rh, offset[0], offset[1], row_stride)
        self._ck(lib().mmb_calc_lines_slice(self._h, ctypes.byref(sl), first_row, last_row, out.ctypes.data, int(floatmap)))
        return out

    def calc_lines_device(self, device_ptr, first_row=0, last_row=None, floatmap=False, stream=0):
        last_row = self.height if last_row is None else last_row
        self._ck(lib().mmb_calc_lines_device(self._h, first_row, last_row, device_ptr, int(floatmap), stream))

    def calc_lines_interleaved_device(self, device_ptr, phase, count, stream=0):
        """8-row blocks b with b % count == phase, stored compactly (row-band sharding across ranks)."""
        self._ck(lib().mmb_calc_lines_interleaved_device(self._h, phase, count, device_ptr, stream))

    def render_frames_device(self, device_ptr, ts, frames=None, stream=0):
        n = len(ts)
        ts_arr = (ctypes.c_float * n)(*ts)
        fr = (ctypes.c_int * n)(*(frames if frames is not None else range(n)))
        self._ck(lib().mmb_render_frames_device(self._h, n, fr, ts_arr, device_ptr, stream))

    def render(self, frame=0, t=0.0, floatmap=False):
        """One whole frame to a host array (init_frame + calc_lines), like one pass of the CLI render loop."""
        self.init_frame(frame, t)
        return self.calc_lines(0, self.height, floatmap)

    def synchronize(self):
        self._ck(lib().mmb_synchronize(self._h))

    @property
    def launch_count(self):
        return lib().mmb_launch_count(self._h)

    @property
    def kernel_name(self):
        s = lib().mmb_kernel_name(self._h)
        return s.decode() if s else ""


def render_file(path, width=None, height=None, uservals=None, t=0.0, frame=0, antialiasing=True, supersampling=False, device=0,
                precise=True):
    """Convenience mirror of `mathmap [-i] [-o] -f script.mm [-s WxH] [-Dname=value ...]`: returns uint8 [H, W, 4].
    The size defaults to that of the first image argument (mathmap_cmdline.c:717-752)."""
    m = Module.from_file(path)
    uservals = uservals or {}
    if width is None or height is None:
        for v in uservals.values():
            if hasattr(v, "shape") and len(v.shape) == 3:
                height, width = v.shape[0], v.shape[1]
                break
    if width is None:
        raise ValueError("no size given and no image argument to take it from")
    inv = Invocation(m, width, height, device, antialiasing, supersampling, precise)
    for k, v in uservals.items():
        inv.set(k, v)
    return inv.render(frame, t)
