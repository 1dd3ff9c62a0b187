"""ORACLE driver (test infrastructure, not product code).

Takes the optimised IR text of a filter module, emits host C shaped like the
reference cc backend's output (oracle/emit_c.py), compiles it with the
reference's own compiler line `gcc -O2 -fPIC` / `gcc -shared` (reference
Makefile:58-60: CGEN_CC / CGEN_LD), links it to the plain-C runtime restatement
(oracle/runtime/*.c) and renders on the host CPU through ctypes.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
import ctypes
import hashlib
import os
import subprocess
import threading

import numpy as np

from . import emit_c

HERE = os.path.dirname(os.path.abspath(__file__))
RUNTIME_DIR = os.path.join(HERE, "runtime")
BUILD_DIR = os.path.join(HERE, "_build")
RUNTIME_SOURCES = ["images.c", "driver.c", "gauss.c", "noise.c", "spec_funcs.c", "elliptic.c", "convolve.c"]
CGEN_CC = ["gcc", "-O2", "-c", "-fPIC"]  # reference Makefile:58
CGEN_LD = ["gcc", "-shared"]             # reference Makefile:59

_build_lock = threading.Lock()


class _UservalUnion(ctypes.Union):
    _fields_ = [("int_const", ctypes.c_int), ("float_const", ctypes.c_float), ("bool_const", ctypes.c_int),
                ("color", ctypes.c_uint), ("curve", ctypes.c_void_p), ("gradient", ctypes.c_void_p), ("image", ctypes.c_void_p)]


class _Userval(ctypes.Structure):
    _fields_ = [("v", _UservalUnion)]


class _RenderParams(ctypes.Structure):
    _fields_ = [("img_width", ctypes.c_int), ("img_height", ctypes.c_int), ("antialiasing", ctypes.c_int),
                ("supersampling", ctypes.c_int), ("edge_behaviour_x", ctypes.c_int), ("edge_behaviour_y", ctypes.c_int),
                ("edge_color_x", ctypes.c_uint), ("edge_color_y", ctypes.c_uint), ("output_bpp", ctypes.c_int),
                ("frame", ctypes.c_int), ("t", ctypes.c_float), ("num_threads", ctypes.c_int), ("floatmap", ctypes.c_int),
                ("num_uservals", ctypes.c_int), ("uservals", ctypes.POINTER(_Userval)), ("output", ctypes.c_void_p),
                ("taps", ctypes.c_long), ("sample_rows", ctypes.c_void_p), ("num_sample_rows", ctypes.c_int)]


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed: %s\n%s" % (" ".join(cmd), r.stdout))


def build_runtime():
    """Compiles oracle/runtime/*.c once into oracle/_build/*.o; returns the object paths."""
    os.makedirs(BUILD_DIR, exist_ok=True)
    objs = []
    with _build_lock:
        for src in RUNTIME_SOURCES:
            path = os.path.join(RUNTIME_DIR, src)
            obj = os.path.join(BUILD_DIR, src[:-2] + ".o")
            deps = [path, os.path.join(RUNTIME_DIR, "mmo_runtime.h"), os.path.join(RUNTIME_DIR, "noise_table.inc")]
            if not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
                tmp = obj + ".%d.tmp" % os.getpid()
                _run(CGEN_CC + ["-I", RUNTIME_DIR, "-o", tmp, path])
                os.replace(tmp, obj)
            objs.append(obj)
    return objs


def default_curve():
    """Identity curve, reference userval.c:281-309."""
    return (np.arange(1024, dtype=np.float32) / np.float32(1023)).astype(np.float32)


def default_gradient():
    """Gray ramp, reference mathmap.c:356-361: (int)(v*255) per channel, alpha 255, packed R<<24|G<<16|B<<8|A."""
    v = (np.arange(1024, dtype=np.float64) / 1023.0 * 255.0).astype(np.uint32)
    return ((v << 24) | (v << 16) | (v << 8) | np.uint32(255)).astype(np.uint32)


class OracleFilter:
    def __init__(self, ir_text):
        self.ir_text = ir_text
        src, module = emit_c.emit_module(ir_text)
        self.module = module
        self.main = module.filters[module.main]
        self.c_source = src
        objs = build_runtime()
        key = hashlib.sha1((src + "".join(sorted(objs))).encode()).hexdigest()[:20]
        so = os.path.join(BUILD_DIR, "filter_%s.so" % key)
        with _build_lock:
            newest_rt = max(os.path.getmtime(o) for o in objs)
            if not os.path.exists(so) or os.path.getmtime(so) < newest_rt:
                cfile = os.path.join(BUILD_DIR, "filter_%s.c" % key)
                with open(cfile, "w") as f:
                    f.write(src)
                obj = cfile[:-2] + ".o"
                _run(CGEN_CC + ["-I", RUNTIME_DIR, "-o", obj, cfile])
                tmp = so + ".%d.tmp" % os.getpid()
                _run(CGEN_LD + ["-o", tmp, obj] + objs + ["-lm", "-lpthread"])
                os.replace(tmp, so)
        self.lib = ctypes.CDLL(so)
        self.lib.mmo_render.argtypes = [ctypes.POINTER(_RenderParams)]
        self.lib.mmo_render.restype = ctypes.c_int
        self.lib.mmo_make_drawable.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        self.lib.mmo_make_drawable.restype = ctypes.c_void_p
        self.lib.mmo_free_drawable.argtypes = [ctypes.c_void_p]
        self.last_taps = 0

    def userval_defaults(self):
        d = {}
        for typ, name, rest in self.main.uservals:
            if typ == "int":
                d[name] = int(rest[2])
            elif typ == "float":
                d[name] = float(np.float32(float(rest[2])))
            elif typ == "bool":
                d[name] = int(rest[0])
            elif typ == "color":
                d[name] = (0.0, 0.0, 0.0, 1.0)
        return d

    def render(self, width, height, uservals=None, t=0.0, frame=0, antialiasing=True, supersampling=False, threads=1,
               edge_behaviour=(0, 0), edge_colors=(0, 0), bpp=4, floatmap=False, sample_rows=None):
        """Renders one frame; returns uint8 [H, W, bpp] (or float32 [H, W, 4] when floatmap).
        sample_rows: optional list of row indices; only those rows are rendered (output has len(sample_rows) rows)."""
        vals = self.userval_defaults()
        if uservals:
            for k, v in uservals.items():
                if k not in [u[1] for u in self.main.uservals]:
                    raise KeyError("filter %s has no argument %r" % (self.main.name, k))
                vals[k] = v
        n = len(self.main.uservals)
        arr = (_Userval * max(n, 1))()
        keep = []
        drawables = []
        for i, (typ, name, rest) in enumerate(self.main.uservals):
            if typ == "int":
                arr[i].v.int_const = int(vals[name])
            elif typ == "float":
                arr[i].v.float_const = float(vals[name])
            elif typ == "bool":
                arr[i].v.bool_const = 1 if vals[name] else 0
            elif typ == "color":
                r, g, b, a = [min(1.0, max(0.0, float(c))) for c in vals[name]]
                arr[i].v.color = (int(r * 255) << 24) | (int(g * 255) << 16) | (int(b * 255) << 8) | int(a * 255)
            elif typ == "curve":
                c = np.ascontiguousarray(vals.get(name, default_curve()), dtype=np.float32)
                keep.append(c)
                arr[i].v.curve = c.ctypes.data
            elif typ == "gradient":
                g = np.ascontiguousarray(vals.get(name, default_gradient()), dtype=np.uint32)
                keep.append(g)
                arr[i].v.gradient = g.ctypes.data
            elif typ == "image":
                img = vals.get(name)
                if img is None:
                    arr[i].v.image = self.lib.mmo_make_drawable(None, 0, 0)
                else:
                    img = np.ascontiguousarray(img, dtype=np.uint8)
                    assert img.ndim == 3 and img.shape[2] == 4, "images are RGBA8 [H, W, 4]"
                    keep.append(img)
                    arr[i].v.image = self.lib.mmo_make_drawable(img.ctypes.data, img.shape[1], img.shape[0])
                drawables.append(arr[i].v.image)
        rows_arr = None
        out_rows = height
        if sample_rows is not None:
            rows_arr = np.ascontiguousarray(sample_rows, dtype=np.int32)
            out_rows = len(rows_arr)
        if floatmap:
            out = np.zeros((out_rows, width, 4), dtype=np.float32)
        else:
            out = np.zeros((out_rows, width, bpp), dtype=np.uint8)
        p = _RenderParams()
        p.img_width, p.img_height = width, height
        p.antialiasing, p.supersampling = int(antialiasing), int(supersampling)
        p.edge_behaviour_x, p.edge_behaviour_y = edge_behaviour
        p.edge_color_x, p.edge_color_y = edge_colors
        p.output_bpp = bpp
        p.frame, p.t = frame, t
        p.num_threads = threads
        p.floatmap = int(floatmap)
        p.num_uservals = n
        p.uservals = arr
        p.output = out.ctypes.data
        if rows_arr is not None:
            p.sample_rows = rows_arr.ctypes.data
            p.num_sample_rows = len(rows_arr)
        rc = self.lib.mmo_render(ctypes.byref(p))
        for d in drawables:
            self.lib.mmo_free_drawable(d)
        if rc != 0:
            raise RuntimeError("oracle render failed (%d)" % rc)
        self.last_taps = p.taps
        return out
