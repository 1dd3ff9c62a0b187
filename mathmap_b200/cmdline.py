"""Command line front end with the reference's flags (mathmap_cmdline.c:425-521, 595-871):

    python -m mathmap_b200.cmdline [-i] [-o] [-s WxH] [-f SCRIPT.mm | EXPRESSION] [-D name=value ...]
                                   [-F FRAMES] [--bench-render-count N] [--bench-no-output] [--bench-only-compile]
                                   [--bench-no-compile-time-limit] [--device N] [--fast-math] OUTFILE

Behaviour kept from the reference: `-D` binds int/float/bool arguments by value and image arguments by file
name (mathmap_cmdline.c:772-795); images are read as RGB with alpha forced to 255 (rwimg, mathmap_cmdline.c:181-183);
without `-s` the output takes the size of the first image argument (:717-752); the PNG written is RGB only
(rwimg/rwpng.c:172-242); `-F n` renders n frames with t = frame / n (:835) into OUTFILE with a %d pattern or a
numeric suffix.  `-c` (input cache size) is accepted and ignored; `-g` (plug-in generators), `-M` (QuickTime input)
and `--htmldoc` are outside the render path.  Rendering happens on the GPU; there is no CPU path.
"""
import argparse
import os
import sys
import time

import numpy as np


def read_image_rgba(path):
    from PIL import Image
    rgb = np.array(Image.open(path).convert("RGB"))
    return np.ascontiguousarray(np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)]))


def write_image(path, rgba):
    from PIL import Image
    Image.fromarray(rgba[:, :, :3], "RGB").save(path)


def main(argv=None):
    import mathmap_b200 as mb
    ap = argparse.ArgumentParser(prog="mathmap", add_help=True)
    ap.add_argument("-f", "--script-file")
    ap.add_argument("-i", "--intersampling", action="store_true")
    ap.add_argument("-o", "--oversampling", action="store_true")
    ap.add_argument("-s", "--size")
    ap.add_argument("-c", "--cache", type=int, default=8)
    ap.add_argument("-D", action="append", default=[], metavar="name=value")
    ap.add_argument("-F", "--frames", type=int, default=1)
    ap.add_argument("--bench-render-count", type=int, default=1)
    ap.add_argument("--bench-only-compile", action="store_true")
    ap.add_argument("--bench-no-output", action="store_true")
    ap.add_argument("--bench-no-compile-time-limit", action="store_true")  # always the case here: no wall-clock pass budget
    ap.add_argument("--bench-no-backend", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--fast-math", action="store_true")
    ap.add_argument("--fast-compile", action="store_true", help="shorter filter compile, a few percent slower kernel (mmb_set_fast_compile)")
    ap.add_argument("--filter-path", default=None, help="directory tree searched for the node types of a .mmc composition")
    ap.add_argument("--version", action="store_true")
    ap.add_argument("rest", nargs="*")
    a = ap.parse_args(argv)
    if a.version:
        print(mb.lib().mmb_version().decode())
        return 0
    if a.script_file:
        if len(a.rest) != 1:
            ap.error("expected OUTFILE")
        source, outfile = open(a.script_file).read(), a.rest[0]
    else:
        if len(a.rest) != 2:
            ap.error("expected EXPRESSION OUTFILE")
        source, outfile = a.rest
    t0 = time.perf_counter()
    try:
        if a.script_file and a.script_file.endswith(".mmc"):  # a composition (not on the reference's command line; GIMP-side there)
            module = mb.Module.from_file(a.script_file, filter_path=a.filter_path)
        else:
            module = mb.Module(source=source)
    except mb.MathMapError as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    if a.bench_no_backend or a.bench_only_compile:
        if not a.bench_no_backend:
            module.compile_check(a.intersampling, not a.fast_math, a.fast_compile)
        print("compiled %s in %.3f s" % (module.name, time.perf_counter() - t0), file=sys.stderr)
        return 0

    infos = {u[0]: u for u in module.uservals()}
    defines = {}
    for d in a.D:
        if "=" not in d:
            ap.error("-D needs name=value")
        name, value = d.split("=", 1)
        if name not in infos:
            print("Error: filter %s has no argument `%s'." % (module.name, name), file=sys.stderr)
            return 1
        typ = infos[name][1]
        if typ == mb.USERVAL_INT:
            defines[name] = int(value)
        elif typ == mb.USERVAL_FLOAT:
            defines[name] = float(value)
        elif typ == mb.USERVAL_BOOL:
            defines[name] = int(value) != 0
        elif typ == mb.USERVAL_IMAGE:
            defines[name] = read_image_rgba(value)
        else:
            print("Error: can only define int, float, bool and image arguments on the command line.", file=sys.stderr)
            return 1
    size = None
    if a.size:
        w, h = a.size.lower().split("x")
        size = (int(w), int(h))
    else:
        for u in module.uservals():
            if u[1] == mb.USERVAL_IMAGE and u[0] in defines:
                size = (defines[u[0]].shape[1], defines[u[0]].shape[0])
                break
    if size is None:
        print("Error: image size not set and no input images given.", file=sys.stderr)
        return 1
    try:
        inv = mb.Invocation(module, size[0], size[1], device=a.device, antialiasing=a.intersampling, supersampling=a.oversampling,
                            precise=not a.fast_math, fast_compile=a.fast_compile)
        for k, v in defines.items():
            inv.set(k, v)
        for _ in range(max(1, a.bench_render_count)):
            for frame in range(a.frames):
                out = inv.render(frame, frame / float(a.frames))
                if not a.bench_no_output:
                    if a.frames > 1:
                        root, ext = os.path.splitext(outfile)
                        name = outfile % frame if "%" in outfile else "%s_%05d%s" % (root, frame, ext)
                    else:
                        name = outfile
                    write_image(name, out)
    except mb.MathMapError as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
