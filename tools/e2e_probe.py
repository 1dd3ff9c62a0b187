"""Timing probe for the host-facing path (16384x16384 Mandelbrot): whole-frame kernel, chunked kernels, a bare 1 GiB device->host
copy and mmb_calc_lines into pinned memory.  Usage (GPU box): python tools/e2e_probe.py"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import mathmap_b200 as mb
src = open("tests/golden/filters/examples/Render/Mandelbrot.mm").read()
W = H = 16384
m = mb.Module(source=src)
inv = mb.Invocation(m, W, H, antialiasing=False)
inv.set("num_iterations", 256) if "num_iterations" in dict((u[0], 1) for u in m.uservals()) else None
dev = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1000 / n
inv.init_frame(0, 0.0)
print("kernel whole frame (device out):", ev(lambda: (inv.calc_lines_device(dev.data_ptr()), inv.synchronize())))
def chunks(k):
    rows = H // k
    for i in range(k):
        inv.calc_lines_device(dev.data_ptr() + i * rows * W * 4, i * rows, (i + 1) * rows)
    inv.synchronize()
for k in (1, 8, 32):
    print("kernel in %d chunks:" % k, ev(lambda: chunks(k)))
print("D2H 1 GiB pinned:", ev(lambda: host.copy_(dev, non_blocking=True)))
out = host.numpy()
print("calc_lines to pinned host:", ev(lambda: inv.calc_lines(0, H, out=out)))
