#!/usr/bin/env python
"""Benchmark of the per-pixel render path (BASELINE.json: megapixels/s per filter, with the
fraction of the FP32 / HBM roofline, next to the reference's CPU path on the host cores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step is one pass of the hot path over one synthetic batch: one frame (or this rank's share
of it / of the frame list) rendered by the generated kernel.  The headline workload is
BASELINE.json configs[1], Render/Mandelbrot at 16384x16384 with 256 iterations.  A default run
(no --workload) measures every other config of BASELINE.json the same way -- twirl, droste (default
uservals and -DNoTransparency=1), gauss, sea -- plus ident / invert (the most memory-bound filters)
and perlin (the libnoise showcase), and carries them in the same JSON line under "per_workload".
`--workload NAME` measures that workload alone (what the ncu captures under profiles/ run).

N > 1 is launched by torchrun, one rank per GPU.  Single-frame workloads split the frame into
8-row blocks interleaved over the ranks (strong scaling, no data-path collective; input
drawables: every rank uploads one band of rows, one NCCL all-gather replicates); sea renders its 240 frames round-robin over the
ranks.  The blur of `gauss` produces a whole-image intermediate and is computed by every rank
("replicas only" for that stage).

Timing: CUDA events on the launching stream (the library launches on the legacy default
stream, which is torch's current stream), W >= 3 warm-up steps, max over ranks.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FILTERS = os.path.join(ROOT, "tests", "golden", "filters", "examples")
IR_DIR = os.path.join(ROOT, "tests", "golden", "ir")

# bytes_px: algorithmic bytes per output pixel; input: does the filter sample a drawable
WORKLOADS = {
    "mandelbrot": dict(script="Render/Mandelbrot.mm", ir="mandelbrot.mmir", w=16384, h=16384, uv={"num_iterations": 256}, aa=False, frames=1,
                       input=False, bytes_px=4, desc="Render/Mandelbrot.mm -s 16384x16384 -Dnum_iterations=256"),
    "twirl": dict(script="Distorts/Twirl.mm", ir="twirl.mmir", w=8192, h=8192, uv={}, aa=True, frames=1, input=True, bytes_px=8,
                  desc="Distorts/Twirl.mm -i, synthetic 8192x8192 RGBA8 input"),
    "droste": dict(script="Map/Droste.mm", ir="droste.mmir", w=8192, h=8192, uv={}, aa=True, frames=1, input=True, bytes_px=8,
                   desc="Map/Droste.mm -i, synthetic 8192x8192 RGBA8 input"),
    "droste_nt": dict(script="Map/Droste.mm", ir="droste.mmir", w=8192, h=8192, uv={"NoTransparency": 1}, aa=True, frames=1, input=True, bytes_px=8,
                      desc="Map/Droste.mm -i -DNoTransparency=1, synthetic 8192x8192 RGBA8 input"),
    "gauss": dict(script="Blur/Gaussian Blur.mm", ir="gauss.mmir", w=8192, h=8192, uv={"dev": 0.0078134}, aa=True, frames=1, input=True, bytes_px=40,
                  bands="contiguous", desc="Blur/Gaussian Blur.mm -i -Ddev=0.0078134 (sigma 32 px), synthetic 8192x8192 RGBA8 input"),
    "sea": dict(script="Distorts/Sea.mm", ir="sea.mmir", w=3840, h=2160, uv={}, aa=True, frames=240, input=True, bytes_px=4,
                desc="Distorts/Sea.mm -i, synthetic 3840x2160 RGBA8 input, 240 frames t=f/240"),
    "ident": dict(script="Utilities/Ident.mm", ir="ident.mmir", w=8192, h=8192, uv={}, aa=True, frames=1, input=True, bytes_px=8,
                  desc="Utilities/Ident.mm -i, synthetic 8192x8192 RGBA8 input"),
    "invert": dict(script="Colors/Invert.mm", ir="invert.mmir", w=8192, h=8192, uv={}, aa=False, frames=1, input=True, bytes_px=8,
                   desc="Colors/Invert.mm (nearest), synthetic 8192x8192 RGBA8 input"),
    "perlin": dict(script="Render/Perlin Noise.mm", ir="perlin.mmir", w=8192, h=8192, uv={}, aa=False, frames=1, input=False, bytes_px=4,
                   desc="Render/Perlin Noise.mm -s 8192x8192 (5 octaves, libnoise in double)"),
}
HEADLINE = "mandelbrot"
# Render/Mandelbrot.mm iterates a QUATERNION square: c*c + p, then |c| < 2.  Distinct float operations per iteration, which is
# what the reference's arithmetic needs and what the kernel's loop executes: 10 products (the 16 of the quaternion product
# hold 6 commutative pairs a*b / b*a; the 4 squares for |c| are the next iteration's a*a, b*b, c*c, d*d, carried in registers
# by the loop-carried value pass, csrc/ir/passes.cpp) + 19 sums (12 + 4 for c*c + p, 3 for |c|^2); the 6 NEGs are operand
# sign modifiers in SASS and the SQRT goes exactly (sqrt(s) < 2 <=> s < 4 for a correctly rounded sqrt).  No FMA credit:
# --fmad=false is required for bit parity.  The loop is 33 SASS instructions: these 29 + counter, two compares, branch, so
# an all-issue-slots-busy kernel reaches 29/33 = 0.879 of this roofline.  (Rounds 1-2 counted the IR's 20 MUL + 19 ADD = 39,
# of which the loop then executed 14 + 19 in 37 instructions; SURVEY.md section 8d's 45 includes the NEGs.)
MANDELBROT_FLOPS_PER_ITERATION = 29
MANDELBROT_LOOP_INSTRUCTIONS = 33
# The blur's recursion (gauss.c:175-196): per step 9 DMUL + 4 DSUB + 5 DADD = 18 double operations, + 1 DADD for vp + vm per
# output sample (half a DADD per sweep step); 2 passes x 2 sweeps x 4 channels steps per pixel.  No FMA (bit parity with the host).
GAUSS_FP64_OPS_PER_PIXEL = 2 * 2 * 4 * 18.5
# Render/Perlin Noise.mm, 5 octaves: per octave 8 lattice corners x 4 products + 3 axes x 9 (7th-order blend) + 7 interpolations x 2
# + 5 (octave bookkeeping) = 78 DMUL, and 42 DADD (mm_noise.cuh; profiles/r02_perlin_sass_hist.txt: 390 DMUL + 210 DADD per pixel executed)
PERLIN_FP64_OPS_PER_PIXEL = 5 * (78 + 42)
B200_SMS, FP32_LANES_PER_SM, FP64_LANES_PER_SM = 148, 128, 64
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def synthetic_input(width, height, seed=1234):
    import numpy as np
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    r = np.hypot(xx - width / 2, yy - height / 2)
    rings = (np.sin(r / 37.0) * 0.5 + 0.5) * 255
    img = np.empty((height, width, 4), dtype=np.uint8)
    noise = rng.integers(0, 256, (height, width, 3), dtype=np.uint8)
    for c in range(3):
        img[:, :, c] = (0.5 * rings + 0.5 * noise[:, :, c]).astype(np.uint8)
    img[:, :, 3] = 255
    return img


def workload_ir(name):
    """The committed optimised IR of a workload (tools/make_golden_ir.py): what the oracle compiles, so that the CPU legs do
    not need the CUDA library."""
    with open(os.path.join(IR_DIR, WORKLOADS[name]["ir"])) as f:
        return f.read()


def ncu_traffic(name):
    """DRAM bytes per launch of the workload's dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) as extracted by
    tools/ncu_traffic.py from the committed `ncu --set full` capture: (bytes or None, provenance string)."""
    try:
        with open(TRAFFIC_FILE, "rb") as f:
            raw = f.read()
        entry = json.loads(raw).get(name)
    except (OSError, ValueError):
        return None, "no profiles/ncu_traffic.json"
    if not entry:
        return None, "no ncu capture of this workload in profiles/ncu_traffic.json"
    return entry["bytes"], "%s kernel %s (profiles/ncu_traffic.json sha256 %s)" % (entry["source"], entry["kernel"], hashlib.sha256(raw).hexdigest()[:12])


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for l in self.lines:
            parts = [p.strip() for p in l.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = the upper half of the samples (idle samples at the edges of the window drag the median down)
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- the CPU path
def oracle_rate(f, width, height, uservals, antialiasing, t, budget_s, threads):
    """Times the oracle on a bounded sample of rows of the frame: (MP/s of a whole frame, sample description,
    pixels actually rendered, seconds of the timed sample).  Frame time = per-frame work (init_frame: frame constants,
    native filters -- single-threaded in the reference too, mathmap_common.c:798-816) + rows x per-row time with
    `threads` band threads (the reference's split, mathmap_common.c:991-1003)."""
    import numpy as np
    rng = np.random.default_rng(7)
    rows_all = rng.permutation(height)  # uniform sample of rows, shuffled so thread chunks are balanced
    # one row, one thread: dominated by the per-frame work when there is any (e.g. the blur of config 4)
    t0 = time.perf_counter()
    f.render(width, height, uservals, t=t, antialiasing=antialiasing, threads=1, sample_rows=rows_all[:1])
    t_one = time.perf_counter() - t0
    frame_const = t_one if t_one > 0.05 else 0.0
    if frame_const >= 0.5:
        # the per-frame work dominates (the single-threaded blur of config 4): the whole frame is rendered once
        t0 = time.perf_counter()
        f.render(width, height, uservals, t=t, antialiasing=antialiasing, threads=threads, sample_rows=rows_all)
        dt = time.perf_counter() - t0
        desc = ("all %d rows of a %dx%d frame on %d thread%s in %.2f s, of which about %.2f s are per-frame work (init_frame: the "
                "native filter, 1 thread, like the reference)" % (height, width, height, threads, "s" * (threads > 1), dt, frame_const))
        return width * height / 1e6 / dt, desc, width * height, dt
    n = min(height, max(threads, 16))
    t0 = time.perf_counter()
    f.render(width, height, uservals, t=t, antialiasing=antialiasing, threads=threads, sample_rows=rows_all[:n])
    per_row = max(1e-7, (time.perf_counter() - t0 - frame_const) / n)
    n2 = int(min(height, max(n, (budget_s - frame_const) / per_row)))
    t0 = time.perf_counter()
    f.render(width, height, uservals, t=t, antialiasing=antialiasing, threads=threads, sample_rows=rows_all[:n2])
    dt = time.perf_counter() - t0
    per_row = max(1e-9, (dt - frame_const) / n2)
    frame_s = frame_const + per_row * height
    desc = ("%d of %d rows of the %dx%d frame (uniform random rows) on %d thread%s in %.2f s; whole frame = %.2f s per-frame "
            "work (init_frame, 1 thread) + %d rows x %.3g s" % (n2, height, width, height, threads, "s" * (threads > 1), dt, frame_const, height, per_row))
    return width * height / 1e6 / frame_s, desc, n2 * width, dt


def cpu_case(name, host_inputs):
    """(width, height, uservals, note) the CPU legs render for a workload.  The blur's CPU time is all per-frame work
    (one thread, about 20 s at 8192^2), so its bounded sample is the same filter at 2048^2 with the same sigma in pixels."""
    wl = WORKLOADS[name]
    W, H, uv = wl["w"], wl["h"], dict(wl["uv"])
    note = ""
    if name == "gauss":
        W = H = 2048
        uv["dev"] = 32.0 / ((W - 1) / 2.0)
        note = " [bounded sample: 2048x2048 frame, same sigma = 32 px; the IIR is linear in the pixel count]"
    if wl["input"]:
        key = (W, H)
        if key not in host_inputs:
            host_inputs[key] = synthetic_input(W, H)
        uv["in"] = host_inputs[key]
    return W, H, uv, note


def cpu_baseline(name, host_inputs, budget_nt, budget_1t):
    """The oracle port on this box's host cores: all online cores with the reference's band split (what the GIMP path does,
    mathmap_common.c:973-1006), and one thread (what the CLI does, mathmap_cmdline.c:844)."""
    from oracle.oracle import OracleFilter
    wl = WORKLOADS[name]
    f = OracleFilter(workload_ir(name))
    W, H, uv, note = cpu_case(name, host_inputs)
    threads = os.cpu_count() or 1
    rn, sn, _, _ = oracle_rate(f, W, H, uv, wl["aa"], 0.0, budget_nt, threads)
    r1, s1, _, _ = oracle_rate(f, W, H, uv, wl["aa"], 0.0, budget_1t, 1)
    return {"value": rn, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sn + note,
            "one_thread": {"value": r1, "unit": "MP/s", "cores": 1, "kind": "port", "sample": s1 + note}}


def workload_config(name, world=1, fast_math=False):
    wl = WORKLOADS[name]
    W, H, frames = wl["w"], wl["h"], wl["frames"]
    cfg = {"workload": wl["desc"], "filter": wl["script"], "width": W, "height": H, "frames_per_step": frames,
           "math": "float libm" if fast_math else "libm evaluated in double and narrowed (parity mode)",
           "sharding": (("one frame, contiguous row bands (mathmap_common.c:997-998): every rank runs the blur's vertical pass over the whole picture "
                         "(replicated work) and its horizontal pass over its own band" if WORKLOADS[name].get("bands") == "contiguous"
                         else "one frame, 8-row blocks interleaved over ranks") if frames == 1 else "frames round-robin over ranks") if world > 1 else "single GPU",
           "l2": "no L2 flush needed: each step writes %d MiB of output%s, larger than the 126 MB L2"
                 % (W * H * 4 * frames >> 20, " and samples a %d MiB input" % (W * H * 4 >> 20) if wl["input"] else "")}
    return cfg


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (the oracle port: the reference cannot be built here) on the
    host cores with all the threads it can use, on the same workload configs; rank 0 only.  Reads the committed IR, not the
    CUDA library."""
    if rank != 0:
        return
    from oracle.oracle import OracleFilter
    threads = os.cpu_count() or 1
    host_inputs = {}
    names = [args.workload] if args.workload else [HEADLINE] + [n for n in WORKLOADS if n != HEADLINE]

    def one(name, steps, warmup, budget):
        wl = WORKLOADS[name]
        f = OracleFilter(workload_ir(name))
        W, H, uv, note = cpu_case(name, host_inputs)
        frames = max(1, wl["frames"])
        rates, times, sample = [], [], ""
        for i in range(steps + warmup):
            r, sample, px, dt = oracle_rate(f, W, H, uv, wl["aa"], (i % frames) / frames, budget, threads)
            if i >= warmup:
                rates.append(r)
                times.append(dt)
        value = sum(rates) / len(rates)
        return {"value": value, "unit": "MP/s", "steps": steps, "warmup": warmup, "ms_per_step": sum(times) / len(times) * 1e3,
                "config": workload_config(name, max(1, args.gpus)), "cpu_baseline": {"value": value, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample + note},
                "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    total = args.steps + args.warmup
    head = one(names[0], args.steps, args.warmup, max(1.0, min(20.0, 100.0 / max(1, total))))
    line = {"impl": "reference", "metric": "megapixels_per_sec", "value": head["value"], "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": head["config"],
            "note": "CPU path on host cores; each step is a bounded row sample of the frame; value is the whole-frame rate it implies, ms_per_step the time of the sample itself",
            "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"]}
    if len(names) > 1:
        line["per_workload"] = {n: one(n, 1, 0, 4.0) for n in names[1:]}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- the GPU path
class Context:
    pass


def shared_host_frame(ctx, tag, nbytes):
    """One pinned host frame shared by all ranks of the box (POSIX shared memory registered with CUDA by every rank): each
    rank's device->host copies land in its rows of the SAME image, so that at N > 1 the end-to-end result is one assembled
    frame in host memory, like the reference's threads writing their bands into one output buffer (mathmap_common.c:991-1003)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    path = "/dev/shm/mathmap_b200_%s_%s" % (os.environ.get("MASTER_PORT", "0"), tag)
    if ctx.rank == 0:
        with open(path, "wb") as f:
            f.truncate(nbytes)
    if ctx.world > 1:
        dist.barrier()
    arr = np.memmap(path, dtype=np.uint8, mode="r+", shape=(nbytes,))
    torch.cuda.check_error(torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, nbytes, 0))
    if ctx.world > 1:
        dist.barrier()
    if ctx.rank == 0:
        os.unlink(path)  # the mappings keep it alive
    return arr


def release_host_frame(arr):
    import torch
    torch.cuda.cudart().cudaHostUnregister(arr.ctypes.data)


def measure(name, ctx, args, with_cpu):
    import numpy as np
    import torch
    import torch.distributed as dist
    import mathmap_b200 as mb
    from mathmap_b200 import sharding

    wl = WORKLOADS[name]
    W, H, frames, aa = wl["w"], wl["h"], wl["frames"], wl["aa"]
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    m = mb.Module.from_file(os.path.join(FILTERS, wl["script"]))
    inv = mb.Invocation(m, W, H, device=ctx.local_rank, antialiasing=aa, precise=not args.fast_math, warp_width=args.warp_width, rows_per_thread=args.rows)
    for k, v in wl["uv"].items():
        inv.set(k, v)
    h2d_bytes = 0
    host_input = d_in = None
    in_band = sharding.band_for_rank(0, H, rank, world)
    if wl["input"]:
        # every rank can read the (synthetic, seeded) host image: it pins and uploads its own band of rows, and one all-gather
        # over NVLink replicates the drawable (samplers read arbitrary coordinates); at N = 1 that is the whole image
        d_in = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        key = (W, H)
        if key not in ctx.host_inputs:
            ctx.host_inputs[key] = synthetic_input(W, H)
        if key not in ctx.pinned_inputs:
            ctx.pinned_inputs[key] = torch.from_numpy(np.ascontiguousarray(ctx.host_inputs[key][in_band[0]:in_band[1]])).pin_memory()
        host_input = ctx.pinned_inputs[key]
        d_in[in_band[0]:in_band[1]].copy_(host_input)
        sharding.replicate_drawable_bands(d_in)
        inv.set("in", d_in)
        h2d_bytes = W * H * 4

    # this rank's share of a step
    if frames == 1:
        contiguous = wl.get("bands") == "contiguous"  # uniform cost per row: the reference's own band split
        band = sharding.band_for_rank(0, H, rank, world)
        my_rows = band[1] - band[0] if contiguous else len(sharding.interleaved_rows_for_rank(H, rank, world))
        out = torch.empty((max(1, (my_rows + 7) // 8 * 8), W, 4), dtype=torch.uint8, device=dev)
        my_frames = [0]
        pixels_per_step_all = W * H

        def step(i):
            inv.init_frame(0, 0.0)
            if world == 1:
                inv.calc_lines_device(out.data_ptr(), 0, H)
            elif contiguous:
                inv.calc_lines_device(out.data_ptr(), band[0], band[1])
            else:
                inv.calc_lines_interleaved_device(out.data_ptr(), rank, world)
    else:
        my_frames = sharding.frames_for_rank(frames, rank, world)
        out = torch.empty((len(my_frames), H, W, 4), dtype=torch.uint8, device=dev)  # every frame of this rank is kept
        pixels_per_step_all = W * H * frames
        my_ts = [f / frames for f in my_frames]  # t = frame / num_frames, mathmap_cmdline.c:835

        def step(i):
            # the batched entry point: one C call renders this rank's frames into consecutive device buffers
            inv.render_frames_device(out.data_ptr(), my_ts, my_frames)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(ctx.local_rank).start()
    # warm-up: W steps, and at least ~0.5 s of the same work so clocks settle and nvidia-smi gets samples under load
    t_w = time.perf_counter()
    i = 0
    while i < args.warmup or (time.perf_counter() - t_w < 0.5 and i < 2000):
        step(i)
        i += 1
        if i >= args.warmup:
            torch.cuda.synchronize()
    barrier()
    launches0 = inv.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = inv.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    # device time of the K steps on this rank: from the first start event to the last end event
    total_ms = ev[0][0].elapsed_time(ev[-1][1])
    total_ms = sharding.max_over_ranks(total_ms, dev)
    value = pixels_per_step_all * args.steps / (total_ms / 1e3) / 1e6
    kernel_ms = sum(step_ms) / len(step_ms)

    res = {"value": value, "unit": "MP/s", "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
           "config": workload_config(name, world, args.fast_math), "kernel": inv.kernel_name, "gpu_launches": launches, "clocks": clocks, "wall_s": wall}

    # ---- roofline of the dominant kernel(s), from the live event times of this rank (rank 0 reports)
    peaks = ctx.peaks
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    launches_per_step = max(1, launches // args.steps)
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    traffic, traffic_src = ncu_traffic(name)
    if name == "mandelbrot":
        # iterations per pixel are recoverable from the output: gray = trunc(iter/256*255) = iter - 1 for iter >= 1
        my_iters = int((out[:my_rows, :, 0].to(torch.int64) + 1).sum().item())
        peak_tflops = B200_SMS * FP32_LANES_PER_SM * sm_mhz * 1e6 / 1e12
        achieved = MANDELBROT_FLOPS_PER_ITERATION * my_iters / (kernel_ms / 1e3) / 1e12
        res["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                           # the same time counted in issue slots: the loop's 33 instructions per iteration (tests/test_loop_carried.py
                           # reads them from the cubin) against one warp instruction per SM sub-partition and clock
                           "issue_slots_frac": achieved / peak_tflops * MANDELBROT_LOOP_INSTRUCTIONS / MANDELBROT_FLOPS_PER_ITERATION,
                           "traffic": traffic,
                           "note": "non-FMA FP32 issue roofline: 148 SMs x 128 lanes x %.0f MHz (median SM clock under load); "
                                   "29 distinct float operations (10 MUL + 19 ADD) per iteration of the quaternion loop, 33 SASS instructions with the loop control (issue-slot ceiling 29/33 = 0.879) x %d iterations per launch on this rank; output writes are %.1f GB/s of the %s %.0f GB/s HBM peak; "
                                   "traffic = DRAM bytes per launch, %s (algorithmic: %.3e output bytes per rank)"
                                   % (sm_mhz, my_iters, my_rows * W * 4 / (kernel_ms / 1e3) / 1e9, hbm_src, hbm_peak, traffic_src, my_rows * W * 4)}
    else:
        bytes_px = wl["bytes_px"]
        if frames > 1:
            px = W * H * len(my_frames)
            alg = bytes_px * px + W * H * 4  # outputs of this rank's frames + the input once
        elif name == "gauss":
            # every rank blurs the whole image (36 B/px: u8 in, f32x4 intermediate out and in) and quantises its rows (4 B/px)
            alg = 36 * W * H + 4 * W * my_rows
        else:
            alg = bytes_px * W * my_rows
        achieved = alg / (kernel_ms / 1e3) / 1e9
        res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                           "note": "%d algorithmic bytes/pixel (SURVEY.md section 8d), %.4g bytes per step on this rank; peak is %s; %d launches per step; traffic: %s"
                                   % (bytes_px, alg, hbm_src, launches_per_step, traffic_src)}
        if name == "gauss":
            ops = GAUSS_FP64_OPS_PER_PIXEL * W * H
            peak64 = B200_SMS * FP64_LANES_PER_SM * sm_mhz * 1e6 / 1e12
            ach64 = ops / (kernel_ms / 1e3) / 1e12
            res["roofline_fp64"] = {"bound": "fp64", "achieved": ach64, "peak": peak64, "unit": "TFLOP/s", "frac": ach64 / peak64,
                                    "note": "non-FMA FP64 issue roofline: 148 SMs x 64 lanes x %.0f MHz; 18.5 double operations per recursion step x "
                                            "2 passes x 2 sweeps x 4 channels per pixel (gauss.c:175-196; recomputed steps are not counted); the exact recursion "
                                            "makes this the binding roofline, the HBM one above is what SURVEY.md section 8d asks for" % sm_mhz}
        if name == "perlin":
            ops = PERLIN_FP64_OPS_PER_PIXEL * W * my_rows
            peak64 = B200_SMS * FP64_LANES_PER_SM * sm_mhz * 1e6 / 1e12
            ach64 = ops / (kernel_ms / 1e3) / 1e12
            res["roofline_fp64"] = {"bound": "fp64", "achieved": ach64, "peak": peak64, "unit": "TFLOP/s", "frac": ach64 / peak64,
                                    "note": "non-FMA FP64 issue roofline: 148 SMs x 64 lanes x %.0f MHz; 5 octaves x (78 products + 42 sums) in double per pixel "
                                            "(libnoise 1.0.0 gradient noise with the 7th-order blend: csrc/runtime/mm_noise.cuh; the executed DMUL / DADD counts of "
                                            "profiles/r02_perlin_sass_hist.txt are exactly these, nothing is recomputed); compares and int<->double conversions, which "
                                            "share the pipe, are not counted" % sm_mhz}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region:
    # H2D of the step's input (at N > 1 one band of rows per rank, then one NCCL all-gather), mmb_init_frame, mmb_calc_lines into host memory
    if not args.no_e2e:
        e2e_steps = args.steps
        frame_bytes = W * H * 4

        def upload():
            if d_in is None:
                return
            if world == 1:
                inv.set("in", host_input.numpy())  # H2D of the step's input through the public API
            else:
                d_in[in_band[0]:in_band[1]].copy_(host_input, non_blocking=True)
                sharding.replicate_drawable_bands(d_in)
                torch.cuda.current_stream().synchronize()
                inv.set("in", d_in)

        if frames == 1:
            rows = sharding.band_for_rank(0, H, rank, world)
            shared = shared_host_frame(ctx, name, frame_bytes)
            frame_arr = shared.reshape(H, W, 4)
            arr = frame_arr[rows[0]:rows[1]]

            def e2e_step():
                upload()
                inv.init_frame(0, 0.0)
                inv.calc_lines(rows[0], rows[1], out=arr)
            px_e2e = W * H
            d2h = frame_bytes
            note = ("mmb_set_userval_image_host (N > 1: each rank uploads its band of the input, one NCCL all-gather) + mmb_init_frame + mmb_calc_lines; contiguous bands per rank, every rank "
                    "copies its band into ONE pinned host frame shared by the ranks")
        else:
            shared = None
            host_out = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            arr = host_out.numpy()

            def e2e_step():
                upload()
                for f in my_frames:
                    inv.init_frame(f, f / frames)
                    inv.calc_lines(0, H, out=arr)
            px_e2e = W * H * frames
            d2h = frame_bytes * frames
            e2e_steps = min(e2e_steps, 2)
            note = "input uploaded once per step; mmb_init_frame + mmb_calc_lines per frame into pinned host memory, frames round-robin over ranks"
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step()
        barrier()
        dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
        res["e2e"] = {"value": px_e2e * e2e_steps / dt / 1e6, "unit": "MP/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                      "steps": e2e_steps, "note": note}
        # What the box's device->host path takes at all: every rank copies as many bytes as its share of a step's result from
        # device memory into the same pinned host memory with plain cudaMemcpyAsync, all ranks at once.  The end-to-end rate
        # cannot exceed this (nothing of the library runs here); at N = 8 on this pool it is the binding limit.
        try:
            share = d2h // world if frames == 1 else frame_bytes
            dst = torch.from_numpy(arr.reshape(-1)[:share]) if frames == 1 else torch.from_numpy(arr.reshape(-1))
            src = torch.empty(dst.numel(), dtype=torch.uint8, device=dev)
            reps = 3 if frames == 1 else 3 * max(1, len(my_frames))
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            barrier()
            dtc = sharding.max_over_ranks(time.perf_counter() - t0, dev)
            ceiling_bytes_s = dst.numel() * reps * world / dtc
            res["e2e"]["d2h_ceiling"] = {"value": ceiling_bytes_s / 1e9, "unit": "GB/s", "as_metric": ceiling_bytes_s / 4 / 1e6, "metric_unit": "MP/s",
                                         "note": "aggregate pinned cudaMemcpyAsync device->host of all ranks at once, same destination buffers"}
            del src, dst
        except Exception as e:  # the probe must never cost the measurement
            res["e2e"]["d2h_ceiling"] = {"error": str(e)[:200]}
        if shared is not None:
            if rank == 0 and frames == 1:
                # the assembled frame: rows of every rank's band are there (checked against this rank's device rows where it has them)
                probe = [sharding.band_for_rank(0, H, r, world)[0] for r in range(world)]
                res["e2e"]["assembled_rows_nonzero"] = bool(all(frame_arr[p].any() for p in probe)) if name != "mandelbrot" else bool(frame_arr[:, :, 3].min() == 255)
            barrier()
            release_host_frame(shared)
            del arr, frame_arr, shared

    # ---- the reference's CPU path on this box's host cores, bounded sample (rank 0, N = 1 only)
    if with_cpu:
        head = name == HEADLINE or args.workload
        res["cpu_baseline"] = cpu_baseline(name, ctx.host_inputs, 15.0 if head else 4.0, 8.0 if head else 3.0)
    del inv, out, d_in
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS), help="measure this workload alone (default: the headline + all others under per_workload)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fast-math", action="store_true", help="CUDA float libm instead of double-evaluated libm (parity mode is the default)")
    ap.add_argument("--warp-width", type=int, default=None)
    ap.add_argument("--rows", type=int, default=None, help="32x8 tiles one block renders in sequence (mmb_set_rows_per_thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--headline-only", action="store_true", help="skip per_workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    ctx = Context()
    ctx.rank, ctx.world, ctx.local_rank = rank, world, local_rank
    ctx.dev = torch.device("cuda", local_rank)
    ctx.host_inputs, ctx.pinned_inputs = {}, {}
    ctx.peaks = {}
    try:
        ctx.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    if world > 1:
        # stdout carries the one JSON line: while the communicator comes up, file descriptor 1 points at stderr, so that NCCL's
        # own banner ("NCCL version ..." at NCCL_DEBUG=VERSION / WARN, written by the C library) does not land in front of it
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=ctx.dev)
            probe = torch.zeros(1, device=ctx.dev)
            dist.all_reduce(probe)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    with_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    head_name = args.workload or HEADLINE
    head = measure(head_name, ctx, args, with_cpu)
    line = {"metric": "megapixels_per_sec", "value": head["value"], "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    for k in ("config", "kernel", "gpu_launches", "clocks", "wall_s", "roofline", "roofline_fp64", "e2e", "cpu_baseline"):
        if k in head:
            line[k] = head[k]
    if not args.workload and not args.headline_only:
        per = {}
        for name in WORKLOADS:
            if name == HEADLINE:
                continue
            ok = 1
            try:
                per[name] = measure(name, ctx, args, with_cpu)
            except Exception as e:  # one workload failing must not lose the headline line; the failure is reported in its place
                per[name] = {"error": "%s: %s" % (type(e).__name__, e)}
                ok = 0
            if world > 1:  # the ranks agree on the outcome, so that a failure on one of them is not reported as a number by rank 0
                flag = torch.tensor([ok], device=ctx.dev, dtype=torch.int32)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if ok and int(flag.item()) == 0:
                    per[name] = {"error": "failed on another rank"}
        line["per_workload"] = per

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
