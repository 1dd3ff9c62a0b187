#!/usr/bin/env python
"""Benchmark of the per-pixel render path (BASELINE.json: megapixels/s per filter, with the
fraction of the FP32 / HBM roofline, next to the reference's CPU path on the host cores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step is one pass of the hot path over one synthetic batch: one frame (or this rank's share
of it / of the frame list) rendered by the generated kernel.  The default workload is
BASELINE.json configs[1], Render/Mandelbrot at 16384x16384 with 256 iterations; the others
(`--workload twirl|droste|gauss|sea`) are the remaining configs, measured the same way.

N > 1 is launched by torchrun, one rank per GPU.  Mandelbrot/twirl/droste/gauss split one
frame into 8-row blocks interleaved over the ranks (strong scaling, no data-path collective;
input drawables are replicated by one NCCL broadcast before the timed region); sea renders
its 240 frames round-robin over the ranks.

Timing: CUDA events on the launching stream (the library launches on the legacy default
stream, which is torch's current stream), W >= 3 warm-up steps, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FILTERS = os.path.join(ROOT, "tests", "golden", "filters", "examples")

WORKLOADS = {
    # name: (script, width, height, uservals, antialiasing, frames per step, description)
    "mandelbrot": ("Render/Mandelbrot.mm", 16384, 16384, {"num_iterations": 256}, False, 1,
                   "Render/Mandelbrot.mm -s 16384x16384 -Dnum_iterations=256"),
    "twirl": ("Distorts/Twirl.mm", 8192, 8192, {}, True, 1, "Distorts/Twirl.mm -i, synthetic 8192x8192 RGBA8 input"),
    "droste": ("Map/Droste.mm", 8192, 8192, {}, True, 1, "Map/Droste.mm -i, synthetic 8192x8192 RGBA8 input"),
    "gauss": ("Blur/Gaussian Blur.mm", 8192, 8192, {"dev": 0.0078134}, True, 1,
              "Blur/Gaussian Blur.mm -i -Ddev=0.0078134 (sigma 32 px), synthetic 8192x8192 RGBA8 input"),
    "ident": ("Utilities/Ident.mm", 8192, 8192, {}, True, 1, "Utilities/Ident.mm -i, synthetic 8192x8192 RGBA8 input"),
    "invert": ("Colors/Invert.mm", 8192, 8192, {}, False, 1, "Colors/Invert.mm (nearest), synthetic 8192x8192 RGBA8 input"),
    "sea": ("Distorts/Sea.mm", 3840, 2160, {}, True, 240, "Distorts/Sea.mm -i, synthetic 3840x2160 RGBA8 input, 240 frames t=f/240"),
}
# Optimised IR per loop iteration: 20 MUL + 19 ADD (+ 6 NEG, which are operand sign modifiers in SASS, not instructions,
# + 1 SQRT that the emitter removes exactly: sqrt(s) < 2 <=> s < 4 for correctly rounded sqrt).  No FMA credit:
# --fmad=false is required for bit parity.  SURVEY.md section 8d counts 45 (with the NEGs); 39 is the instruction-level figure.
MANDELBROT_FLOPS_PER_ITERATION = 39
# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu --set full captures, profiles/r01_*_ncu_full.txt
# (gauss: the dominant kernel of its four launches, the row pass of the IIR)
NCU_TRAFFIC_BYTES = {"mandelbrot": 7.99e6 + 1.0168e9, "twirl": 241.9e6 + 228.7e6, "droste": 211.2e6 + 228.8e6, "gauss": 4.258e9 + 3.196e9,
                     "sea": 32.6e6 + 0.9e6, "ident": 268.6e6 + 232.3e6}
B200_SMS, FP32_LANES_PER_SM = 148, 128


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


def oracle_rate(ir, width, height, uservals, antialiasing, t, budget_s, threads):
    """Times the oracle on a bounded sample of rows of the same frame and extrapolates to the whole frame:
    frame time = per-frame work (init_frame: frame constants, native filters -- single-threaded in the
    reference too, mathmap_common.c:798-816) + rows x per-row time with `threads` band threads.
    Returns (MP/s of a whole frame, description, seconds spent)."""
    import numpy as np
    from oracle.oracle import OracleFilter
    f = OracleFilter(ir)
    rng = np.random.default_rng(7)
    rows_all = rng.permutation(height)  # uniform sample of rows, shuffled so thread chunks are balanced
    spent = time.perf_counter()
    # one row, one thread: dominated by the per-frame work when there is any (e.g. the blur of config 4)
    t0 = time.perf_counter()
    f.render(width, height, uservals, t=t, antialiasing=antialiasing, threads=1, sample_rows=rows_all[:1])
    t_one = time.perf_counter() - t0
    frame_const = t_one if t_one > 0.05 else 0.0
    if frame_const >= 1.0:
        # the per-frame work dominates (the single-threaded blur of config 4): a difference of two such timings says
        # nothing about the rows, so the whole frame is rendered once and its time taken as it is
        t0 = time.perf_counter()
        f.render(width, height, uservals, t=t, antialiasing=antialiasing, threads=threads, sample_rows=rows_all)
        dt = time.perf_counter() - t0
        spent = time.perf_counter() - spent
        desc = ("all %d rows of the %dx%d frame on %d threads in %.1f s, of which about %.1f s are per-frame work (init_frame: the "
                "native filter, 1 thread, like the reference)" % (height, width, height, threads, dt, frame_const))
        return width * height / 1e6 / dt, desc, spent
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
    spent = time.perf_counter() - spent
    desc = ("%d of %d rows of the %dx%d frame (uniform random rows) on %d threads in %.1f s; whole frame extrapolated as "
            "%.2f s per-frame work (init_frame, 1 thread) + %d rows x %.3g s" % (n2, height, width, height, threads, dt, frame_const, height, per_row))
    return width * height / 1e6 / frame_s, desc, spent


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (the oracle port: the reference cannot be built here,
    DESIGN.md) on the host cores, on the same workload config; rank 0 only."""
    if rank != 0:
        return
    import mathmap_b200 as mb
    script, W, H, uv, aa, frames, desc = WORKLOADS[args.workload]
    m = mb.Module.from_file(os.path.join(FILTERS, script))
    uservals = dict(uv)
    if args.workload != "mandelbrot":
        uservals["in"] = synthetic_input(W, H)
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    budget = max(2.0, min(20.0, 150.0 / max(1, total)))
    rates = []
    sample = ""
    for i in range(total):
        r, sample, _ = oracle_rate(m.ir, W, H, uservals, aa, (i % max(1, frames)) / max(1, frames), budget, threads)
        if i >= args.warmup:
            rates.append(r)
    value = sum(rates) / len(rates)
    line = {"impl": "reference", "metric": "megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": W * H * max(1, frames) / (value * 1e6) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "note": "CPU path on host cores; each step is a bounded row sample of the frame; ms_per_step is the whole step (all pixels) extrapolated from it"},
            "cpu_baseline": {"value": value, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="mandelbrot", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--fast-math", action="store_true", help="CUDA float libm instead of double-evaluated libm (parity mode is the default)")
    ap.add_argument("--warp-width", type=int, default=None)
    ap.add_argument("--rows", type=int, default=None, help="32x8 tiles one block renders in sequence (mmb_set_rows_per_thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import mathmap_b200 as mb
    from mathmap_b200 import sharding

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    script, W, H, uv, aa, frames, desc = WORKLOADS[args.workload]
    m = mb.Module.from_file(os.path.join(FILTERS, script))
    inv = mb.Invocation(m, W, H, device=local_rank, antialiasing=aa, precise=not args.fast_math, warp_width=args.warp_width, rows_per_thread=args.rows)
    for k, v in uv.items():
        inv.set(k, v)
    h2d_bytes = 0
    host_input = None
    if args.workload != "mandelbrot":
        # rank 0 makes the drawable; one NCCL broadcast replicates it (samplers read arbitrary coordinates)
        d_in = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
        if rank == 0:
            host_input = torch.from_numpy(synthetic_input(W, H)).pin_memory()
            d_in.copy_(host_input)
        sharding.broadcast_drawable(d_in, src=0)
        inv.set("in", d_in)
        h2d_bytes = W * H * 4

    # this rank's share of a step
    if frames == 1:
        my_rows = len(sharding.interleaved_rows_for_rank(H, rank, world))
        out = torch.empty((max(1, (my_rows + 7) // 8 * 8), W, 4), dtype=torch.uint8, device=dev)
        my_frames = [0]
        pixels_per_step_all = W * H

        def step(i):
            inv.init_frame(0, 0.0)
            if world == 1:
                inv.calc_lines_device(out.data_ptr(), 0, H)
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

    sampler = ClockSampler(local_rank)
    sampler.start()
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

    line = {"metric": "megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "filter": script, "width": W, "height": H, "frames_per_step": frames,
                       "math": "float libm" if args.fast_math else "libm evaluated in double and narrowed (parity mode)",
                       "sharding": ("one frame, 8-row blocks interleaved over ranks" if frames == 1 else "frames round-robin over ranks") if world > 1 else "single GPU",
                       "l2": "no L2 flush needed: each step writes %d MiB of output%s, larger than the 126 MB L2"
                             % (W * H * 4 * (1 if frames == 1 else 1) >> 20, "" if args.workload == "mandelbrot" else " and samples a %d MiB input" % (W * H * 4 >> 20)),
                       "kernel": inv.kernel_name},
            "gpu_launches": launches, "clocks": clocks, "wall_s": wall}

    if rank == 0:
        # ---- roofline of the dominant kernel, from the live event times
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        launches_per_step = max(1, launches // args.steps)
        if args.workload == "mandelbrot":
            # iterations per pixel are recoverable from the output: gray = trunc(iter/256*255) = iter - 1 for iter >= 1
            if world == 1:
                iters = int((out[:H, :, 0].to(torch.int64) + 1).sum().item())
            else:
                iters = None
            sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
            peak_tflops = B200_SMS * FP32_LANES_PER_SM * sm_mhz * 1e6 / 1e12
            if iters is not None:
                achieved = MANDELBROT_FLOPS_PER_ITERATION * iters / (kernel_ms / 1e3) / 1e12
                line["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                                    "traffic": NCU_TRAFFIC_BYTES.get(args.workload),
                                    "note": "non-FMA FP32 issue roofline: 148 SMs x 128 lanes x %.0f MHz (median SM clock under load); "
                                            "39 flops (20 MUL + 19 ADD) per iteration x %d iterations per launch; output writes are %.1f GB/s of the %s %.0f GB/s HBM peak; "
                                            "traffic = DRAM bytes per launch from profiles/r01_mandelbrot_ncu_full.txt (algorithmic: 1.074e9 output bytes)"
                                            % (sm_mhz, iters, W * H * 4 / (kernel_ms / 1e3) / 1e9, hbm_src, hbm_peak)}
        else:
            bytes_per_px = {"twirl": 8, "droste": 8, "gauss": 40, "sea": 4, "ident": 8, "invert": 8}[args.workload]
            px = W * H * (len(my_frames) if frames > 1 else 1) / (1 if frames > 1 else world)
            achieved = bytes_per_px * px / (kernel_ms / 1e3) / 1e9
            line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                                "traffic": NCU_TRAFFIC_BYTES.get(args.workload),
                                "note": "%d algorithmic bytes/pixel (SURVEY.md section 8d); peak is %s; %d launches per step" % (bytes_per_px, hbm_src, launches_per_step)}

    # ---- end to end through the C ABI with host buffers (pinned), copies inside the timed region
    if not args.no_e2e:
        e2e_steps = args.steps
        if frames == 1:
            rows = sharding.band_for_rank(0, H, rank, world)
            host_out = torch.empty((rows[1] - rows[0], W, 4), dtype=torch.uint8).pin_memory()
            arr = host_out.numpy()

            def e2e_step():
                if host_input is not None:
                    inv.set("in", host_input.numpy())  # H2D of the step's input through the public API
                inv.init_frame(0, 0.0)
                inv.calc_lines(rows[0], rows[1], out=arr)
            px_e2e = W * H
            d2h = W * H * 4
        else:
            host_out = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            arr = host_out.numpy()

            def e2e_step():
                if host_input is not None:
                    inv.set("in", host_input.numpy())
                for f in my_frames:
                    inv.init_frame(f, f / frames)
                    inv.calc_lines(0, H, out=arr)
            px_e2e = W * H * frames
            d2h = W * H * 4 * frames
            e2e_steps = min(e2e_steps, 2)
        if host_input is None and args.workload != "mandelbrot":
            host_input = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory()
            host_input.copy_(d_in)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step()
        barrier()
        dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
        line["e2e"] = {"value": px_e2e * e2e_steps / dt / 1e6, "unit": "MP/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
                       "steps": e2e_steps, "note": "mmb_init_frame + mmb_calc_lines into pinned host memory; contiguous bands per rank"}

    # ---- the reference's CPU path on this box's host cores, bounded sample (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        uservals = dict(uv)
        if args.workload != "mandelbrot":
            uservals["in"] = host_input.numpy() if host_input is not None else synthetic_input(W, H)
        threads = os.cpu_count() or 1
        r, sample, _ = oracle_rate(m.ir, W, H, uservals, aa, 0.0, 15.0, threads)
        line["cpu_baseline"] = {"value": r, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
