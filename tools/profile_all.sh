#!/bin/sh
# Measurement pass for profiles/: bench lines, ncu launch lists and one `ncu --set full` capture per workload.
# Usage (GPU box, repo root): sh tools/profile_all.sh OUTDIR   -- every bench run exits before its ncu run starts
set -u
OUT=${1:-gpurun_out/final}
mkdir -p "$OUT"
for w in mandelbrot twirl droste droste_nt gauss sea ident invert perlin; do
    python bench.py --workload $w --steps 10 --warmup 3 > "$OUT/bench_$w.log" 2>&1
    tail -1 "$OUT/bench_$w.log" > "$OUT/bench_$w.json"
done
for w in mandelbrot twirl droste gauss sea perlin; do
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/${w}_launches.csv" \
        python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/ncu_launches_$w.log" 2>&1
done
for w in mandelbrot twirl droste sea ident invert perlin; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_kernel -c 1 -s 3 -o "$OUT/prof_$w" -f \
        python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > "$OUT/ncu_full_$w.log" 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gauss_iir -c 2 -s 2 -o "$OUT/prof_gauss" -f \
    python bench.py --workload gauss --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > "$OUT/ncu_full_gauss.log" 2>&1
# A/B of the loop-carried value pass on the headline kernel (one process, same box)
python tools/time_kernel.py "tests/golden/filters/examples/Render/Mandelbrot.mm" 16384 16384 num_iterations=256 --launches 5 --env MMB_LOOP_CARRY=0,1 \
    > "$OUT/loop_carry_ab.json" 2> "$OUT/loop_carry_ab.err"
ls -la "$OUT"
