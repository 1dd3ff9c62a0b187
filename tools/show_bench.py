#!/usr/bin/env python
"""Prints a bench.py JSON line (the last line starting with '{' of the given log) as a table."""
import json
import sys


def show(n, v):
    if "error" in v:
        print(n, v["error"])
        return
    r, c, e = v.get("roofline", {}), v.get("cpu_baseline", {}), v.get("e2e", {})
    print("%-10s %9.0f MP/s %8.3f ms  %s %.3f%s  e2e %8.0f  cpuN %6.1f (%s) cpu1 %6.2f  launches %s" % (
        n, v["value"], v["ms_per_step"], r.get("bound"), r.get("frac", 0),
        "  fp64 %.3f" % v["roofline_fp64"]["frac"] if "roofline_fp64" in v else "", e.get("value", 0), c.get("value", 0), c.get("cores"),
        c.get("one_thread", {}).get("value", 0), v.get("gpu_launches")))


def main():
    line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
    d = json.loads(line)
    show(d["config"]["workload"].split()[0].split("/")[-1][:10], d)
    for k, v in d.get("per_workload", {}).items():
        show(k, v)


if __name__ == "__main__":
    main()
