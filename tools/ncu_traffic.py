#!/usr/bin/env python
"""Extracts the DRAM traffic of each workload's dominant kernel from the committed ncu summaries
(profiles/rNN_<workload>_ncu_full.txt, written by tools/ncu_summary.py from one `ncu --set full` capture)
into profiles/ncu_traffic.json, which bench.py reads for `roofline.traffic`.  The newest round's summary
of a workload wins; the dominant kernel is the one with the largest gpu__time_duration.sum in the file.

    python tools/ncu_traffic.py
"""
import glob
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def parse(path):
    kernels, cur = [], None
    for line in open(path):
        m = re.match(r"kernel: (.*?)\s+grid ", line)
        if m:
            cur = {"kernel": m.group(1).split("(")[0].strip(), "read": 0.0, "write": 0.0, "ms": 0.0}
            kernels.append(cur)
            continue
        f = line.split()
        if cur is None or len(f) < 3:
            continue
        if f[0] == "dram__bytes_read.sum":
            cur["read"] = float(f[1]) * UNITS[f[2]]
        elif f[0] == "dram__bytes_write.sum":
            cur["write"] = float(f[1]) * UNITS[f[2]]
        elif f[0] == "gpu__time_duration.sum":
            cur["ms"] = float(f[1]) * TIME[f[2]]
    return kernels


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full.txt"))):
        m = re.match(r"r(\d+)_(.*)_ncu_full\.txt", os.path.basename(path))
        if not m:
            continue
        rnd, name = int(m.group(1)), m.group(2)
        ks = parse(path)
        if not ks:
            continue
        top = max(ks, key=lambda k: k["ms"])
        if name not in out or out[name]["round"] <= rnd:
            out[name] = {"bytes": top["read"] + top["write"], "read": top["read"], "write": top["write"], "kernel": top["kernel"],
                         "kernel_ms_under_ncu": top["ms"], "source": "profiles/" + os.path.basename(path), "round": rnd}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    for k, v in sorted(out.items()):
        print("%-12s %10.1f MB  %s  (%s)" % (k, v["bytes"] / 1e6, v["kernel"], v["source"]))


if __name__ == "__main__":
    main()
