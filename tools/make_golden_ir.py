#!/usr/bin/env python
"""Writes the optimised IR ("mmir 1" text) of bench.py's workloads to tests/golden/ir/<name>.mmir.

`bench.py --impl reference` reads these files, so the reference arm (the oracle: IR -> host C -> gcc -O2) does not load
libmathmap_b200.so.  tests/test_frontend.py checks that the committed text still equals what the front end produces.

    python tools/make_golden_ir.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    import mathmap_b200 as mb
    out_dir = os.path.join(ROOT, "tests", "golden", "ir")
    os.makedirs(out_dir, exist_ok=True)
    for name, wl in bench.WORKLOADS.items():
        m = mb.Module.from_file(os.path.join(bench.FILTERS, wl["script"]))
        path = os.path.join(out_dir, wl["ir"])
        with open(path, "w") as f:
            f.write(m.ir)
        print(path, len(m.ir))


if __name__ == "__main__":
    main()
