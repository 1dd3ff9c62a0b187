"""Per-instruction view of an `ncu --set full --import-source on` capture: the executed path with counts per warp-pixel,
or the instructions with the most stall samples (and how many of those wait on memory / on fixed latencies).
Usage: python tools/ncu_hot.py REPORT.ncu-rep path [WARPS]     executed instructions, counts divided by WARPS (default: max count)
       python tools/ncu_hot.py REPORT.ncu-rep stalls [N]        the N (default 30) instructions with the most samples"""
import csv
import io
import subprocess
import sys


def rows_of(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # one section per profiled kernel ("Kernel Name" row, header row, instructions); MMB_NCU_KERNEL picks one (default 0)
    import os
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    sections = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
    sections = sections[::2] if len(sections) > 1 and rows[sections[0][0]] == rows[sections[1][0]] else sections  # ncu lists every kernel twice
    k = int(os.environ.get("MMB_NCU_KERNEL", "0"))
    rows = rows[sections[k][0]:sections[k][1]]
    print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?")
    hdr = rows[1]
    col = {name: hdr.index(name) for name in ("Address", "Source", "Instructions Executed", "# Samples", "stall_long_sb", "stall_wait")}
    body = [r for r in rows[2:] if len(r) > col["# Samples"] and r[col["Instructions Executed"]].isdigit()]
    return col, body


def main():
    report, mode = sys.argv[1], sys.argv[2]
    col, body = rows_of(report)
    base = int(body[0][col["Address"]], 16)
    if mode == "path":
        warps = float(sys.argv[3]) if len(sys.argv) > 3 else max(int(r[col["Instructions Executed"]]) for r in body)
        total = 0.0
        for r in body:
            n = int(r[col["Instructions Executed"]]) / warps
            if n < 0.2:
                continue
            total += n
            print("%05x %6.2f %6s  %s" % (int(r[col["Address"]], 16) - base, n, r[col["# Samples"]], r[col["Source"]].strip()))
        print("instructions per warp on paths taken by at least 20 %% of the warps: %.1f" % total)
    else:
        top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
        samples = sum(int(r[col["# Samples"]] or 0) for r in body)
        print("total samples", samples)
        for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
            print("%05x %7s  long_scoreboard %6s  wait %6s  %s" % (int(r[col["Address"]], 16) - base, r[col["# Samples"]], r[col["stall_long_sb"]],
                                                                   r[col["stall_wait"]], r[col["Source"]].strip()[:90]))


if __name__ == "__main__":
    main()
