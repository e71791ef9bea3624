#!/usr/bin/env python
"""Per-source-line cost table from an ncu report captured with --import-source on (and -lineinfo):
warp-stall samples and executed warp instructions aggregated by CUDA source line.

    python tools/ncu_by_line.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, agg = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":           # a source-line summary row
            d = dict(zip(hdr, r))
            try:
                samples = int(d["# Samples"])
                inst = int(d["Instructions Executed"])
            except ValueError:
                continue
            if samples or inst:
                agg.append((samples, inst, fname, int(r[0]), r[1].strip()))
    tot_s = sum(a[0] for a in agg) or 1
    tot_i = sum(a[1] for a in agg) or 1
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    print(f"{'%smp':>6} {'%inst':>6}  location")
    for s, i, f, ln, src in sorted(agg, reverse=True)[:top]:
        print(f"{100 * s / tot_s:6.2f} {100 * i / tot_i:6.2f}  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main()
