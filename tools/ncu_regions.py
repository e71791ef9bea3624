#!/usr/bin/env python
"""Opcode histogram per address region of a kernel, per 32 walker*pixels (= warp instructions per pixel of a lane),
from the SASS source page of an ncu report.

    python tools/ncu_regions.py rep.ncu-rep <walker*pixels of the launch> name:lo:hi [name:lo:hi ...]   (hex offsets)
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, wpx = sys.argv[1], float(sys.argv[2])
    regions = []
    for a in sys.argv[3:]:
        n, lo, hi = a.split(":")
        regions.append((n, int(lo, 16), int(hi, 16)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ix = {h: i for i, h in enumerate(hdr)}
    inst = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    base = int(inst[0][0], 16)
    unit = wpx / 32.0
    tot_s = sum(int(r[ix["# Samples"]]) for r in inst) or 1
    grand = 0.0
    for name, lo, hi in regions:
        c, smp = collections.Counter(), 0
        for r in inst:
            a = int(r[0], 16) - base
            if lo <= a < hi:
                op = r[1].strip().split()
                if op[0].startswith("@"):
                    op = op[1:]
                c[op[0].split(".")[0]] += int(r[ix["Instructions Executed"]])
                smp += int(r[ix["# Samples"]])
        n = sum(c.values())
        grand += n / unit
        print(f"{name:14s} {n / unit:7.2f} instr/px {100 * smp / tot_s:5.1f} % samples  " +
              " ".join(f"{k}:{v / unit:.2f}" for k, v in c.most_common(12)))
    print(f"{'sum':14s} {grand:7.2f}")


if __name__ == "__main__":
    main()
