#!/usr/bin/env python
"""Warp-stall samples and executed warp instructions of an ncu report (--import-source on) aggregated over SASS
ranges: the main body split at its CALL targets (= the out-of-line device functions) plus user-given split points.

    python tools/ncu_by_range.py gpurun_out/prof.ncu-rep [extra hex offsets ...]
"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    extra = [int(x, 16) for x in sys.argv[2:]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    ix = {h: i for i, h in enumerate(hdr)}
    inst = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
    base = int(inst[0][0], 16)
    cuts = set(extra)
    for r in inst:
        m = re.search(r"CALL\.REL\.NOINC\s+(0x[0-9a-f]+)", r[1])
        if m:
            t = int(m.group(1), 16)
            cuts.add(t - base if t >= base else t)
        if re.search(r"\bEXIT\b", r[1]) and not r[1].strip().startswith("@"):
            cuts.add(int(r[0], 16) - base + 16)
    cuts = sorted(c for c in cuts if c > 0)
    bounds = [0] + cuts + [1 << 40]
    tot_s = sum(int(r[ix["# Samples"]]) for r in inst) or 1
    tot_i = sum(int(r[ix["Instructions Executed"]]) for r in inst) or 1
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        sel = [r for r in inst if lo <= int(r[0], 16) - base < hi]
        if not sel:
            continue
        s = sum(int(r[ix["# Samples"]]) for r in sel)
        i = sum(int(r[ix["Instructions Executed"]]) for r in sel)
        print(f"[{lo:#07x}, {min(hi, int(sel[-1][0], 16) - base + 16):#07x})  {len(sel):5d} SASS  "
              f"{100 * s / tot_s:6.2f} % samples  {100 * i / tot_i:6.2f} % warp instructions")


if __name__ == "__main__":
    main()
