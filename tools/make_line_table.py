#!/usr/bin/env python
"""Build rbvfit_b200/data/atomic_lines.json from the reference's atomic line list.

Source: /root/reference/src/rbvfit/lines/atom_full.dat -- a plain 4-column table
(ion, rest wavelength [A], oscillator strength f, damping gamma [1/s]; Morton-2003-style
atomic data), which is DATA, not code.  The values are kept as the decimal strings of the
source so that loaders can reproduce the reference's float32 rounding of f and gamma
(rb_setline.py:42,44) bit for bit.  Run once in the build container; the JSON is committed.
"""
import json
import os
import sys

SRC = "/root/reference/src/rbvfit/lines/atom_full.dat"


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "rbvfit_b200", "data", "atomic_lines.json")
    rows = []
    with open(SRC) as fh:
        for line in fh:
            parts = line.split()
            if len(parts) != 4:
                continue
            rows.append(parts)
    doc = {
        "source": "rbvfit lines/atom_full.dat (ion, wrest_A, f, gamma_per_s as decimal strings)",
        "columns": ["ion", "wrest", "f", "gamma"],
        "lines": rows,
    }
    with open(dst, "w") as fh:
        fh.write('{"source": %s,\n "columns": %s,\n "lines": [\n' % (json.dumps(doc["source"]),
                                                                  json.dumps(doc["columns"])))
        fh.write(",\n".join("  " + json.dumps(r) for r in rows))
        fh.write("\n]}\n")
    print(f"wrote {os.path.normpath(dst)}: {len(rows)} lines")


if __name__ == "__main__":
    sys.exit(main())
