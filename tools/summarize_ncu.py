#!/usr/bin/env python
"""Turn an ncu report (--set full, one kernel) into the small summaries committed under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_c5a  [--latest]

Writes <out>.json (key metrics, stall ratios, phase breakdown, opcode mix) and <out>.md (the same, readable).
--latest also refreshes profiles/latest_ncu_summary.json, which bench.py reads for roofline.traffic.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = ncu_csv(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    ix = {h: i for i, h in enumerate(hdr)}
    summ = {"report": os.path.basename(rep), "kernel": vals[ix["Kernel Name"]] if "Kernel Name" in ix else None}
    metrics = {}
    for k in KEYS:
        if k in ix:
            v, u = vals[ix[k]], units[ix[k]]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            metrics[k] = {"value": v, "unit": u}
    summ["metrics"] = metrics
    stalls = {}
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(vals[ix[h]])
    summ["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))

    def to_bytes(k):
        m = metrics.get(k)
        return None if m is None else m["value"] * UNIT_SCALE.get(m["unit"], 1)
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    summ["dram_bytes_per_launch"] = None if rd is None else rd + (wr or 0)
    dur = metrics.get("gpu__time_duration.sum")
    summ["duration_ms"] = None if dur is None else dur["value"] * UNIT_SCALE.get(dur["unit"], 1)

    src = ncu_csv(rep, "source")
    if len(src) > 3:
        sh, data = src[1], src[2:]
        six = {h: i for i, h in enumerate(sh)}
        tot = sum(int(r[six["# Samples"]]) for r in data) or 1
        st_cols = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        regions, cur, start = [], [0, 0, collections.Counter()], 0
        ops = collections.Counter()
        for k, r in enumerate(data):
            cur[0] += int(r[six["# Samples"]])
            cur[1] += int(r[six["Instructions Executed"]])
            for h in st_cols:
                cur[2][h] += int(r[six[h]])
            s = r[six["Source"]].strip()
            parts = s.split()
            op = (parts[1] if s.startswith("@") and len(parts) > 1 else parts[0]).split(".")[0] if parts else "?"
            ops[op] += int(r[six["Instructions Executed"]])
            if "BAR.SYNC" in s or "EXIT" in s:
                regions.append((start, k, cur))
                cur, start = [0, 0, collections.Counter()], k + 1
        regions.append((start, len(data) - 1, cur))
        summ["phases_between_barriers"] = [
            {"sass_range": [a, b], "warp_time_pct": round(100 * c[0] / tot, 1), "warp_instructions": c[1],
             "top_stalls_pct": {h[6:]: round(100 * v / tot, 1) for h, v in c[2].most_common(4)}}
            for a, b, c in regions if c[0] > 0.005 * tot]
        n_inst = sum(ops.values()) or 1
        summ["opcode_mix_pct"] = {k: round(100 * v / n_inst, 1) for k, v in ops.most_common(10)}

    with open(out + ".json", "w") as fh:
        json.dump(summ, fh, indent=1)
    lines = [f"# ncu summary: {summ['report']}", "", f"kernel: `{summ['kernel']}`", "", "| metric | value | unit |", "|---|---|---|"]
    for k, m in metrics.items():
        lines.append(f"| {k} | {m['value']} | {m['unit']} |")
    lines += ["", "Stalled warps per issued instruction (smsp__average_warps_issue_stalled_*):", ""]
    lines += [f"- {k}: {v:.2f}" for k, v in summ["stall_warps_per_issue"].items() if v >= 0.05]
    if "phases_between_barriers" in summ:
        lines += ["", "Warp time between barriers (sampling):", ""]
        for ph in summ["phases_between_barriers"]:
            lines.append(f"- SASS {ph['sass_range']}: {ph['warp_time_pct']} % of warp time, "
                         f"{ph['warp_instructions']:.3e} warp instructions, stalls {ph['top_stalls_pct']}")
        lines += ["", f"Opcode mix (% of warp instructions): {summ['opcode_mix_pct']}"]
    with open(out + ".md", "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if "--latest" in sys.argv:
        with open(os.path.join(os.path.dirname(out) or ".", "latest_ncu_summary.json"), "w") as fh:
            note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else None
            json.dump({"report": summ["report"], "note": note, "dram_bytes_per_launch": summ["dram_bytes_per_launch"],
                       "duration_ms": summ["duration_ms"],
                       "fp64_pipe_pct": metrics.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                                    {}).get("value")}, fh, indent=1)
    print(f"wrote {out}.json / .md")


if __name__ == "__main__":
    main()
