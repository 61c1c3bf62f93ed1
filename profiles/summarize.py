#!/usr/bin/env python
"""Summarise gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep into
profiles/<tag>_summary.md (run here, no GPU needed).  Usage: python profiles/summarize.py <tag>"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = ["# ncu summary `%s`" % tag, "",
       "Command: `bash profiles/run_ncu.sh %s <samples>` under gpurun (B200, one GPU); numbers under ncu are "
       "cold-cache and serialised -- compare SHARES, not absolutes.  Bench values come from the plain run." % tag, ""]

lp = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = {}
    for x in csv.DictReader(lines):
        n = x["Kernel Name"].split("(")[0].replace("void ", "").replace("mmsig::", "")
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out += ["## Launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
            "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for n, a in sorted(agg.items(), key=lambda t: -t[1][1]):
        out.append("| `%s` | %d | %.3f | %.1f%% |" % (n, a[0], a[1], 100 * a[1] / tot))
    out.append("")

rp = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
rc = os.path.join(ROOT, "gpurun_out", "prof_%s_raw.csv" % tag)          # written on the GPU box when the report is too large to pull
if os.path.exists(rp) or os.path.exists(rc):
    raw = open(rc).read() if os.path.exists(rc) else subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in data]
    out += ["## `--set full` capture of one iteration's hot kernels", "",
            "| metric | unit | " + " | ".join("`%s`" % n for n in names) + " |", "|---|---|" + "---:|" * len(names)]
    for w in want:
        if w in idx:
            out.append("| %s | %s | %s |" % (w, units[idx[w]], " | ".join(r[idx[w]] for r in data)))
    out.append("")
    # top stall reasons per kernel
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if stall:
        out += ["### Warp stall reasons (warps stalled per issue-active cycle, top 6)", ""]
        for n, r in zip(names, data):
            vals = sorted(((float(r[idx[s]] or 0), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stall), reverse=True)[:6]
            out.append("- `%s`: " % n + ", ".join("%s %.2f" % (s, v) for v, s in vals))
        out.append("")
open(os.path.join(ROOT, "profiles", "%s_summary.md" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out))
