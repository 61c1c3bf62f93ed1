#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line (file:line) of one kernel, from an ncu report with
--import-source on (needs -lineinfo).  Usage: python profiles/ncu_line_hot.py <report.ncu-rep> <kernel-regex> [n]"""
import collections
import csv
import re
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fpath, func, hdr, cur, seen = None, None, None, None, set()
acc = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])      # (file, line) -> samples, inst, text, stall kinds
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    elif hdr and func and re.search(kre, func) and len(r) > iI:
        if r[0].isdigit():             # a CUDA line; its SASS rows follow (a preview, "...", then the full list)
            cur = (fpath, int(r[0]))
            acc[cur][2] = r[1].strip()
            seen = set()
        elif r[0] == "" and r[2].startswith("0x") and cur is not None and r[2] not in seen:
            seen.add(r[2])
            a = acc[cur]
            a[0] += int(r[iS] or 0)
            a[1] += int(r[iI] or 0)
            for i, h in stall:
                if r[i] not in ("", "0", "-"):
                    a[3][h] += int(r[i])
ts = sum(a[0] for a in acc.values()) or 1
ti = sum(a[1] for a in acc.values()) or 1
print("kernel /%s/: %d stall samples, %d warp-instructions" % (kre, ts, ti))
for (f, l), a in sorted(acc.items(), key=lambda kv: -kv[1][0])[:n]:
    top = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(a[0], 1)) for k, v in a[3].most_common(3))
    print("%5.2f%% smp %5.2f%% inst  %s:%d  %s   [%s]" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, f, l, a[2][:70], top))
