#!/usr/bin/env python
"""Opcode mix + hottest source lines of one kernel from an ncu report's source page.
Usage: python profiles/sass_mix.py <report.ncu-rep> <kernel-regex> [n]"""
import collections
import csv
import re
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 28
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
first_kernel = True
data = []
seen_hdr = 0
for r in rows:
    if r == hdr:
        seen_hdr += 1
        continue
    if seen_hdr == 1 and len(r) > iSamp and r[iE].isdigit():
        data.append(r)
tot = sum(int(r[iE]) for r in data)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[iE])
    samp[op] += int(r[iSamp])
ts = max(sum(samp.values()), 1)
print("kernel /%s/: %d SASS instructions, %d warp-instructions executed" % (kre, len(data), tot))
print("%-10s %8s %9s" % ("opcode", "inst %", "samples %"))
for op, c in ops.most_common(n):
    print("%-10s %7.2f%% %8.2f%%" % (op, 100.0 * c / tot, 100.0 * samp[op] / ts))
