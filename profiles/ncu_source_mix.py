#!/usr/bin/env python
"""Per-kernel dynamic opcode mix, stall samples by opcode and hottest source lines from the source page of an ncu report.
Usage: python profiles/ncu_source_mix.py <report.ncu-rep> [kernel-regex] [n]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, hdr, tables, seen = None, None, collections.OrderedDict(), set()
for r in csv.reader(raw.splitlines()):
    if r and r[0] == "Kernel Name":
        kern = r[1]
        tables.setdefault(kern, [])
        hdr = None
    elif r and r[0] == "Address":
        hdr = r
    elif kern and hdr and len(r) >= len(hdr) - 2 and (kern, r[0]) not in seen:
        seen.add((kern, r[0]))
        tables[kern].append(dict(zip(hdr, r)))
for kern, rows in tables.items():
    if not re.search(kre, kern) or not rows:
        continue
    tot = sum(int(r["Instructions Executed"]) for r in rows)
    samp = sum(int(r["# Samples"]) for r in rows)
    ops, osamp = collections.Counter(), collections.Counter()
    for r in rows:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r["Source"])
        op = m.group(2) if m else "?"
        op = "IMAD.MOV" if op.startswith("IMAD.MOV") else op.split(".")[0]
        ops[op] += int(r["Instructions Executed"])
        osamp[op] += int(r["# Samples"])
    fp = sum(ops[k] for k in ("DFMA", "DADD", "DMUL", "DSETP"))
    print("== %s\n   %d SASS lines, %d warp-instructions executed, FP64 %.1f %%, %d stall samples" % (kern, len(rows), tot, 100.0 * fp / tot, samp))
    print("   %-10s %8s %9s" % ("opcode", "inst %", "samples %"))
    for op, c in ops.most_common(n):
        print("   %-10s %7.2f%% %8.2f%%" % (op, 100.0 * c / tot, 100.0 * osamp[op] / max(samp, 1)))
