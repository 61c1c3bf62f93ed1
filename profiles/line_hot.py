#!/usr/bin/env python
"""Attribute executed warp-instructions of one kernel to CUDA source lines: joins the ncu source
page (per-SASS counts) with nvdisasm -g line info of the same cubin (same instruction order).
Usage: python profiles/line_hot.py <report.ncu-rep> <kernel-regex> <mangled-substring> [libmmsig.so] [n]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
so = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "multimodalmusig.jl_b200", "libmmsig.so")
n = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data, seen = [], 0
for r in rows:
    if r == hdr:
        seen += 1
        continue
    if seen == 1 and len(r) > iSamp and r[iE].isdigit():
        data.append((r[iS].strip(), int(r[iE]), int(r[iSamp])))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
if len(lines) != len(data):
    print("warning: %d SASS in report vs %d in cubin (stale .so?)" % (len(data), len(lines)))
agg, sagg = collections.Counter(), collections.Counter()
for (src, e, s), ln in zip(data, lines):
    agg[ln] += e
    sagg[ln] += s
tot, ts = sum(agg.values()), max(sum(sagg.values()), 1)
srcs = {}
print("%-26s %7s %9s  source" % ("file:line", "inst %", "samples %"))
for ln, c in agg.most_common(n):
    text = ""
    if ln:
        p = os.path.join(os.path.dirname(so), "csrc", ln[0])
        if p not in srcs and os.path.exists(p):
            srcs[p] = open(p).read().splitlines()
        if p in srcs and ln[1] - 1 < len(srcs[p]):
            text = srcs[p][ln[1] - 1].strip()[:90]
    print("%-26s %6.2f%% %8.2f%%  %s" % ("%s:%d" % ln if ln else "?", 100.0 * c / tot, 100.0 * sagg[ln] / ts, text))

if os.environ.get("DUMP_LINE"):
    f, l = os.environ["DUMP_LINE"].split(":")
    ops = collections.Counter()
    for (src, e, s), ln in zip(data, lines):
        if ln and ln[0] == f and ln[1] == int(l):
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
            ops[m.group(2) if m else "?"] += e
    print("opcode mix attributed to", os.environ["DUMP_LINE"])
    for op, c in ops.most_common(25):
        print("  %-28s %6.2f%%" % (op, 100.0 * c / tot))
