#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full report: DRAM bytes (read + write) per launch of the
hot kernels.  Usage: python profiles/make_traffic.py <tag> <D>   (reads gpurun_out/prof_<tag>.ncu-rep)"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, D = sys.argv[1], int(sys.argv[2])
csvp = os.path.join(ROOT, "gpurun_out", "prof_%s_raw.csv" % tag)          # written on the GPU box when the report is too large to pull
if os.path.exists(csvp):
    raw = open(csvp).read()
else:
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag), "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
iN, iR, iW = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"source": "ncu --set full --clock-control none, profiles/run_ncu.sh %s %d (B200, D = %d, K=[10,8,6])" % (tag, D, D),
       "D": D, "kernels": {}}
for r in rows[2:]:
    name = r[iN].split("(")[0].replace("void ", "").replace("mmsig::", "")
    b = float(r[iR].replace(",", "")) * scale[units[iR]] + float(r[iW].replace(",", "")) * scale[units[iW]]
    k = out["kernels"].setdefault(name, {"dram_bytes_per_launch": [], "dram_bytes_per_sample": []})
    k["dram_bytes_per_launch"].append(b)
    k["dram_bytes_per_sample"].append(b / D)
# the solve is one logical kernel per step launched as two phases (nu, then lambda): bench.py's dominant kernel "k_solve"
ph = [v for k, v in out["kernels"].items() if k.startswith("k_solve_lean")]
if len(ph) == 2:
    tot = sum(v["dram_bytes_per_launch"][0] for v in ph)
    out["kernels"] = dict([("k_solve", {"dram_bytes_per_launch": [tot], "dram_bytes_per_sample": [tot / D],
                                        "what": "k_solve_lean nu phase + lambda phase of one iteration"})] + list(out["kernels"].items()))
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps({k: v["dram_bytes_per_sample"] for k, v in out["kernels"].items()}))
