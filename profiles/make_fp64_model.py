#!/usr/bin/env python
"""profiles/fp64_model.json: FP64 lane-instructions the solve kernels issue per LD_MMA evaluation and coordinate, from an
ncu --set full capture of k_solve_lean (nu and lambda phase of ONE iteration) and the bench line of the same command
(evaluation counts of that iteration), plus the measured FP64 pipe rate (profiles/micro/fp64_peak.cu).
Usage: python profiles/make_fp64_model.py <report.ncu-rep> <plain bench log with the JSON line> <fp64_peak.jsonl>"""
import csv
import json
import re
import subprocess
import sys

rep, plain, peakf = sys.argv[1:4]
line = [l for l in open(plain) if l.startswith("{")][-1]
b = json.loads(line)
D, MK = b["config"]["samples"], sum(b["config"]["K"])
ev = b["mma_evaluations_per_sample_last_iteration"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, hdr, fp64, tot, seen = None, None, {}, {}, set()
for r in csv.reader(raw.splitlines()):
    if r and r[0] == "Kernel Name":
        kern, hdr = r[1], None
    elif r and r[0] == "Address":
        hdr = r
    elif kern and hdr and len(r) >= len(hdr) - 2:
        d = dict(zip(hdr, r))
        if (kern, d["Address"]) in seen:          # the page lists every instruction twice
            continue
        seen.add((kern, d["Address"]))
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
        op = m.group(2).split(".")[0] if m else "?"
        n = int(d["Instructions Executed"])
        tot[kern] = tot.get(kern, 0) + n
        if op in ("DFMA", "DADD", "DMUL", "DSETP"):
            fp64[kern] = fp64.get(kern, 0) + n
out = {}
for k in fp64:
    ph = "nu" if re.search(r"\(int\)1>|, 1>", k) else "lambda"
    e = ev["nu_mean"] if ph == "nu" else ev["lambda_mean"]
    out[ph] = {"kernel": k, "warp_inst": tot[k], "fp64_warp_inst": fp64[k], "fp64_share": fp64[k] / tot[k],
               "lane_inst_per_coordinate_eval": fp64[k] * 32.0 / (D * MK * e)}
pk = [json.loads(l) for l in open(peakf) if l.startswith("{")]
rate = max(p["warp_inst_per_clk_per_sm"] for p in pk if "op" in p)
dev = pk[0]
res = {"fp64_lane_inst_per_coordinate_eval": {"nu": out["nu"]["lane_inst_per_coordinate_eval"], "lambda": out["lambda"]["lane_inst_per_coordinate_eval"]},
       "detail": out, "D": D, "MK": MK, "evaluations": ev,
       "peak_warp_inst_per_clk_per_sm": rate, "sms": dev["sms"], "clock_ghz": dev["clock_ghz_nominal"],
       "peak_source": "profiles/micro/fp64_peak.cu on the box's B200: %.3f FP64 warp-instructions / clock / SM (DFMA = DADD = DMUL), %d SMs, %.3f GHz" % (rate, dev["sms"], dev["clock_ghz_nominal"]),
       "model_source": "ncu source page of %s: executed DFMA + DADD + DMUL + DSETP warp-instructions of k_solve_lean at D=%d, divided by the evaluations of that iteration" % (rep.split("/")[-1], D)}
json.dump(res, open("profiles/fp64_model.json", "w"), indent=1)
print(json.dumps(res["fp64_lane_inst_per_coordinate_eval"]), {k: round(v["fp64_share"], 3) for k, v in out.items()})
