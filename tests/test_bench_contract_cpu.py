"""bench.py's reference arm runs without a GPU: it must print exactly one JSON line on stdout with
the keys the driver reads, and nothing else (libraries' banners go to stderr)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--samples", "3000", "--cpu-samples", "3000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "mmctm_em_iterations_per_sec" and j["unit"] == "iterations/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and "model" not in j["config"]
    # the sample is the whole (3000-sample) workload here: nothing extrapolated, and the line says so
    assert j["extrapolated"] is False and j["sample_D"] == 3000 and j["cpu_baseline"]["extrapolated"] is False


def test_reference_arm_other_configs():
    for cfgno, metric in ((2, "lda_em_iterations_per_sec"), (3, "mmctm_em_iterations_per_sec")):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(cfgno), "--steps", "1",
                            "--warmup", "0", "--samples", "40000", "--cpu-budget-s", "2"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        j = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
        assert j["metric"] == metric and j["value"] > 0 and j["sample_D"] <= 40000
        assert j["extrapolated"] == (j["sample_D"] < 40000)


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--samples", "2000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_gpu_bench_line_keeps_the_contract():
    """The product arm needs a B200; its last line from the GPU box is committed (profiles/bench_r02p_c4_final.json).  It has
    to carry what the driver and the judge read, with consistent arithmetic: value = 1000 / ms_per_step, roofline.frac =
    achieved / peak with achieved = algorithmic bytes / the dominant kernel's duration, e2e with real copies, a CPU arm."""
    j = json.load(open(os.path.join(ROOT, "profiles", "bench_r02p_c4_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in j, k
    assert j["metric"] == "mmctm_em_iterations_per_sec" and j["unit"] == "iterations/s" and j["dtype"] == "f64"
    assert j["n_gpus"] == 1 and j["warmup"] >= 3 and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert abs(j["value"] - 1000.0 / j["ms_per_step"]) <= 1e-9 * j["value"]
    assert "workload" in j["config"] and "model" not in j["config"] and j["config"]["samples"] == 1000000
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "k_solve"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["kernel_ms_per_launch"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert r["kernel_ms_per_launch"] <= j["ms_per_step"] and r["traffic"] > 0
    f = j["roofline_fp64"]
    assert f["bound"] == "fp64_pipe" and 0 < f["frac"] < 1 and abs(f["frac"] - f["achieved"] / f["peak"]) <= 1e-12
    e = j["e2e"]
    assert e["unit"] == j["unit"] and e["h2d_bytes_per_step"] > 1e9 and e["d2h_bytes_per_step"] > 5e8 and 0 < e["value"] < j["value"]
    c = j["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == j["unit"] and c["sample"]
    assert j["clocks"]["sm_mhz"] > 0 and j["clocks"]["sm_max_mhz"] >= j["clocks"]["sm_mhz"]
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert j["gpu_launches"] >= 13 * j["steps"]               # θ x3, solve x2, combine x2, mstep1, moments, LL x3, mstep2 per iteration
    assert sum(v["ms_per_step"] for v in j["kernels"].values()) <= 1.05 * j["ms_per_step_with_kernel_timing"]
    m = j["fp32_mode"]                                        # reported beside the headline, never as it
    assert m["ms_per_step"] < j["ms_per_step"] and m["ll_rel_diff_to_fp64"] < 1e-5
