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
