"""N>1 path on real GPUs: torchrun with one rank per GPU, NCCL all-gather of the packed partials.
Skipped on a single-GPU box (the host-side sharding logic is covered on CPU by test_shard_cpu.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


MODES = ["mmctm", "lda", "mmctm_host", "mmctm_fit"]


@pytest.mark.parametrize("mode", MODES)
def test_two_ranks_match_full_data_oracle(mode):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + MODES.index(mode)), os.path.join(ROOT, "tests", "mp_worker.py"), mode, "3000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI-RANK PARITY OK" in r.stdout
