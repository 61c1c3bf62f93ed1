"""N>1 host logic on CPU (world_size 2, gloo): contiguous nnz-balanced sharding, and the
exchange protocol of the multi-GPU path -- every rank contributes double-double partial sums of
its shard, the partials are ALL-GATHERED (not all-reduced) and summed locally in rank order --
reproduces the full-data result of the pinned-arithmetic oracle bit for bit."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import orc
import mmsig
from mmsig.counts import shard_rows, slice_csr

K, V, D = [3, 2], [12, 7], 90


def _data():
    counts = mmsig.synth.generate(D, K, V, key=3)
    g0 = mmsig.synth.init_gamma(K, V)
    return counts, g0


def _dd(addends):
    hi = math.fsum(addends)
    lo = math.fsum(list(addends) + [-hi])
    return hi, lo


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counts, g0 = _data()
    b = shard_rows([c[0] for c in counts], world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    shard = [slice_csr(c, lo, hi) for c in counts]
    o = orc.OracleMMCTM(K, [0.1, 0.1], V, shard, g0, arith=orc.ARITH_DET)
    for d in range(o.D):                                   # E-step of this shard
        o.L.orc_mmctm_fitdoc(o.p, d)
    G, MK = o.G, o.MK
    part = np.zeros((G + MK, 2))
    goff = np.cumsum([0] + [k * v for k, v in zip(K, V)])
    for m in range(2):
        rp, term, cnt = shard[m]
        th = o.theta(m)
        for k in range(K[m]):
            for v in range(V[m]):
                sel = term == v
                part[goff[m] + k * V[m] + v] = _dd(list(th[sel, k] * cnt[sel].astype(float)))
    for j in range(MK):
        part[G + j] = _dd(list(o.lam[:, j]))
    t = torch.from_numpy(part)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    tot = np.array([math.fsum([float(g[i, 0]) for g in gathered] + [float(g[i, 1]) for g in gathered])
                    for i in range(G + MK)])
    if rank == 0:
        np.save(out, tot)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_partials_reproduce_full_data_oracle(tmp_path):
    out = str(tmp_path / "tot.npy")
    mp.spawn(_worker, args=(2, 29731, out), nprocs=2, join=True)
    tot = np.load(out)
    counts, g0 = _data()
    o = orc.OracleMMCTM(K, [0.1, 0.1], V, counts, g0, arith=orc.ARITH_DET)
    for d in range(D):
        o.L.orc_mmctm_fitdoc(o.p, d)
    o.L.orc_mmctm_update_mu(o.p)
    o.L.orc_mmctm_update_gamma(o.p)
    G = o.G
    alpha = np.repeat([0.1, 0.1], [K[0] * V[0], K[1] * V[1]])
    # gamma = exact_round(alpha + stats): recompute from the gathered partial sums
    # (the partials were rounded to dd, so compare the statistics, then gamma to 1 ulp)
    np.testing.assert_array_equal(tot[G:] / D, o.mu)
    np.testing.assert_allclose(tot[:G] + alpha, o.gamma, rtol=3e-16)


def test_shard_rows_and_slices():
    counts, _ = _data()
    for n in (1, 2, 3, 8):
        b = shard_rows([c[0] for c in counts], n)
        assert b[0] == 0 and b[-1] == D and np.all(np.diff(b) >= 0) and len(b) == n + 1
        rows = 0
        for r in range(n):
            s = [slice_csr(c, int(b[r]), int(b[r + 1])) for c in counts]
            rows += len(s[0][0]) - 1
            assert s[0][0][0] == 0 and s[0][0][-1] == len(s[0][1])
        assert rows == D
    # balanced by nonzeros within 2x of ideal
    b = shard_rows([c[0] for c in counts], 3)
    nnz = [sum(int(c[0][b[r + 1]] - c[0][b[r]]) for c in counts) for r in range(3)]
    assert max(nnz) < 2 * (sum(nnz) / 3)


def test_synthetic_shards_do_not_depend_on_world_size():
    full = mmsig.synth.generate(1000, [3], [12], key=9)[0]
    for lo, hi in ((0, 400), (400, 1000), (123, 777)):
        part = mmsig.synth.generate(1000, [3], [12], key=9, lo=lo, hi=hi)[0]
        ref = slice_csr(full, lo, hi)
        assert all(np.array_equal(a, c) for a, c in zip(part, ref))
