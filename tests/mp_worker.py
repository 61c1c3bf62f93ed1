"""Multi-rank parity worker (launched by torchrun from test_gpu_multi.py / by hand):
every rank owns a contiguous shard of the samples on its own GPU; rank 0 checks the gathered
result against the full-data oracle -- bit-exactly, because every cross-sample sum is exactly
rounded and therefore independent of the sharding."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mmsig  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "mmctm"
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl")
    uid = [mmsig.capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    comm = (uid[0], rank, world)
    if mode == "mmctm_host":
        # the one-call fit from host buffers (chunked, transfers pipelined), every rank on its shard
        os.environ["MMSIG_PIPE_CHUNKS"] = "3"
        K, V = [10, 8, 6], [96, 32, 83]
        full = mmsig.synth.generate(D, K, V, key=6)
        b = mmsig.counts.shard_rows([c[0] for c in full], world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        shard = [mmsig.counts.slice_csr(c, lo, hi) for c in full]
        g0 = mmsig.synth.init_gamma(K, V)
        m = mmsig.MMCTM(K, [0.1] * 3, shard, V=V, gamma0=g0, device=local, comm=comm, D_total=D)
        hist, s = m.fit_host(shard, g0, maxiter=3, D_total=D)
        elbo = m.calculate_elbo()[0]
        parts = [None] * world
        dist.gather_object((lo, hi, s["lam"], s["nu"], s["props"], s["zeta"]), parts if rank == 0 else None, 0)
        if rank == 0:
            import orc
            o = orc.OracleMMCTM(K, [0.1] * 3, V, full, g0, arith=orc.ARITH_DET, nthreads=os.cpu_count() or 1)
            llo = o.fit(maxiter=3, tol=1e-4)
            for i, k in enumerate(("lam", "nu", "props", "zeta")):
                assert np.array_equal(np.concatenate([p[2 + i] for p in parts]), getattr(o, k)), k
            for k in ("gamma", "mu", "Sigma", "invSigma", "phi"):
                assert np.array_equal(s[k], getattr(o, k)), k
            assert np.array_equal(hist, np.asarray(llo))
            eo = o.elbo()[0]
            assert abs(elbo - eo) <= 1e-12 * abs(eo), (elbo, eo)
            print("MULTI-RANK PARITY OK mmctm_host world=%d D=%d" % (world, D), flush=True)
        m.close()
    elif mode == "mmctm_fit":
        # a whole fit!: ten iterations without host round trips, then batches with the stopping rule on the device, the
        # second half of every M-step (and its all-gather) on the side stream; stops where the full-data oracle stops
        K, V = [10, 8, 6], [96, 32, 83]
        full = mmsig.synth.generate(D, K, V, key=8)
        b = mmsig.counts.shard_rows([c[0] for c in full], world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        shard = [mmsig.counts.slice_csr(c, lo, hi) for c in full]
        g0 = mmsig.synth.init_gamma(K, V)
        m = mmsig.MMCTM(K, [0.1] * 3, shard, V=V, gamma0=g0, device=local, comm=comm, D_total=D)
        tol = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5e-3       # the rule fires in iteration 15: three enqueued iterations are skipped
        hist = m.fit(maxiter=22, tol=tol, verbose=False)
        s = m.state()
        parts = [None] * world
        dist.gather_object((lo, hi, s["lam"], s["nu"]), parts if rank == 0 else None, 0)
        if rank == 0:
            import orc
            o = orc.OracleMMCTM(K, [0.1] * 3, V, full, g0, arith=orc.ARITH_DET, nthreads=os.cpu_count() or 1)
            llo = o.fit(maxiter=22, tol=tol)
            assert np.asarray(hist).shape == llo.shape, (np.asarray(hist).shape, llo.shape)
            assert np.array_equal(np.asarray(hist), llo)
            assert np.array_equal(np.concatenate([p[2] for p in parts]), o.lam) and np.array_equal(np.concatenate([p[3] for p in parts]), o.nu)
            for k in ("gamma", "mu", "Sigma", "invSigma", "phi"):
                assert np.array_equal(s[k], getattr(o, k)), k
            eo = o.elbo()[0]
            print("mmctm_fit: elbo device %.17g oracle %.17g rel %.3e" % (m.elbo, eo, abs(m.elbo - eo) / abs(eo)), flush=True)
            assert abs(m.elbo - eo) <= 1e-12 * abs(eo), (m.elbo, eo)
            print("MULTI-RANK PARITY OK mmctm_fit world=%d D=%d iterations=%d converged=%s" % (world, D, len(llo), o.converged), flush=True)
        m.close()
    elif mode == "mmctm":
        K, V = [10, 8, 6], [96, 32, 83]
        full = mmsig.synth.generate(D, K, V, key=5)
        b = mmsig.counts.shard_rows([c[0] for c in full], world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        shard = [mmsig.counts.slice_csr(c, lo, hi) for c in full]
        g0 = mmsig.synth.init_gamma(K, V)
        m = mmsig.MMCTM(K, [0.1] * 3, shard, V=V, gamma0=g0, device=local, comm=comm, D_total=D)
        lls = [m.iterate() for _ in range(3)]
        s = m.state()
        elbo = m.calculate_elbo()[0]
        parts = [None] * world
        dist.gather_object((lo, hi, s["lam"], s["nu"], s["props"]), parts if rank == 0 else None, 0)
        if rank == 0:
            import orc
            o = orc.OracleMMCTM(K, [0.1] * 3, V, full, g0, arith=orc.ARITH_DET, nthreads=os.cpu_count() or 1)
            llo = [o.iterate() for _ in range(3)]
            lam = np.concatenate([p[2] for p in parts]); nu = np.concatenate([p[3] for p in parts])
            props = np.concatenate([p[4] for p in parts])
            assert [p[0] for p in parts] == [int(x) for x in b[:-1]]
            assert np.array_equal(lam, o.lam) and np.array_equal(nu, o.nu) and np.array_equal(props, o.props)
            for k in ("gamma", "mu", "Sigma", "invSigma", "phi"):
                assert np.array_equal(s[k], getattr(o, k)), k
            assert np.array_equal(np.asarray(lls), np.asarray(llo))
            eo = o.elbo()[0]
            assert abs(elbo - eo) <= 1e-12 * abs(eo), (elbo, eo)
            print("MULTI-RANK PARITY OK mmctm world=%d D=%d shards=%s" % (world, D, b.tolist()), flush=True)
        m.close()
    else:
        K, V = 20, 96
        full = mmsig.synth.generate(D, [K], [V], key=5)[0]
        b = mmsig.counts.shard_rows([full[0]], world)
        lo, hi = int(b[rank]), int(b[rank + 1])
        lam0 = mmsig.synth.init_lda_lambda(K, V)
        m = mmsig.LDA(K, 0.1, 0.1, mmsig.counts.slice_csr(full, lo, hi), V=V, lambda0=lam0, device=local, comm=comm, D_total=D)
        lls = [m.iterate() for _ in range(3)]
        s = m.state()
        elbo = m.calculate_elbo()[0]
        parts = [None] * world
        dist.gather_object((lo, hi, s["gamma"]), parts if rank == 0 else None, 0)
        if rank == 0:
            import orc
            o = orc.OracleLDA(K, 0.1, 0.1, V, full, lam0, nthreads=os.cpu_count() or 1)
            llo = [o.iterate() for _ in range(3)]
            gam = np.concatenate([p[2] for p in parts])
            rel = lambda a, c: float(np.max(np.abs(a - c) / np.abs(c)))
            assert rel(gam, o.gamma) < 1e-12 and rel(s["lam"], o.lam) < 1e-12 and rel(np.asarray(lls), np.asarray(llo)) < 1e-12
            eo = o.elbo()[0]
            assert abs(elbo - eo) <= 1e-11 * abs(eo), (elbo, eo)
            print("MULTI-RANK PARITY OK lda world=%d D=%d" % (world, D), flush=True)
        m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
