#!/usr/bin/env python
"""bench.py -- MMCTM E+M iterations/sec on synthetic Poisson counts (BASELINE.json metric).

Default workload (config.workload): configs[3] of BASELINE.json -- MMCTM 3-modality K=[10,8,6] on synthetic
1M samples (SNV96 / SV32 / ID83), FP64, total sample count fixed and sharded over the N GPUs (strong scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--samples D] [--impl reference] [--config 1|2|3|4|5]

N > 1 runs either as one process per GPU under torchrun (RANK / WORLD_SIZE in the environment; NCCL all-gathers)
or, started plainly, as ONE process that drives the N GPUs through the library's group API (peer-memory exchange).
--config selects another BASELINE.json configuration (1: brca-eu fit, 2: LDA K=20, 3: CTM K=10, 5: 64 restarts of
MMCTM([7,7]) on 100k samples); the default and the driver's line are config 4.

One JSON line on stdout (rank 0).  `value`: iterations/sec with counts and state resident in HBM, per-kernel event
timing OFF; `kernels`: the per-kernel split from a second, profiled pass; `e2e`: the same through the public API
with host buffers (counts + state H2D, one iteration, state D2H inside the timed region), pinned and pageable;
`roofline`: the dominant kernel against the measured HBM peak, `roofline_fp64`: the same kernel against the
measured FP64 pipe rate (the roof that actually binds the exact-LD_MMA E-step); `cpu_baseline`: the oracle on the
host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "mmctm_em_iterations_per_sec"
UNIT = "iterations/s"
CONFIGS = {
    4: dict(model="mmctm", K=[10, 8, 6], V=[96, 32, 83], D=1_000_000,
            name="MMCTM 3-modality K=[10,8,6], synthetic Poisson counts, D=%d samples (SNV96/SV32/ID83), FP64, exact LD_MMA "
                 "E-step; BASELINE.json configs[3]"),
    3: dict(model="mmctm", K=[10], V=[96], D=1_000_000,
            name="CTM (MMCTM single modality, K=10), synthetic Poisson counts, D=%d samples x 96 SNV terms, FP64, exact LD_MMA "
                 "E-step; BASELINE.json configs[2]"),
    2: dict(model="lda", K=20, V=96, D=1_000_000,
            name="LDA(20, 0.1, 0.1), synthetic Poisson counts, D=%d samples x 96 SNV terms, FP64; BASELINE.json configs[1]"),
    5: dict(model="restarts", K=[7, 7], V=[96, 32], D=100_000, R=64, maxiter=30,
            name="64 random-restart MMCTM([7,7]) fits (30 iterations each) on synthetic D=%d samples (SNV96/SV32), restarts "
                 "dealt over the GPUs, best-ELBO selection; BASELINE.json configs[4]"),
    1: dict(model="brca", K=[7, 7], V=[96, 48], D=560,
            name="MMCTM([7,7],[0.1,0.1]) on the bundled brca-eu SNV+SV counts (D=%d), fit!(tol=1e-5); BASELINE.json configs[0]"),
}


def algorithmic_bytes(nnz_total, D, MK, M):
    """B_iter = 16 nnz + D (40 MK + 24 M)."""
    return 16.0 * nnz_total + D * (40.0 * MK + 24.0 * M)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []
        self.t0 = self.t1 = None            # the timed region, set by the caller

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) >= 6:
                    self.rows.append([time.perf_counter()] + p)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows, window = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= self.t1], "timed region"
        if not rows:                        # region shorter than one nvidia-smi call
            rows = [r for r in self.rows if self.t0 is not None and r[0] >= self.t0]
            window = "timed region + identical untimed iterations run right after it (timed region shorter than one sample)"
        if not rows:
            rows, window = self.rows, "warm-up + timed region"
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "window": window}


def oracle_model(cfg, counts, nthreads):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    import mmsig
    if cfg["model"] == "lda":
        lam0 = mmsig.synth.init_lda_lambda(cfg["K"], cfg["V"])
        return orc.OracleLDA(cfg["K"], 0.1, 0.1, cfg["V"], counts[0], lam0, arith=orc.ARITH_LITERAL, nthreads=nthreads)
    g0 = mmsig.synth.init_gamma(cfg["K"], cfg["V"])
    return orc.OracleMMCTM(cfg["K"], [0.1] * len(cfg["K"]), cfg["V"], counts, g0, arith=orc.ARITH_LITERAL, nthreads=nthreads)


def cpu_baseline(cfg, D_total, nthreads, steps, warmup, budget_s, single_thread=True, max_samples=None):
    """Oracle (literal restatement of the reference) on the host cores: `warmup` untimed + `steps` timed iterations
    on the first sample_D samples of the SAME corpus, sample_D = the whole corpus if that fits the time budget, else
    the largest prefix that does (the value is then extrapolated linearly in D and flagged)."""
    import mmsig
    Ks = cfg["K"] if isinstance(cfg["K"], list) else [cfg["K"]]
    Vs = cfg["V"] if isinstance(cfg["V"], list) else [cfg["V"]]
    gen = lambda lo, hi: mmsig.synth.generate(D_total, Ks, Vs, lo=lo, hi=hi)            # noqa: E731
    probe_D = min(D_total, 20_000)
    m = oracle_model(cfg, gen(0, probe_D), nthreads)
    m.iterate()
    t = time.perf_counter()
    m.iterate()
    per_sample = (time.perf_counter() - t) / probe_D           # seconds per sample-iteration on all threads
    del m
    # later iterations cost up to ~1.6x the second one (LD_MMA needs more evaluations as the fit proceeds)
    sample_D = int(min(D_total, max(probe_D, budget_s / (1.6 * per_sample * (steps + warmup)))))
    if max_samples:
        sample_D = max(1, min(sample_D, max_samples))
    sample_D -= sample_D % 32 if (sample_D < D_total and sample_D > 64) else 0
    m = oracle_model(cfg, gen(0, sample_D), nthreads)
    for _ in range(warmup):
        m.iterate()
    t = time.perf_counter()
    for _ in range(steps):
        m.iterate()
    dt = (time.perf_counter() - t) / steps
    its = (1.0 / dt) * (sample_D / float(D_total))
    out = {"value": its, "unit": UNIT, "cores": nthreads, "kind": "port", "extrapolated": sample_D < D_total,
           "sample": "oracle (C restatement of src/MMCTM.jl / src/LDA.jl + NLopt LD_MMA, literal arithmetic; the Julia reference "
                     "cannot run here: no julia / libnlopt in the image), OpenMP over samples on %d threads, first %d of the %d "
                     "samples, %d warm-up + %d timed iterations, %.3f s per iteration of the sample%s; the reference itself is "
                     "single-threaded (src/MMCTM.jl:463-465)"
                     % (nthreads, sample_D, D_total, warmup, steps, dt,
                        "" if sample_D == D_total else ", scaled linearly in D (the loop is O(D))"),
           "seconds_per_sample_iteration": dt, "sample_D": sample_D}
    if single_thread:
        # the faithful analogue of the reference's serial loop: one thread, a small prefix, same iterations
        sD = int(min(sample_D, max(2_000, 3.0 / (per_sample * nthreads * (steps + warmup)))))
        m1 = oracle_model(cfg, gen(0, sD), 1)
        for _ in range(warmup):
            m1.iterate()
        t = time.perf_counter()
        for _ in range(steps):
            m1.iterate()
        dt1 = (time.perf_counter() - t) / steps
        out["single_thread"] = {"value": (1.0 / dt1) * (sD / float(D_total)), "unit": UNIT, "cores": 1, "sample_D": sD,
                                "seconds_per_sample_iteration": dt1, "extrapolated": sD < D_total}
    return out


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D = args.samples
    nthreads = os.cpu_count() or 1
    if cfg["model"] in ("restarts", "brca"):
        emit({"impl": "reference", "unavailable": "the reference arm times configs 2, 3 and 4 (one iteration per step)"})
        return
    cb = cpu_baseline(cfg, D, nthreads, steps=max(args.steps, 1), warmup=args.warmup, budget_s=args.cpu_budget_s, single_thread=False,
                      max_samples=args.cpu_samples)
    line = {"impl": "reference", "metric": metric_name(cfg), "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / cb["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, D, args.gpus), "cpu_baseline": cb,
            "measured_ms_per_step_on_sample": cb["seconds_per_sample_iteration"] * 1e3, "sample_D": cb["sample_D"],
            "extrapolated": cb["extrapolated"],
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def metric_name(cfg):
    return "lda_em_iterations_per_sec" if cfg["model"] == "lda" else METRIC


def workload_config(cfg, D, n, mode=""):
    return {"workload": cfg["name"] % D, "samples": D, "K": cfg["K"], "V": cfg["V"],
            "parallelism": "samples sharded over %d GPU(s)%s" % (n, mode),
            "l2": "inputs per iteration (>= 2.6 GB at D=1e6) exceed the 126 MB L2; no flush needed"}


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else a library prints to
    fd 1 meanwhile (NCCL's version banner, for one) has been routed to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def fp64_roofline(dom_ms, evals, Dl, MK):
    """The dominant kernel against the FP64 pipe: FP64 warp-instructions it has to issue (per LD_MMA evaluation and
    coordinate, from the ncu capture in profiles/fp64_model.json) over its duration, against the measured pipe rate
    (profiles/micro/fp64_peak.cu: 1.97 warp-instructions / clock / SM at 1965 MHz)."""
    try:
        fm = json.load(open(os.path.join(ROOT, "profiles", "fp64_model.json")))
    except Exception:
        return None
    lanes = (fm["fp64_lane_inst_per_coordinate_eval"]["nu"] * evals["nu_mean"] +
             fm["fp64_lane_inst_per_coordinate_eval"]["lambda"] * evals["lambda_mean"]) * MK * Dl
    warp_inst = lanes / 32.0
    peak = fm["peak_warp_inst_per_clk_per_sm"] * fm["sms"] * fm["clock_ghz"] * 1e9
    ach = warp_inst / (dom_ms * 1e-3)
    return {"bound": "fp64_pipe", "achieved": ach / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s", "frac": ach / peak,
            "fp64_warp_instructions_per_launch": warp_inst, "peak_source": fm["peak_source"], "model_source": fm["model_source"]}


_T0 = time.perf_counter()


def stage(msg):
    """BENCH_TRACE=<seconds>: progress marks on stderr, and every thread's Python stack after that many seconds
    (where a multi-rank run is waiting, should it ever wait)."""
    if os.environ.get("BENCH_TRACE"):
        sys.stderr.write("[bench rank %s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter() - _T0, msg))
        sys.stderr.flush()


def main():
    if os.environ.get("BENCH_TRACE"):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BENCH_TRACE"]), repeat=True, file=sys.stderr)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=None)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--cpu-budget-s", type=float, default=None, help="wall-clock budget of the CPU arm (default 25 s inside "
                    "the GPU line, 200 s for --impl reference)")
    ap.add_argument("--cpu-samples", type=int, default=None, help="upper bound on the CPU arm's sample (default: what the budget allows)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pageable", action="store_true")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp32: the optional FP32 mode of the tile passes (mmsig_config.precision); the headline is always fp64")
    ap.add_argument("--no-fast", action="store_true", help="skip the extra FP32-mode measurement of the default line")
    ap.add_argument("--e2e-unpipelined", action="store_true", help="e2e through set_data/set_state/iterate/get_state")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.samples is None:
        args.samples = cfg["D"]
    if args.cpu_budget_s is None:
        args.cpu_budget_s = 200.0 if args.impl == "reference" else 25.0
    if args.impl == "reference":
        return run_reference(args, cfg)
    if cfg["model"] == "lda":
        return bench_lda(args, cfg)
    if cfg["model"] == "restarts":
        return bench_restarts(args, cfg)
    if cfg["model"] == "brca":
        return bench_brca(args, cfg)
    return bench_mmctm(args, cfg)


def _torch_setup():
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    return torch, world, rank, local


def bench_mmctm(args, cfg):
    import mmsig
    torch, world, rank, local = _torch_setup()
    K_CFG, V_CFG = cfg["K"], cfg["V"]
    ALPHA = [0.1] * len(K_CFG)
    grouped = world == 1 and args.gpus > 1                # one process drives the N GPUs (mmsig_group_*)
    ngpu = args.gpus if grouped else world
    dist = None
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(mmsig.capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        comm = (bytes(uid.cpu().tolist()), rank, world)

    D = args.samples
    per = -(-D // world)
    lo, hi = min(D, rank * per), min(D, (rank + 1) * per)
    stage("process group up, generating samples %d..%d" % (lo, hi))
    counts = mmsig.synth.generate(D, K_CFG, V_CFG, lo=lo, hi=hi)
    stage("counts generated")
    Dl = hi - lo
    nnz_local = sum(int(c[0][-1]) for c in counts)
    MK, M = sum(K_CFG), len(K_CFG)
    g0 = mmsig.synth.init_gamma(K_CFG, V_CFG)

    # pinned host buffers (the e2e leg copies from / to these)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    keep = []
    counts_p = []
    for r, t, c in counts:
        trip = []
        for a in (r, t, c):
            n, tt = pin(a)
            keep.append(tt)
            trip.append(n)
        counts_p.append(tuple(trip))

    if grouped:
        model = mmsig.MMCTMGroup(K_CFG, ALPHA, counts_p, list(range(ngpu)), V=V_CFG, gamma0=g0, profile=False, precision=args.precision)
        stream = None
    else:
        stream = torch.cuda.Stream()
        model = mmsig.MMCTM(K_CFG, ALPHA, counts_p, V=V_CFG, gamma0=g0, device=local, profile=False, comm=comm, D_total=D,
                            precision=args.precision)
        model.h.set_stream(stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        if grouped:
            for d in range(ngpu):
                torch.cuda.synchronize(d)

    def launches():
        return model.grp.launch_count() if grouped else model.h.launch_count()

    stage("model on the device")
    sampler = ClockSampler(local)
    sampler.start()
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.device(0)
    with ctx:
        for _ in range(args.warmup):
            model.iterate()
        stage("warm-up done")
        barrier()
        sampler.t0 = time.perf_counter()
        l0 = launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if stream is not None:
            e0.record(stream)
        w0 = time.perf_counter()
        # fit!(model; maxiter=steps) continuing from the warmed-up state: exactly `steps` iterations of the loop of
        # src/MMCTM.jl:462-489 (tol = 0 never fires), the closing ELBO left out
        ll = model.fit(maxiter=args.steps, tol=0.0, verbose=False, elbo=False)[-1]
        if stream is not None:
            e1.record(stream)
        stage("timed fit returned")
        barrier()
        sampler.t1 = time.perf_counter()
        # one process per GPU: CUDA events on the launching stream; the single-process group blocks in each call until
        # every member's stream is synchronised, so its step time is the host clock around the calls
        ms = e0.elapsed_time(e1) if stream is not None else (sampler.t1 - w0) * 1e3
        n_launch = launches() - l0
    nev_nu, nev_lam = model.evals()
    evals = {"nu_mean": float(nev_nu.mean()), "nu_max": int(nev_nu.max()), "lambda_mean": float(nev_lam.mean()),
             "lambda_max": int(nev_lam.max())}
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nz = torch.tensor([float(nnz_local)], dtype=torch.float64, device="cuda")
        dist.all_reduce(nz)
        nnz_total = float(nz.item())
    else:
        nnz_total = float(nnz_local)
    ms_step = ms / args.steps
    value = 1000.0 / ms_step
    stage("timed region reduced over ranks")
    st = model.state()              # the e2e leg below repeats the iteration that follows the timed ones
    # second pass with per-kernel event timing ON: the split of the step over the kernels (and what the timing costs)
    ktimes, ms_prof = {}, None
    if not grouped:
        model.h.set_profile(True)
        model.h.kernel_times(reset=True)
        with torch.cuda.stream(stream):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(stream)
            model.fit(maxiter=args.steps, tol=0.0, verbose=False, elbo=False)
            p1.record(stream)
            torch.cuda.synchronize()
            ms_prof = p0.elapsed_time(p1) / args.steps
        ktimes = model.h.kernel_times(reset=True)
        model.h.set_profile(False)
    # the optional FP32 mode of the tile passes on the same workload, reported beside the headline, never as it
    fp32_mode = None
    if world == 1 and not grouped and args.precision == "fp64" and not args.no_fast:
        m32 = mmsig.MMCTM(K_CFG, ALPHA, counts_p, V=V_CFG, gamma0=g0, device=local, profile=False, precision="fp32")
        m32.h.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            for _ in range(args.warmup):
                m32.iterate()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            ll32 = m32.fit(maxiter=args.steps, tol=0.0, verbose=False, elbo=False)[-1]
            f1.record(stream)
            torch.cuda.synchronize()
            ms32 = f0.elapsed_time(f1) / args.steps
            m32.h.set_profile(True)
            m32.h.kernel_times(reset=True)
            m32.fit(maxiter=args.steps, tol=0.0, verbose=False, elbo=False)
            torch.cuda.synchronize()
            k32 = m32.h.kernel_times(reset=True)
        fp32_mode = {"ms_per_step": ms32, "value": 1000.0 / ms32, "unit": UNIT, "ll": [float(x) for x in ll32],
                     "ll_rel_diff_to_fp64": float(np.max(np.abs((np.asarray(ll32) - np.asarray(ll)) / np.asarray(ll)))),
                     "kernels": {k: {"ms_per_step": v[0] / args.steps} for k, v in k32.items() if v[1] > 0},
                     "what": "mmsig_config.precision = MMSIG_PRECISION_FP32: theta / log-likelihood tile passes in float, sums over samples, "
                             "LD_MMA solves and M-step in double; same warm-up and step count from the same constructor state"}
        m32.close()
        del m32
    # a timed region shorter than ~1.5 s can fall between two nvidia-smi samples: keep the identical
    # load running, untimed, on every rank (same count everywhere: the iterations are collective)
    # ms is already the max over ranks; ms_prof is this rank's own clock, so it must not enter a count of collective calls
    n_extra = min(500, int(np.ceil(max(0.0, 1500.0 - 2.0 * ms) / ms_step)))
    stage("per-kernel pass done, %d extra iterations" % n_extra)
    for _ in range(n_extra):
        model.iterate()
    sampler.stop_flag.set()
    stage("extra iterations done")

    # ---- e2e: public API with host buffers; H2D of counts + state, one iteration, D2H of state ----
    lam_h, t1 = pin(np.zeros((Dl, MK)))
    nu_h, t2 = pin(np.ones((Dl, MK)))
    keep += [t1, t2]
    mu_h, Sg_h, iS_h, gam_h = st["mu"], st["Sigma"], st["invSigma"], st["gamma"]
    lam_h[:] = st["lam"]
    nu_h[:] = st["nu"]
    out = {k: pin(np.empty_like(v))[0] for k, v in st.items()}
    h2d = sum(a.nbytes for trip in counts_p for a in trip) + lam_h.nbytes + nu_h.nbytes + mu_h.nbytes + \
        Sg_h.nbytes + iS_h.nbytes + gam_h.nbytes + 8 * M
    d2h = sum(v.nbytes for v in out.values()) + 8 * M
    order = ("lam", "nu", "zeta", "mu", "Sigma", "invSigma", "gamma", "Elnphi", "phi", "props")

    def e2e_step(cnt, lam_, nu_, out_):
        if args.e2e_unpipelined and not grouped:            # the four separate calls, every copy serialised with the kernels
            model._set_data(cnt, D)                                               # counts H2D (+ row packing)
            model.set_state(gam_h, lam=lam_, nu=nu_, mu=mu_h, Sigma=Sg_h, invSigma=iS_h)     # state H2D
            ll_ = model.iterate()                                                 # one E+M iteration, LL D2H
            model.h.check(model.h.lib.mmsig_mmctm_get_state(model.h.h, *[mmsig.capi.dp(out_[k]) for k in order]))
            return ll_
        # what fit!(model; maxiter=1) does through the Julia shim: ONE call, host buffers in and out
        if grouped:
            hist, _ = model.fit_host(cnt, gam_h, lam=lam_, nu=nu_, mu=mu_h, Sigma=Sg_h, invSigma=iS_h, maxiter=1, out=out_)
        else:
            hist, _ = model.fit_host(cnt, gam_h, lam=lam_, nu=nu_, mu=mu_h, Sigma=Sg_h, invSigma=iS_h, maxiter=1,
                                     out=out_, D_total=D)
        return hist[-1]

    def time_e2e(cnt, lam_, nu_, out_, n):
        e2e_step(cnt, lam_, nu_, out_)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            e2e_step(cnt, lam_, nu_, out_)
        barrier()
        ms_ = (time.perf_counter() - t0) * 1000.0 / n
        if dist is not None:
            tt = torch.tensor([ms_], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_ = float(tt.item())
        return ms_

    # the link this leg runs over: one plain pinned-host -> device copy of the largest count array
    big = max((t for t in keep), key=lambda t: t.numel() * t.element_size())
    dev_buf = torch.empty_like(big, device="cuda")
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_buf.copy_(big, non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    dev_buf.copy_(big, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = big.numel() * big.element_size() / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del dev_buf
    # the form julia/MMSigB200.jl hands over: 4-byte records (term | count << 10) in page-locked memory
    counts_pk = []
    for r, t, c in counts_p:
        rec, tt = pin(np.zeros(t.size, np.uint32))
        keep.append(tt)
        mmsig.capi.pack_records(t, c, out=rec)
        counts_pk.append((r, rec))
    h2d_packed = h2d - sum(c.nbytes for _, _, c in counts_p)
    stage("e2e buffers ready")
    e2e_unpacked_ms = time_e2e(counts_p, lam_h, nu_h, out, max(1, args.e2e_steps - 1))
    stage("e2e (unpacked) done")
    e2e_ms = time_e2e(counts_pk, lam_h, nu_h, out, args.e2e_steps)
    stage("e2e (packed) done")
    e2e_pageable_ms = None
    if not args.no_pageable:
        # what a caller with ordinary (pageable) arrays gets, e.g. Julia Vectors that were not allocated through
        # mmsig_host_alloc: the driver stages every copy; same call, same bytes
        cnt_pg = [tuple(np.array(a, copy=True) for a in trip) for trip in counts_pk]
        out_pg = {k: np.empty_like(v) for k, v in st.items()}
        e2e_pageable_ms = time_e2e(cnt_pg, np.array(lam_h, copy=True), np.array(nu_h, copy=True), out_pg, max(1, args.e2e_steps - 1))
        del cnt_pg, out_pg

    if rank != 0:
        model.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    per_kernel = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                  for k, v in ktimes.items() if v[1] > 0}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"]) if per_kernel else None
    alg_local = algorithmic_bytes(nnz_local, Dl, MK, M)
    traffic, traffic_src = None, None
    try:        # DRAM bytes of the dominant kernel from the committed ncu capture, per sample x local samples
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # the dominant kernel's own entry ("k_solve" = the two launches of the lean solver of one iteration), else the first
        # entry that carries its name
        pick = dom if dom in tj["kernels"] else next((n for n in tj["kernels"] if dom and n.startswith(dom)), None)
        if pick is not None:
            traffic = tj["kernels"][pick]["dram_bytes_per_sample"][0] * Dl
            traffic_src = "ncu dram__bytes_read.sum + dram__bytes_write.sum of %s at D=%d (%s), scaled by the local sample count" % (pick, tj["D"], "profiles/traffic.json")
    except Exception:
        pass
    roof, roof64 = None, None
    if dom:
        # the solve is one logical kernel per step (its nu and lambda phases are two launches under one name)
        dms = per_kernel[dom]["ms_per_step"]
        ach = alg_local / (dms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_local, "kernel_ms_per_launch": dms,
                "launches_per_step": per_kernel[dom]["launches_per_step"],
                "iteration_gbs_all_kernels": algorithmic_bytes(nnz_total, D, MK, M) / (ms_step * 1e-3) / 1e9,
                "note": "exact-LD_MMA FP64 mode is FP64-pipe bound, not HBM bound (DESIGN.md): see roofline_fp64"}
        if dom == "k_solve":
            roof64 = fp64_roofline(dms, evals, Dl, MK)
            if roof64:
                roof64["kernel"] = dom
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, D, ngpu, " of ONE process (mmsig_group_*, peer-memory exchange)" if grouped else
                                      (" (one process per GPU, NCCL all-gather)" if world > 1 else "")),
            "timing": ("host clock around the blocking group calls" if grouped else "CUDA events on the launching stream, max over ranks") +
                      "; per-kernel event timing off in the timed region; the steps are one fit!(maxiter=steps) call (the first 10 "
                      "iterations of a fit cannot end it, src/MMCTM.jl:485, so they run without host round trips)",
            "ms_per_step_with_kernel_timing": ms_prof,
            "samples_iterations_per_sec": value * D,
            "nnz_per_sample": nnz_total / D,
            "ll": [float(x) for x in ll],
            "mma_evaluations_per_sample_last_iteration": evals,
            "clocks": sampler.summary(),
            "e2e": {"value": 1000.0 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d_packed), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "h2d_link_gbs_measured": h2d_gbs,
                    "unpacked": {"value": 1000.0 / e2e_unpacked_ms, "ms_per_step": e2e_unpacked_ms, "h2d_bytes_per_step": int(h2d),
                                 "what": "mmsig_mmctm_fit_host: separate int32 term / count arrays (8 bytes per nonzero), pinned"},
                    "pageable": None if e2e_pageable_ms is None else {"value": 1000.0 / e2e_pageable_ms, "ms_per_step": e2e_pageable_ms,
                                                                       "what": "the same call from / to ordinary (pageable) host arrays"},
                    "what": ("set_data + set_state (pinned host -> device), iterate, get_state (device -> pinned host)" if args.e2e_unpipelined else
                             "mmsig_mmctm_fit_host_packed(maxiter=1): counts (4-byte records) + state from pinned host buffers, one E+M "
                             "iteration, state back to pinned host buffers; copies pipelined behind the E-step chunk by chunk")},
            "gpu_launches": int(n_launch),
            "kernels": per_kernel,
            "roofline": roof,
            "roofline_fp64": roof64,
            "fp32_mode": fp32_mode}
    if args.precision == "fp32":
        line["dtype"] = "f32 tile passes, f64 solves and sums (optional FP32 mode; not the headline)"
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(cfg, D, os.cpu_count() or 1, steps=min(max(args.steps, 1), 3), warmup=min(args.warmup, 3),
                                            budget_s=args.cpu_budget_s, max_samples=args.cpu_samples)
    emit(line)
    model.close()
    if dist is not None:
        dist.destroy_process_group()


def bench_lda(args, cfg):
    """config 2: LDA(20, 0.1, 0.1), one GPU (or sharded under torchrun)."""
    import mmsig
    torch, world, rank, local = _torch_setup()
    K, V, D = cfg["K"], cfg["V"], args.samples
    dist, comm = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(mmsig.capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        comm = (bytes(uid.cpu().tolist()), rank, world)
    per = -(-D // world)
    lo, hi = min(D, rank * per), min(D, (rank + 1) * per)
    csr = mmsig.synth.generate(D, [K], [V], lo=lo, hi=hi)[0]
    nnz_local = int(csr[0][-1])
    stream = torch.cuda.Stream()
    m = mmsig.LDA(K, 0.1, 0.1, csr, V=V, lambda0=mmsig.synth.init_lda_lambda(K, V), device=local, comm=comm, D_total=D,
                  precision=args.precision)
    m.h.set_stream(stream.cuda_stream)
    sampler = ClockSampler(local)
    sampler.start()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            m.iterate()
        barrier()
        sampler.t0 = time.perf_counter()
        l0 = m.h.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ll = m.fit(maxiter=args.steps, tol=0.0, verbose=False, elbo=False)[-1]      # fit!(model; maxiter=steps), the closing ELBO left out
        e1.record(stream)
        barrier()
        sampler.t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        n_launch = m.h.launch_count() - l0
        m.h.set_profile(True)
        m.h.kernel_times(reset=True)
        m.fit(maxiter=args.steps, tol=0.0, verbose=False, elbo=False)
        kt = m.h.kernel_times(reset=True)
        m.h.set_profile(False)
        for _ in range(min(2000, int(1500.0 / max(ms / args.steps, 1e-3)))):
            m.iterate()
    sampler.stop_flag.set()
    # e2e: fit!(model::LDA; maxiter=1) from / to host arrays through ONE call (mmsig_lda_fit_host), pinned buffers
    st = m.state()

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    keep = [pin(a) for a in csr] + [pin(st["lam"]), pin(st["gamma"])]
    csr_p, lam_p, gam_p = tuple(k[0] for k in keep[:3]), keep[3][0], keep[4][0]
    h2d = sum(a.nbytes for a in csr_p) + lam_p.nbytes + gam_p.nbytes
    d2h = 3 * lam_p.nbytes + 3 * gam_p.nbytes
    m.fit_host(csr_p, lam_p, gamma_next=gam_p, maxiter=1, D_total=D)
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.e2e_steps)):
        m.fit_host(csr_p, lam_p, gamma_next=gam_p, maxiter=1, D_total=D)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, args.e2e_steps)
    nnz_total = float(nnz_local)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nz = torch.tensor([nnz_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(nz)
        nnz_total = float(nz.item())
    if rank != 0:
        m.close()
        dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_kernel = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in kt.items() if v[1] > 0}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])
    alg = 16.0 * nnz_total + 24.0 * K * D            # counts twice, gamma read + write + read by the LL pass
    alg_local = 16.0 * nnz_local + 24.0 * K * (hi - lo)
    line = {"metric": metric_name(cfg), "value": 1000.0 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(cfg, D, world), "ll": [float(ll)],
            "nnz_per_sample": nnz_total / D, "clocks": sampler.summary(), "gpu_launches": int(n_launch), "kernels": per_kernel,
            "e2e": {"value": 1000.0 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "mmsig_lda_fit_host(maxiter=1): counts + lambda + gamma from pinned host buffers, one iteration, the six state "
                            "arrays back (results land in pageable arrays of the Python mirror)"},
            "roofline": {"bound": "hbm", "kernel": "iteration (all kernels)", "achieved": alg / (ms_step * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms_step * 1e-3) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": alg, "dominant_kernel": dom,
                         "dominant_kernel_gbs": alg_local / (per_kernel[dom]["ms_per_step"] * 1e-3) / 1e9}}
    if args.precision == "fp32":
        line["dtype"] = "f32 tile passes, f64 sums over samples (optional FP32 mode; not the headline)"
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(cfg, D, os.cpu_count() or 1, steps=3, warmup=1, budget_s=args.cpu_budget_s,
                                            max_samples=args.cpu_samples)
    emit(line)
    m.close()
    if dist is not None:
        dist.destroy_process_group()


def bench_restarts(args, cfg):
    """config 5: R restarts of MMCTM([7,7]) on 100k samples, dealt over the GPUs of this process (mmsig_group_mmctm_restarts);
    a step = the whole batch of restarts with a fixed number of iterations each."""
    import mmsig
    torch, world, rank, local = _torch_setup()
    if world > 1:
        raise SystemExit("config 5 runs as one process (python bench.py --config 5 --gpus N): restarts need no rendezvous")
    K, V, D, R, maxiter = cfg["K"], cfg["V"], args.samples, cfg["R"], cfg["maxiter"]
    n = args.gpus
    counts = mmsig.synth.generate(D, K, V)
    rng = np.random.Generator(np.random.Philox(key=7))
    g0s = rng.integers(1, 101, size=(R, sum(k * v for k, v in zip(K, V)))).astype(float)
    sampler = ClockSampler(0)
    sampler.start()
    if n > 1:
        m = mmsig.MMCTMGroup(K, [0.1, 0.1], counts, list(range(n)), V=V, gamma0=g0s[0])
    else:
        m = mmsig.MMCTM(K, [0.1, 0.1], counts, V=V, gamma0=g0s[0])
    m.fit_restarts(g0s[:n], maxiter=3, tol=0.0)          # warm-up: plans, allocations
    steps = max(1, min(args.steps, 2))
    torch.cuda.synchronize()
    sampler.t0 = time.perf_counter()
    for _ in range(steps):
        elbo, ll, nit, best = m.fit_restarts(g0s, maxiter=maxiter, tol=0.0)
    sampler.t1 = time.perf_counter()
    sampler.stop_flag.set()
    dt = (sampler.t1 - sampler.t0) / steps
    its = float(np.sum(nit))
    line = {"metric": METRIC, "value": its / dt, "unit": UNIT, "n_gpus": n, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, D, n, ": restarts dealt over the GPUs, no communication"),
            "timing": "host clock around the blocking restart calls (counts resident per restart batch upload included)",
            "restarts": R, "iterations_per_restart": maxiter, "restarts_per_sec": R / dt, "best_restart": int(best),
            "best_elbo": float(elbo[best]), "clocks": sampler.summary(), "e2e": None, "roofline": None,
            "gpu_launches": int(m.grp.launch_count() if n > 1 else m.h.launch_count())}
    emit(line)
    m.close()


def bench_brca(args, cfg):
    """config 1: the README example on the bundled data (a correctness config; far too small to load a GPU)."""
    import mmsig
    torch, world, rank, local = _torch_setup()
    z = np.load(os.path.join(ROOT, "tests", "golden", "brca_eu_counts.npz"))
    brca = [(z["rowptr0"], z["term0"], z["count0"]), (z["rowptr1"], z["term1"], z["count1"])]
    g0 = mmsig.synth.init_gamma(cfg["K"], cfg["V"])
    m = mmsig.MMCTM(cfg["K"], [0.1, 0.1], brca, V=cfg["V"], gamma0=g0)
    m.fit(maxiter=100, tol=1e-5, verbose=False)
    m.set_state(g0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    hist = m.fit(maxiter=100, tol=1e-5, verbose=False)
    dt = time.perf_counter() - t
    line = {"metric": METRIC, "value": len(hist) / dt, "unit": UNIT, "n_gpus": 1, "steps": len(hist), "warmup": 1,
            "ms_per_step": dt * 1e3 / len(hist), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "bundled brca-eu counts (tests/golden/brca_eu_counts.npz)", "config": workload_config(cfg, 560, 1),
            "converged": bool(m.converged), "elbo": float(m.elbo), "ll": [float(x) for x in hist[-1]], "e2e": None, "roofline": None,
            "gpu_launches": int(m.h.launch_count())}
    emit(line)
    m.close()


if __name__ == "__main__":
    main()
