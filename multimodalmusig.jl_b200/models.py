"""Host-side mirror of the reference's model API for the hot path, over the C ABI.

    reference (Julia)                              here
    MMCTM(K, α, X) / MMCTM(K, α, V, X)             MMCTM(K, alpha, counts, V=None, gamma0=...)
    fit!(model; maxiter, tol, verbose, updateΣ)    model.fit(maxiter=100, tol=1e-4, verbose=True, updateSigma=True)
    model.ϕ / .props / .λ / .ν / .μ / .Σ / .γ      model.phi / .props / .lam / .nu / .mu / .Sigma / .gamma
    model.elbo / .ll / .converged                  same names
    LDA(K, α, η, X) / fit! / .β / .θ               LDA(K, alpha, eta, counts) / .fit / .beta / .theta

Model construction (including the random γ₀ / λ₀, src/MMCTM.jl:59-63, src/LDA.jl:36) stays on the
host, as it does in Julia; the library receives state and runs the iterations.  All state lives
on the GPU between calls; the properties download it.
"""
import ctypes as C

import numpy as np

from . import capi
from .counts import infer_V


class MMCTM:
    """src/MMCTM.jl:1-108.  counts: list over modalities of (rowptr, term0, count)."""

    def __init__(self, K, alpha, counts, V=None, gamma0=None, rng=None, device=0,
                 stop_rule=capi.STOP_NLOPT27, profile=False, comm=None, D_total=None, dense=None,
                 dense_layout=capi.DENSE_TERM_MAJOR, precision=capi.PRECISION_FP64):
        """counts: CSR triples per modality (format_counts_mmctm); or counts=None and dense = list of
        dense integer matrices ((V_m, D) term-major as the TSV files, or (D, V_m) with
        dense_layout=DENSE_SAMPLE_MAJOR), turned into CSR on the GPU (mmsig_mmctm_set_data_dense)."""
        self.K = [int(k) for k in K]
        self.M = len(self.K)
        self.alpha = np.asarray(alpha, dtype=np.float64).copy()
        if len(dense if dense is not None else counts) != self.M or self.alpha.size != self.M:
            raise ValueError("K, alpha and counts must have one entry per modality")
        if dense is not None:
            lay = dense_layout
            dense = [np.ascontiguousarray(x if x.dtype in (np.int32, np.int64) else np.asarray(x, np.int64)) for x in map(np.asarray, dense)]
            if len({x.dtype for x in dense}) != 1:
                dense = [x.astype(np.int64) for x in dense]
            if V is None:
                raise ValueError("V is required with dense counts (scripts/run_mmctm.jl:262 passes it explicitly)")
            self.V = [int(v) for v in V]
            self.D = int(dense[0].shape[1] if lay == capi.DENSE_TERM_MAJOR else dense[0].shape[0])
            counts = [None] * len(dense)
        else:
            self.V = infer_V(counts) if V is None else [int(v) for v in V]      # src/MMCTM.jl:94-108
            self.D = len(counts[0][0]) - 1
        self.MK = sum(self.K)
        self.G = sum(k * v for k, v in zip(self.K, self.V))
        if gamma0 is None:                                                  # init=:random, :59-63
            rng = np.random.default_rng() if rng is None else rng
            gamma0 = rng.integers(1, 101, size=self.G).astype(np.float64)
        self.h = capi.Handle(device=device, stop_rule=stop_rule, profile=profile, precision=precision)   # "fp32": the optional fast mode
        if comm is not None:
            uid, rank, nranks = comm
            self.h.comm_init(uid, rank, nranks)
        if dense is not None:
            self._set_data_dense(dense, dense_layout, self.D if D_total is None else D_total)
        else:
            self._set_data(counts, self.D if D_total is None else D_total)
        self.set_state(gamma=gamma0)
        self.converged = False
        self.elbo = float("nan")
        self.ll = None

    def _set_data_dense(self, dense, layout, D_total):
        M = self.M
        ptrs = (C.c_void_p * M)(*[x.ctypes.data_as(C.c_void_p) for x in dense])
        K = np.asarray(self.K, np.int32)
        V = np.asarray(self.V, np.int32)
        self.h.check(self.h.lib.mmsig_mmctm_set_data_dense(self.h.h, self.D, D_total, M, K.ctypes.data_as(capi.c_i32p),
                                                           V.ctypes.data_as(capi.c_i32p), ptrs, dense[0].dtype.itemsize, layout))
        self.nnz = None

    def _set_data(self, counts, D_total):
        lib = self.h.lib
        keep = [(np.ascontiguousarray(r, np.int64), np.ascontiguousarray(t, np.int32),
                 np.ascontiguousarray(c, np.int32)) for r, t, c in counts]
        M = self.M
        rp = (capi.c_i64p * M)(*[k[0].ctypes.data_as(capi.c_i64p) for k in keep])
        tp = (capi.c_i32p * M)(*[k[1].ctypes.data_as(capi.c_i32p) for k in keep])
        cp = (capi.c_i32p * M)(*[k[2].ctypes.data_as(capi.c_i32p) for k in keep])
        K = np.asarray(self.K, np.int32)
        V = np.asarray(self.V, np.int32)
        self.h.check(lib.mmsig_mmctm_set_data(self.h.h, self.D, D_total, M, K.ctypes.data_as(capi.c_i32p),
                                              V.ctypes.data_as(capi.c_i32p), rp, tp, cp))
        self.nnz = [int(k[0][-1]) for k in keep]

    def set_state(self, gamma, lam=None, nu=None, mu=None, Sigma=None, invSigma=None, alpha=None):
        """Upload variational state; None -> the constructor's value (src/MMCTM.jl:44-46,82-83)."""
        if alpha is not None:
            self.alpha = np.asarray(alpha, dtype=np.float64).copy()
        D, MK = self.D, self.MK
        a = [capi.f64(self.alpha, self.M), capi.f64(gamma, self.G), capi.f64(lam, D * MK), capi.f64(nu, D * MK),
             capi.f64(mu, MK), capi.f64(Sigma, MK * MK), capi.f64(invSigma, MK * MK)]
        self.h.check(self.h.lib.mmsig_mmctm_set_state(self.h.h, *[capi.dp(x) for x in a]))

    def iterate(self, updateSigma=True, flags=None):
        """One body of fit!'s loop (src/MMCTM.jl:463-479); returns the M log-likelihoods.
        flags (capi.FLAG_*) selects the frozen / unsmoothed loop bodies of fit_heldout / transform."""
        ll = np.zeros(self.M)
        if flags is None:
            flags = capi.FLAG_UPDATE_SIGMA if updateSigma else 0
        self.h.check(self.h.lib.mmsig_mmctm_iterate(self.h.h, flags, capi.dp(ll)))
        return ll

    def set_phi(self, phi_flat):
        self.h.check(self.h.lib.mmsig_mmctm_set_phi(self.h.h, capi.dp(capi.f64(phi_flat, self.G))))

    def _loop(self, flags, maxiter, tol, verbose):
        hist = []
        for it in range(1, maxiter + 1):
            hist.append(self.iterate(flags=flags))
            if verbose:
                print("%d\tLog-likelihoods: %s" % (it, ", ".join(repr(float(x)) for x in hist[-1])))
            if len(hist) > 10 and _converged(hist[-2], hist[-1], tol):
                self.converged = True
                break
        self.ll = hist[-1].copy()
        return np.asarray(hist)

    def fit_heldout(self, counts_heldout, maxiter=100, verbose=False, device=0):
        """fit_heldout(Xheldout, model; maxiter=100) (src/MMCTM.jl:554-586): a model on the held-out
        samples with μ, Σ, invΣ, γ, Elnϕ, ϕ of this one; per iteration E-step, props, LL."""
        s = self.state(props=False)
        new = MMCTM(self.K, self.alpha, counts_heldout, V=self.V, gamma0=s["gamma"], device=device)
        new.set_state(s["gamma"], mu=s["mu"], Sigma=s["Sigma"], invSigma=s["invSigma"])
        new.set_phi(s["phi"])
        new.ll_history = new._loop(capi.FLAG_FREEZE_TOPICS | capi.FLAG_FREEZE_MU, maxiter, 1e-4, verbose)
        return new

    def transform(self, counts, maxiter=1000, tol=1e4, fit_gaussian=False, verbose=False, device=0, rng=None):
        """transform(model, X; maxiter=1000, tol=1e4, fit_gaussian=false) (src/MMCTM.jl:511-552):
        unsmoothed θ ∝ exp(λ)·ϕ with this model's ϕ; μ, Σ copied unless fit_gaussian -- and, as in
        the reference, invΣ is NOT copied (it stays the constructor's identity, :517-520)."""
        s = self.state(props=False)
        rng = np.random.default_rng() if rng is None else rng
        g0 = rng.integers(1, 101, size=self.G).astype(np.float64)       # the fresh model's γ (unused by θ)
        new = MMCTM(self.K, self.alpha, counts, V=self.V, gamma0=g0, device=device)
        if not fit_gaussian:
            new.set_state(g0, mu=s["mu"], Sigma=s["Sigma"])
        new.set_phi(s["phi"])
        flags = capi.FLAG_FREEZE_TOPICS | capi.FLAG_UNSMOOTHED | \
            (capi.FLAG_UPDATE_SIGMA if fit_gaussian else capi.FLAG_FREEZE_MU)
        new.ll_history = new._loop(flags, maxiter, tol, verbose)
        return new

    def predict_modality_eta(self, counts_obs, m, maxiter=100, device=0, esteps=None):
        """predict_modality_η(Xobs, m, model; maxiter=100) (src/MMCTM.jl:588-634): fit λ on the observed
        modalities with everything else frozen, then η_u = μ_u + Σ_uo invΣ[o,o] (λ - μ_o).

        How many E-steps: the reference's stopping test (:609-618) reads `props` it never computes -- uninitialised
        but CONSTANT memory (:47-50) -- so its log-likelihood is constant and `length(ll) > 10 && check_convergence`
        fires at iteration 11 whenever that garbage is finite (the usual case), and never when it is NaN / Inf.
        LD_MMA stops early (xtol 1e-4), so λ after 11 sweeps differs from λ after 100.  `esteps` states the count:
        default min(maxiter, 11), the reference's usual behaviour; esteps=maxiter is its behaviour when the
        uninitialised values are not finite."""
        if esteps is None:
            esteps = min(maxiter, 11)
        s = self.state(props=False)
        obsM = [i for i in range(self.M) if i != m]
        ko = np.cumsum([0] + self.K)
        go = np.cumsum([0] + [k * v for k, v in zip(self.K, self.V)])
        un = np.arange(ko[m], ko[m + 1])
        ob = np.concatenate([np.arange(ko[i], ko[i + 1]) for i in obsM])
        g_obs = np.concatenate([s["gamma"][go[i]:go[i + 1]] for i in obsM])
        om = MMCTM([self.K[i] for i in obsM], self.alpha[obsM], counts_obs, V=[self.V[i] for i in obsM],
                   gamma0=g_obs, device=device)
        om.set_state(g_obs, mu=s["mu"][ob], Sigma=s["Sigma"][np.ix_(ob, ob)], invSigma=s["invSigma"][np.ix_(ob, ob)])
        for _ in range(esteps):
            om.iterate(flags=capi.FLAG_FREEZE_TOPICS | capi.FLAG_FREEZE_MU)
        lam = om.lam
        om.close()
        A = s["Sigma"][np.ix_(un, ob)] @ s["invSigma"][np.ix_(ob, ob)]
        return s["mu"][un] + (lam - s["mu"][ob]) @ A.T

    def fit(self, maxiter=100, tol=1e-4, verbose=True, autoalpha=False, updateSigma=True, elbo=True):
        """fit!(model; maxiter=100, tol=1e-4, verbose=true, autoα=false, updateΣ=true), src/MMCTM.jl:457-494.
        elbo=False skips the closing calculate_elbo (:490) -- for timing the loop alone."""
        flags = (capi.FLAG_UPDATE_SIGMA if updateSigma else 0) | (capi.FLAG_AUTO_ALPHA if autoalpha else 0)
        if verbose:
            hist = []
            for it in range(1, maxiter + 1):
                ll = self.iterate(flags=flags)
                hist.append(ll)
                print("%d\tLog-likelihoods: %s" % (it, ", ".join(repr(float(x)) for x in ll)))   # :482
                if len(hist) > 10 and _converged(hist[-2], hist[-1], tol):
                    self.converged = True
                    break
            hist = np.asarray(hist)
        else:
            buf = np.zeros((maxiter, self.M))
            n, conv = C.c_int32(), C.c_int32()
            self.h.check(self.h.lib.mmsig_mmctm_fit(self.h.h, maxiter, tol, flags, capi.dp(buf),
                                                    C.byref(n), C.byref(conv)))
            hist = buf[:n.value].copy()
            self.converged = bool(conv.value)
        if autoalpha:
            a = np.zeros(self.M)
            self.h.check(self.h.lib.mmsig_mmctm_get_alpha(self.h.h, capi.dp(a)))
            self.alpha = a
        if elbo:
            self.elbo = self.calculate_elbo()[0]      # :490
        self.ll = hist[-1].copy()                     # :491
        return hist

    def fit_host(self, counts, gamma, lam=None, nu=None, mu=None, Sigma=None, invSigma=None, maxiter=100, tol=1e-4,
                 updateSigma=True, flags=None, out=None, D_total=None):
        """fit! in ONE library call from / to host arrays (mmsig_mmctm_fit_host): the same results as
        _set_data + set_state + fit + state(), with the host<->device copies pipelined behind the
        E-step chunk by chunk.  counts / gamma / lam ...: what _set_data and set_state take.
        out: optional dict of preallocated (e.g. page-locked) result arrays keyed as state().
        Returns (ll_history, out)."""
        packed = len(counts[0]) == 2            # (rowptr, rec) with rec = capi.pack_records(term, count): 4-byte records
        M = self.M
        if packed:
            keep = [(np.ascontiguousarray(r, np.int64), np.ascontiguousarray(x, np.uint32)) for r, x in counts]
            tp = (capi.c_u32p * M)(*[k[1].ctypes.data_as(capi.c_u32p) for k in keep])
        else:
            keep = [(np.ascontiguousarray(r, np.int64), np.ascontiguousarray(t, np.int32),
                     np.ascontiguousarray(c, np.int32)) for r, t, c in counts]
            tp = (capi.c_i32p * M)(*[k[1].ctypes.data_as(capi.c_i32p) for k in keep])
            cp = (capi.c_i32p * M)(*[k[2].ctypes.data_as(capi.c_i32p) for k in keep])
        D = len(keep[0][0]) - 1
        rp = (capi.c_i64p * M)(*[k[0].ctypes.data_as(capi.c_i64p) for k in keep])
        K = np.asarray(self.K, np.int32)
        V = np.asarray(self.V, np.int32)
        MK, G = self.MK, self.G
        a = [capi.f64(self.alpha, M), capi.f64(gamma, G), capi.f64(lam, D * MK), capi.f64(nu, D * MK),
             capi.f64(mu, MK), capi.f64(Sigma, MK * MK), capi.f64(invSigma, MK * MK)]
        if out is None:
            out = dict(lam=np.empty((D, MK)), nu=np.empty((D, MK)), zeta=np.empty((D, M)), mu=np.empty(MK),
                       Sigma=np.empty((MK, MK)), invSigma=np.empty((MK, MK)), gamma=np.empty(G), Elnphi=np.empty(G),
                       phi=np.empty(G), props=np.empty((D, MK)))
        order = ("lam", "nu", "zeta", "mu", "Sigma", "invSigma", "gamma", "Elnphi", "phi", "props")
        if flags is None:
            flags = capi.FLAG_UPDATE_SIGMA if updateSigma else 0
        hist = np.zeros((maxiter, M))
        n, conv = C.c_int32(), C.c_int32()
        fn = self.h.lib.mmsig_mmctm_fit_host_packed if packed else self.h.lib.mmsig_mmctm_fit_host
        self.h.check(fn(
            self.h.h, D, D if D_total is None else D_total, M, K.ctypes.data_as(capi.c_i32p),
            V.ctypes.data_as(capi.c_i32p), rp, *((tp,) if packed else (tp, cp)), *[capi.dp(x) for x in a], maxiter, tol, flags,
            capi.dp(hist), C.byref(n), C.byref(conv), *[capi.dp(out.get(k)) for k in order]))
        self.D = D
        self.nnz = [int(k[0][-1]) for k in keep]
        self.converged = bool(conv.value)
        self.ll = hist[n.value - 1].copy()
        return hist[:n.value].copy(), out

    def fit_restarts(self, gamma0s, maxiter=100, tol=1e-4, updateSigma=True):
        """R independent restarts from the constructor state with gamma0s[r] on the resident counts;
        keeps the best-ELBO fit in the model (README.md:42; scripts/run_mmctm.jl:77-111).
        Returns (elbo[R], ll[R, M], n_iter[R], best)."""
        g = capi.f64(np.asarray(gamma0s, dtype=np.float64).reshape(-1, self.G))
        R = g.shape[0]
        elbo, ll = np.zeros(R), np.zeros((R, self.M))
        nit, best = np.zeros(R, np.int32), C.c_int32()
        self.h.check(self.h.lib.mmsig_mmctm_restarts(
            self.h.h, R, capi.dp(g), maxiter, tol, capi.FLAG_UPDATE_SIGMA if updateSigma else 0,
            capi.dp(elbo), capi.dp(ll), nit.ctypes.data_as(capi.c_i32p), C.byref(best)))
        self.elbo, self.ll = float(elbo[best.value]), ll[best.value].copy()
        return elbo, ll, nit, best.value

    def calculate_elbo(self):
        e = C.c_double()
        t = np.zeros(7)
        self.h.check(self.h.lib.mmsig_mmctm_elbo(self.h.h, C.byref(e), capi.dp(t)))
        return e.value, t

    def state(self, props=True):
        D, MK, M, G = self.D, self.MK, self.M, self.G
        out = dict(lam=np.empty((D, MK)), nu=np.empty((D, MK)), zeta=np.empty((D, M)), mu=np.empty(MK),
                   Sigma=np.empty((MK, MK)), invSigma=np.empty((MK, MK)), gamma=np.empty(G), Elnphi=np.empty(G),
                   phi=np.empty(G), props=np.empty((D, MK)) if props else None)
        order = ("lam", "nu", "zeta", "mu", "Sigma", "invSigma", "gamma", "Elnphi", "phi", "props")
        self.h.check(self.h.lib.mmsig_mmctm_get_state(self.h.h, *[capi.dp(out[k]) for k in order]))
        return out

    def _get(self, name):
        return self.state(props=(name == "props"))[name]

    lam = property(lambda s: s._get("lam"))
    nu = property(lambda s: s._get("nu"))
    zeta = property(lambda s: s._get("zeta"))
    mu = property(lambda s: s._get("mu"))
    Sigma = property(lambda s: s._get("Sigma"))
    invSigma = property(lambda s: s._get("invSigma"))
    gamma = property(lambda s: s._get("gamma"))
    Elnphi = property(lambda s: s._get("Elnphi"))

    def _split(self, flat):
        out, o = [], 0
        for k, v in zip(self.K, self.V):
            out.append(flat[o:o + k * v].reshape(k, v))
            o += k * v
        return out

    @property
    def phi(self):
        """model.ϕ[m][k] -> list over m of (K_m, V_m) arrays."""
        return self._split(self._get("phi"))

    @property
    def props(self):
        """model.props[d][m] -> list over m of (D, K_m) arrays."""
        p = self._get("props")
        o = np.cumsum([0] + self.K)
        return [p[:, o[m]:o[m + 1]] for m in range(self.M)]

    def theta(self, m):
        """model.θ[d][m] for all d: (nnz_m, K_m), recomputed from the last E-step's inputs."""
        out = np.empty((self.nnz[m], self.K[m]))
        self.h.check(self.h.lib.mmsig_mmctm_get_theta(self.h.h, m, capi.dp(out)))
        return out

    def evals(self):
        a = np.zeros(self.D, np.int32)
        b = np.zeros(self.D, np.int32)
        self.h.check(self.h.lib.mmsig_mmctm_get_evals(self.h.h, a.ctypes.data_as(capi.c_i32p),
                                                      b.ctypes.data_as(capi.c_i32p)))
        return a, b

    def close(self):
        self.h.close()


def _converged(prev, cur, tol):
    """check_convergence, src/common.jl:48-51."""
    with np.errstate(all="ignore"):
        r = np.max(np.abs(np.asarray(prev) - np.asarray(cur)) / np.abs(np.asarray(cur)))
    return bool(r < tol)


class IMMCTM(MMCTM):
    """src/IMMCTM.jl:1-108: an MMCTM whose topics factorise over features.
    features: list over modalities of (V_m, I_m) integer arrays, 0-BASED feature values
    (model.features[m] .- 1); alpha: one value per modality (src/IMMCTM.jl:93-100) or per (modality,
    feature); gammaf0: flat [m][k][i][j] table (the constructor draws rand(1:100), :60-67).
    fit / iterate / calculate_elbo / state are the MMCTM's; state()['gamma'] is a placeholder."""

    def __init__(self, K, alpha, features, counts, gammaf0=None, rng=None, **kw):
        self.features = [np.ascontiguousarray(f, dtype=np.int32) for f in features]
        self.I = [f.shape[1] for f in self.features]
        self.J = [[int(f[:, i].max()) + 1 for i in range(f.shape[1])] for f in self.features]
        self.alphaf = np.concatenate([np.full(n, float(a)) if np.ndim(a) == 0 else np.asarray(a, float)
                                      for a, n in zip(alpha, self.I)])
        self.T = sum(int(k) * sum(j) for k, j in zip(K, self.J))
        if gammaf0 is None:
            rng = np.random.default_rng() if rng is None else rng
            gammaf0 = rng.integers(1, 101, size=self.T).astype(np.float64)
        self._gammaf0 = np.ascontiguousarray(gammaf0, dtype=np.float64)
        V = [f.shape[0] for f in self.features]
        self._in_ctor = True          # MMCTM.__init__ calls set_state with its K x V placeholder
        super().__init__(K, [1.0] * len(V), counts, V=V, gamma0=np.ones(sum(int(k) * v for k, v in zip(K, V))), **kw)
        self._in_ctor = False

    def _set_data(self, counts, D_total):
        super()._set_data(counts, D_total)
        M = self.M
        nf = np.asarray(self.I, np.int32)
        fp = (capi.c_i32p * M)(*[f.ctypes.data_as(capi.c_i32p) for f in self.features])
        self.h.check(self.h.lib.mmsig_immctm_set_features(self.h.h, nf.ctypes.data_as(capi.c_i32p), fp))

    def set_state(self, gamma=None, lam=None, nu=None, mu=None, Sigma=None, invSigma=None, alpha=None):
        """gamma: the flat feature table [m][k][i][j] (None / the constructor's placeholder -> gammaf0)."""
        g = self._gammaf0 if gamma is None or self._in_ctor else capi.f64(gamma, self.T)
        if alpha is not None:
            self.alphaf = np.asarray(alpha, dtype=np.float64).copy()
        D, MK = self.D, self.MK
        a = [capi.f64(self.alphaf, sum(self.I)), capi.f64(g, self.T), capi.f64(lam, D * MK), capi.f64(nu, D * MK),
             capi.f64(mu, MK), capi.f64(Sigma, MK * MK), capi.f64(invSigma, MK * MK)]
        self.h.check(self.h.lib.mmsig_immctm_set_state(self.h.h, *[capi.dp(x) for x in a]))

    def fit_heldout(self, counts_heldout, maxiter=100, verbose=False, device=0):
        """fit_heldout(Xheldout, model::IMMCTM; maxiter=100) (src/IMMCTM.jl:547-579): a model on the
        held-out samples with this one's μ, Σ, invΣ and feature tables; per iteration E-step and LL."""
        s, t = self.state(props=False), self.tables()
        new = IMMCTM(self.K, [t["alphaf"][sum(self.I[:m]):sum(self.I[:m + 1])] for m in range(self.M)], self.features,
                     counts_heldout, gammaf0=t["gammaf"], device=device)
        new.set_state(t["gammaf"], mu=s["mu"], Sigma=s["Sigma"], invSigma=s["invSigma"])
        new.ll_history = new._loop(capi.FLAG_FREEZE_TOPICS | capi.FLAG_FREEZE_MU, maxiter, 1e-4, verbose)
        return new

    def transform(self, *a, **k):
        raise NotImplementedError("the reference defines no transform for the IMMCTM")

    def fit_restarts(self, gammaf0s, maxiter=100, tol=1e-4, updateSigma=True):
        """R independent restarts from the constructor state with the feature tables gammaf0s[r]."""
        G, self.G = self.G, self.T            # MMCTM.fit_restarts sizes a restart's table by self.G
        try:
            return super().fit_restarts(gammaf0s, maxiter=maxiter, tol=tol, updateSigma=updateSigma)
        finally:
            self.G = G

    def _table_slices(self):
        """(start, stop) of every modality in the flat [m][k][i][j] tables and in alphaf."""
        t = np.cumsum([0] + [int(k) * sum(j) for k, j in zip(self.K, self.J)])
        a = np.cumsum([0] + self.I)
        return t, a

    def predict_modality_eta(self, counts_obs, m, maxiter=100, device=0, esteps=None):
        """predict_modality_η(Xobs, m, model::IMMCTM; maxiter=100) (src/IMMCTM.jl:581-627): fit λ on the
        observed modalities with their feature tables and the Gaussian prior frozen, then
        η_u = μ_u + Σ_uo invΣ[o,o] (λ - μ_o).  `esteps` as for the MMCTM's (default min(maxiter, 11))."""
        if esteps is None:
            esteps = min(maxiter, 11)
        s, t = self.state(props=False), self.tables()
        obsM = [i for i in range(self.M) if i != m]
        ko = np.cumsum([0] + self.K)
        ts, asl = self._table_slices()
        un = np.arange(ko[m], ko[m + 1])
        ob = np.concatenate([np.arange(ko[i], ko[i + 1]) for i in obsM])
        g_obs = np.concatenate([t["gammaf"][ts[i]:ts[i + 1]] for i in obsM])
        om = IMMCTM([self.K[i] for i in obsM], [t["alphaf"][asl[i]:asl[i + 1]] for i in obsM],
                    [self.features[i] for i in obsM], counts_obs, gammaf0=g_obs, device=device)
        om.set_state(g_obs, mu=s["mu"][ob], Sigma=s["Sigma"][np.ix_(ob, ob)], invSigma=s["invSigma"][np.ix_(ob, ob)])
        for _ in range(esteps):
            om.iterate(flags=capi.FLAG_FREEZE_TOPICS | capi.FLAG_FREEZE_MU)
        lam = om.lam
        om.close()
        A = s["Sigma"][np.ix_(un, ob)] @ s["invSigma"][np.ix_(ob, ob)]
        return s["mu"][un] + (lam - s["mu"][ob]) @ A.T

    def tables(self):
        g, e, a = np.empty(self.T), np.empty(self.T), np.empty(sum(self.I))
        self.h.check(self.h.lib.mmsig_immctm_get_tables(self.h.h, capi.dp(g), capi.dp(e), capi.dp(a)))
        return dict(gammaf=g, Elnphif=e, alphaf=a)


class LDA:
    """src/LDA.jl:1-67.  counts: (rowptr, term0, count)."""

    def __init__(self, K, alpha, eta, counts, V=None, lambda0=None, rng=None, device=0, profile=False,
                 comm=None, D_total=None, dense=None, dense_layout=capi.DENSE_TERM_MAJOR, precision=capi.PRECISION_FP64):
        """counts: CSR triple (format_counts_lda); or counts=None and dense = the (V, D) term-major
        (or (D, V) sample-major) integer matrix, turned into CSR on the GPU (mmsig_lda_set_data_dense)."""
        self.K = int(K)
        self.alpha, self.eta = float(alpha), float(eta)
        if dense is not None:
            dense = np.asarray(dense)
            dense = np.ascontiguousarray(dense if dense.dtype in (np.int32, np.int64) else dense.astype(np.int64))
            Vd, Dd = dense.shape if dense_layout == capi.DENSE_TERM_MAJOR else dense.shape[::-1]
            self.V = int(Vd) if V is None else int(V)
            self.D, self.nnz = int(Dd), None
        else:
            r, t, c = counts
            self.V = (int(np.max(t)) + 1 if len(t) else 0) if V is None else int(V)   # src/LDA.jl:57-67
            self.D = len(r) - 1
            self.nnz = int(r[-1])
        if lambda0 is None:                                                       # src/LDA.jl:36
            rng = np.random.default_rng() if rng is None else rng
            lambda0 = rng.integers(1, 101, size=self.K * self.V).astype(np.float64)
        self.h = capi.Handle(device=device, profile=profile, precision=precision)
        if comm is not None:
            uid, rank, nranks = comm
            self.h.comm_init(uid, rank, nranks)
        if dense is not None:
            self.h.check(self.h.lib.mmsig_lda_set_data_dense(self.h.h, self.D, self.D if D_total is None else D_total,
                                                             self.K, self.V, dense.ctypes.data_as(C.c_void_p),
                                                             dense.dtype.itemsize, dense_layout))
        else:
            keep = (np.ascontiguousarray(r, np.int64), np.ascontiguousarray(t, np.int32), np.ascontiguousarray(c, np.int32))
            self.h.check(self.h.lib.mmsig_lda_set_data(self.h.h, self.D, self.D if D_total is None else D_total,
                                                       self.K, self.V, keep[0].ctypes.data_as(capi.c_i64p),
                                                       keep[1].ctypes.data_as(capi.c_i32p), keep[2].ctypes.data_as(capi.c_i32p)))
        self._init_state(lambda0)
        self.converged = False
        self.elbo = float("nan")
        self.ll = float("nan")

    def _init_state(self, lambda0):
        self.set_state(lambda0)

    def set_state(self, lam, gamma_next=None):
        self.h.check(self.h.lib.mmsig_lda_set_state(self.h.h, self.alpha, self.eta,
                                                    capi.dp(capi.f64(lam, self.K * self.V)),
                                                    capi.dp(capi.f64(gamma_next, self.D * self.K))))

    def iterate(self, flags=0):
        ll = C.c_double()
        self.h.check(self.h.lib.mmsig_lda_iterate_flags(self.h.h, flags, C.byref(ll)))
        return ll.value

    def set_beta(self, beta):
        self.h.check(self.h.lib.mmsig_lda_set_beta(self.h.h, capi.dp(capi.f64(beta, self.K * self.V))))

    def _loop(self, flags, maxiter, tol, verbose):
        hist = []
        for it in range(1, maxiter + 1):
            hist.append(self.iterate(flags))
            if verbose:
                print("%d\tLog-likelihood: %r" % (it, hist[-1]))
            if len(hist) > 10 and _converged([hist[-2]], [hist[-1]], tol):
                self.converged = True
                break
        self.ll = float(hist[-1])
        return np.asarray(hist)

    def transform(self, counts, maxiter=1000, tol=1e-4, verbose=False, device=0):
        """transform(model, X) (src/LDA.jl:233-263): θ of new samples under this model's β
        (unsmoothed ϕ ∝ exp(Elnθ)·β); returns θ as (D, K)."""
        s = self.state()
        new = LDA(self.K, self.alpha, self.eta, counts, V=self.V, lambda0=np.ones(self.K * self.V), device=device)
        new.set_beta(s["beta"])
        new._loop(capi.FLAG_FREEZE_TOPICS | capi.FLAG_UNSMOOTHED, maxiter, tol, verbose)
        th = new.theta
        new.close()
        return th

    def fit_heldout(self, counts_heldout, maxiter=100, verbose=False, device=0):
        """fit_heldout(Xheldout, model; maxiter=100) (src/LDA.jl:265-295): λ, β, Elnβ frozen."""
        s = self.state()
        new = LDA(self.K, self.alpha, self.eta, counts_heldout, V=self.V, lambda0=s["lam"], device=device)
        new.set_beta(s["beta"])
        new.ll_history = new._loop(capi.FLAG_FREEZE_TOPICS, maxiter, 1e-4, verbose)
        new.elbo = new.calculate_elbo()[0]
        return new

    def fit_host(self, counts, lambda0, gamma_next=None, maxiter=1000, tol=1e-4, D_total=None):
        """fit!(model::LDA) in ONE library call from / to host arrays (mmsig_lda_fit_host).  Returns (ll_history, state)."""
        r, t, c = (np.ascontiguousarray(counts[0], np.int64), np.ascontiguousarray(counts[1], np.int32),
                   np.ascontiguousarray(counts[2], np.int32))
        D, K, V = len(r) - 1, self.K, self.V
        out = dict(lam=np.empty(K * V), Elnbeta=np.empty(K * V), beta=np.empty(K * V), gamma=np.empty((D, K)),
                   Elntheta=np.empty((D, K)), theta=np.empty((D, K)))
        hist = np.zeros(maxiter)
        n, conv = C.c_int32(), C.c_int32()
        self.h.check(self.h.lib.mmsig_lda_fit_host(
            self.h.h, D, D if D_total is None else D_total, K, V, r.ctypes.data_as(capi.c_i64p), t.ctypes.data_as(capi.c_i32p),
            c.ctypes.data_as(capi.c_i32p), self.alpha, self.eta, capi.dp(capi.f64(lambda0, K * V)),
            capi.dp(capi.f64(gamma_next, D * K)), maxiter, tol, capi.dp(hist), C.byref(n), C.byref(conv),
            *[capi.dp(out[k]) for k in ("lam", "Elnbeta", "beta", "gamma", "Elntheta", "theta")]))
        self.D, self.nnz = D, int(r[-1])
        self.converged, self.ll = bool(conv.value), float(hist[n.value - 1])
        return hist[:n.value].copy(), out

    def fit(self, maxiter=1000, tol=1e-4, verbose=True, elbo=True):
        """fit!(model::LDA; maxiter=1000, tol=1e-4, verbose=true), src/LDA.jl:198-224."""
        if verbose:
            hist = []
            for it in range(1, maxiter + 1):
                hist.append(self.iterate())
                print("%d\tLog-likelihood: %r" % (it, hist[-1]))               # :212
                if len(hist) > 10 and _converged([hist[-2]], [hist[-1]], tol):
                    self.converged = True
                    break
            hist = np.asarray(hist)
        else:
            buf = np.zeros(maxiter)
            n, conv = C.c_int32(), C.c_int32()
            self.h.check(self.h.lib.mmsig_lda_fit(self.h.h, maxiter, tol, capi.dp(buf), C.byref(n), C.byref(conv)))
            hist = buf[:n.value].copy()
            self.converged = bool(conv.value)
        if elbo:
            self.elbo = self.calculate_elbo()[0]
        self.ll = float(hist[-1])
        return hist

    def calculate_elbo(self):
        e = C.c_double()
        t = np.zeros(7)
        self.h.check(self.h.lib.mmsig_lda_elbo(self.h.h, C.byref(e), capi.dp(t)))
        return e.value, t

    def state(self):
        K, V, D = self.K, self.V, self.D
        out = dict(lam=np.empty((K, V)), Elnbeta=np.empty((K, V)), beta=np.empty((K, V)),
                   gamma=np.empty((D, K)), Elntheta=np.empty((D, K)), theta=np.empty((D, K)))
        order = ("lam", "Elnbeta", "beta", "gamma", "Elntheta", "theta")
        self.h.check(self.h.lib.mmsig_lda_get_state(self.h.h, *[capi.dp(out[k]) for k in order]))
        return out

    beta = property(lambda s: s.state()["beta"])       # [k][v] (Julia: V x K)
    theta = property(lambda s: s.state()["theta"])     # [d][k] (Julia: K x D)
    lam = property(lambda s: s.state()["lam"])
    gamma = property(lambda s: s.state()["gamma"])

    def phi(self):
        out = np.empty((self.nnz, self.K))
        self.h.check(self.h.lib.mmsig_lda_get_phi(self.h.h, capi.dp(out)))
        return out

    def close(self):
        self.h.close()


class ILDA(LDA):
    """src/ILDA.jl:1-58: LDA whose topics factorise over features, β_kv = Π_i β_i[f(v,i), k].
    features: (V, I) integer array, 0-BASED values (model.features .- 1); eta: scalar or one per feature;
    lambdaf0: flat [k][i][j] table (the constructor's rand(1:100, J[i], K), :38)."""

    def __init__(self, K, alpha, eta, features, counts, lambdaf0=None, rng=None, **kw):
        f = np.ascontiguousarray(features, dtype=np.int32)
        if f.ndim != 2:
            raise ValueError("features must be a (V, I) matrix")
        self._features = f
        self.I = int(f.shape[1])
        self.J = [int(f[:, i].max()) + 1 for i in range(self.I)]
        self.T = int(K) * sum(self.J)
        self.etaf = np.full(self.I, float(eta)) if np.ndim(eta) == 0 else np.ascontiguousarray(eta, dtype=np.float64)
        if self.etaf.size != self.I:
            raise ValueError("one eta per feature")
        if lambdaf0 is None:
            rng = np.random.default_rng() if rng is None else rng
            lambdaf0 = rng.integers(1, 101, size=self.T).astype(np.float64)
        self._lambdaf0 = lambdaf0
        super().__init__(K, alpha, float(self.etaf[0]), counts, V=f.shape[0], lambda0=np.ones(int(K) * f.shape[0]), **kw)

    def _init_state(self, lambda0):
        self.h.check(self.h.lib.mmsig_ilda_set_features(self.h.h, self.I, self._features.ctypes.data_as(capi.c_i32p)))
        self.set_state(self._lambdaf0)

    def set_state(self, lambdaf, gamma_next=None):
        self.h.check(self.h.lib.mmsig_ilda_set_state(self.h.h, self.alpha, capi.dp(self.etaf),
                                                     capi.dp(capi.f64(lambdaf, self.T)),
                                                     capi.dp(capi.f64(gamma_next, self.D * self.K))))

    def tables(self):
        """(λ, Elnβ) feature tables, flat [k][i][j]."""
        lf, ef = np.empty(self.T), np.empty(self.T)
        self.h.check(self.h.lib.mmsig_ilda_get_tables(self.h.h, capi.dp(lf), capi.dp(ef)))
        return lf, ef

    def table(self, flat, k, i):
        o = k * sum(self.J) + sum(self.J[:i])
        return flat[o:o + self.J[i]]

    def set_beta(self, beta):
        raise NotImplementedError("the ILDA's β follows its feature tables")

    def transform(self, *a, **k):
        raise NotImplementedError("the reference's ILDA has no working transform (src/ILDA.jl)")

    def fit_heldout(self, counts_heldout, maxiter=100, verbose=False, device=0):
        """fit_heldout(Xheldout, model::ILDA; maxiter=100): the feature tables frozen."""
        lf, _ = self.tables()
        new = ILDA(self.K, self.alpha, self.etaf, self._features, counts_heldout, lambdaf0=lf, device=device)
        new.ll_history = new._loop(capi.FLAG_FREEZE_TOPICS, maxiter, 1e-4, verbose)
        new.elbo = new.calculate_elbo()[0]
        return new


class MMCTMGroup:
    """MMCTM over several GPUs of ONE process (mmsig_group_*): `fit!(model; devices=0:7)` of the Julia shim.
    Same arrays as MMCTM for the whole corpus; the library shards the samples (contiguous, balanced by
    nonzeros) and every device ends each iteration with bit-identical tables."""

    def __init__(self, K, alpha, counts, devices, V=None, gamma0=None, rng=None, stop_rule=capi.STOP_NLOPT27, profile=False,
                 precision=capi.PRECISION_FP64):
        self.K = [int(k) for k in K]
        self.M = len(self.K)
        self.alpha = np.asarray(alpha, dtype=np.float64).copy()
        self.V = infer_V(counts) if V is None else [int(v) for v in V]
        self.D = len(counts[0][0]) - 1
        self.MK = sum(self.K)
        self.G = sum(k * v for k, v in zip(self.K, self.V))
        if gamma0 is None:
            rng = np.random.default_rng() if rng is None else rng
            gamma0 = rng.integers(1, 101, size=self.G).astype(np.float64)
        self.grp = capi.Group(devices, stop_rule=stop_rule, profile=profile, precision=precision)
        self._csr(counts)
        lib = self.grp.lib
        self.grp.check(lib.mmsig_group_mmctm_set_data(self.grp.g, self.D, self.M, *self._kv, *self._ptrs))
        self.set_state(gamma0)
        self.converged, self.elbo, self.ll = False, float("nan"), None

    def _csr(self, counts):
        M = self.M
        self._packed = len(counts[0]) == 2      # (rowptr, rec): 4-byte records (capi.pack_records)
        if self._packed:
            keep = [(np.ascontiguousarray(r, np.int64), np.ascontiguousarray(x, np.uint32)) for r, x in counts]
            self._ptrs = ((capi.c_i64p * M)(*[k[0].ctypes.data_as(capi.c_i64p) for k in keep]),
                          (capi.c_u32p * M)(*[k[1].ctypes.data_as(capi.c_u32p) for k in keep]))
        else:
            keep = [(np.ascontiguousarray(r, np.int64), np.ascontiguousarray(t, np.int32), np.ascontiguousarray(c, np.int32))
                    for r, t, c in counts]
            self._ptrs = ((capi.c_i64p * M)(*[k[0].ctypes.data_as(capi.c_i64p) for k in keep]),
                          (capi.c_i32p * M)(*[k[1].ctypes.data_as(capi.c_i32p) for k in keep]),
                          (capi.c_i32p * M)(*[k[2].ctypes.data_as(capi.c_i32p) for k in keep]))
        self._keep = keep
        self._K32, self._V32 = np.asarray(self.K, np.int32), np.asarray(self.V, np.int32)
        self._kv = (self._K32.ctypes.data_as(capi.c_i32p), self._V32.ctypes.data_as(capi.c_i32p))

    def set_state(self, gamma, lam=None, nu=None, mu=None, Sigma=None, invSigma=None):
        D, MK = self.D, self.MK
        a = [capi.f64(self.alpha, self.M), capi.f64(gamma, self.G), capi.f64(lam, D * MK), capi.f64(nu, D * MK),
             capi.f64(mu, MK), capi.f64(Sigma, MK * MK), capi.f64(invSigma, MK * MK)]
        self.grp.check(self.grp.lib.mmsig_group_mmctm_set_state(self.grp.g, *[capi.dp(x) for x in a]))

    def iterate(self, updateSigma=True, flags=None):
        ll = np.zeros(self.M)
        if flags is None:
            flags = capi.FLAG_UPDATE_SIGMA if updateSigma else 0
        self.grp.check(self.grp.lib.mmsig_group_mmctm_iterate(self.grp.g, flags, capi.dp(ll)))
        return ll

    def fit(self, maxiter=100, tol=1e-4, updateSigma=True, verbose=False, elbo=True):
        hist = np.zeros((maxiter, self.M))
        n, conv = C.c_int32(0), C.c_int32(0)
        flags = capi.FLAG_UPDATE_SIGMA if updateSigma else 0
        self.grp.check(self.grp.lib.mmsig_group_mmctm_fit(self.grp.g, maxiter, tol, flags, capi.dp(hist), C.byref(n), C.byref(conv)))
        hist = hist[:n.value].copy()
        self.converged, self.ll = bool(conv.value), hist[-1].copy()
        if elbo:
            self.elbo = self.calculate_elbo()[0]
        return hist

    def calculate_elbo(self):
        e, t = C.c_double(0.0), np.zeros(7)
        self.grp.check(self.grp.lib.mmsig_group_mmctm_elbo(self.grp.g, C.byref(e), capi.dp(t)))
        return e.value, t

    def _state_buffers(self, props=True):
        D, MK, M, G = self.D, self.MK, self.M, self.G
        out = {"lam": np.empty((D, MK)), "nu": np.empty((D, MK)), "zeta": np.empty((D, M)), "mu": np.empty(MK),
               "Sigma": np.empty((MK, MK)), "invSigma": np.empty((MK, MK)), "gamma": np.empty(G), "Elnphi": np.empty(G),
               "phi": np.empty(G)}
        if props:
            out["props"] = np.empty((D, MK))
        return out

    def state(self, props=True):
        out = self._state_buffers(props)
        order = ("lam", "nu", "zeta", "mu", "Sigma", "invSigma", "gamma", "Elnphi", "phi", "props")
        self.grp.check(self.grp.lib.mmsig_group_mmctm_get_state(self.grp.g, *[capi.dp(out.get(k)) for k in order]))
        return out

    def evals(self):
        a, b = np.zeros(self.D, np.int32), np.zeros(self.D, np.int32)
        self.grp.check(self.grp.lib.mmsig_group_mmctm_get_evals(self.grp.g, a.ctypes.data_as(capi.c_i32p), b.ctypes.data_as(capi.c_i32p)))
        return a, b

    def fit_host(self, counts, gamma, lam=None, nu=None, mu=None, Sigma=None, invSigma=None, maxiter=100, tol=1e-4,
                 updateSigma=True, out=None, props=True):
        """mmsig_group_mmctm_fit_host: the whole fit from / to host buffers over all devices of the group."""
        self._csr(counts)
        self.D = len(counts[0][0]) - 1
        D, MK = self.D, self.MK
        out = self._state_buffers(props) if out is None else out
        a = [capi.f64(self.alpha, self.M), capi.f64(gamma, self.G), capi.f64(lam, D * MK), capi.f64(nu, D * MK),
             capi.f64(mu, MK), capi.f64(Sigma, MK * MK), capi.f64(invSigma, MK * MK)]
        hist = np.zeros((maxiter, self.M))
        n, conv = C.c_int32(0), C.c_int32(0)
        order = ("lam", "nu", "zeta", "mu", "Sigma", "invSigma", "gamma", "Elnphi", "phi", "props")
        flags = capi.FLAG_UPDATE_SIGMA if updateSigma else 0
        fn = self.grp.lib.mmsig_group_mmctm_fit_host_packed if self._packed else self.grp.lib.mmsig_group_mmctm_fit_host
        self.grp.check(fn(
            self.grp.g, D, self.M, *self._kv, *self._ptrs, *[capi.dp(x) for x in a], maxiter, tol, flags, capi.dp(hist),
            C.byref(n), C.byref(conv), *[capi.dp(out.get(k)) for k in order]))
        hist = hist[:n.value].copy()
        self.converged, self.ll = bool(conv.value), hist[-1].copy()
        return hist, out

    def fit_restarts(self, gamma0s, maxiter=100, tol=1e-4, updateSigma=True):
        """R restarts dealt over the devices (scripts/run_mmctm.jl:99-111); returns (elbo[R], ll[R, M], n_iter[R], best);
        state() / calculate_elbo() afterwards read the best restart."""
        g0 = capi.f64(np.asarray(gamma0s, float).reshape(-1))
        R = g0.size // self.G
        elbo, ll, nit, best = np.zeros(R), np.zeros((R, self.M)), np.zeros(R, np.int32), C.c_int32(-1)
        flags = capi.FLAG_UPDATE_SIGMA if updateSigma else 0
        self.grp.check(self.grp.lib.mmsig_group_mmctm_restarts(
            self.grp.g, self.D, self.M, *self._kv, *self._ptrs, capi.dp(capi.f64(self.alpha, self.M)), R, capi.dp(g0), maxiter, tol,
            flags, capi.dp(elbo), capi.dp(ll), nit.ctypes.data_as(capi.c_i32p), C.byref(best)))
        self.ll = ll[best.value].copy()
        return elbo, ll, nit, best.value

    def close(self):
        self.grp.close()
