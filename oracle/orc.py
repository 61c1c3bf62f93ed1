"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("mmsig_oracle.c", "mmsig_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


class MMCTM(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("MK", C.c_int), ("D", C.c_int64),
        ("K", c_ip), ("V", c_ip), ("koff", c_ip), ("goff", c_i64p),
        ("rowptr", C.POINTER(c_i64p)), ("term", C.POINTER(c_i32p)), ("cnt", C.POINTER(c_i32p)),
        ("N", c_i64p), ("alpha", c_dp),
        ("mu", c_dp), ("Sigma", c_dp), ("invSigma", c_dp),
        ("lambda_", c_dp), ("nu", c_dp), ("zeta", c_dp), ("props", c_dp),
        ("gamma", c_dp), ("Elnphi", c_dp), ("phi", c_dp),
        ("theta", C.POINTER(c_dp)),
        ("stop_rule", C.c_int), ("arith", C.c_int), ("nthreads", C.c_int), ("converged", C.c_int),
        ("elbo", C.c_double), ("ll", c_dp),
        ("nev_nu", c_i32p), ("nev_lambda", c_i32p),
        ("rz", C.POINTER(c_dp)), ("expl", c_dp), ("sumtheta_e", c_dp), ("theta_unsm", C.c_int),
        ("factored", C.c_int), ("nfeat", c_ip), ("feat", C.POINTER(c_ip)), ("J", C.POINTER(c_ip)),
        ("foff", c_i64p), ("aoff", c_ip), ("alphaf", c_dp), ("gammaf", c_dp), ("Elnphif", c_dp),
    ]


class LDA(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("V", C.c_int), ("D", C.c_int64),
        ("rowptr", c_i64p), ("term", c_i32p), ("cnt", c_i32p), ("N", c_i64p),
        ("alpha", C.c_double), ("eta", C.c_double),
        ("lambda_", c_dp), ("Elnbeta", c_dp), ("beta", c_dp),
        ("gamma", c_dp), ("Elntheta", c_dp), ("theta", c_dp), ("phi", c_dp),
        ("arith", C.c_int), ("nthreads", C.c_int), ("converged", C.c_int),
        ("elbo", C.c_double), ("ll", C.c_double),
        ("factored", C.c_int), ("nfeat", C.c_int), ("feat", c_ip), ("J", c_ip), ("T", C.c_int64),
        ("etaf", c_dp), ("lambdaf", c_dp), ("Elnbetaf", c_dp),
    ]


ORC_FUNC = C.CFUNCTYPE(C.c_double, C.c_uint, c_dp, c_dp, C.c_void_p)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    d, i, i64, vp = C.c_double, C.c_int, C.c_int64, C.c_void_p
    pm, pl = C.POINTER(MMCTM), C.POINTER(LDA)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("orc_ilda_enable", None, pl, i, c_ip, c_dp, c_dp)
    sig("orc_ilda_compose", None, pl)
    sig("orc_ilda_update_Elnbeta", None, pl)
    sig("orc_immctm_enable", None, pm, c_ip, C.POINTER(c_ip), c_dp, c_dp)
    sig("orc_immctm_update_Elnphi", None, pm)
    sig("orc_immctm_update_gamma", None, pm)
    sig("orc_immctm_table_size", i64, pm)
    sig("orc_make_count_csr", i64, i64, i, c_i64p, i, c_i64p, c_i32p, c_i32p)
    sig("orc_digamma", d, d)
    sig("orc_digamma_det", d, d)
    sig("orc_lgamma", d, d)
    sig("orc_exp", d, d)
    sig("orc_log", d, d)
    sig("orc_logmvbeta", d, c_dp, i)
    sig("orc_mma_minimize", i, C.c_uint, ORC_FUNC, vp, c_dp, c_dp, c_dp, c_dp, d, d, i, i, c_ip)
    sig("orc_lambda_objective", d, i, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, i)
    sig("orc_nu_objective", d, i, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, i)
    sig("orc_alpha_objective", d, d, c_dp, d, i, i)
    sig("orc_inv", i, i, c_dp, c_dp)
    sig("orc_logabsdet", d, i, c_dp)
    sig("orc_mmctm_new", pm, i, c_ip, c_ip, i64, C.POINTER(c_i64p), C.POINTER(c_i32p),
        C.POINTER(c_i32p), c_dp, c_dp)
    sig("orc_mmctm_free", None, pm)
    for n in ("zeta", "theta", "nu", "lambda", "fitdoc"):
        sig("orc_mmctm_%s" % ("update_" + n if n != "fitdoc" else n), None, pm, i64)
    sig("orc_mmctm_calc_sumtheta", None, pm, i64, c_dp)
    sig("orc_mmctm_calc_Ndivzeta", None, pm, i64, c_dp)
    for n in ("mu", "Sigma", "Elnphi", "gamma", "props", "phi", "alpha"):
        sig("orc_mmctm_update_" + n, None, pm)
    sig("orc_mmctm_loglikelihoods", None, pm, c_dp)
    sig("orc_mmctm_elbo", d, pm, c_dp)
    sig("orc_mmctm_iterate", None, pm, i, i, c_dp)
    sig("orc_mmctm_fit", i, pm, i, d, i, i, c_dp)
    sig("orc_mmctm_iterate_flags", None, pm, C.c_uint, c_dp)
    sig("orc_mmctm_unsmoothed_update_theta", None, pm, i64)
    sig("orc_lda_iterate_flags", d, pl, C.c_uint)
    sig("orc_lda_unsmoothed_update_phi", None, pl)
    sig("orc_lda_new", pl, i, i, i64, c_i64p, c_i32p, c_i32p, d, d, c_dp)
    sig("orc_lda_free", None, pl)
    for n in ("Elntheta", "gamma", "phi", "Elnbeta", "lambda", "beta", "theta"):
        sig("orc_lda_update_" + n, None, pl)
    sig("orc_lda_loglikelihood", d, pl)
    sig("orc_lda_elbo", d, pl, c_dp)
    sig("orc_lda_iterate", d, pl)
    sig("orc_lda_fit", i, pl, i, d, c_dp)
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _view(ptr, shape):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape)
    return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(shape)


ARITH_LITERAL, ARITH_DET = 0, 1
FLAG_UPDATE_SIGMA, FLAG_FREEZE_TOPICS, FLAG_FREEZE_MU, FLAG_UNSMOOTHED = 1, 2, 4, 8
STOP_NLOPT27, STOP_NLOPT26 = 0, 1


class OracleMMCTM:
    """Flat-array view of the oracle's MMCTM (reference src/MMCTM.jl:1-108).

    counts: list over modalities of (rowptr int64[D+1], term int32[nnz] 0-based, cnt int32[nnz]).
    """

    def __init__(self, K, alpha, V, counts, gamma0, arith=ARITH_LITERAL, stop_rule=STOP_NLOPT27,
                 nthreads=1):
        L = lib()
        self.L = L
        M = len(K)
        self.K = np.asarray(K, dtype=np.int32)
        self.V = np.asarray(V, dtype=np.int32)
        self.M = M
        self.D = len(counts[0][0]) - 1
        self._keep = [(np.ascontiguousarray(r, np.int64), np.ascontiguousarray(t, np.int32),
                       np.ascontiguousarray(c, np.int32)) for r, t, c in counts]
        rp = (c_i64p * M)(*[k[0].ctypes.data_as(c_i64p) for k in self._keep])
        tp = (c_i32p * M)(*[k[1].ctypes.data_as(c_i32p) for k in self._keep])
        cp = (c_i32p * M)(*[k[2].ctypes.data_as(c_i32p) for k in self._keep])
        al = np.asarray(alpha, dtype=np.float64)
        g0 = np.ascontiguousarray(gamma0, dtype=np.float64)
        assert g0.size == int((self.K * self.V).sum())
        self.p = L.orc_mmctm_new(M, self.K.ctypes.data_as(c_ip), self.V.ctypes.data_as(c_ip),
                                 self.D, rp, tp, cp, _dp(al), _dp(g0))
        s = self.p.contents
        s.arith = arith
        s.stop_rule = stop_rule
        s.nthreads = nthreads
        self.MK = s.MK
        self.G = int((self.K * self.V).sum())
        if arith != ARITH_LITERAL:          # ctor-time Elnphi/zeta depend on the arithmetic
            L.orc_mmctm_update_Elnphi(self.p)
            for d in range(self.D):
                L.orc_mmctm_update_zeta(self.p, d)

    def __del__(self):
        try:
            self.L.orc_mmctm_free(self.p)
        except Exception:
            pass

    # live numpy views on the oracle's state
    def _arr(self, name, shape):
        return _view(getattr(self.p.contents, name), shape)

    lam = property(lambda s: s._arr("lambda_", (s.D, s.MK)))
    nu = property(lambda s: s._arr("nu", (s.D, s.MK)))
    zeta = property(lambda s: s._arr("zeta", (s.D, s.M)))
    props = property(lambda s: s._arr("props", (s.D, s.MK)))
    mu = property(lambda s: s._arr("mu", (s.MK,)))
    Sigma = property(lambda s: s._arr("Sigma", (s.MK, s.MK)))
    invSigma = property(lambda s: s._arr("invSigma", (s.MK, s.MK)))
    gamma = property(lambda s: s._arr("gamma", (s.G,)))
    Elnphi = property(lambda s: s._arr("Elnphi", (s.G,)))
    phi = property(lambda s: s._arr("phi", (s.G,)))
    alpha = property(lambda s: s._arr("alpha", (s.M,)))
    nev_nu = property(lambda s: np.ctypeslib.as_array(s.p.contents.nev_nu, shape=(max(s.D, 1),))[:s.D])
    nev_lambda = property(lambda s: np.ctypeslib.as_array(s.p.contents.nev_lambda, shape=(max(s.D, 1),))[:s.D])

    def theta(self, m):
        nnz = int(self._keep[m][0][-1])
        return _view(self.p.contents.theta[m], (nnz, int(self.K[m])))

    def set_state(self, gamma, lam, nu, mu, Sigma, invSigma):
        """Overwrite the model state (what `fit!` reads from the struct, src/MMCTM.jl:1-27) and re-derive
        Elnphi (:78-79) and zeta (:85-86) in this oracle's arithmetic: two oracles (or an oracle and the
        device) can then take one iteration from an IDENTICAL state at any point of a fit."""
        self.gamma[:] = np.asarray(gamma, float).reshape(-1)
        self.lam[:] = np.asarray(lam, float).reshape(self.D, self.MK)
        self.nu[:] = np.asarray(nu, float).reshape(self.D, self.MK)
        self.mu[:] = np.asarray(mu, float).reshape(-1)
        self.Sigma[:] = np.asarray(Sigma, float).reshape(self.MK, self.MK)
        self.invSigma[:] = np.asarray(invSigma, float).reshape(self.MK, self.MK)
        self.L.orc_mmctm_update_Elnphi(self.p)
        for d in range(self.D):
            self.L.orc_mmctm_update_zeta(self.p, d)

    def N(self):
        return np.ctypeslib.as_array(self.p.contents.N, shape=(self.D * self.M,)).reshape(self.D, self.M)

    def iterate(self, updateSigma=True, autoalpha=False):
        ll = np.zeros(self.M)
        self.L.orc_mmctm_iterate(self.p, int(updateSigma), int(autoalpha), _dp(ll))
        return ll

    def iterate_flags(self, flags):
        ll = np.zeros(self.M)
        self.L.orc_mmctm_iterate_flags(self.p, int(flags), _dp(ll))
        return ll

    def fit(self, maxiter=100, tol=1e-4, updateSigma=True, autoalpha=False):
        hist = np.zeros((maxiter, self.M))
        n = self.L.orc_mmctm_fit(self.p, maxiter, tol, int(updateSigma), int(autoalpha), _dp(hist))
        return hist[:n].copy()

    def elbo(self):
        t = np.zeros(7)
        v = self.L.orc_mmctm_elbo(self.p, _dp(t))
        return v, t

    def loglikelihoods(self):
        ll = np.zeros(self.M)
        self.L.orc_mmctm_loglikelihoods(self.p, _dp(ll))
        return ll

    converged = property(lambda s: bool(s.p.contents.converged))


class OracleIMMCTM(OracleMMCTM):
    """IMMCTM (reference src/IMMCTM.jl): MMCTM whose topics factorise over features.
    features: list over modalities of (V_m, I_m) integer arrays with 0-BASED feature values;
    alphaf: list over modalities of per-feature alphas (or one float per modality, :93-100);
    gammaf0: flat [m][k][i][j] table (the constructor's rand(1:100), :60-67)."""

    def __init__(self, K, alphaf, features, counts, gammaf0, arith=ARITH_LITERAL, stop_rule=STOP_NLOPT27, nthreads=1):
        feats = [np.ascontiguousarray(f, dtype=np.int32) for f in features]
        V = [f.shape[0] for f in feats]
        self.I = [f.shape[1] for f in feats]
        self.J = [[int(f[:, i].max()) + 1 for i in range(f.shape[1])] for f in feats]
        al = [np.full(n, float(a)) if np.ndim(a) == 0 else np.asarray(a, float) for a, n in zip(alphaf, self.I)]
        G = sum(k * v for k, v in zip(K, V))
        super().__init__(K, [a[0] for a in al], V, counts, np.ones(G), arith=arith, stop_rule=stop_rule, nthreads=nthreads)
        self._feats = feats
        self.T = sum(k * sum(j) for k, j in zip(K, self.J))
        g0 = np.ascontiguousarray(gammaf0, dtype=np.float64)
        assert g0.size == self.T
        nf = np.asarray(self.I, dtype=np.int32)
        fp = (c_ip * self.M)(*[f.ctypes.data_as(c_ip) for f in feats])
        alf = np.ascontiguousarray(np.concatenate(al))
        self.L.orc_immctm_enable(self.p, nf.ctypes.data_as(c_ip), fp, _dp(alf), _dp(g0))
        for d in range(self.D):
            self.L.orc_mmctm_update_zeta(self.p, d)

    gammaf = property(lambda s: s._arr("gammaf", (s.T,)))
    Elnphif = property(lambda s: s._arr("Elnphif", (s.T,)))
    alphaf = property(lambda s: s._arr("alphaf", (sum(s.I),)))

    def table(self, flat, m, k, i):
        """View of feature i of topic k of modality m in a flat [m][k][i][j] table."""
        o = sum(int(self.K[mm]) * sum(self.J[mm]) for mm in range(m)) + k * sum(self.J[m]) + sum(self.J[m][:i])
        return flat[o:o + self.J[m][i]]


class OracleLDA:
    """Flat-array view of the oracle's LDA (reference src/LDA.jl:1-67)."""

    def __init__(self, K, alpha, eta, V, counts, lambda0, arith=ARITH_LITERAL, nthreads=1):
        L = lib()
        self.L = L
        r, t, c = counts
        self._keep = (np.ascontiguousarray(r, np.int64), np.ascontiguousarray(t, np.int32),
                      np.ascontiguousarray(c, np.int32))
        self.K, self.V, self.D = int(K), int(V), len(r) - 1
        l0 = np.ascontiguousarray(lambda0, dtype=np.float64)     # [k*V+v]  (Julia V x K column-major)
        assert l0.size == self.K * self.V
        self.p = L.orc_lda_new(self.K, self.V, self.D, self._keep[0].ctypes.data_as(c_i64p),
                               self._keep[1].ctypes.data_as(c_i32p), self._keep[2].ctypes.data_as(c_i32p),
                               float(alpha), float(eta), _dp(l0))
        s = self.p.contents
        s.arith = arith
        s.nthreads = nthreads
        if arith != ARITH_LITERAL:
            L.orc_lda_update_Elnbeta(self.p)
            L.orc_lda_update_Elntheta(self.p)

    def __del__(self):
        try:
            self.L.orc_lda_free(self.p)
        except Exception:
            pass

    def _arr(self, name, shape):
        return _view(getattr(self.p.contents, name), shape)

    lam = property(lambda s: s._arr("lambda_", (s.K, s.V)))          # [k][v]
    Elnbeta = property(lambda s: s._arr("Elnbeta", (s.K, s.V)))
    beta = property(lambda s: s._arr("beta", (s.K, s.V)))
    gamma = property(lambda s: s._arr("gamma", (s.D, s.K)))          # [d][k]
    Elntheta = property(lambda s: s._arr("Elntheta", (s.D, s.K)))
    theta = property(lambda s: s._arr("theta", (s.D, s.K)))
    phi = property(lambda s: s._arr("phi", (int(s._keep[0][-1]), s.K)))

    def iterate(self):
        return self.L.orc_lda_iterate(self.p)

    def iterate_flags(self, flags):
        return self.L.orc_lda_iterate_flags(self.p, int(flags))

    def fit(self, maxiter=1000, tol=1e-4):
        hist = np.zeros(maxiter)
        n = self.L.orc_lda_fit(self.p, maxiter, tol, _dp(hist))
        return hist[:n].copy()

    def elbo(self):
        t = np.zeros(7)
        v = self.L.orc_lda_elbo(self.p, _dp(t))
        return v, t

    converged = property(lambda s: bool(s.p.contents.converged))


def make_count_csr(dense, layout=0):
    """format_counts_* (src/utils.jl:1-36) by the oracle.  dense: (V, D) array for layout 0
    (term-major), (D, V) for layout 1 (sample-major).  Returns (rowptr, term0, count)."""
    a = np.ascontiguousarray(dense, dtype=np.int64)
    V, D = (a.shape if layout == 0 else a.shape[::-1])
    L = lib()
    rowptr = np.zeros(D + 1, dtype=np.int64)
    nnz = L.orc_make_count_csr(D, V, a.ctypes.data_as(c_i64p), layout, rowptr.ctypes.data_as(c_i64p), None, None)
    if nnz < 0:
        raise OverflowError("a count exceeds int32")
    term, cnt = np.zeros(nnz, dtype=np.int32), np.zeros(nnz, dtype=np.int32)
    L.orc_make_count_csr(D, V, a.ctypes.data_as(c_i64p), layout, rowptr.ctypes.data_as(c_i64p),
                         term.ctypes.data_as(c_i32p), cnt.ctypes.data_as(c_i32p))
    return rowptr, term, cnt


class OracleILDA(OracleLDA):
    """ILDA (reference src/ILDA.jl): LDA whose topics factorise over features.
    features: (V, I) integer array with 0-BASED feature values; eta: scalar or one per feature (:54-58);
    lambdaf0: flat [k][i][j] table (the constructor's rand(1:100), :38)."""

    def __init__(self, K, alpha, eta, features, counts, lambdaf0, arith=ARITH_LITERAL, nthreads=1):
        f = np.ascontiguousarray(features, dtype=np.int32)
        V, I = f.shape
        self.I = I
        self.J = [int(f[:, i].max()) + 1 for i in range(I)]
        et = np.full(I, float(eta)) if np.ndim(eta) == 0 else np.asarray(eta, dtype=np.float64)
        super().__init__(K, alpha, float(et[0]), V, counts, np.ones(int(K) * V), arith=arith, nthreads=nthreads)
        self._feat = f
        self.T = int(K) * sum(self.J)
        l0 = np.ascontiguousarray(lambdaf0, dtype=np.float64)
        assert l0.size == self.T
        self.L.orc_ilda_enable(self.p, I, f.ctypes.data_as(c_ip), _dp(np.ascontiguousarray(et)), _dp(l0))

    lambdaf = property(lambda s: s._arr("lambdaf", (s.T,)))
    Elnbetaf = property(lambda s: s._arr("Elnbetaf", (s.T,)))
    etaf = property(lambda s: s._arr("etaf", (s.I,)))

    def table(self, flat, k, i):
        o = k * sum(self.J) + sum(self.J[:i])
        return flat[o:o + self.J[i]]
