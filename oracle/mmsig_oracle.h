/*
 * mmsig_oracle.h -- CPU restatement (the ORACLE) of the variational-EM inner
 * loop of shahcompbio/MultiModalMuSig.jl (MMCTM / CTM / LDA `fit!`).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path (multimodalmusig.jl_b200/csrc) never does.
 *
 * PARITY STATUS: "parity unpinned" at the NLopt boundary.  The reference needs
 * Julia + NLopt (LD_MMA), neither of which exists in this container or on the
 * GPU box, and no reference test pins the optimiser's numeric output
 * (reference test/mmctm.jl:92-101,150-155 only check "changed / not NaN / >0").
 * Everything the reference's own tests DO pin (closed-form updates, the
 * lambda/nu objectives and gradients, log-likelihoods) is checked against
 * golden vectors in tests/golden/ (see tests/golden/make_golden.py).
 *
 * Third-party arithmetic that is not under /root/reference and is restated
 * here from its published algorithm:
 *   - NLopt LD_MMA (Svanberg MMA / CCSA, NLopt src/algs/mma/mma.c), pinned by
 *     the reference only as NLopt.jl "0.5.1, ~0.6" (Project.toml:15), i.e.
 *     libnlopt 2.5-2.7.  With zero constraints the dual problem is empty and
 *     each inner iteration is closed form.  Both x-tolerance stop rules
 *     (NLopt >= 2.7: L1 norm test; NLopt <= 2.6: per-coordinate) are provided.
 *   - SpecialFunctions.jl digamma (asymptotic series after shifting x >= 7),
 *     lgamma (openlibm lgamma_r == fdlibm; glibc lgamma used here).
 *   - LinearAlgebra inv / logdet (LAPACK getrf/getri): LU with partial
 *     pivoting restated.
 *   - Julia Base pairwise `sum` (block size 1024) for `mean(model.lambda)` and
 *     `sum(diagm.(0 .=> model.nu))`.
 *
 * Layout conventions (flat restatement of the reference's nested vectors):
 *   counts of modality m: CSR  rowptr[m][0..D], term[m][w] (0-BASED), cnt[m][w]
 *   lambda, nu : D x MK row-major   (reference model.lambda[d][j])
 *   zeta       : D x M
 *   gamma/Elnphi/phi : concatenated [m][k][v] row-major, modality offsets
 *                      goff[m] = sum_{m'<m} K[m']*V[m']
 *   theta      : per modality, nnz_m x K_m  (theta[m][w*K_m + k])
 *   props      : D x MK (block m of row d is props[d][m])
 */
#ifndef MMSIG_ORACLE_H
#define MMSIG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- special functions -------------------------------------------------- */
double orc_digamma(double x);
double orc_digamma_det(double x);   /* same recipe on det_log */
double orc_exp(double x);           /* det_exp: pinned algorithm, < 1 ulp */
double orc_log(double x);           /* det_log: pinned algorithm, < 1 ulp */
double orc_lgamma(double x);
double orc_logmvbeta(const double *vals, int n);      /* src/common.jl:1-9 */

/* ---- NLopt LD_MMA, m = 0 constraints ------------------------------------ */
typedef double (*orc_func)(unsigned n, const double *x, double *grad, void *data);

/* arithmetic modes, see mmsig_oracle.c "Arithmetic modes" */
#define ORC_ARITH_LITERAL 0  /* reference operation order, glibc exp/log      */
#define ORC_ARITH_DET     1  /* same addends, every rounding pinned            */

/* iteration flags (same values as MMSIG_FLAG_*) */
#define ORC_FLAG_UPDATE_SIGMA   1u
#define ORC_FLAG_FREEZE_TOPICS  2u
#define ORC_FLAG_FREEZE_MU      4u
#define ORC_FLAG_UNSMOOTHED     8u

#define ORC_STOP_NLOPT27 0   /* NLopt >= 2.7 x-tolerance rule (default) */
#define ORC_STOP_NLOPT26 1   /* NLopt <= 2.6 x-tolerance rule           */

/* minimises f; returns number of objective evaluations; *nouter = outer iters */
int orc_mma_minimize(unsigned n, orc_func f, void *fdata,
                     const double *lb, const double *ub,
                     double *x, double *minf,
                     double xtol_rel, double xtol_abs,
                     int stop_rule, int arith, int *nouter);

/* ---- objectives (src/common.jl:11-46); return the MAXIMISED value ------- */
double orc_lambda_objective(int MK, const double *lam, double *grad,
                            const double *nu, const double *Ndivzeta,
                            const double *sumtheta, const double *mu,
                            const double *invSigma, int arith);
double orc_nu_objective(int MK, const double *nu, double *grad,
                        const double *lam, const double *Ndivzeta,
                        const double *mu, const double *invSigma, int arith);
double orc_alpha_objective(double alpha, double *grad, double sum_Elnphi,
                           int K, int V);

/* ---- dense helpers ------------------------------------------------------- */
int    orc_inv(int n, const double *A, double *Ainv);   /* LU, partial pivoting */
double orc_logabsdet(int n, const double *A);

/* ---- MMCTM model --------------------------------------------------------- */
typedef struct {
    int M, MK;
    int64_t D;
    int *K, *V, *koff;            /* koff[m] = offset of block m in MK-vector */
    int64_t *goff;                /* gamma offsets, goff[M] = total           */
    int64_t **rowptr;             /* [M][D+1]                                  */
    int32_t **term, **cnt;        /* [M][nnz_m]                                */
    int64_t *N;                   /* D x M  (src/MMCTM.jl:38)                  */
    double *alpha;                /* M                                         */
    double *mu, *Sigma, *invSigma;
    double *lambda, *nu, *zeta, *props;
    double *gamma, *Elnphi, *phi;
    double **theta;               /* [M] nnz_m x K_m                           */
    int stop_rule;
    int arith;                    /* ORC_ARITH_*                               */
    int nthreads;                 /* OpenMP threads for the per-sample loop    */
    int converged;
    double elbo;
    double *ll;                   /* M */
    /* diagnostics: objective evaluations of the last E-step, per sample      */
    int32_t *nev_nu, *nev_lambda; /* D each */
    /* DET bookkeeping of the last E-step (what the device keeps in registers / HBM):  */
    double **rz;                  /* [M] nnz_m : 1 / Z_w                                   */
    double *expl;                 /* D x MK   : exp(lambda) as update_theta saw it         */
    double *sumtheta_e;           /* D x MK   : sum-theta of the last E-step (:110-117)    */
    int theta_unsm;               /* table of the last E-step: 0 exp(Elnphi), 1 phi        */
    /* IMMCTM (src/IMMCTM.jl): the topic-term distribution of modality m factorises over I_m
       features, phi_kv = prod_i phi_k,i,f(v,i).  factored != 0: gamma/Elnphi/phi (K x V) above are
       the COMPOSITE tables derived from the feature tables below.                           */
    int factored;
    int *nfeat;                   /* [M]   I_m                                              */
    int **feat;                   /* [M]   V_m x I_m row-major, 0-BASED feature values      */
    int **J;                      /* [M][I_m] values per feature                            */
    int64_t *foff;                /* [M+1] offsets of the flat feature tables, layout [m][k][i][j] */
    int *aoff;                    /* [M+1] offsets into alphaf ([m][i])                     */
    double *alphaf, *gammaf, *Elnphif;
} orc_mmctm;

orc_mmctm *orc_mmctm_new(int M, const int *K, const int *V, int64_t D,
                         const int64_t *const *rowptr, const int32_t *const *term,
                         const int32_t *const *cnt, const double *alpha,
                         const double *gamma0);   /* ctor: src/MMCTM.jl:29-91 */
void orc_mmctm_free(orc_mmctm *m);

void orc_mmctm_update_zeta(orc_mmctm *m, int64_t d);       /* :172-181 */
void orc_mmctm_update_theta(orc_mmctm *m, int64_t d);      /* :183-198 */
void orc_mmctm_unsmoothed_update_theta(orc_mmctm *m, int64_t d); /* :496-509 */
void orc_mmctm_iterate_flags(orc_mmctm *m, unsigned flags, double *ll); /* fit_heldout :566-573, transform :523-538 */
void orc_mmctm_calc_sumtheta(const orc_mmctm *m, int64_t d, double *out); /* :110-117 */
void orc_mmctm_calc_Ndivzeta(const orc_mmctm *m, int64_t d, double *out); /* :119-125 */
void orc_mmctm_update_nu(orc_mmctm *m, int64_t d);         /* :156-170 */
void orc_mmctm_update_lambda(orc_mmctm *m, int64_t d);     /* :127-143 */
void orc_mmctm_fitdoc(orc_mmctm *m, int64_t d);            /* :450-455 */
void orc_mmctm_update_mu(orc_mmctm *m);                    /* :200-202 */
void orc_mmctm_update_Sigma(orc_mmctm *m);                 /* :204-212 */
void orc_mmctm_update_Elnphi(orc_mmctm *m);                /* :214-222 */
void orc_mmctm_update_gamma(orc_mmctm *m);                 /* :224-242 */
void orc_mmctm_update_props(orc_mmctm *m);                 /* :145-154 */
void orc_mmctm_update_phi(orc_mmctm *m);                   /* :244-250 */
void orc_mmctm_update_alpha(orc_mmctm *m);                 /* :252-269 */
void orc_mmctm_loglikelihoods(const orc_mmctm *m, double *ll); /* :384-448 */
/* ELBO terms, :271-382.  terms[7] = PPhi, PEta, PZ, PX, QPhi, QEta, QZ */
double orc_mmctm_elbo(const orc_mmctm *m, double *terms);
/* one body of the fit! loop, :463-479 */
void orc_mmctm_iterate(orc_mmctm *m, int updateSigma, int autoalpha, double *ll);
/* fit!, :457-494; ll_hist is maxiter x M; returns iterations done */
int orc_mmctm_fit(orc_mmctm *m, int maxiter, double tol, int updateSigma,
                  int autoalpha, double *ll_hist);

/* ---- LDA model ----------------------------------------------------------- */
typedef struct {
    int K, V;
    int64_t D;
    int64_t *rowptr; int32_t *term, *cnt;
    int64_t *N;
    double alpha, eta;
    double *lambda, *Elnbeta, *beta;     /* V x K, column-major like Julia: [k*V+v] */
    double *gamma, *Elntheta, *theta;    /* K x D, column-major like Julia: [d*K+k] */
    double *phi;                         /* nnz x K  (phi[w*K+k])                   */
    int arith;
    int nthreads;
    int converged;
    double elbo, ll;
    /* ILDA (src/ILDA.jl): beta_kv = prod_i beta_i[f(v,i), k].  factored != 0: lambda / Elnbeta / beta
       (K x V) above are COMPOSITE tables derived from the feature tables below ([k][i][j] flat). */
    int factored, nfeat;
    int *feat;                           /* V x I row-major, 0-BASED feature values */
    int *J;                              /* [I] */
    int64_t T;                           /* K * sum_i J_i */
    double *etaf;                        /* [I] */
    double *lambdaf, *Elnbetaf;
} orc_lda;

orc_lda *orc_lda_new(int K, int V, int64_t D, const int64_t *rowptr,
                     const int32_t *term, const int32_t *cnt,
                     double alpha, double eta, const double *lambda0); /* src/LDA.jl:24-54 */
void orc_lda_free(orc_lda *m);
void orc_lda_update_Elntheta(orc_lda *m);   /* :78-80  */
void orc_lda_update_gamma(orc_lda *m);      /* :82-90  */
void orc_lda_update_phi(orc_lda *m);        /* :69-76  */
void orc_lda_update_Elnbeta(orc_lda *m);    /* :96-98  */
void orc_lda_update_lambda(orc_lda *m);     /* :100-108 */
void orc_lda_update_beta(orc_lda *m);       /* :110-112 */
void orc_lda_update_theta(orc_lda *m);      /* :92-94  */
double orc_lda_loglikelihood(const orc_lda *m); /* :174-188 */
double orc_lda_elbo(const orc_lda *m, double *terms); /* :114-172; PBeta,PTheta,PZ,PX,QBeta,QTheta,QZ */
double orc_lda_iterate(orc_lda *m);         /* :202-209 */
void orc_lda_unsmoothed_update_phi(orc_lda *m);               /* :226-231 */
double orc_lda_iterate_flags(orc_lda *m, unsigned flags);     /* fit_heldout :275-280, transform :242-246 */
int orc_lda_fit(orc_lda *m, int maxiter, double tol, double *ll_hist); /* :198-224 */

/* ILDA (src/ILDA.jl:1-62): switch a freshly constructed LDA model to feature-factorised topics.
 * feat V x I (0-based), etaf [I], lambdaf0 [k][i][j] (ctor: rand 1:100) */
void orc_ilda_enable(orc_lda *m, int nfeat, const int *feat, const double *etaf, const double *lambdaf0);
void orc_ilda_compose(orc_lda *m);            /* composite Elnbeta / beta from the feature tables */
void orc_ilda_update_Elnbeta(orc_lda *m);     /* src/ILDA.jl:97-103 */

/* IMMCTM (src/IMMCTM.jl:1-108): switch a freshly constructed model to feature-factorised topics.
 * nfeat[M], feat[m] V_m x I_m (0-based values), alphaf [m][i], gammaf0 [m][k][i][j] (ctor: rand 1:100) */
void orc_immctm_enable(orc_mmctm *m, const int *nfeat, const int *const *feat,
                       const double *alphaf, const double *gammaf0);
void orc_immctm_update_Elnphi(orc_mmctm *m);   /* src/IMMCTM.jl:186-195 + composite K x V tables */
void orc_immctm_update_gamma(orc_mmctm *m);    /* :197-221 */
int64_t orc_immctm_table_size(const orc_mmctm *m);

/* format_counts_* (src/utils.jl:1-36): dense count matrix -> CSR; see mmsig_oracle.c */
int64_t orc_make_count_csr(int64_t D, int V, const int64_t *dense, int layout,
                           int64_t *rowptr, int32_t *term, int32_t *cnt);

#ifdef __cplusplus
}
#endif
#endif
