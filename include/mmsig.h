/*
 * mmsig.h -- C ABI of libmmsig.so: the B200-native (sm_100a) variational-EM inner loop of the
 * MMCTM / CTM / LDA topic models of shahcompbio/MultiModalMuSig.jl.
 *
 * The reference has no FFI of its own (it is pure Julia); the boundary this library replaces
 * is the Julia method level
 *     fit!(model::MMCTM; maxiter, tol, verbose, autoα, updateΣ)      reference src/MMCTM.jl:457-494
 *     fit!(model::LDA;   maxiter, tol, verbose)                      reference src/LDA.jl:198-224
 * and, underneath, fitdoc!/update_*!/calculate_* (src/MMCTM.jl:110-455, src/LDA.jl:69-196).
 * julia/MMSigB200.jl `ccall`s these entry points; tests and bench.py bind them with ctypes.
 *
 * Conventions
 *   - every call blocks and returns 0 on success, a negative MMSIG_E* code otherwise; the
 *     message is at mmsig_last_error(h).  No exceptions / longjmp cross the boundary.
 *   - the caller owns every host buffer; the library never keeps a host pointer after return.
 *     Device memory lives behind the opaque handle.  A handle is not thread-safe; distinct
 *     handles are independent.
 *   - any output pointer may be NULL (= not wanted).
 *   - flat layouts (the Julia shim flattens / scatters the nested vectors):
 *       counts, modality m : CSR  rowptr[m][0..D] int64, term[m][w] int32 0-BASED, count[m][w] int32 > 0
 *       lambda, nu, props  : D x MK row-major (MK = sum K[m]; block m of row d = model.λ[d][off_m+1 : off_m+K_m])
 *       zeta               : D x M
 *       gamma/Elnphi/phi   : concatenated [m][k][v] row-major
 *       mu MK ; Sigma, invSigma MK x MK row-major
 *       LDA lambda/Elnbeta/beta : [k][v] (= Julia's V x K column-major) ; gamma/theta : [d][k] (= K x D column-major)
 *   - there is NO CPU fallback: without a CUDA device mmsig_create fails with MMSIG_ENODEV.
 */
#ifndef MMSIG_H
#define MMSIG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMSIG_OK        0
#define MMSIG_EINVAL   -1   /* bad argument / call order                  */
#define MMSIG_ENODEV   -2   /* no usable CUDA device                      */
#define MMSIG_ECUDA    -3   /* CUDA runtime error                         */
#define MMSIG_ENOMEM   -4
#define MMSIG_ENCCL    -5   /* NCCL missing or failed                     */
#define MMSIG_ELIMIT   -6   /* unsupported size (see mmsig_limits)        */

#define MMSIG_STOP_NLOPT27 0   /* NLopt >= 2.7 x-tolerance rule (default)  */
#define MMSIG_STOP_NLOPT26 1   /* NLopt <= 2.6 rule                         */

#define MMSIG_FLAG_UPDATE_SIGMA  1u  /* fit!'s updateΣ=true (src/MMCTM.jl:468-470)                        */
#define MMSIG_FLAG_FREEZE_TOPICS 2u  /* keep γ, Elnϕ, ϕ (LDA: λ, Elnβ, β): fit_heldout / transform / predict */
#define MMSIG_FLAG_FREEZE_MU     4u  /* keep μ (with UPDATE_SIGMA clear: the Gaussian prior is frozen)     */
#define MMSIG_FLAG_AUTO_ALPHA   16u  /* fit!'s autoα=true: update_α! after update_γ! (src/MMCTM.jl:252-269,472-474), on the host */
#define MMSIG_FLAG_UNSMOOTHED    8u  /* θ ∝ exp(λ)·ϕ (unsmoothed_update_θ!, src/MMCTM.jl:496-509); LDA: ϕ ∝ exp(Elnθ)·β (src/LDA.jl:226-231) */

typedef struct mmsig_handle mmsig_handle;

/* Arithmetic of the tile passes (θ / log-likelihood pass of the MMCTM and IMMCTM, E / log-likelihood pass of the LDA
 * and ILDA).  FP64 (default) matches Julia's Float64 and the pinned oracle bit for bit.  FP32 is the optional fast
 * mode: float arithmetic per nonzero, sums over samples in double, the LD_MMA solves and the M-step tables stay FP64;
 * state keeps its FP64 layout.  Its tolerance against the FP64 path from the same state: one iteration 1e-5 on ϕ / β,
 * 1e-6 on the log-likelihood; a 20-iteration fit 1e-6 (LDA) / 1e-3 (MMCTM: LD_MMA's stop decisions amplify any
 * perturbation) on the log-likelihood (tests/test_gpu_fp32.py). */
#define MMSIG_PRECISION_FP64 0
#define MMSIG_PRECISION_FP32 1

typedef struct {
    int32_t device;        /* CUDA device ordinal                                          */
    int32_t stop_rule;     /* MMSIG_STOP_*: which NLopt x-tolerance rule LD_MMA follows    */
    int32_t profile;       /* != 0: time every kernel with CUDA events (mmsig_kernel_times) */
    int32_t precision;     /* MMSIG_PRECISION_*                                            */
    int32_t reserved[4];   /* must be 0                                                    */
} mmsig_config;

int32_t     mmsig_version(void);
/* the sizes beyond which set_data answers MMSIG_ELIMIT: modalities, sum of K_m, K_m, V_m of MMCTM
 * (V_m additionally has to fit the theta tile in shared memory), V of LDA.  Needs no device. */
int32_t     mmsig_limits(int32_t *max_modalities, int32_t *max_sum_K, int32_t *max_K, int32_t *max_V_mmctm,
                         int32_t *max_V_lda);
int32_t     mmsig_create(const mmsig_config *cfg, mmsig_handle **out);
int32_t     mmsig_destroy(mmsig_handle *h);
const char *mmsig_last_error(const mmsig_handle *h);           /* h may be NULL: last create error */
/* run every kernel of this handle on an existing CUDA stream (a cudaStream_t, e.g. torch's) */
int32_t     mmsig_set_stream(mmsig_handle *h, void *cuda_stream);
int32_t     mmsig_synchronize(mmsig_handle *h);
/* switch the per-kernel CUDA-event timing of cfg.profile on / off (mmsig_kernel_times) */
int32_t     mmsig_set_profile(mmsig_handle *h, int32_t on);
/* page-locked ("pinned") host memory for the caller's count / state buffers: transfers from and to it run
 * at link speed and overlap with kernels; mmsig_mmctm_fit_host accepts pageable buffers too (slower). */
int32_t     mmsig_host_alloc(uint64_t bytes, void **out);
int32_t     mmsig_host_free(void *p);

/* ---- multi-GPU: one handle (process) per GPU, samples sharded over ranks -----------------
 * Per iteration the ranks exchange one small packed buffer of double-double partial sums
 * (topic-term statistics, Σλ, Σν, then ΣΔΔᵀ and the LL numerators) with ncclAllGather and
 * each rank reduces it in rank order, so all ranks hold bit-identical globals. */
int32_t mmsig_comm_unique_id(uint8_t id_out[128]);
int32_t mmsig_comm_init(mmsig_handle *h, const uint8_t id[128], int32_t rank, int32_t nranks);

/* ---- multi-GPU from ONE process: a group of devices -------------------------------------------
 * The reference's only parallelism is `addprocs` + `pmap` over independent restarts
 * (scripts/run_mmctm.jl:8-11,99-111); fit! itself is serial (src/MMCTM.jl:463-465).  A group serves a single
 * caller (julia/MMSigB200.jl: fit!(model; devices=0:7), fit_restarts(...; devices=0:7)) with both patterns:
 *   - one fit with the samples sharded over the devices (contiguous shards balanced by nonzeros); per iteration
 *     the members exchange the same two packed buffers of double-double partial sums as above, by peer stores
 *     over NVLink into each other's exchange arenas + stream-ordered events (no NCCL, no host staging), and
 *     every member reduces them in rank order: all devices hold bit-identical globals, and the result is
 *     bit-identical to the one-GPU fit;
 *   - restarts dealt over the devices, each device holding the whole corpus, no communication.
 * The group runs one host thread per device inside each call; calls block; a group is not thread-safe.
 * All arrays are those of the one-handle calls for the WHOLE corpus (the library shards them).  The same
 * device may be listed more than once (several shards on one GPU: how a one-GPU box tests this path);
 * distinct devices need peer access to one another. */
typedef struct mmsig_group mmsig_group;
int32_t     mmsig_group_create(const mmsig_config *cfg, int32_t n_devices, const int32_t *device_ids, mmsig_group **out);
int32_t     mmsig_group_destroy(mmsig_group *g);
const char *mmsig_group_last_error(const mmsig_group *g);     /* g may be NULL: last create error */
int32_t     mmsig_group_size(const mmsig_group *g);
/* member i's handle, for the instrumentation calls only (mmsig_launch_count, mmsig_kernel_times) */
mmsig_handle *mmsig_group_member(mmsig_group *g, int32_t i);
int32_t mmsig_group_mmctm_set_data(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                   const int64_t *const *rowptr, const int32_t *const *term,
                                   const int32_t *const *count);
int32_t mmsig_group_mmctm_set_state(mmsig_group *g, const double *alpha, const double *gamma, const double *lambda,
                                    const double *nu, const double *mu, const double *Sigma, const double *invSigma);
int32_t mmsig_group_mmctm_iterate(mmsig_group *g, uint32_t flags, double *ll_out);
int32_t mmsig_group_mmctm_fit(mmsig_group *g, int32_t maxiter, double tol, uint32_t flags, double *ll_hist,
                              int32_t *n_iter, int32_t *converged);
int32_t mmsig_group_mmctm_elbo(mmsig_group *g, double *elbo, double *terms);
int32_t mmsig_group_mmctm_get_state(mmsig_group *g, double *lambda, double *nu, double *zeta, double *mu,
                                    double *Sigma, double *invSigma, double *gamma, double *Elnphi, double *phi,
                                    double *props);
int32_t mmsig_group_mmctm_get_evals(mmsig_group *g, int32_t *nev_nu, int32_t *nev_lambda);
int32_t mmsig_group_mmctm_get_theta(mmsig_group *g, int32_t m, double *theta_out);
/* mmsig_mmctm_fit_host over the group: every device pipelines its shard's uploads behind its own E-step */
int32_t mmsig_group_mmctm_fit_host(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                   const int64_t *const *rowptr, const int32_t *const *term,
                                   const int32_t *const *count, const double *alpha, const double *gamma,
                                   const double *lambda, const double *nu, const double *mu, const double *Sigma,
                                   const double *invSigma, int32_t maxiter, double tol, uint32_t flags,
                                   double *ll_hist, int32_t *n_iter, int32_t *converged, double *lambda_out,
                                   double *nu_out, double *zeta_out, double *mu_out, double *Sigma_out,
                                   double *invSigma_out, double *gamma_out, double *Elnphi_out, double *phi_out,
                                   double *props_out);
int32_t mmsig_group_mmctm_fit_host_packed(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                          const int64_t *const *rowptr, const uint32_t *const *rec, const double *alpha,
                                          const double *gamma, const double *lambda, const double *nu, const double *mu,
                                          const double *Sigma, const double *invSigma, int32_t maxiter, double tol,
                                          uint32_t flags, double *ll_hist, int32_t *n_iter, int32_t *converged,
                                          double *lambda_out, double *nu_out, double *zeta_out, double *mu_out,
                                          double *Sigma_out, double *invSigma_out, double *gamma_out, double *Elnphi_out,
                                          double *phi_out, double *props_out);
/* R independent restarts dealt over the devices (restart r on member r mod n; every member holds the whole
 * corpus; scripts/run_mmctm.jl:99-111's pmap): outputs as mmsig_mmctm_restarts; afterwards
 * mmsig_group_mmctm_get_state / _elbo read the device that holds the best restart. */
int32_t mmsig_group_mmctm_restarts(mmsig_group *g, int64_t D, int32_t M, const int32_t *K, const int32_t *V,
                                   const int64_t *const *rowptr, const int32_t *const *term,
                                   const int32_t *const *count, const double *alpha, int32_t R, const double *gamma0,
                                   int32_t maxiter, double tol, uint32_t flags, double *elbo_out, double *ll_out,
                                   int32_t *n_iter_out, int32_t *best);
/* LDA (src/LDA.jl:198-224) over the group */
int32_t mmsig_group_lda_set_data(mmsig_group *g, int64_t D, int32_t K, int32_t V, const int64_t *rowptr,
                                 const int32_t *term, const int32_t *count);
int32_t mmsig_group_lda_set_state(mmsig_group *g, double alpha, double eta, const double *lambda,
                                  const double *gamma_next);
int32_t mmsig_group_lda_fit(mmsig_group *g, int32_t maxiter, double tol, double *ll_hist, int32_t *n_iter,
                            int32_t *converged);
int32_t mmsig_group_lda_elbo(mmsig_group *g, double *elbo, double *terms);
int32_t mmsig_group_lda_get_state(mmsig_group *g, double *lambda, double *Elnbeta, double *beta, double *gamma,
                                  double *Elntheta, double *theta);

/* ---- count ingest: format_counts_mmctm / _ctm / _lda (reference src/utils.jl:1-36) on the device ----
 * make_count_matrix (src/utils.jl:1-7) for every sample of one modality's dense count matrix:
 * entries > 0 become (term, count) rows in ascending term order, entries <= 0 are dropped.
 * dense: HOST pointer to D*V integers of elem_bytes (4: int32, 8: int64 = Julia Int) in layout
 *   MMSIG_DENSE_TERM_MAJOR   dense[v*D + d]  one row per term: the TSV files (data/brca-eu_snv_counts.tsv), a C-order (V, D) array
 *   MMSIG_DENSE_SAMPLE_MAJOR dense[d*V + v]  Julia's column-major V x D Matrix / one DataFrame column per sample
 * A count above 2^31-1 is refused (MMSIG_ELIMIT). */
#define MMSIG_DENSE_TERM_MAJOR   0
#define MMSIG_DENSE_SAMPLE_MAJOR 1
/* -> rowptr_out[D+1], *nnz_out; the records stay on the device until _fetch copies them out
 * (term 0-BASED, count), which also releases them */
int32_t mmsig_format_counts(mmsig_handle *h, int64_t D, int32_t V, const void *dense, int32_t elem_bytes,
                            int32_t layout, int64_t *rowptr_out, int64_t *nnz_out);
int32_t mmsig_format_counts_fetch(mmsig_handle *h, int32_t *term_out, int32_t *count_out);
/* mmsig_mmctm_set_data / mmsig_lda_set_data from dense matrices: the CSR is built on the device and
 * never visits the host (dense[m] as above, V[m] terms each) */
int32_t mmsig_mmctm_set_data_dense(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                   const int32_t *V, const void *const *dense, int32_t elem_bytes, int32_t layout);
int32_t mmsig_lda_set_data_dense(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V,
                                 const void *dense, int32_t elem_bytes, int32_t layout);

/* the count TSV files the reference reads (README.md:14-16, data/brca-eu_snv_counts.tsv: a `term`
 * column, one column per sample): dimensions, then the integer body as a V x D term-major int32
 * matrix (what the calls above take with MMSIG_DENSE_TERM_MAJOR).  Host only, no handle; errors
 * are reported through mmsig_last_error(NULL).  Term and sample names stay with the caller. */
int32_t mmsig_tsv_dims(const char *path, int64_t *V, int64_t *D);
int32_t mmsig_tsv_read(const char *path, int64_t V, int64_t D, int32_t *dense_term_major);

/* ---- MMCTM / CTM  (reference src/MMCTM.jl) -------------------------------------------- */
/* model.X, K, V (src/MMCTM.jl:29-40); with mmsig_comm_init, D and the CSR are this rank's shard
 * and D_total is the global sample count (else pass D_total = D). */
int32_t mmsig_mmctm_set_data(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M,
                             const int32_t *K, const int32_t *V,
                             const int64_t *const *rowptr, const int32_t *const *term,
                             const int32_t *const *count);
/* model.α, γ, λ, ν, μ, Σ, invΣ.  NULL => the constructor's value (src/MMCTM.jl:44-46,82-83):
 * λ=0, ν=1, μ=0, Σ=invΣ=I.  alpha and gamma are required.  Elnϕ is derived (src/MMCTM.jl:78-79). */
int32_t mmsig_mmctm_set_state(mmsig_handle *h, const double *alpha, const double *gamma,
                              const double *lambda, const double *nu, const double *mu,
                              const double *Sigma, const double *invSigma);
/* model.α as update_α! left it */
int32_t mmsig_mmctm_get_alpha(mmsig_handle *h, double *alpha_out);
/* model.ϕ override (fit_heldout / transform copy the fitted model's ϕ, src/MMCTM.jl:515,563) */
int32_t mmsig_mmctm_set_phi(mmsig_handle *h, const double *phi);
/* one body of fit!'s loop (src/MMCTM.jl:463-479): E-step over all samples, μ, [Σ, invΣ], γ, Elnϕ,
 * props, ϕ, per-modality log-likelihoods -> ll_out[M].  flags = MMSIG_FLAG_UPDATE_SIGMA for fit!;
 * FREEZE_TOPICS|FREEZE_MU is the loop body of fit_heldout (src/MMCTM.jl:566-573) and
 * predict_modality_η (:604-609); adding UNSMOOTHED gives transform's (:523-538). */
int32_t mmsig_mmctm_iterate(mmsig_handle *h, uint32_t flags, double *ll_out);
/* fit! (src/MMCTM.jl:457-494): loop + `length(ll) > 10 && check_convergence` (src/common.jl:48-51);
 * ll_hist is maxiter x M.  The ELBO of :490 is mmsig_mmctm_elbo.  The loop runs without a host round trip per
 * iteration: the rule is evaluated on the device and the iterations are enqueued in batches; the state left behind
 * is that of the iteration where the rule fired, as in the reference (autoα: one round trip per iteration). */
int32_t mmsig_mmctm_fit(mmsig_handle *h, int32_t maxiter, double tol, uint32_t flags,
                        double *ll_hist, int32_t *n_iter, int32_t *converged);
/* fit! from and to HOST buffers in one call: exactly mmsig_mmctm_set_data + _set_state + _fit +
 * _get_state (same arguments, same results bit for bit), with the transfers hidden behind the
 * E-step: the samples are cut into chunks, chunk c's counts / λ / ν are copied while chunk c-1
 * runs, and when the loop ends by maxiter the chunk's λ, ν, ζ, props are copied back while the next
 * chunk runs.  This is what julia/MMSigB200.jl's fit! calls.  Page-locked host buffers overlap
 * fully; pageable ones still pipeline chunk by chunk.  The counts and the final state stay
 * resident (mmsig_mmctm_elbo, _get_theta, further _iterate calls work afterwards). */
int32_t mmsig_mmctm_fit_host(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                             const int32_t *V, const int64_t *const *rowptr, const int32_t *const *term,
                             const int32_t *const *count, const double *alpha, const double *gamma,
                             const double *lambda, const double *nu, const double *mu, const double *Sigma,
                             const double *invSigma, int32_t maxiter, double tol, uint32_t flags,
                             double *ll_hist, int32_t *n_iter, int32_t *converged, double *lambda_out,
                             double *nu_out, double *zeta_out, double *mu_out, double *Sigma_out,
                             double *invSigma_out, double *gamma_out, double *Elnphi_out, double *phi_out,
                             double *props_out);
/* The same call with 4-byte count records, which halves the host -> device traffic of the counts (the transfer an
 * end-to-end fit! of a million samples is bound by): rec[m][w] = term | count << 10 (term < 1024, count < 2^22;
 * mmsig_pack_records builds them and answers MMSIG_ELIMIT when a value does not fit).  julia/MMSigB200.jl's
 * flatten_counts writes this form directly. */
int32_t mmsig_mmctm_fit_host_packed(mmsig_handle *h, int64_t D, int64_t D_total, int32_t M, const int32_t *K,
                                    const int32_t *V, const int64_t *const *rowptr, const uint32_t *const *rec,
                                    const double *alpha, const double *gamma, const double *lambda, const double *nu,
                                    const double *mu, const double *Sigma, const double *invSigma, int32_t maxiter,
                                    double tol, uint32_t flags, double *ll_hist, int32_t *n_iter, int32_t *converged,
                                    double *lambda_out, double *nu_out, double *zeta_out, double *mu_out,
                                    double *Sigma_out, double *invSigma_out, double *gamma_out, double *Elnphi_out,
                                    double *phi_out, double *props_out);
int32_t mmsig_pack_records(int64_t nnz, const int32_t *term, const int32_t *count, uint32_t *rec);
/* calculate_elbo (src/MMCTM.jl:271-382) with the staleness of :490: θ, ζ, sumθ from the last
 * E-step, everything else current.  terms[7] = ElnPϕ, ElnPη, ElnPZ, ElnPX, ElnQϕ, ElnQη, ElnQZ. */
int32_t mmsig_mmctm_elbo(mmsig_handle *h, double *elbo, double *terms);
int32_t mmsig_mmctm_get_state(mmsig_handle *h, double *lambda, double *nu, double *zeta,
                              double *mu, double *Sigma, double *invSigma, double *gamma,
                              double *Elnphi, double *phi, double *props);
/* model.θ[d][m] for all d of modality m, nnz_m x K_m row-major ([w][k]); recomputed lazily from
 * the λ / Elnϕ the last E-step used (the library never stores θ). */
int32_t mmsig_mmctm_get_theta(mmsig_handle *h, int32_t m, double *theta_out);
/* Independent random restarts on the resident counts (README.md:42 "fit many models and pick the
 * best one"; scripts/run_mmctm.jl:77-111): for r < R, reset the state to the constructor's with
 * gamma0[r] (R x sum K_m V_m), run mmsig_mmctm_fit, take the ELBO.  elbo_out[R], ll_out[R x M],
 * n_iter_out[R]; *best = argmax ELBO and the handle is left holding that restart's final state.
 * Restarts need no communication: with several GPUs give each handle its own slice of restarts
 * (each handle holds the full counts) and take the arg-max of the returned ELBOs on the host. */
int32_t mmsig_mmctm_restarts(mmsig_handle *h, int32_t R, const double *gamma0, int32_t maxiter, double tol,
                             uint32_t flags, double *elbo_out, double *ll_out, int32_t *n_iter_out,
                             int32_t *best);
/* diagnostics: objective evaluations LD_MMA spent per sample in the last E-step */
int32_t mmsig_mmctm_get_evals(mmsig_handle *h, int32_t *nev_nu, int32_t *nev_lambda);

/* ---- IMMCTM  (reference src/IMMCTM.jl): topics factorised over features --------------------------
 * phi_kv = prod_i phi_k,i,f(v,i).  The per-sample E-step (src/IMMCTM.jl:105-172, :518-523) is the
 * MMCTM's over composite K x V tables; the M-step (:174-221) runs over the feature tables.
 * Call order: mmsig_mmctm_set_data, mmsig_immctm_set_features, mmsig_immctm_set_state, then
 * mmsig_mmctm_iterate / _fit / _elbo / _get_state (lambda, nu, zeta, mu, Sigma, invSigma, props; gamma is
 * a placeholder, Elnphi / phi are the composite tables), mmsig_immctm_get_tables.
 * features[m]: V_m x I_m row-major, 0-BASED feature values (model.features[m] - 1); tables are flat
 * [m][k][i][j] (model.γ[m][k][i][j]), alpha [m][i] (model.α[m][i]).  MMSIG_FLAG_AUTO_ALPHA is
 * update_α! of :223-241; MMSIG_FLAG_UNSMOOTHED is refused (the IMMCTM has no transform).
 * mmsig_mmctm_restarts takes R feature tables ([r][m][k][i][j]) as gamma0 in this mode; mmsig_mmctm_fit_host
 * is the MMCTM's only. */
int32_t mmsig_immctm_set_features(mmsig_handle *h, const int32_t *nfeat, const int32_t *const *features);
int32_t mmsig_immctm_set_state(mmsig_handle *h, const double *alphaf, const double *gammaf, const double *lambda,
                               const double *nu, const double *mu, const double *Sigma, const double *invSigma);
int32_t mmsig_immctm_get_tables(mmsig_handle *h, double *gammaf, double *Elnphif, double *alphaf);

/* ---- LDA  (reference src/LDA.jl) -------------------------------------------------------- */
int32_t mmsig_lda_set_data(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V,
                           const int64_t *rowptr, const int32_t *term, const int32_t *count);
/* model.α, η, λ (required, [k][v]); gamma [d][k] or NULL => constructor state (γ=1, ϕ=1/K,
 * src/LDA.jl:41-49), i.e. the first update_γ! gives α + N_d/K. */
int32_t mmsig_lda_set_state(mmsig_handle *h, double alpha, double eta, const double *lambda,
                            const double *gamma_next);
/* one body of fit!'s loop (src/LDA.jl:202-209) -> *ll_out */
int32_t mmsig_lda_iterate(mmsig_handle *h, double *ll_out);
/* model.β override, and the loop bodies of fit_heldout (src/LDA.jl:275-280: FREEZE_TOPICS) and
 * transform (:242-246: FREEZE_TOPICS | UNSMOOTHED) */
int32_t mmsig_lda_set_beta(mmsig_handle *h, const double *beta);
int32_t mmsig_lda_iterate_flags(mmsig_handle *h, uint32_t flags, double *ll_out);
int32_t mmsig_lda_fit(mmsig_handle *h, int32_t maxiter, double tol, double *ll_hist,
                      int32_t *n_iter, int32_t *converged);
/* fit!(model::LDA) from and to host buffers in one call: mmsig_lda_set_data + _set_state + _fit + _get_state */
int32_t mmsig_lda_fit_host(mmsig_handle *h, int64_t D, int64_t D_total, int32_t K, int32_t V, const int64_t *rowptr,
                           const int32_t *term, const int32_t *count, double alpha, double eta, const double *lambda,
                           const double *gamma_next, int32_t maxiter, double tol, double *ll_hist, int32_t *n_iter,
                           int32_t *converged, double *lambda_out, double *Elnbeta_out, double *beta_out,
                           double *gamma_out, double *Elntheta_out, double *theta_out);
/* calculate_elbo (src/LDA.jl:114-172); terms[7] = ElnPβ, ElnPθ, ElnPZ, ElnPX, ElnQβ, ElnQθ, ElnQZ */
int32_t mmsig_lda_elbo(mmsig_handle *h, double *elbo, double *terms);
int32_t mmsig_lda_get_state(mmsig_handle *h, double *lambda, double *Elnbeta, double *beta,
                            double *gamma, double *Elntheta, double *theta);
/* model.ϕ[d] for all d, nnz x K row-major ([w][k]), recomputed lazily */
int32_t mmsig_lda_get_phi(mmsig_handle *h, double *phi_out);

/* ---- ILDA  (reference src/ILDA.jl): LDA whose topics factorise over features ---------------------
 * beta_kv = prod_i beta_i[f(v,i), k].  The per-sample passes (update_phi / update_gamma / loglikelihood,
 * src/ILDA.jl:60-103, :183-199) are the LDA's over composite K x V tables; the M-step (:105-129) runs over
 * the feature tables.  Call order: mmsig_lda_set_data, mmsig_ilda_set_features, mmsig_ilda_set_state, then
 * mmsig_lda_iterate / _iterate_flags / _fit / _elbo / _get_state (lambda there = the statistics sum n phi of
 * the last M-step; Elnbeta / beta are the composite tables), mmsig_ilda_get_tables.
 * features: V x I row-major, 0-BASED feature values (model.features - 1); tables are flat [k][i][j]
 * (model.lambda[i][j, k]); eta [i] (model.eta[i], :54-58 fills a scalar).  mmsig_lda_elbo evaluates
 * calculate_ElnQbeta as written at :174-181 (only the last feature contributes). */
int32_t mmsig_ilda_set_features(mmsig_handle *h, int32_t nfeat, const int32_t *features);
int32_t mmsig_ilda_set_state(mmsig_handle *h, double alpha, const double *eta, const double *lambdaf,
                             const double *gamma_next);
int32_t mmsig_ilda_get_tables(mmsig_handle *h, double *lambdaf, double *Elnbetaf);

/* ---- test hook: the pinned device math on arrays (fn 0 = exp, 1 = log, 2 = digamma; the branch-free
 * IEEE sequences of det_math.cuh: 3 = x[i] / x[n+i] (x holds 2n values), 4 = 1 / x, 5 = sqrt) ---- */
int32_t mmsig_debug_math(mmsig_handle *h, int32_t fn, int64_t n, const double *x, double *y);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t mmsig_launch_count(const mmsig_handle *h);
/* with cfg.profile: per-kernel accumulated device time.  names_out: n_max pointers to static
 * strings; returns the number of kernels filled.  reset != 0 clears the accumulators. */
int32_t mmsig_kernel_times(mmsig_handle *h, int32_t n_max, const char **names_out,
                           double *ms_total_out, int64_t *launches_out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* MMSIG_H */
