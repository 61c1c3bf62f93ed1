// mmctm_kernels.cuh -- sm_100a kernels for one MMCTM / CTM variational-EM iteration
// (reference src/MMCTM.jl:450-479; dataflow in DESIGN.md).  FP64, pinned arithmetic.
//
//   k_theta_tile    (theta_tile.cuh) the θ pass of one modality as skinny products over sample
//                   tiles: sumθ per sample and the K x V statistics Σ_d exp(λ) n / Z
//   k_solve         per-sample ζ, then LD_MMA for ν and for λ (one warp per sample,
//                   lane j = coordinate j), Σλ / Σν partials
//   k_combine       block partials -> one double-double vector
//   k_mstep1        γ, Elnϕ, ϕ, μ
//   k_moments       ΣΔΔᵀ partials (k_loglik_tile, theta_tile.cuh: props and the log-likelihood pass)
//   k_mstep2        Σ, invΣ (LU with partial pivoting, one warp), LL
#pragma once
#include "det_math.cuh"

namespace mmsig {

constexpr int MAXM = 8;          // modalities
constexpr int MAXMK = 64;        // ΣK_m: one coordinate per lane up to 32, two per lane up to 64 (mmctm_wide.cuh)
constexpr int MMA_MAXEVAL = 10000;
#ifndef MATVEC_UNROLL
#define MATVEC_UNROLL 4
#endif
constexpr int kMatvecUnroll = MATVEC_UNROLL;
constexpr int SROW_STRIDE = 34;   // doubles per padded row of invΣ in shared memory (>= 32, 16-byte multiple, odd multiple of 16 B / 8)
#ifndef SOLVE_MIN_BLOCKS
#define SOLVE_MIN_BLOCKS 3
#endif

struct MmctmDev {
    int M, MK;
    long long D, D_total;
    int K[MAXM], V[MAXM], koff[MAXM + 1], goff[MAXM + 1];
    const long long *rowptr[MAXM];
    const int2 *rec[MAXM];        // (term, count) per nonzero
    const int *cnt[MAXM];         // dense count tiles [ceil(D / 32)][32][V] of a dense modality (tile_stage.cuh), else null
    const double *N;              // D x M
    double Ntot[MAXM];            // Σ_d N_dm over ALL ranks
    double *lam, *lam_prev, *nu, *zeta, *sumtheta;
    double *gamma, *Elnphi, *Elnphi_prev, *phi, *stats, *alpha;
    double *mu, *Sigma, *invSigma;
    double2 *nusum;               // MK, rank-summed Σ_d ν (dd) from mstep1 for mstep2
    int *nev_nu, *nev_lam;
    int stop_rule;
    int accum;                    // != 0: E-step kernels add their block partials to what the slot holds (chunked launches)
    unsigned long long *work;     // sample counter of the solve kernels (zeroed before each launch)
    int *ctl;                     // [0] != 0: the convergence rule fired in an earlier iteration of the enqueued batch -- every
                                  // kernel of an iteration returns at once; [1] last iteration k_mstep2 completed
    // IMMCTM (reference src/IMMCTM.jl): topics factorised over features.  factored != 0: Elnphi / phi
    // (K x V) are COMPOSITE tables derived from the feature tables gammaf / Elnphif ([m][k][i][j] flat)
    int factored, T, R;           // T entries, R rows (m, k, i) of the feature tables
    int nfeat[MAXM], foff[MAXM + 1], aoff[MAXM + 1];
    const int *feat[MAXM];        // V_m x I_m row-major, 0-based feature values
    const int *ent_row;           // [T] row of an entry
    const int *row_off, *row_len, *row_alpha;   // [R] first entry, J, index into alphaf
    double *gammaf, *Elnphif, *alphaf;
};

// Dynamic distribution of samples over the warps (lane groups) of a solve kernel: the cost of a
// sample is its number of MMA evaluations (40-90 here), so a static stride leaves the kernel
// waiting for its unluckiest warp (+5 % at 280 samples per warp, +13 % at 40, i.e. in the chunked
// launches of mmsig_mmctm_fit_host).  One atomicAdd per sample against ~16 k instructions of work.
__device__ __forceinline__ long long next_sample(unsigned long long *work, unsigned mask, int leader, bool is_leader) {
    unsigned long long d = 0;
    if (is_leader) d = atomicAdd(work, 1ULL);
    const unsigned hi = (unsigned)__shfl_sync(mask, (int)(d >> 32), leader);
    const unsigned lo = (unsigned)__shfl_sync(mask, (int)(d & 0xffffffffu), leader);
    return (long long)(((unsigned long long)hi << 32) | lo);
}

// block partial -> its slot; chunked E-steps (mmsig_mmctm_fit_host) launch the same grid once per
// chunk of samples on one stream and accumulate
__device__ __forceinline__ void put_partial(double2 *slot, double hi, double lo, int accum) {
    if (accum) {
        const double2 o = *slot;
        dd_merge(hi, lo, o.x, o.y);
    }
    *slot = make_double2(hi, lo);
}

// ------------------------------------------------------------------------------------------
// NLopt LD_MMA with zero constraints, one warp per problem, lane j = coordinate j.
// Follows the pinned-arithmetic (DET) specification of the recurrence op for op.
// ------------------------------------------------------------------------------------------
struct SolveCtx {
    double Sjj;        // invΣ[j][j]
    double muj;        // μ[j]
    double c;          // N_dm / ζ_dm of this lane's modality
    double s;          // sumθ[j]
    double other;      // ν solve: λ_j ; λ solve: 0.5 ν_j
    bool active;       // lane < MK
};

// Lane-local part of one objective evaluation: the per-coordinate term t (to be tree-summed
// into the maximised objective) and the gradient.  Inactive lanes (>= MK) carry a dummy
// problem with a non-zero gradient so that no division ever sees a zero numerator (IEEE double
// division takes a ~65-instruction slow path for zero / subnormal operands, and a warp pays
// for it if any lane does); their terms are masked out of every reduction.
template <int MKP, bool IS_NU>
__device__ __forceinline__ void mma_eval_local(double x, const SolveCtx &c, const double *__restrict__ ST,
                                               double *dsh, int lane, double &t, double &g) {
    double grad;
    if (IS_NU) {
        // src/common.jl:25-36, maximised; fused per-coordinate term (DET)
        const double e = det_exp(c.other + 0.5 * x);
        grad = (-0.5 * c.Sjj - (c.c / 2) * e) + fast_rcp(2 * x);
        t = (-0.5 * (x * c.Sjj) - c.c * e) + det_log(x) / 2;
    } else {
        // src/common.jl:11-23
        const double diff = x - c.muj;
        const double e = det_exp(x + c.other);
        dsh[lane] = c.active ? diff : 0.0;          // all 32 lanes store: no divergent branch around the STS
        __syncwarp();
        double q0 = 0.0, q1 = 0.0;                  // DET: even / odd index fma chains, then one add
        // row `lane` of invΣ, padded to SROW_STRIDE doubles (208 B): 128-bit loads of neighbouring
        // lanes fall into distinct bank groups, so both operands come in as LDS.128
        const double2 *srow = reinterpret_cast<const double2 *>(ST + lane * SROW_STRIDE);
        const double2 *dv2 = reinterpret_cast<const double2 *>(dsh);
#pragma unroll kMatvecUnroll
        for (int i = 0; i < MKP / 2; ++i) {
            const double2 sv = srow[i], dv = dv2[i];
            q0 = fma(sv.x, dv.x, q0);
            q1 = fma(sv.y, dv.y, q1);
        }
        const double q = q0 + q1;
        __syncwarp();
        const double ce = c.c * e;
        grad = (-q + c.s) - ce;
        const double a = q * diff, b = x * c.s;
        t = (b - 0.5 * a) - ce;
    }
    if (!c.active) { t = 0.0; grad = -1.0; }
    g = -grad;                       // NLopt minimises the negated objective
}

template <int MKP, bool IS_NU>
__device__ __forceinline__ int mma_solve(double &x, const SolveCtx &c, const double *__restrict__ ST,
                                         double *dsh, int lane, int stop_rule) {
    const double lb = IS_NU ? 1e-7 : -__longlong_as_double(0x7ff0000000000000LL);
    const double xtol_rel = 1e-4, xtol_abs = 1e-4;
    double sigma = 1.0, isig = 1.0;  // a bound is infinite -> sigma = 1; isig = 1 / sigma, refreshed when sigma moves
    double rho = 1.0;
    double g, fmin, fcur, gcur;
    {
        double t;
        mma_eval_local<MKP, IS_NU>(x, c, ST, dsh, lane, t, g);
        fmin = -warp_tree_sum(t);
    }
    int nev = 1;
    double xcur = x, xprev = x, xprevprev = x;
    int k = 0;
    while (true) {
        if (++k > 1) xprevprev = xprev;
        xprev = xcur;
        while (true) {
            double u = g;
            const double v = fabs(g) * sigma + 0.5 * rho;
            const double sigma2 = sigma * sigma;
            u *= sigma2;
            const double qv = fast_div(u, v);
            const double r = qv * isig;               // DET: u / (v sigma) as (u / v)(1 / sigma): one division per evaluation less
            const double om = fabs(1 - r * r);
            const double sq = fast_sqrt(om < 0x1p-200 ? 0x1p-200 : om);   // om is 0 or >= 2^-53: sqrt(0) -> 2^-100, and -1 - 2^-100 == -1
            double dx = fast_div(qv, -1 - sq);
            double xc = x + dx;
            const double mv = 0.9 * sigma, xhi = x + mv, xlo = x - mv;          // move limits: selects, no branches
            xc = xc > xhi ? xhi : (xc < xlo ? xlo : xc);
            if (xc < lb) xc = lb;
            if (!c.active) xc = x;                   // dummy lanes stay put
            dx = xc - x;
            const double dx2 = dx * dx;
            const double denominv = fast_rcp(sigma2 - dx2);       // |dx| <= 0.9 sigma
            const double cc = sigma2 * dx;
            double gterm = (g * cc + (fabs(g) * sigma + 0.5 * rho) * dx2) * denominv;
            double wterm = 0.5 * dx2 * denominv;
            xcur = xc;
            // the new point's lane-local work first, then ONE 3-way butterfly for gval, wval, f
            double t;
            mma_eval_local<MKP, IS_NU>(xcur, c, ST, dsh, lane, t, gcur);
            warp_tree_sum3h(gterm, wterm, t, lane);
            const double gval = fmin + gterm, wval = wterm;
            fcur = -t;
            ++nev;
            const bool inner_done = __all_sync(FULLMASK, gval >= fcur);   // identical in all lanes; the vote tells the compiler so
            if (fcur < fmin) { fmin = fcur; x = xcur; g = gcur; }
            if (nev >= MMA_MAXEVAL) return nev;           // nev is lane-invariant
            if (inner_done) break;
            if (__all_sync(FULLMASK, fcur > gval)) {
                const double r1 = 10 * rho, r2 = 1.1 * (rho + guarded_div(fcur - gval, wval));
                rho = r1 < r2 ? r1 : r2;
            }
        }
        // x-tolerance (NLopt stop.c) on (xcur, xprev)
        const double ad = c.active ? fabs(xcur - xprev) : 0.0;
        bool stop;
        if (stop_rule == 1) {       // NLopt <= 2.6: per coordinate
            const bool ok = ad < xtol_abs || ad < xtol_rel * (fabs(xcur) + fabs(xprev)) * 0.5 || xcur == xprev;
            stop = __all_sync(FULLMASK, ok || !c.active);
        } else {                    // NLopt >= 2.7: L1 norms, else all |dx| <= xtol_abs
            double dn = ad, xn = c.active ? fabs(xcur) : 0.0;
            warp_tree_sum2h(dn, xn, lane);
            stop = __all_sync(FULLMASK, dn <= xtol_rel * xn) || __all_sync(FULLMASK, !(ad > xtol_abs));
        }
        if (stop) break;
        rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;
        if (k > 1) {
            const double s2 = (xcur - xprev) * (xprev - xprevprev);
            const double gam = s2 < 0 ? 0.7 : (s2 > 0 ? 1.2 : 1.0);
            sigma *= gam;
            isig = fast_rcp(sigma);
        }
    }
    return nev;
}

// sequential sum over this lane's modality block [lo, hi) of a per-lane value
__device__ __forceinline__ double block_sum_seq(double e, int lo, int hi, int MK) {
    double s = 0.0;
    for (int i = 0; i < MK; ++i) {
        const double v = shfl_d(e, i);
        if (i >= lo && i < hi) s += v;
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// fitdoc! minus the θ pass (src/MMCTM.jl:450-455): ζ from the old λ, ν (:172-181), then
// update_ν! (:156-170) and update_λ! (:127-143, with the new ν, old ζ, old sumθ).
// partial: [gridDim.x][2*MK] dd of Σ_d λ_new and Σ_d ν_new.
// ------------------------------------------------------------------------------------------
template <int MKP>
__global__ void __launch_bounds__(256, SOLVE_MIN_BLOCKS) k_solve(MmctmDev p, double2 *partial) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    __shared__ __align__(16) double dsh_all[8][32];
    __shared__ double2 red[8][2][32];
    __shared__ __align__(16) double ST[32 * SROW_STRIDE];    // invΣ rows, zero padded: ST[j*SROW_STRIDE+i] = invΣ[j][i]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = p.MK, M = p.M;
    double *dsh = dsh_all[warp];
    dsh[lane] = 0.0;
    const bool active = lane < MK;
    int mod = 0;
    for (int m = 0; m < M; ++m)
        if (lane >= p.koff[m]) mod = m;
    const int blo = p.koff[mod], bhi = p.koff[mod + 1];

    for (int t = threadIdx.x; t < 32 * SROW_STRIDE; t += blockDim.x) {
        const int j = t / SROW_STRIDE, i = t % SROW_STRIDE;
        ST[t] = (i < MK && j < MK) ? p.invSigma[j * MK + i] : 0.0;
    }
    __syncthreads();
    SolveCtx c;
    c.active = active;
    c.Sjj = active ? p.invSigma[lane * MK + lane] : 0.0;
    c.muj = active ? p.mu[lane] : 0.0;

    double lsh = 0.0, lsl = 0.0, nsh = 0.0, nsl = 0.0;
    for (long long d = next_sample(p.work, FULLMASK, 0, lane == 0); d < p.D; d = next_sample(p.work, FULLMASK, 0, lane == 0)) {
        const long long base = d * MK + lane;
        double lam = active ? p.lam_prev[base] : 0.0;
        double nu = active ? p.nu[base] : 1.5;
        c.s = active ? p.sumtheta[base] : 0.0;
        // ζ_dm = Σ_{k in block m} exp(λ + ν/2), index order
        const double e0 = active ? det_exp(lam + 0.5 * nu) : 0.0;
        const double zeta = block_sum_seq(e0, blo, bhi, MK);
        const double Ndm = active ? p.N[d * M + mod] : 0.0;
        c.c = active ? Ndm / zeta : 0.0;
        if (active && lane == blo) p.zeta[d * M + mod] = zeta;
        // ν
        c.other = lam;
        const int nev_nu = mma_solve<MKP, true>(nu, c, ST, dsh, lane, p.stop_rule);
        // λ (new ν, old ζ)
        c.other = 0.5 * nu;
        const int nev_lam = mma_solve<MKP, false>(lam, c, ST, dsh, lane, p.stop_rule);
        if (active) {
            p.lam[base] = lam;
            p.nu[base] = nu;
            dd_add(lsh, lsl, lam);
            dd_add(nsh, nsl, nu);
        }
        if (lane == 0) { p.nev_nu[d] = nev_nu; p.nev_lam[d] = nev_lam; }
    }
    red[warp][0][lane] = make_double2(lsh, lsl);
    red[warp][1][lane] = make_double2(nsh, nsl);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int which = threadIdx.x >> 5;
        double hi = 0.0, lo = 0.0;
        for (int wv = 0; wv < 8; ++wv) dd_merge(hi, lo, red[wv][which][lane].x, red[wv][which][lane].y);
        if (lane < MK) put_partial(partial + (size_t)blockIdx.x * 2 * MK + which * MK + lane, hi, lo, p.accum);
    }
}

// ------------------------------------------------------------------------------------------
// block partials -> one dd vector.  Segment s: dst[dst_off + i] = Σ_part src[part*n + i].
// ------------------------------------------------------------------------------------------
struct CombineSegs {
    int nseg;
    const double2 *src[MAXM + 2];
    int nparts[MAXM + 2], n[MAXM + 2], dst_off[MAXM + 2];
    int stride[MAXM + 2];        // distance between parts in src (0 = n)
};
__global__ void __launch_bounds__(256) k_combine(CombineSegs s, double2 *dst, const int *ctl = nullptr) {
    if (ctl && ctl[0]) return;
    // one warp per output entry; lanes stride over the parts, then a dd butterfly
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    int base = 0;
    for (int g = 0; g < s.nseg; ++g) {
        if (t < base + s.n[g]) {
            const int i = t - base;
            const int stride = s.stride[g] ? s.stride[g] : s.n[g];
            double hi = 0.0, lo = 0.0;
            // four loads in flight per lane, merged in part order (the order of the one-at-a-time loop)
            const double2 *src = s.src[g] + i;
            const int np = s.nparts[g];
            int part = lane;
            for (; part + 96 < np; part += 128) {
                const double2 v0 = src[(size_t)part * stride], v1 = src[(size_t)(part + 32) * stride];
                const double2 v2 = src[(size_t)(part + 64) * stride], v3 = src[(size_t)(part + 96) * stride];
                dd_merge(hi, lo, v0.x, v0.y);
                dd_merge(hi, lo, v1.x, v1.y);
                dd_merge(hi, lo, v2.x, v2.y);
                dd_merge(hi, lo, v3.x, v3.y);
            }
            for (; part < np; part += 32) {
                const double2 v = src[(size_t)part * stride];
                dd_merge(hi, lo, v.x, v.y);
            }
            warp_dd_allreduce(hi, lo);
            if (lane == 0) dst[s.dst_off[g] + i] = make_double2(hi, lo);
            return;
        }
        base += s.n[g];
    }
}

// ------------------------------------------------------------------------------------------
// M-step, part 1 (single block).  gathered: [nranks][G + 2 MK] dd.
// γ = exact_round(α + Σ n θ) (src/MMCTM.jl:224-240), Elnϕ = ψ(γ) - ψ(Σ_v γ) (:214-222),
// ϕ = γ / Σ_v γ (:244-250), μ = exact_round(Σ_d λ) / D (:200-202).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_mstep1(MmctmDev p, const double2 *gathered, int nranks, int freeze_topics,
                                                 int freeze_mu, int unsmoothed) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    __shared__ double rowsum[MAXMK], rowdig[MAXMK];
    const int G = p.goff[p.M], MK = p.MK, P1 = G + 2 * MK;
    for (int i = threadIdx.x; i < P1; i += blockDim.x) {
        double hi = 0.0, lo = 0.0;
        for (int r = 0; r < nranks; ++r) {
            const double2 v = gathered[(size_t)r * P1 + i];
            dd_merge(hi, lo, v.x, v.y);
        }
        if (i < G) {
            if (freeze_topics) continue;
            int m = 0;
            while (i >= p.goff[m + 1]) ++m;
            // gathered: S_kv = Σ_d exp(λ_dk) n_dv / Z_dv (k_theta_tile); Σ n θ_kv = E_kv S_kv with the
            // table the E-step used, which is still in place here
            const double S = dd_round(hi, lo);
            const double E = unsmoothed ? p.phi[i] : det_exp(p.Elnphi[i]);
            p.stats[i] = E * S;
            p.gamma[i] = fma(E, S, p.alpha[m]);
            p.Elnphi_prev[i] = p.Elnphi[i];
        } else if (i < G + MK) {
            if (!freeze_mu) p.mu[i - G] = dd_round(hi, lo) / (double)p.D_total;
        } else {
            p.nusum[i - G - MK] = make_double2(hi, lo);
        }
    }
    if (freeze_topics) return;
    __syncthreads();
    if (threadIdx.x < MK) {
        int m = 0;
        while ((int)threadIdx.x >= p.koff[m + 1]) ++m;
        const int k = threadIdx.x - p.koff[m];
        const double *g = p.gamma + p.goff[m] + k * p.V[m];
        double s = 0.0;
        for (int v = 0; v < p.V[m]; ++v) s += g[v];
        rowsum[threadIdx.x] = s;
        rowdig[threadIdx.x] = det_digamma(s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        int m = 0;
        while (i >= p.goff[m + 1]) ++m;
        const int row = p.koff[m] + (i - p.goff[m]) / p.V[m];
        const double gv = p.gamma[i];
        p.Elnphi[i] = det_digamma(gv) - rowdig[row];
        p.phi[i] = gv / rowsum[row];
    }
}

// Elnϕ from γ only (constructor, src/MMCTM.jl:78-79)
__global__ void __launch_bounds__(1024) k_elnphi(MmctmDev p) {
    __shared__ double rowdig[MAXMK];
    const int G = p.goff[p.M], MK = p.MK;
    if (threadIdx.x < MK) {
        int m = 0;
        while ((int)threadIdx.x >= p.koff[m + 1]) ++m;
        const int k = threadIdx.x - p.koff[m];
        const double *g = p.gamma + p.goff[m] + k * p.V[m];
        double s = 0.0;
        for (int v = 0; v < p.V[m]; ++v) s += g[v];
        rowdig[threadIdx.x] = det_digamma(s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        int m = 0;
        while (i >= p.goff[m + 1]) ++m;
        const int row = p.koff[m] + (i - p.goff[m]) / p.V[m];
        p.Elnphi[i] = det_digamma(p.gamma[i]) - rowdig[row];
        p.Elnphi_prev[i] = p.Elnphi[i];
        p.phi[i] = p.gamma[i];                  // model.ϕ = deepcopy(model.γ), src/MMCTM.jl:80
    }
}

// ------------------------------------------------------------------------------------------
// IMMCTM M-step, part 1 (single block; reference src/IMMCTM.jl:186-221).  From the feature tables
// γf: Elnϕf = ψ(γf) - ψ(Σ_j γf), then the composite tables every per-sample kernel reads,
// Elnϕ_kv = Σ_i Elnϕf_k,i,f(v,i) (index order, from 0) and ϕ_kv = Π_i γf / Σ_j γf (index order,
// from 1).  rowsum / rowdig: R doubles each in shared memory.
// ------------------------------------------------------------------------------------------
__device__ inline void immctm_compose(const MmctmDev &p, double *rowsum, double *rowdig) {
    for (int r = threadIdx.x; r < p.R; r += blockDim.x) {
        const double *g = p.gammaf + p.row_off[r];
        double s = 0.0;
        for (int j = 0; j < p.row_len[r]; ++j) s += g[j];
        rowsum[r] = s;
        rowdig[r] = det_digamma(s);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) p.Elnphif[t] = det_digamma(p.gammaf[t]) - rowdig[p.ent_row[t]];
    __syncthreads();
    const int G = p.goff[p.M];
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        int m = 0;
        while (i >= p.goff[m + 1]) ++m;
        const int V = p.V[m], nf = p.nfeat[m], k = (i - p.goff[m]) / V, v = (i - p.goff[m]) % V;
        // rows of (m, k) are consecutive: first row index = aoff-based count
        int r = 0;
        for (int mm = 0; mm < m; ++mm) r += p.K[mm] * p.nfeat[mm];
        r += k * nf;
        double e = 0.0, ph = 1.0;
        for (int f = 0; f < nf; ++f, ++r) {
            const int t = p.row_off[r] + p.feat[m][v * nf + f];
            e += p.Elnphif[t];
            ph *= p.gammaf[t] / rowsum[r];
        }
        p.Elnphi[i] = e;
        p.phi[i] = ph;
    }
}

// constructor / set_state: composite tables from γf (src/IMMCTM.jl:69-70)
__global__ void __launch_bounds__(1024) k_icompose(MmctmDev p) {
    extern __shared__ double ism[];
    immctm_compose(p, ism, ism + p.R);
    __syncthreads();
    const int G = p.goff[p.M];
    for (int i = threadIdx.x; i < G; i += blockDim.x) { p.Elnphi_prev[i] = p.Elnphi[i]; p.gamma[i] = p.phi[i]; }
}

// gathered: [nranks][G + 2 MK] dd as for k_mstep1.  Σ n θ_kv = E_kv S_kv, then
// γf_k,i,j = α_i + Σ_{v: f(v,i) = j} Σ n θ_kv (ascending v), Elnϕf, the composite tables, μ.
__global__ void __launch_bounds__(1024) k_imstep1(MmctmDev p, const double2 *gathered, int nranks, int freeze_topics,
                                                  int freeze_mu) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ double ism[];
    const int G = p.goff[p.M], MK = p.MK, P1 = G + 2 * MK;
    for (int i = threadIdx.x; i < P1; i += blockDim.x) {
        double hi = 0.0, lo = 0.0;
        for (int r = 0; r < nranks; ++r) {
            const double2 v = gathered[(size_t)r * P1 + i];
            dd_merge(hi, lo, v.x, v.y);
        }
        if (i < G) {
            if (freeze_topics) continue;
            const double S = dd_round(hi, lo);
            p.stats[i] = det_exp(p.Elnphi[i]) * S;
            p.Elnphi_prev[i] = p.Elnphi[i];
        } else if (i < G + MK) {
            if (!freeze_mu) p.mu[i - G] = dd_round(hi, lo) / (double)p.D_total;
        } else {
            p.nusum[i - G - MK] = make_double2(hi, lo);
        }
    }
    if (freeze_topics) return;
    __syncthreads();
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
        const int r = p.ent_row[t], j = t - p.row_off[r];
        // (m, k, f) of row r
        int m = 0, base = 0;
        while (r >= base + p.K[m] * p.nfeat[m]) { base += p.K[m] * p.nfeat[m]; ++m; }
        const int nf = p.nfeat[m], k = (r - base) / nf, f = (r - base) % nf, V = p.V[m];
        const double *st = p.stats + p.goff[m] + k * V;
        const int *ft = p.feat[m];
        double acc = p.alphaf[p.row_alpha[r]];
        for (int v = 0; v < V; ++v)
            if (ft[v * nf + f] == j) acc += st[v];
        p.gammaf[t] = acc;
    }
    __syncthreads();
    immctm_compose(p, ism, ism + p.R);
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += blockDim.x) p.gamma[i] = p.phi[i];      // model.γ has no K x V form here
}

// ------------------------------------------------------------------------------------------
// Moments pass: ΣΔΔᵀ partials with the new μ (src/MMCTM.jl:204-210); lane j owns row j.
// partial: [gridDim.x][MK*MK + M] dd (the last M entries belong to the log-likelihood tiles).
// ------------------------------------------------------------------------------------------
template <int MKP>
__global__ void __launch_bounds__(256) k_moments(MmctmDev p, double2 *partial) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ double smem[];
    const int MK = p.MK, M = p.M;
    double2 *red = reinterpret_cast<double2 *>(smem);      // 256 double2 = 512 doubles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool active = lane < MK;
    const double muj = active ? p.mu[lane] : 0.0;
    double mhi[MKP], mlo[MKP];
#pragma unroll
    for (int i = 0; i < MKP; ++i) { mhi[i] = 0.0; mlo[i] = 0.0; }
    const long long nw = (long long)gridDim.x * 8;
    // the next sample's λ is loaded while the current one is accumulated (the load was 20 % of the kernel's stall
    // samples, profiles/r02e_summary.md); same samples in the same order per warp
    long long d = (long long)blockIdx.x * 8 + warp;
    double lam_next = (active && d < p.D) ? p.lam[d * MK + lane] : 0.0;
    for (; d < p.D; d += nw) {
        const double diff = lam_next - muj;
        const long long dn = d + nw;
        lam_next = (active && dn < p.D) ? p.lam[dn * MK + lane] : 0.0;
#pragma unroll
        for (int i = 0; i < MKP; ++i) {
            const double di = shfl_d(diff, i);
            if (i < MK) dd_add(mhi[i], mlo[i], diff * di);
        }
    }
    // block combine through shared memory, one moment column at a time
    __syncthreads();
    double2 *out = partial + (size_t)blockIdx.x * (MK * MK + M);
#pragma unroll
    for (int i = 0; i < MKP; ++i) {
        red[warp * 32 + lane] = make_double2(mhi[i], mlo[i]);
        __syncthreads();
        if (warp == 0 && active && i < MK) {
            double hi = 0.0, lo = 0.0;
            for (int wv = 0; wv < 8; ++wv) dd_merge(hi, lo, red[wv * 32 + lane].x, red[wv * 32 + lane].y);
            out[lane * MK + i] = make_double2(hi, lo);          // Σ_d diff_lane * diff_i
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// LU with partial pivoting + inverse, one warp, n <= 32; restates the oracle's lu_factor /
// orc_inv (LAPACK getrf/getri as used by Julia's inv, src/MMCTM.jl:211) op for op.
// A: n x n row-major in shared memory (destroyed); B: n x n scratch; out: global row-major.
// returns log|det A| via *logabsdet if not null (plain sum of det_log |u_ii|).
// ------------------------------------------------------------------------------------------
__device__ inline bool warp_lu_inverse(int n, double *A, double *B, int *piv, double *out, double *logabsdet) {
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < n; ++k) {
        // first index of the maximum |A[i][k]|, i >= k (each lane scans rows lane, lane+32)
        double best = -1.0;
        int bi = lane;
        for (int i = lane; i < n; i += 32)
            if (i >= k) {
                const double a = fabs(A[i * n + k]);
                if (a > best) { best = a; bi = i; }
            }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double ob = shfl_xor_d(best, off);
            const int oi = __shfl_xor_sync(FULLMASK, bi, off);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) piv[k] = bi;
        if (best == 0.0) return false;
        __syncwarp();
        if (bi != k)
            for (int j = lane; j < n; j += 32) {
                const double t = A[k * n + j];
                A[k * n + j] = A[bi * n + j];
                A[bi * n + j] = t;
            }
        __syncwarp();
        const double inv = 1.0 / A[k * n + k];
        __syncwarp();
        for (int i = lane; i < n; i += 32)
            if (i > k) A[i * n + k] = A[i * n + k] * inv;
        __syncwarp();
        for (int j = lane; j < n; j += 32)
            if (j > k)
                for (int i = k + 1; i < n; ++i) A[i * n + j] -= A[i * n + k] * A[k * n + j];
        __syncwarp();
    }
    if (logabsdet) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += det_log(fabs(A[i * n + i]));
        *logabsdet = r;
    }
    if (out)
        for (int c = lane; c < n; c += 32) {
            for (int i = 0; i < n; ++i) B[i * n + c] = (i == c) ? 1.0 : 0.0;
            for (int k = 0; k < n; ++k)
                if (piv[k] != k) {
                    const double t = B[k * n + c];
                    B[k * n + c] = B[piv[k] * n + c];
                    B[piv[k] * n + c] = t;
                }
            for (int i = 0; i < n; ++i) {
                double s = B[i * n + c];
                for (int j = 0; j < i; ++j) s -= A[i * n + j] * B[j * n + c];
                B[i * n + c] = s;
            }
            for (int i = n - 1; i >= 0; --i) {
                double s = B[i * n + c];
                for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * B[j * n + c];
                B[i * n + c] = s / A[i * n + i];
            }
            for (int i = 0; i < n; ++i) out[i * n + c] = B[i * n + c];
        }
    __syncwarp();
    return true;
}

// The same factorisation and inverse for n <= NP <= 32 with the matrix in registers: lane j owns column j
// (a[i] = A[i][j]), every loop is unrolled, rows >= n and lanes >= n are inert.  Operation for operation the
// arithmetic of warp_lu_inverse (same pivot rule, same multiply-then-subtract per element, same order of the
// substitution sums), so the results are bit-identical; what changes is where the operands live: no shared-memory
// round trip and no __syncwarp between dependent steps (62 us -> a few us at n = 24).
// M-step, part 2 (one block; the inverse on its first warp).  gathered: [nranks][MK*MK + M] dd.
// Σ = exact_round(diag Σ_d ν + Σ_d ΔΔᵀ) / D, invΣ = inv(Σ) (src/MMCTM.jl:204-212); ll_m (:417).
// (A register-resident, fully unrolled LU -- lane j owning column j -- was tried in round 2: 25 k straight-line SASS
// instructions at MK = 24 run at instruction-fetch speed, 105 us against 64 us for this loop over shared memory.)
// iter > 0: iteration `iter` of a fit whose later iterations are already enqueued: the block evaluates the reference's
// stopping rule (src/MMCTM.jl:485, src/common.jl:48-51: more than 10 entries and max_m |ll_prev - ll| / |ll| < tol, the
// same IEEE operations as converged_vec on the host) and raises p.ctl[0].  ll_prev: the previous iteration's vector.
__global__ void __launch_bounds__(256) k_mstep2(MmctmDev p, const double2 *gathered, int nranks, int do_sigma,
                                                double *ll_out, int *status, int iter, double tol, double *ll_prev) {
    if (p.ctl && p.ctl[0]) return;
    extern __shared__ double lu_smem[];                  // A, B: MK x MK each; piv: MK ints
    const int MK = p.MK, M = p.M, P2 = MK * MK + M;
    double *A = lu_smem, *B = lu_smem + MK * MK;
    int *piv = reinterpret_cast<int *>(B + MK * MK);
    for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        double hi = 0.0, lo = 0.0;
        for (int r = 0; r < nranks; ++r) {
            const double2 v = gathered[(size_t)r * P2 + i];
            dd_merge(hi, lo, v.x, v.y);
        }
        if (i < MK * MK) {
            if (do_sigma) {
                const int a = i / MK, b = i % MK;
                if (a == b) dd_merge(hi, lo, p.nusum[a].x, p.nusum[a].y);
                const double sv = dd_round(hi, lo) / (double)p.D_total;
                p.Sigma[i] = sv;
                A[i] = sv;
            }
        } else {
            const int m = i - MK * MK;
            ll_out[m] = dd_round(hi, lo) / p.Ntot[m];
        }
    }
    __syncthreads();
    if (do_sigma && threadIdx.x < 32) {
        const bool ok = warp_lu_inverse(MK, A, B, piv, p.invSigma, nullptr);
        if (threadIdx.x == 0) *status = ok ? 0 : 1;
    }
    if (threadIdx.x == 32 && ll_prev) {                  // a warp the inverse does not use
        double r = 0.0;
        for (int m = 0; m < M; ++m) {
            const double v = fabs(ll_prev[m] - ll_out[m]) / fabs(ll_out[m]);
            if (v > r || v != v) r = v;
            ll_prev[m] = ll_out[m];
        }
        if (p.ctl) {
            if (iter > 10 && r < tol) p.ctl[0] = 1;
            if (iter > 0) p.ctl[1] = iter;
        }
    }
}

// props only (get_state): softmax of each λ block
__global__ void __launch_bounds__(256) k_props(MmctmDev p, double *props_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = p.MK;
    const bool active = lane < MK;
    int mod = 0;
    for (int m = 0; m < p.M; ++m)
        if (lane >= p.koff[m]) mod = m;
    const int blo = p.koff[mod], bhi = p.koff[mod + 1];
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const double lam = active ? p.lam[d * MK + lane] : 0.0;
        const double e = active ? det_exp(lam) : 0.0;
        const double s = block_sum_seq(e, blo, bhi, MK);
        if (active) props_out[d * MK + lane] = e / s;
    }
}

// ζ from the current λ, ν (constructor, src/MMCTM.jl:85-86)
__global__ void __launch_bounds__(256) k_zeta(MmctmDev p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = p.MK;
    const bool active = lane < MK;
    int mod = 0;
    for (int m = 0; m < p.M; ++m)
        if (lane >= p.koff[m]) mod = m;
    const int blo = p.koff[mod], bhi = p.koff[mod + 1];
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const double e = active ? det_exp(p.lam[d * MK + lane] + 0.5 * p.nu[d * MK + lane]) : 0.0;
        const double s = block_sum_seq(e, blo, bhi, MK);
        if (active && lane == blo) p.zeta[d * p.M + mod] = s;
    }
}

// θ of one modality, materialised on request (model.θ[d][m]); uses the λ / Elnϕ of the last
// E-step (lam_prev, Elnphi_prev).  out: nnz x K.
__global__ void __launch_bounds__(256) k_theta_out(MmctmDev p, int m, double *out, int unsmoothed) {
    extern __shared__ double smem[];
    const int K = p.K[m], V = p.V[m], KV = K * V, off = p.koff[m];
    double *Eln = smem;
    const double *Eg = (unsmoothed ? p.phi : p.Elnphi_prev) + p.goff[m];
    for (int i = threadIdx.x; i < KV; i += blockDim.x) Eln[i] = unsmoothed ? Eg[i] : det_exp(Eg[i]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const long long beg = p.rowptr[m][d], end = p.rowptr[m][d + 1];
        for (long long w = beg + lane; w < end; w += 32) {
            const int v = p.rec[m][w].x & 0xffff;                 // low half: term (high half: tile slot tag)
            double Z = 0.0;
            for (int k = 0; k < K; ++k) {
                const double lk = p.lam_prev[d * p.MK + off + k];
                const double e = det_exp(lk) * Eln[k * V + v];
                out[w * K + k] = e;
                Z += e;
            }
            const double rz = 1.0 / Z;
            for (int k = 0; k < K; ++k) out[w * K + k] = out[w * K + k] * rz;
        }
    }
}

}  // namespace mmsig
