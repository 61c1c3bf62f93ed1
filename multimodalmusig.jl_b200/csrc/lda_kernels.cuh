// lda_kernels.cuh -- LDA variational-EM iteration (reference src/LDA.jl:69-224), FP64.
//
// Iteration t of fit! (src/LDA.jl:201-209):
//   γ_t = α + Σ_w n ϕ_{t-1}   (update_γ!, :82-90; ϕ_0 = 1/K)      <- produced by pass t-1 (γ_next)
//   Elnθ_t = ψ(γ_t) - ψ(Σ_k γ_t)                                   (:78-80)
//   ϕ_t[k,w] ∝ exp(Elnθ_t[k] + Elnβ_{t-1}[v_w,k])                  (update_ϕ!, :69-76)
//   λ_t = η + Σ_d n ϕ_t ; Elnβ_t ; β_t ; θ_t                       (:92-112)
//   ll_t = Σ n log(θ_t[:,d]·β_t[v,:]) / N                          (:174-188)
// ϕ (K x nnz) is never stored: one pass computes ϕ_t on the fly, adds n ϕ_t into the K x V
// statistics and emits γ_{t+1}.  LDA has no data-dependent branches, so its tolerance is the
// plain 1e-12 of north_star and the exponentials are hoisted: ϕ ∝ e^{Elnθ_k} · e^{Elnβ_kv}
// (K exps per sample + a K x V table instead of K·nnz exps).
#pragma once
#include "det_math.cuh"

namespace mmsig {

struct LdaDev {
    int K, V;
    long long D, D_total;
    const long long *rowptr;
    const int2 *rec;
    const int *cnt;              // dense count tiles (tile_stage.cuh) when the corpus is dense and the t32 kernels run, else null
    const double *N;             // D
    double Ntot;                 // Σ_d N_d over all ranks
    double alpha, eta;
    double *lam, *Elnbeta, *Elnbeta_prev, *beta, *expElnbeta, *expElnbeta_prev;   // K x V, [k][v]
    double *gamma, *gamma_next;  // D x K, [d][k]
    // ILDA (reference src/ILDA.jl): beta_kv = prod_i beta_i[f(v,i), k].  factored != 0: Elnbeta / expElnbeta /
    // beta (K x V) are COMPOSITE tables derived from the feature tables lambdaf / Elnbetaf ([k][i][j] flat);
    // lam (K x V) then holds the statistics sum_d n phi of the last M-step
    int factored, nfeat, T, R;
    const int *feat;             // V x I row-major, 0-based feature values
    const int *ent_row;          // [T] row (k, i) of an entry
    const int *row_off, *row_len, *row_eta;   // [R] first entry, J_i, feature index i
    double *lambdaf, *Elnbetaf, *etaf;
};

// Elnθ of one sample on lanes k < K (all lanes get ψ(Σγ) consistently)
__device__ __forceinline__ double lda_elntheta(double gk, int K, int lane) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += shfl_d(gk, k);
    const double ds = det_digamma(s);
    return (lane < K) ? det_digamma(gk) - ds : 0.0;
}

// γ_1 = α + Σ_w (1/K) n_w  (update_γ! on the constructor's ϕ = 1/K, src/LDA.jl:46-49,82-90)
__global__ void __launch_bounds__(256) k_lda_gamma_init(LdaDev p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * 8;
    const double invK = 1.0 / p.K;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        double s = 0.0;
        for (long long w = p.rowptr[d] + lane; w < p.rowptr[d + 1]; w += 32) s += invK * (double)p.rec[w].y;
        s = warp_tree_sum(s);
        if (lane < p.K) p.gamma_next[d * p.K + lane] = p.alpha + s;
    }
}

// One E pass.  partial: [gridDim.x][K*V] (as double2 {sum, 0} so that k_combine can be shared).
// Shared-memory tables are term-major, [v][KP + 2] (topics of one term contiguous, padded so that
// the 128-bit accesses of neighbouring lanes -- neighbouring terms -- fall into distinct bank
// groups): per nonzero the lane streams its term's row with LDS.128 / STS.128 and immediate
// offsets.  Topics k >= K are zero padding (e^{Elnθ} = 0), so the unrolled loops need no guards.
template <int KP, int NP>
__global__ void __launch_bounds__(256) k_lda_estep(LdaDev p, double2 *partial, int nwarps_blk, const double *Etab,
                                                   int want_stats) {
    extern __shared__ double smem[];
    constexpr int KPAD = KP + 2;
    const int K = p.K, V = p.V, TS = V * KPAD;
    double *E = smem;                                  // e^{Elnβ} (or β), [v][KPAD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *tab = smem + TS + (size_t)warp * TS;       // this warp's statistics, [v][KPAD]
    for (int i = threadIdx.x; i < TS; i += blockDim.x) {
        const int v = i / KPAD, k = i % KPAD;
        E[i] = k < K ? Etab[k * V + v] : 0.0;
    }
    for (int i = lane; i < TS; i += 32) tab[i] = 0.0;
    __syncthreads();
    const long long nw = (long long)gridDim.x * nwarps_blk;
    for (long long d = (long long)blockIdx.x * nwarps_blk + warp; d < p.D; d += nw) {
        const double gk = (lane < K) ? p.gamma[d * K + lane] : 0.0;
        const double elt = lda_elntheta(gk, K, lane);
        const double mine = (lane < K) ? det_exp(elt) : 0.0;
        double et[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) et[k] = shfl_d(mine, k);
        double g[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) g[k] = 0.0;
        for (long long w = p.rowptr[d] + lane; w < p.rowptr[d + 1]; w += 32) {
            const int2 r = p.rec[w];
            const double2 *Ev = reinterpret_cast<const double2 *>(E + (r.x & 0xffff) * KPAD);
            double2 *Tv = reinterpret_cast<double2 *>(tab + (r.x & 0xffff) * KPAD);
            double pk[KP];
            double Z = 0.0;
#pragma unroll
            for (int k = 0; k < KP; k += 2) {
                const double2 e2 = Ev[k / 2];
                pk[k] = et[k] * e2.x;
                Z += pk[k];
                pk[k + 1] = et[k + 1] * e2.y;
                Z += pk[k + 1];
            }
            const double scale = (double)r.y / Z;
#pragma unroll
            for (int k = 0; k < KP; k += 2) {
                const double a0 = pk[k] * scale, a1 = pk[k + 1] * scale;
                if (want_stats) {
                    double2 t2 = Tv[k / 2];
                    t2.x += a0;
                    t2.y += a1;
                    Tv[k / 2] = t2;
                }
                g[k] += a0;
                g[k + 1] += a1;
            }
        }
        __syncwarp();
        warp_multi_reduce<NP>(g, lane);
        const int idx = warp_multi_index<NP>(lane);
        constexpr int GROUP = 32 / NP;
        if (idx < K && (lane & (GROUP - 1)) == 0) p.gamma_next[d * K + idx] = p.alpha + g[0];
    }
    __syncthreads();
    double2 *out = partial + (size_t)blockIdx.x * K * V;
    for (int i = threadIdx.x; i < K * V; i += blockDim.x) {
        const int k = i / V, v = i % V;
        double s = 0.0;
        for (int wv = 0; wv < nwarps_blk; ++wv) s += smem[TS + (size_t)wv * TS + v * KPAD + k];
        out[i] = make_double2(s, 0.0);
    }
}

// ------------------------------------------------------------------------------------------
// The same E pass as three skinny products over a tile of TS = 32·NW samples (NW = ⌈V/32⌉ warps):
//   Z[d][v]   = Σ_k e^{Elnθ}[d][k] · E[k][v]                   (lane <-> term, E column in registers)
//   R[d][v]   = n[d][v] / Z[d][v]                               (dense tile in shared memory)
//   S[k][v]   = E[k][v] · Σ_d e^{Elnθ}[d][k] · R[d][v]          (lane <-> term: no reduction until the end)
//   γ'[d][k]  = α + e^{Elnθ}[d][k] · Σ_v E[k][v] · R[d][v]      (lane <-> sample: no reduction at all)
// i.e. the (D x V)ᵀ(D x K) form of the statistics: every multiply-add is a DFMA on operands that
// sit in registers or are broadcast from shared memory, and neither output needs a cross-lane
// reduction per sample (the per-nonzero kernel above spends most of its instructions on those
// and on shared-memory read-modify-writes).  Results differ from it by reassociation only.
// ------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_lda_estep_tile(LdaDev p, double2 *partial, const double *Etab, int want_stats,
                                                        int NW) {
    extern __shared__ double smem[];
    constexpr int KPAD = KP + 2;
    const int K = p.K, V = p.V, TS = 32 * NW, VP = V | 1;
    double *Esm = smem;                         // [v][KPAD]
    double *rt = Esm + V * KPAD;                // [t][VP]   n, then R
    double *et = rt + TS * VP;                  // [t][KP]   e^{Elnθ}
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int v = warp * 32 + lane;
    const bool vok = v < V;
    for (int i = tid; i < V * KPAD; i += blockDim.x) {
        const int vv = i / KPAD, k = i % KPAD;
        Esm[i] = k < K ? Etab[k * V + vv] : 0.0;
    }
    double Ereg[KP], acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        Ereg[k] = (vok && k < K) ? Etab[k * V + v] : 0.0;
        acc[k] = 0.0;
    }
    __syncthreads();
    const long long ntiles = (p.D + TS - 1) / TS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * TS;
        // ---- per-sample preparation, thread <-> sample: dense count row of the tile (every lane
        // walks its own CSR row, so 32 rows are in flight per load instruction), Elnθ (:78-80), e^{Elnθ}
        {
            const long long d = d0 + tid;
            double *row = rt + tid * VP;
            for (int i = 0; i < V; ++i) row[i] = 0.0;
            double *er = et + tid * KP;
            if (d < p.D) {
                const long long beg = p.rowptr[d], end = p.rowptr[d + 1];
#pragma unroll 8
                for (long long w = beg; w < end; ++w) {
                    const int2 r = p.rec[w];
                    row[r.x & 0xffff] = (double)r.y;
                }
                double gk[KP];
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    gk[k] = k < K ? p.gamma[d * K + k] : 1.0;
                    if (k < K) s += gk[k];
                }
                const double ds = det_digamma(s);
                // independent digamma / exp chains, interleaved four at a time by the unroll
#pragma unroll 4
                for (int k = 0; k < KP; ++k) {
                    const double e = det_exp(det_digamma(gk[k]) - ds);
                    er[k] = k < K ? e : 0.0;
                }
            } else
                for (int k = 0; k < KP; ++k) er[k] = 0.0;
        }
        __syncthreads();
        // ---- Z, R and the statistics, lane <-> term
        if (vok) {
            for (int t = 0; t < TS; ++t) {
                const double2 *e2 = reinterpret_cast<const double2 *>(et + t * KP);
                double ek[KP];
                double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;       // four partial chains for ILP
#pragma unroll
                for (int k = 0; k < KP; k += 4) {
                    const double2 x2 = e2[k / 2], y2 = e2[k / 2 + 1];
                    ek[k] = x2.x;
                    ek[k + 1] = x2.y;
                    ek[k + 2] = y2.x;
                    ek[k + 3] = y2.y;
                    z0 = fma(x2.x, Ereg[k], z0);
                    z1 = fma(x2.y, Ereg[k + 1], z1);
                    z2 = fma(y2.x, Ereg[k + 2], z2);
                    z3 = fma(y2.y, Ereg[k + 3], z3);
                }
                const double Z = (z0 + z1) + (z2 + z3);
                const double n = rt[t * VP + v];
                const double r = (Z > 0.0) ? n * (1.0 / Z) : 0.0;      // padded samples have Z = 0
                rt[t * VP + v] = r;
                if (want_stats)
#pragma unroll
                    for (int k = 0; k < KP; ++k) acc[k] = fma(ek[k], r, acc[k]);
            }
        }
        __syncthreads();
        // ---- γ_{t+1}, thread <-> sample
        {
            const long long d = d0 + tid;
            if (d < p.D) {
                double g[KP];
#pragma unroll
                for (int k = 0; k < KP; ++k) g[k] = 0.0;
                const double *row = rt + tid * VP;
                for (int i = 0; i < V; ++i) {
                    const double r = row[i];
                    const double2 *E2 = reinterpret_cast<const double2 *>(Esm + i * KPAD);
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 x2 = E2[k / 2];
                        g[k] = fma(x2.x, r, g[k]);
                        g[k + 1] = fma(x2.y, r, g[k + 1]);
                    }
                }
                const double *er = et + tid * KP;
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    if (k < K) p.gamma_next[d * K + k] = p.alpha + er[k] * g[k];
            }
        }
        __syncthreads();
    }
    double2 *out = partial + (size_t)blockIdx.x * K * V;
    if (vok) {
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) out[k * V + v] = make_double2(Ereg[k] * acc[k], 0.0);    // S = E ∘ (e^{Elnθ})ᵀ R
    }
}

// M-step (single block): λ = η + Σ n ϕ (:100-105), Elnβ (:96-98), β (:110-112), e^{Elnβ}.
__global__ void __launch_bounds__(1024) k_lda_mstep(LdaDev p, const double2 *gathered, int nranks) {
    __shared__ double rowsum[32], rowdig[32];
    const int K = p.K, V = p.V, KV = K * V;
    for (int i = threadIdx.x; i < KV; i += blockDim.x) {
        double hi = 0.0, lo = 0.0;
        for (int r = 0; r < nranks; ++r) dd_merge(hi, lo, gathered[(size_t)r * KV + i].x, gathered[(size_t)r * KV + i].y);
        dd_add(hi, lo, p.eta);
        p.lam[i] = dd_round(hi, lo);
        p.Elnbeta_prev[i] = p.Elnbeta[i];
        p.expElnbeta_prev[i] = p.expElnbeta[i];
    }
    __syncthreads();
    if ((int)threadIdx.x < K) {
        double s = 0.0;
        for (int v = 0; v < V; ++v) s += p.lam[threadIdx.x * V + v];
        rowsum[threadIdx.x] = s;
        rowdig[threadIdx.x] = det_digamma(s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KV; i += blockDim.x) {
        const int k = i / V;
        const double l = p.lam[i];
        const double e = det_digamma(l) - rowdig[k];
        p.Elnbeta[i] = e;
        p.expElnbeta[i] = det_exp(e);
        p.beta[i] = l / rowsum[k];
    }
}

// constructor: Elnβ, e^{Elnβ} from λ (src/LDA.jl:36-39)
__global__ void __launch_bounds__(1024) k_lda_elnbeta(LdaDev p) {
    __shared__ double rowdig[32];
    const int K = p.K, V = p.V, KV = K * V;
    if ((int)threadIdx.x < K) {
        double s = 0.0;
        for (int v = 0; v < V; ++v) s += p.lam[threadIdx.x * V + v];
        rowdig[threadIdx.x] = det_digamma(s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KV; i += blockDim.x) {
        const double e = det_digamma(p.lam[i]) - rowdig[i / V];
        p.Elnbeta[i] = e;
        p.Elnbeta_prev[i] = e;
        p.expElnbeta[i] = det_exp(e);
        p.expElnbeta_prev[i] = p.expElnbeta[i];
        p.beta[i] = 0.0;
    }
}

// ---- ILDA (src/ILDA.jl:97-129): Elnβ_i = ψ(λ_i) - ψ(Σ_j λ_i), then the composite tables the per-sample
// kernels read: Elnβ_kv = Σ_i Elnβ_i[f(v,i), k] (index order, from 0), e^{Elnβ_kv}, β_kv = Π_i λ_i / Σ_j λ_i.
// rowsum / rowdig: R doubles each in shared memory.
__device__ inline void ilda_compose(const LdaDev &p, double *rowsum, double *rowdig, bool keep_prev) {
    for (int r = threadIdx.x; r < p.R; r += blockDim.x) {
        const double *l = p.lambdaf + p.row_off[r];
        double s = 0.0;
        for (int j = 0; j < p.row_len[r]; ++j) s += l[j];
        rowsum[r] = s;
        rowdig[r] = det_digamma(s);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) p.Elnbetaf[t] = det_digamma(p.lambdaf[t]) - rowdig[p.ent_row[t]];
    __syncthreads();
    const int KV = p.K * p.V, nf = p.nfeat;
    for (int i = threadIdx.x; i < KV; i += blockDim.x) {
        const int k = i / p.V, v = i % p.V;
        int r = k * nf;
        double e = 0.0, b = 1.0;
        for (int f = 0; f < nf; ++f, ++r) {
            const int t = p.row_off[r] + p.feat[v * nf + f];
            e += p.Elnbetaf[t];
            b *= p.lambdaf[t] / rowsum[r];
        }
        if (keep_prev) { p.Elnbeta_prev[i] = p.Elnbeta[i]; p.expElnbeta_prev[i] = p.expElnbeta[i]; }
        p.Elnbeta[i] = e;
        p.expElnbeta[i] = det_exp(e);
        p.beta[i] = b;
        if (!keep_prev) { p.Elnbeta_prev[i] = e; p.expElnbeta_prev[i] = p.expElnbeta[i]; }
    }
}
// constructor / set_state (src/ILDA.jl:38-42)
__global__ void __launch_bounds__(1024) k_ilda_compose(LdaDev p) {
    extern __shared__ double ism[];
    ilda_compose(p, ism, ism + p.R, false);
}
// M-step (single block): statistics S_kv = Σ_d n ϕ (gathered), λ_i[j, k] = η_i + Σ_{v: f(v,i) = j} S_kv
// (ascending v, src/ILDA.jl:105-125), then the tables above
__global__ void __launch_bounds__(1024) k_ilda_mstep(LdaDev p, const double2 *gathered, int nranks) {
    extern __shared__ double ism[];
    const int K = p.K, V = p.V, KV = K * V, nf = p.nfeat;
    for (int i = threadIdx.x; i < KV; i += blockDim.x) {
        double hi = 0.0, lo = 0.0;
        for (int r = 0; r < nranks; ++r) dd_merge(hi, lo, gathered[(size_t)r * KV + i].x, gathered[(size_t)r * KV + i].y);
        p.lam[i] = dd_round(hi, lo);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
        const int r = p.ent_row[t], j = t - p.row_off[r], k = r / nf, f = r % nf;
        const double *st = p.lam + k * V;
        double acc = p.etaf[p.row_eta[r]];
        for (int v = 0; v < V; ++v)
            if (p.feat[v * nf + f] == j) acc += st[v];
        p.lambdaf[t] = acc;
    }
    __syncthreads();
    ilda_compose(p, ism, ism + p.R, true);
}

// log-likelihood pass (src/LDA.jl:174-188) with θ_t = γ_t / Σγ_t and the new β.
// partial: [gridDim.x] double2.
__global__ void __launch_bounds__(256) k_lda_ll(LdaDev p, double2 *partial) {
    extern __shared__ double smem[];
    __shared__ double2 red[8];
    const int K = p.K, V = p.V, KV = K * V;
    double *beta = smem;                 // KV
    double *thsh = smem + KV;            // 8 x 32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *th = thsh + warp * 32;
    for (int i = threadIdx.x; i < KV; i += blockDim.x) beta[i] = p.beta[i];
    __syncthreads();
    double hi = 0.0, lo = 0.0;
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const double gk = (lane < K) ? p.gamma[d * K + lane] : 0.0;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += shfl_d(gk, k);
        __syncwarp();
        th[lane] = gk / s;
        __syncwarp();
        for (long long w = p.rowptr[d] + lane; w < p.rowptr[d + 1]; w += 32) {
            const int2 r = p.rec[w];
            double dot = 0.0;
            for (int k = 0; k < K; ++k) dot += th[k] * beta[k * V + (r.x & 0xffff)];
            dd_add(hi, lo, (double)r.y * det_log(dot));
        }
    }
    __syncwarp();
    warp_dd_allreduce(hi, lo);
    if (lane == 0) red[warp] = make_double2(hi, lo);
    __syncthreads();
    if (threadIdx.x == 0) {
        double h = 0.0, l = 0.0;
        for (int wv = 0; wv < 8; ++wv) dd_merge(h, l, red[wv].x, red[wv].y);
        partial[blockIdx.x] = make_double2(h, l);
    }
}

__global__ void k_lda_ll_final(const double2 *gathered, int nranks, double Ntot, double *ll_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double hi = 0.0, lo = 0.0;
        for (int r = 0; r < nranks; ++r) dd_merge(hi, lo, gathered[r].x, gathered[r].y);
        *ll_out = dd_round(hi, lo) / Ntot;
    }
}

// θ, Elnθ of the current γ (get_state)
__global__ void __launch_bounds__(256) k_lda_theta_out(LdaDev p, double *Elntheta, double *theta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, K = p.K;
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const double gk = (lane < K) ? p.gamma[d * K + lane] : 0.0;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += shfl_d(gk, k);
        const double elt = lda_elntheta(gk, K, lane);
        if (lane < K) {
            if (Elntheta) Elntheta[d * K + lane] = elt;
            if (theta) theta[d * K + lane] = gk / s;
        }
    }
}

// ELBO data terms (src/LDA.jl:120-160) and optional ϕ output; ϕ_T is recomputed from γ_T and the
// Elnβ the last E pass used.  partial: [gridDim.x][8] dd:
//  [0] Σ Elnθ   [1] ElnPZ = Σ ϕ Elnθ n   [2] ElnPX = Σ ϕ Elnβ_T n   [3] ElnQZ = Σ ϕ log ϕ (no n, as :154-160)
//  [4] Σ lgamma(γ)   [5] Σ_d lgamma(Σ_k γ)   [6] Σ (γ-1) Elnθ
__global__ void __launch_bounds__(256) k_lda_elbo(LdaDev p, double2 *partial, double *phi_out, const double *Etab) {
    extern __shared__ double smem[];
    __shared__ double2 red[8 * 7];
    const int K = p.K, V = p.V, KV = K * V;
    double *Eprev = smem;              // e^{Elnβ_{T-1}}
    double *Eln = smem + KV;           // Elnβ_T
    double *sh = smem + 2 * KV;        // 8 x 64 : e^{Elnθ}, Elnθ per warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *et = sh + warp * 64, *el = et + 32;
    for (int i = threadIdx.x; i < KV; i += blockDim.x) { Eprev[i] = Etab[i]; Eln[i] = p.Elnbeta[i]; }
    __syncthreads();
    double hi[7] = {0, 0, 0, 0, 0, 0, 0}, lo[7] = {0, 0, 0, 0, 0, 0, 0};
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const double gk = (lane < K) ? p.gamma[d * K + lane] : 0.0;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += shfl_d(gk, k);
        const double elt = lda_elntheta(gk, K, lane);
        __syncwarp();
        et[lane] = (lane < K) ? det_exp(elt) : 0.0;
        el[lane] = elt;
        __syncwarp();
        if (lane < K) {
            dd_add(hi[0], lo[0], elt);
            dd_add(hi[4], lo[4], lgamma(gk));
            dd_add(hi[6], lo[6], (gk - 1) * elt);
        }
        if (lane == 0) dd_add(hi[5], lo[5], lgamma(s));
        for (long long w = p.rowptr[d] + lane; w < p.rowptr[d + 1]; w += 32) {
            const int2 r = p.rec[w];
            const double n = (double)r.y;
            double Z = 0.0;
            for (int k = 0; k < K; ++k) Z += et[k] * Eprev[k * V + (r.x & 0xffff)];
            double a = 0.0, b = 0.0, c = 0.0;
            for (int k = 0; k < K; ++k) {
                const double ph = et[k] * Eprev[k * V + (r.x & 0xffff)] / Z;
                if (phi_out) phi_out[w * K + k] = ph;
                a += ph * el[k] * n;
                b += ph * Eln[k * V + (r.x & 0xffff)] * n;
                if (ph > 0.0) c += ph * det_log(ph);
            }
            dd_add(hi[1], lo[1], a);
            dd_add(hi[2], lo[2], b);
            dd_add(hi[3], lo[3], c);
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 7; ++i) warp_dd_allreduce(hi[i], lo[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 7; ++i) red[warp * 7 + i] = make_double2(hi[i], lo[i]);
    __syncthreads();
    if (threadIdx.x < 7) {
        double h = 0.0, l = 0.0;
        for (int wv = 0; wv < 8; ++wv) dd_merge(h, l, red[wv * 7 + threadIdx.x].x, red[wv * 7 + threadIdx.x].y);
        partial[(size_t)blockIdx.x * 8 + threadIdx.x] = make_double2(h, l);
    }
}

// table terms of the LDA ELBO (single block): out[0] = ΣElnβ, out[1] = Σ lgamma(λ),
// out[2] = Σ_k lgamma(Σ_v λ), out[3] = Σ (λ-1) Elnβ
__global__ void __launch_bounds__(256) k_lda_elbo_tables(LdaDev p, double *out) {
    __shared__ double2 red[8 * 4];
    __shared__ double2 res[4];
    const int K = p.K, V = p.V, KV = K * V;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < KV; i += blockDim.x) {
        dd_add(hi[0], lo[0], p.Elnbeta[i]);
        dd_add(hi[1], lo[1], lgamma(p.lam[i]));
        dd_add(hi[3], lo[3], (p.lam[i] - 1) * p.Elnbeta[i]);
    }
    if ((int)threadIdx.x < K) {
        double s = 0.0;
        for (int v = 0; v < V; ++v) s += p.lam[threadIdx.x * V + v];
        dd_add(hi[2], lo[2], lgamma(s));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) warp_dd_allreduce(hi[i], lo[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 4; ++i) red[warp * 4 + i] = make_double2(hi[i], lo[i]);
    __syncthreads();
    if (threadIdx.x < 4) {
        double h = 0.0, l = 0.0;
        for (int wv = 0; wv < 8; ++wv) dd_merge(h, l, red[wv * 4 + threadIdx.x].x, red[wv * 4 + threadIdx.x].y);
        res[threadIdx.x] = make_double2(h, l);
        out[threadIdx.x] = dd_round(h, l);
    }
}

}  // namespace mmsig
