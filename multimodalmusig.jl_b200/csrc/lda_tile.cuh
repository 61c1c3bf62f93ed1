// lda_tile.cuh -- LDA's E pass and log-likelihood pass over tiles of 32 samples, one thread per term
// (the layout of theta_tile.cuh).  k_lda_estep_t32 is k_lda_estep_tile (lda_kernels.cuh) with
// 32-sample tiles instead of 32·⌈V/32⌉: 45 KB instead of 107 KB of shared memory per block at
// K = 20, V = 96, i.e. 12 resident warps per SM instead of 6, with the per-sample preparation
// (digamma, exp) spread over all threads as (sample, k) pairs and γ' computed in 1x4 register
// micro-tiles.  Records carry their sample's slot in the tile (k_pack_rows, tag).
#pragma once
#include "lda_kernels.cuh"
#include "theta_tile.cuh"

namespace mmsig {

constexpr int LDA_TS = 32;

// DENSE: dense count tiles staged by bulk copies (tile_stage.cuh), as in k_theta_tile
template <int KP, int NWT, bool DENSE>
__global__ void __launch_bounds__(32 * NWT) k_lda_estep_t32(LdaDev p, double2 *partial, const double *Etab, int want_stats) {
    extern __shared__ __align__(16) double smem[];
    const int K = p.K, V = p.V, VP = V | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    double *Evk = smem;                       // [v][KP]
    double *rt = Evk + V * KP;                // [t][VP]  n, then R = n / Z
    double *et = rt + LDA_TS * VP;            // [t][KP]  γ, then e^{Elnθ}
    double *ssum = et + LDA_TS * KP;          // [t]      ψ(Σ_k γ)
    long long *rp = reinterpret_cast<long long *>(ssum + LDA_TS);
    double *stg = smem + tile_stage_offset((size_t)(ssum - smem) + LDA_TS);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 2);                        // [t][V]
    const int v = tid;
    const bool vok = v < V;
    for (int i = tid; i < V * KP; i += blockDim.x) {
        const int vv = i / KP, k = i % KP;
        Evk[i] = k < K ? Etab[k * V + vv] : 0.0;
    }
    __syncthreads();
    double Ereg[KP], acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        Ereg[k] = vok ? Evk[v * KP + k] : 0.0;
        acc[k] = 0.0;
    }
    const long long ntiles = (p.D + LDA_TS - 1) / LDA_TS;
    unsigned parity = 0;
    if (DENSE) {
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (tid == 0 && blockIdx.x < ntiles) stage_tile(p.cnt, blockIdx.x, V, nt, mbar);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * LDA_TS;
        // ---- phase 1: γ rows, Elnθ (src/LDA.jl:78-80) and e^{Elnθ}; CSR: clear the tile, scatter the counts
        if (!DENSE)
            for (int i = tid; i < LDA_TS * VP; i += blockDim.x) rt[i] = 0.0;
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            et[i] = (k < K && d < p.D) ? p.gamma[d * K + k] : 0.0;
        }
        if (!DENSE && tid == 0) {
            rp[0] = p.rowptr[d0];
            rp[1] = p.rowptr[min(d0 + LDA_TS, p.D)];
            const long long dn = min(d0 + (long long)gridDim.x * LDA_TS, p.D);      // this block's next tile
            rp[2] = p.rowptr[dn];
            rp[3] = p.rowptr[min(dn + LDA_TS, p.D)];
        }
        __syncthreads();
        if (!DENSE) {
#pragma unroll 4
            for (long long w = rp[0] + tid; w < rp[1]; w += blockDim.x) {
                const int2 r = p.rec[w];
                rt[(r.x >> 16) * VP + (r.x & 0xffff)] = (double)r.y;
            }
        }
        if (tid < LDA_TS) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += et[tid * KP + k];
            ssum[tid] = (d0 + tid < p.D) ? det_digamma(s) : 0.0;
        }
        __syncthreads();
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            et[i] = (k < K && d0 + t < p.D) ? det_exp(det_digamma(et[i]) - ssum[t]) : 0.0;
        }
        __syncthreads();
        if (DENSE) {
            mbar_wait(mbar, parity);                 // this tile's counts have landed in nt
            parity ^= 1u;
        } else {
            prefetch_records(p.rec, rp[2], rp[3], tid, blockDim.x);    // next tile's records towards L2
        }
        // ---- phase 2: Z, R and the statistics, lane <-> term
        if (DENSE && vok) {
            // Dense counts: two samples per trip, no branch on n: eight independent Z chains, two reciprocals and two runs
            // of K independent fmas interleave (the kernel sat at 29 % of the FP64 pipe with one chain of work per thread)
#pragma unroll 1
            for (int t = 0; t < LDA_TS; t += 2) {
                const int n0 = nt[t * V + v], n1 = nt[(t + 1) * V + v];
                const double2 *a2 = reinterpret_cast<const double2 *>(et + t * KP), *b2 = a2 + KP / 2;
                double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0, y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
#pragma unroll
                for (int k = 0; k < KP; k += 4) {
                    const double2 xa = a2[k / 2], xb = a2[k / 2 + 1], ya = b2[k / 2], yb = b2[k / 2 + 1];
                    z0 = fma(xa.x, Ereg[k], z0);
                    y0 = fma(ya.x, Ereg[k], y0);
                    z1 = fma(xa.y, Ereg[k + 1], z1);
                    y1 = fma(ya.y, Ereg[k + 1], y1);
                    z2 = fma(xb.x, Ereg[k + 2], z2);
                    y2 = fma(yb.x, Ereg[k + 2], y2);
                    z3 = fma(xb.y, Ereg[k + 3], z3);
                    y3 = fma(yb.y, Ereg[k + 3], y3);
                }
                const double Z0 = (z0 + z1) + (z2 + z3), Z1 = (y0 + y1) + (y2 + y3);
                const double r0 = n0 > 0 ? (double)n0 * (1.0 / Z0) : 0.0, r1 = n1 > 0 ? (double)n1 * (1.0 / Z1) : 0.0;
                rt[t * VP + v] = r0;
                rt[(t + 1) * VP + v] = r1;
                if (want_stats) {
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 x = a2[k / 2], y = b2[k / 2];
                        acc[k] = fma(x.x, r0, acc[k]);
                        acc[k + 1] = fma(x.y, r0, acc[k + 1]);
                        acc[k] = fma(y.x, r1, acc[k]);
                        acc[k + 1] = fma(y.y, r1, acc[k + 1]);
                    }
                }
            }
        } else if (vok) {
            for (int t = 0; t < LDA_TS; ++t) {
                const double n = DENSE ? (double)nt[t * V + v] : rt[t * VP + v];
                if (DENSE && !(n > 0.0)) rt[t * VP + v] = 0.0;       // the tile is not cleared: every cell is written
                if (n > 0.0) {
                    const double2 *e2 = reinterpret_cast<const double2 *>(et + t * KP);
                    double ek[KP];
                    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
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
                    const double r = n * (1.0 / Z);
                    rt[t * VP + v] = r;
                    if (want_stats) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) acc[k] = fma(ek[k], r, acc[k]);
                    }
                }
            }
        }
        __syncthreads();
        if (DENSE && tid == 0 && tile + gridDim.x < ntiles) stage_tile(p.cnt, tile + gridDim.x, V, nt, mbar);   // flies during phase 3
        // ---- phase 3: γ_{t+1} = α + e^{Elnθ} ∘ (R Eᵀ), thread <-> (sample, four consecutive k)
        {
            const int t = lane;
            const long long d = d0 + t;
            const double *row = rt + t * VP;
            for (int kb = 4 * warp; kb < KP; kb += 4 * NW) {
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
#pragma unroll 4
                for (int vv = 0; vv < V; ++vv) {
                    const double R = row[vv];
                    const double2 *E2 = reinterpret_cast<const double2 *>(Evk + vv * KP + kb);
                    const double2 a = E2[0], b = E2[1];
                    g0 = fma(a.x, R, g0);
                    g1 = fma(a.y, R, g1);
                    g2 = fma(b.x, R, g2);
                    g3 = fma(b.y, R, g3);
                }
                if (d < p.D) {
                    const double *L = et + t * KP + kb;
                    double *gn = p.gamma_next + d * K + kb;
                    if (kb + 0 < K) gn[0] = p.alpha + L[0] * g0;
                    if (kb + 1 < K) gn[1] = p.alpha + L[1] * g1;
                    if (kb + 2 < K) gn[2] = p.alpha + L[2] * g2;
                    if (kb + 3 < K) gn[3] = p.alpha + L[3] * g3;
                }
            }
        }
        __syncthreads();
    }
    if (vok && want_stats) {
        double2 *out = partial + (size_t)blockIdx.x * K * V;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) out[k * V + v] = make_double2(Ereg[k] * acc[k], 0.0);    // S = E ∘ (e^{Elnθ})ᵀ R
    }
}

// log-likelihood pass (src/LDA.jl:174-188) with θ_t = γ_t / Σγ_t and the new β, over the same tiles.
// partial: [gridDim.x] double2.
template <int KP, int NWT, bool DENSE>
__global__ void __launch_bounds__(32 * NWT) k_lda_ll_tile(LdaDev p, double2 *partial) {
    extern __shared__ __align__(16) double smem[];
    const int K = p.K, V = p.V, VP = V | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    double *xt = smem;                        // [t][VP]  n, then n log(θ·β)
    double *pt = xt + LDA_TS * VP;            // [t][KP]  γ, then θ
    double *ssum = pt + LDA_TS * KP;          // [t]
    double *bsum = ssum + LDA_TS;             // [NW][32]
    long long *rp = reinterpret_cast<long long *>(bsum + NW * 32);
    double *stg = smem + tile_stage_offset((size_t)(bsum - smem) + NW * 32);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 2);                        // [t][V]
    const int v = tid;
    const bool vok = v < V;
    double Breg[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) Breg[k] = (vok && k < K) ? p.beta[k * V + v] : 0.0;
    double ahi = 0.0, alo = 0.0;
    const long long ntiles = (p.D + LDA_TS - 1) / LDA_TS;
    unsigned parity = 0;
    if (DENSE) {
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (tid == 0 && blockIdx.x < ntiles) stage_tile(p.cnt, blockIdx.x, V, nt, mbar);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * LDA_TS;
        if (!DENSE)
            for (int i = tid; i < LDA_TS * VP; i += blockDim.x) xt[i] = 0.0;
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            pt[i] = (k < K && d < p.D) ? p.gamma[d * K + k] : 0.0;
        }
        if (!DENSE && tid == 0) {
            rp[0] = p.rowptr[d0];
            rp[1] = p.rowptr[min(d0 + LDA_TS, p.D)];
            const long long dn = min(d0 + (long long)gridDim.x * LDA_TS, p.D);      // this block's next tile
            rp[2] = p.rowptr[dn];
            rp[3] = p.rowptr[min(dn + LDA_TS, p.D)];
        }
        __syncthreads();
        if (!DENSE) {
#pragma unroll 4
            for (long long w = rp[0] + tid; w < rp[1]; w += blockDim.x) {
                const int2 r = p.rec[w];
                xt[(r.x >> 16) * VP + (r.x & 0xffff)] = (double)r.y;
            }
        }
        if (tid < LDA_TS) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += pt[tid * KP + k];
            ssum[tid] = s;
        }
        __syncthreads();
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            if (k < K && d0 + t < p.D) pt[i] = pt[i] / ssum[t];
        }
        __syncthreads();
        if (DENSE) {
            mbar_wait(mbar, parity);                 // this tile's counts have landed in nt
            parity ^= 1u;
        } else {
            prefetch_records(p.rec, rp[2], rp[3], tid, blockDim.x);
        }
        if (DENSE && vok) {
            // Dense counts: four samples per trip, no branch on n (four independent dot-product and logarithm chains)
#pragma unroll 1
            for (int t = 0; t < LDA_TS; t += 4) {
                double dot[4] = {0.0, 0.0, 0.0, 0.0};
                const double2 *p2 = reinterpret_cast<const double2 *>(pt + t * KP);
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 x = p2[j * (KP / 2) + k / 2];
                        dot[j] += x.x * Breg[k];
                        dot[j] += x.y * Breg[k + 1];
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ni = nt[(t + j) * V + v];
                    const double lg = det_log(dot[j] > 0.0 ? dot[j] : 1.0);
                    xt[(t + j) * VP + v] = ni > 0 ? (double)ni * lg : 0.0;
                }
            }
        } else if (vok) {
            for (int t = 0; t < LDA_TS; ++t) {
                const double n = DENSE ? (double)nt[t * V + v] : xt[t * VP + v];
                if (DENSE && !(n > 0.0)) xt[t * VP + v] = 0.0;       // the tile is not cleared: every cell is written
                if (n > 0.0) {
                    const double2 *p2 = reinterpret_cast<const double2 *>(pt + t * KP);
                    double dot = 0.0;
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 x = p2[k / 2];
                        dot += x.x * Breg[k];
                        dot += x.y * Breg[k + 1];
                    }
                    xt[t * VP + v] = n * det_log(dot);
                }
            }
        }
        __syncthreads();
        if (DENSE && tid == 0 && tile + gridDim.x < ntiles) stage_tile(p.cnt, tile + gridDim.x, V, nt, mbar);   // flies during the row sums
        {
            const double *row = xt + lane * VP;
            const int vb = 32 * warp, ve = min(V, vb + 32);
            double b = 0.0;
            for (int vv = vb; vv < ve; ++vv) b += row[vv];
            bsum[warp * 32 + lane] = b;
        }
        __syncthreads();
        if (tid < LDA_TS && d0 + tid < p.D) {
            double rs = bsum[tid];
            for (int j = 1; j < NW; ++j) rs += bsum[j * 32 + tid];
            dd_add(ahi, alo, rs);
        }
        __syncthreads();
    }
    if (warp == 0) {
        warp_dd_allreduce(ahi, alo);
        if (lane == 0) partial[blockIdx.x] = make_double2(ahi, alo);
    }
}

}  // namespace mmsig
