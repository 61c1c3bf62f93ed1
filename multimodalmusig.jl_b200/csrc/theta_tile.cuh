// theta_tile.cuh -- the θ pass of one modality (reference src/MMCTM.jl:183-198 update_θ!, :110-117
// calculate_sumθ, :224-240 the Σ n θ part of update_γ!) as three skinny products over tiles of 32
// samples, the (D x K)ᵀ(D x V) form of the statistics:
//   L[d][k]  = exp(λ_dk)                              E[k][v] = exp(Elnϕ_kv)   (ϕ_kv when unsmoothed, :496-509)
//   Z[d][v]  = Σ_k L[d][k] E[k][v]                    (index order, two roundings per term)
//   R[d][v]  = n[d][v] · (1 / Z[d][v])                (dense tile in shared memory, 0 where n = 0)
//   S[k][v]  = Σ_d L[d][k] · R[d][v]                  (exactly rounded: double-double per thread, lane <-> term)
//   sumθ[d][k] = L[d][k] · Σ_v E[k][v] R[d][v]        (one fma chain in term order, thread <-> sample)
// and Σ_d n θ_kv = E[k][v] · S[k][v] (applied in k_mstep1).  θ itself, (L E)(1/Z), is never formed.
// Against the per-nonzero kernel this replaces (warp per sample, lane <-> nonzero, a private
// double-double K x V table per warp in shared memory): the accumulators live in registers (no
// shared-memory read-modify-write per addend, no 15 KB of table per warp capping occupancy at
// 12-20 warps / SM), and neither output needs a cross-lane reduction.  The results differ from it
// by roundings only; DESIGN.md section 2 (pinned arithmetic) states this form.
#pragma once
#include "mmctm_kernels.cuh"
#include "tile_stage.cuh"

namespace mmsig {

constexpr int TILE_S = 32;        // samples per tile

// the records of this block's NEXT tile towards L2 while the current tile computes: the scatter of
// a tile waits on its record loads (24-29 % of the stall samples of both tile kernels in
// profiles/r01j), and an L2 hit costs a third of a DRAM access
__device__ __forceinline__ void prefetch_records(const int2 *rec, long long beg, long long end, int tid, int nthreads) {
    const char *p0 = reinterpret_cast<const char *>(rec + beg), *p1 = reinterpret_cast<const char *>(rec + end);
    for (const char *q = p0 + (size_t)tid * 128; q < p1; q += (size_t)nthreads * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
}

// doubles of shared memory in front of the staging area of the DENSE variants (mbarrier, then the 32 x V int32 counts)
__host__ __device__ inline size_t tile_stage_offset(size_t doubles_before) { return (doubles_before + 1) & ~(size_t)1; }
__host__ __device__ inline size_t tile_stage_doubles(int V) { return 2 + (size_t)TILE_S * V / 2; }

// EREG: the thread's column of E in registers (K <= 16); else read from shared memory [k][v].
// NWT: upper bound of the block's warp count (one thread per term: blockDim = 32 ceil(V / 32)).
// DENSE: the modality keeps dense count tiles (tile_stage.cuh): the next tile's counts arrive by one bulk copy while
// this tile computes; no clear / scatter of records.
template <int KP, bool EREG, int NWT, bool DENSE>
__global__ void __launch_bounds__(32 * NWT) k_theta_tile(MmctmDev p, int m, double2 *partial, int unsmoothed, int want_stats) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ __align__(16) double smem[];
    const int K = p.K[m], V = p.V[m], off = p.koff[m], VP = V | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    double *Evk = smem;                        // [v][KP]  phase 3 (a sample's thread walks v, reads a row of k)
    double *Ekv = Evk + V * KP;                // [k][VP]  phase 2 when !EREG
    double *rt = Ekv + (EREG ? 0 : KP * VP);   // [t][VP]  n, then R
    double *et = rt + TILE_S * VP;             // [t][KP]  L
    long long *rp = reinterpret_cast<long long *>(et + TILE_S * KP);   // [TILE_S + 1] row pointers of the tile (CSR variant)
    double *stg = smem + tile_stage_offset((size_t)(et - smem) + TILE_S * KP);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);                // DENSE: the staging barrier, then the counts
    int *nt = reinterpret_cast<int *>(stg + 2);                        // [t][V]
    const int *cnt = p.cnt[m];
    const int v = tid;
    const bool vok = v < V;
    const double *Eg = (unsmoothed ? p.phi : p.Elnphi) + p.goff[m];
    for (int i = tid; i < V * KP; i += blockDim.x) {
        const int vv = i / KP, k = i % KP;
        const double e = k < K ? (unsmoothed ? Eg[k * V + vv] : det_exp(Eg[k * V + vv])) : 0.0;
        Evk[i] = e;
        if (!EREG) Ekv[k * VP + vv] = e;
    }
    __syncthreads();
    double Ereg[EREG ? KP : 1];
    if (EREG) {
#pragma unroll
        for (int k = 0; k < KP; ++k) Ereg[k] = vok ? Evk[v * KP + k] : 0.0;
    }
    double ahi[KP], alo[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) { ahi[k] = 0.0; alo[k] = 0.0; }

    const long long *rowptr = p.rowptr[m];
    const int2 *rec = p.rec[m];
    const long long ntiles = (p.D + TILE_S - 1) / TILE_S;
    unsigned parity = 0;
    if (DENSE) {
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (tid == 0 && blockIdx.x < ntiles) stage_tile(cnt, blockIdx.x, V, nt, mbar);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * TILE_S;
        // ---- phase 1: L = exp(λ) (thread <-> (sample, k)); CSR: clear the count tile, the tile's row pointers
        if (!DENSE)
            for (int i = tid; i < TILE_S * VP; i += blockDim.x) rt[i] = 0.0;
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            et[i] = (k < K && d < p.D) ? det_exp(p.lam_prev[d * p.MK + off + k]) : 0.0;
        }
        if (!DENSE && tid == 0) {
            rp[0] = rowptr[d0];
            rp[TILE_S] = rowptr[min(d0 + TILE_S, p.D)];
            const long long dn = min(d0 + (long long)gridDim.x * TILE_S, p.D);      // this block's next tile
            rp[1] = rowptr[dn];
            rp[2] = rowptr[min(dn + TILE_S, p.D)];
        }
        __syncthreads();
        if (DENSE) {
            mbar_wait(mbar, parity);                 // this tile's counts have landed in nt
            parity ^= 1u;
        } else {
            // scatter: the tile's records are one contiguous range of rec, streamed by all threads (every
            // thread has its loads in flight at once); a record carries its sample's slot in the tile
#pragma unroll 4
            for (long long w = rp[0] + tid; w < rp[TILE_S]; w += blockDim.x) {
                const int2 r = rec[w];
                rt[(r.x >> 16) * VP + (r.x & 0xffff)] = (double)r.y;        // slot tag | term (k_pack_rows)
            }
            __syncthreads();
            prefetch_records(rec, rp[1], rp[2], tid, blockDim.x);
        }
        // ---- phase 2: Z, R and the statistics, lane <-> term
        if (DENSE && vok) {
            // Dense counts: two samples per trip and no branch on n (88 % of the cells are nonzero): the two Z chains
            // (K dependent additions each), the two reciprocals and the two runs of double-double updates are
            // independent and interleave.  Per cell the same operations in the same order as below; a zero cell adds
            // an exact + 0.0 to the accumulators (R = 0), so the sums keep their bits.
#pragma unroll 1
            for (int t = 0; t < TILE_S; t += 2) {
                const int n0 = nt[t * V + v], n1 = nt[(t + 1) * V + v];
                const double2 *a2 = reinterpret_cast<const double2 *>(et + t * KP), *b2 = a2 + KP / 2;
                double Z0 = 0.0, Z1 = 0.0;
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
                    const double2 x = a2[k / 2], y = b2[k / 2];
                    const double E0 = EREG ? Ereg[k] : Ekv[k * VP + v], E1 = EREG ? Ereg[k + 1] : Ekv[(k + 1) * VP + v];
                    Z0 += x.x * E0;
                    Z1 += y.x * E0;
                    Z0 += x.y * E1;
                    Z1 += y.y * E1;
                }
                const double R0 = n0 > 0 ? (double)n0 * (1.0 / Z0) : 0.0, R1 = n1 > 0 ? (double)n1 * (1.0 / Z1) : 0.0;
                rt[t * VP + v] = R0;
                rt[(t + 1) * VP + v] = R1;
                if (want_stats) {
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 x = a2[k / 2], y = b2[k / 2];
                        dd_add(ahi[k], alo[k], x.x * R0);
                        dd_add(ahi[k + 1], alo[k + 1], x.y * R0);
                        dd_add(ahi[k], alo[k], y.x * R1);
                        dd_add(ahi[k + 1], alo[k + 1], y.y * R1);
                    }
                }
            }
        } else if (vok) {
            for (int t = 0; t < TILE_S; ++t) {
                const double n = DENSE ? (double)nt[t * V + v] : rt[t * VP + v];
                if (DENSE && !(n > 0.0)) rt[t * VP + v] = 0.0;       // the tile is not cleared: every cell is written
                if (n > 0.0) {
                    const double2 *e2 = reinterpret_cast<const double2 *>(et + t * KP);
                    double ek[KP];
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 x = e2[k / 2];
                        ek[k] = x.x;
                        ek[k + 1] = x.y;
                    }
                    double Z = 0.0;
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        const double e = ek[k] * (EREG ? Ereg[k] : Ekv[k * VP + v]);   // padded k: 0 * 0
                        Z += e;
                    }
                    const double R = n * (1.0 / Z);
                    rt[t * VP + v] = R;
                    if (want_stats) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) dd_add(ahi[k], alo[k], ek[k] * R);
                    }
                }
            }
        }
        __syncthreads();
        if (DENSE && tid == 0 && tile + gridDim.x < ntiles) stage_tile(cnt, tile + gridDim.x, V, nt, mbar);   // flies during phase 3
        // ---- phase 3: sumθ, thread <-> (sample t, four consecutive k): per term one load of R, two 128-bit
        // loads of E, four DFMAs.  Warp kq takes the k-quads kq, kq + NW, ...
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
                    double *st = p.sumtheta + d * p.MK + off + kb;
                    if (kb + 0 < K) st[0] = L[0] * g0;
                    if (kb + 1 < K) st[1] = L[1] * g1;
                    if (kb + 2 < K) st[2] = L[2] * g2;
                    if (kb + 3 < K) st[3] = L[3] * g3;
                }
            }
        }
        __syncthreads();
    }
    if (vok && want_stats) {
        double2 *out = partial + (size_t)blockIdx.x * K * V;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) put_partial(out + k * V + v, ahi[k], alo[k], p.accum);
    }
}

}  // namespace mmsig

namespace mmsig {

// ------------------------------------------------------------------------------------------
// The log-likelihood pass of one modality (reference src/MMCTM.jl:384-448 with update_props!
// :145-154 folded in) over the same 32-sample tiles:
//   P[d][k]  = exp(λ_dk) / Σ_k' exp(λ_dk')            (softmax of the block, no max-subtraction)
//   pw[d][v] = Σ_k P[d][k] ϕ[k][v]                    (index order, two roundings per term)
//   x[d][v]  = n[d][v] · log pw[d][v]                 (dense tile in shared memory, 0 where n = 0)
//   row sum  = blocks of 32 terms summed in term order, the blocks added in order  (DET)
//   ll_m     = Σ_d (row/N_dm)·N_dm, exactly rounded over samples, / Σ_d N_dm   (k_mstep2)
// lane <-> term for pw and the logarithm (ϕ column in registers), thread <-> (sample, block) for
// the sums.  One double-double per sample slot, reduced once per block at the end.
// partial[blockIdx.x * pstride] receives the block's sum.
// ------------------------------------------------------------------------------------------
template <int KP, bool PREG, int NWT, bool DENSE>
__global__ void __launch_bounds__(32 * NWT) k_loglik_tile(MmctmDev p, int m, double2 *partial, int pstride) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ __align__(16) double smem[];
    const int K = p.K[m], V = p.V[m], off = p.koff[m], VP = V | 1, M = p.M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    double *Pkv = smem;                              // [k][VP] when !PREG
    double *xt = Pkv + (PREG ? 0 : KP * VP);         // [t][VP]  n, then n log pw
    double *pt = xt + TILE_S * VP;                   // [t][KP]  exp(λ), then props
    double *ssum = pt + TILE_S * KP;                 // [t]
    double *bsum = ssum + TILE_S;                    // [NW][32]
    long long *rp = reinterpret_cast<long long *>(bsum + NW * 32);
    double *stg = smem + tile_stage_offset((size_t)(bsum - smem) + NW * 32);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 2);                        // [t][V]
    const int *cnt = p.cnt[m];
    const int v = tid;
    const bool vok = v < V;
    const double *ph = p.phi + p.goff[m];
    if (!PREG) {
        for (int i = tid; i < KP * V; i += blockDim.x) {
            const int k = i / V, vv = i % V;
            Pkv[k * VP + vv] = k < K ? ph[k * V + vv] : 0.0;
        }
        __syncthreads();
    }
    double Preg[PREG ? KP : 1];
    if (PREG) {
#pragma unroll
        for (int k = 0; k < KP; ++k) Preg[k] = (vok && k < K) ? ph[k * V + v] : 0.0;
    }
    double ahi = 0.0, alo = 0.0;                      // threads 0..31: Σ over this block's tiles of sample slot t
    const long long *rowptr = p.rowptr[m];
    const int2 *rec = p.rec[m];
    const long long ntiles = (p.D + TILE_S - 1) / TILE_S;
    unsigned parity = 0;
    if (DENSE) {
        if (tid == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (tid == 0 && blockIdx.x < ntiles) stage_tile(cnt, blockIdx.x, V, nt, mbar);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * TILE_S;
        if (!DENSE)
            for (int i = tid; i < TILE_S * VP; i += blockDim.x) xt[i] = 0.0;
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            pt[i] = (k < K && d < p.D) ? det_exp(p.lam[d * p.MK + off + k]) : 0.0;
        }
        if (!DENSE && tid == 0) {
            rp[0] = rowptr[d0];
            rp[TILE_S] = rowptr[min(d0 + TILE_S, p.D)];
            const long long dn = min(d0 + (long long)gridDim.x * TILE_S, p.D);      // this block's next tile
            rp[1] = rowptr[dn];
            rp[2] = rowptr[min(dn + TILE_S, p.D)];
        }
        __syncthreads();
        if (!DENSE) {
#pragma unroll 4
            for (long long w = rp[0] + tid; w < rp[TILE_S]; w += blockDim.x) {
                const int2 r = rec[w];
                xt[(r.x >> 16) * VP + (r.x & 0xffff)] = (double)r.y;        // slot tag | term (k_pack_rows)
            }
        }
        if (tid < TILE_S) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += pt[tid * KP + k];      // index order (src/MMCTM.jl:150)
            ssum[tid] = s;
        }
        __syncthreads();
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            if (k < K && d0 + t < p.D) pt[i] = pt[i] / ssum[t];
        }
        __syncthreads();
        if (DENSE) {
            mbar_wait(mbar, parity);                 // this tile's counts have landed in nt
            parity ^= 1u;
        } else {
            prefetch_records(rec, rp[1], rp[2], tid, blockDim.x);
        }
        if (DENSE && vok) {
            // Dense counts: four samples per trip, no branch on n -- four independent chains (K dependent additions, then
            // the logarithm's 30-odd dependent operations) per thread instead of one.  Per cell the same operations in
            // the same order as below.
#pragma unroll 1
            for (int t = 0; t < TILE_S; t += 4) {
                double pw[4] = {0.0, 0.0, 0.0, 0.0};
                const double2 *p2 = reinterpret_cast<const double2 *>(pt + t * KP);
#pragma unroll
                for (int k = 0; k < KP; k += 2) {
                    const double P0 = PREG ? Preg[k] : Pkv[k * VP + v], P1 = PREG ? Preg[k + 1] : Pkv[(k + 1) * VP + v];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 x = p2[j * (KP / 2) + k / 2];
                        pw[j] += x.x * P0;
                        pw[j] += x.y * P1;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ni = nt[(t + j) * V + v];
                    const double lg = det_log(pw[j] > 0.0 ? pw[j] : 1.0);       // a padding sample's row is 0: keep the logarithm in range
                    xt[(t + j) * VP + v] = ni > 0 ? (double)ni * lg : 0.0;
                }
            }
        } else if (vok) {
            for (int t = 0; t < TILE_S; ++t) {
                const double n = DENSE ? (double)nt[t * V + v] : xt[t * VP + v];
                if (DENSE && !(n > 0.0)) xt[t * VP + v] = 0.0;       // the tile is not cleared: every cell is written
                if (n > 0.0) {
                    const double2 *p2 = reinterpret_cast<const double2 *>(pt + t * KP);
                    double pw = 0.0;
#pragma unroll
                    for (int k = 0; k < KP; k += 2) {
                        const double2 x = p2[k / 2];
                        pw += x.x * (PREG ? Preg[k] : Pkv[k * VP + v]);          // padded k: 0 * 0
                        pw += x.y * (PREG ? Preg[k + 1] : Pkv[(k + 1) * VP + v]);
                    }
                    xt[t * VP + v] = n * det_log(pw);
                }
            }
        }
        __syncthreads();
        if (DENSE && tid == 0 && tile + gridDim.x < ntiles) stage_tile(cnt, tile + gridDim.x, V, nt, mbar);   // flies during the row sums
        {
            const double *row = xt + lane * VP;
            const int vb = 32 * warp, ve = min(V, vb + 32);
            double b = 0.0;
            for (int vv = vb; vv < ve; ++vv) b += row[vv];
            bsum[warp * 32 + lane] = b;
        }
        __syncthreads();
        if (tid < TILE_S) {
            const long long d = d0 + tid;
            if (d < p.D) {
                const double docN = p.N[d * M + m];
                if (docN > 0) {
                    double rs = bsum[tid];
                    for (int j = 1; j < NW; ++j) rs += bsum[j * 32 + tid];
                    const double dl = rs / docN;                     // src/MMCTM.jl:399
                    dd_add(ahi, alo, dl * docN);                     // :412
                }
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
        warp_dd_allreduce(ahi, alo);
        if (lane == 0) partial[(size_t)blockIdx.x * pstride] = make_double2(ahi, alo);
    }
}

}  // namespace mmsig
