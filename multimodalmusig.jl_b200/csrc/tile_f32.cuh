// tile_f32.cuh -- the optional FP32 mode (mmsig_config.precision = MMSIG_PRECISION_FP32): the four tile passes
// (θ / log-likelihood pass of the MMCTM, E / log-likelihood pass of the LDA) in single precision.
//
// What is FP32 here and what is not.  The arithmetic per nonzero (exp, Z, 1/Z, the products, the logarithm, digamma)
// and the per-tile partial sums are float; sums over SAMPLES are double (a tile's float partial is added to a double
// accumulator), and everything outside the tile passes is unchanged: the LD_MMA solves stay FP64 -- their accept
// test compares numbers that differ by 1e-12 relative, an FP32 build of them would be a different
// optimiser -- as do the M-step tables.  State in HBM keeps its FP64 layout (λ, ν, sumθ, γ), so the mode is a flag,
// not a second data format.  The kernels read the dense count tiles of tile_stage.cuh (FP32 mode keeps them for every
// modality).  Softmax-type exponentials subtract the row maximum first (the FP64 kernels follow the reference and do
// not): every consumer is a ratio, so the results agree up to rounding and e^{λ} cannot overflow a float.
// Stated tolerance (tests/test_gpu_fp32.py, against the oracle from the same state): one iteration ϕ / β 1e-5,
// log-likelihood 1e-6 relative (measured 2e-7 / 7e-8, every sample on the same LD_MMA trace); a 20-iteration fit: LDA
// 1e-6, MMCTM 1e-3 on log-likelihood and ELBO (measured 2e-4: LD_MMA's stop decisions amplify any perturbation).
#pragma once
#include "theta_tile.cuh"
#include "lda_tile.cuh"

namespace mmsig {

// ψ(x), x > 0: recurrence to x >= 6, then ln x - 1/(2x) - 1/(12x²) + 1/(120x⁴) - 1/(252x⁶)
__device__ __forceinline__ float digamma_f32(float x) {
    float psi = 0.f;
    while (x < 6.f) {
        psi -= __frcp_rn(x);
        x += 1.f;
    }
    const float t = __frcp_rn(x), t2 = t * t;
    float p = fmaf(t2, -1.f / 252.f, 1.f / 120.f);
    p = fmaf(-t2, p, 1.f / 12.f);
    return psi + (__logf(x) - 0.5f * t - t2 * p);
}

__host__ __device__ inline size_t f32_stage_offset_bytes(size_t floats_before) { return (floats_before * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t f32_stage_bytes(int V) { return 16 + (size_t)TILE_S * V * 4; }
// shared memory of the four kernels (bytes)
__host__ __device__ inline size_t f32_theta_smem(int KP, int V) {
    return f32_stage_offset_bytes((size_t)V * KP + (size_t)TILE_S * (V | 1) + (size_t)TILE_S * KP + TILE_S) + f32_stage_bytes(V);
}
__host__ __device__ inline size_t f32_ll_smem(int KP, int V) {
    return f32_stage_offset_bytes((size_t)TILE_S * (V | 1) + (size_t)TILE_S * KP + 2 * TILE_S + (size_t)((V + 31) / 32) * 32) + f32_stage_bytes(V);
}

// ------------------------------------------------------------------------------------------------------------------
// θ pass of one modality (k_theta_tile's three skinny products) in float.  partial: [grid][K V] (hi = Σ, lo = 0).
// ------------------------------------------------------------------------------------------------------------------
template <int KP, int NWT>
__global__ void __launch_bounds__(32 * NWT) k_theta_tile_f32(MmctmDev p, int m, double2 *partial, int unsmoothed, int want_stats) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K[m], V = p.V[m], off = p.koff[m], VP = V | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    float *Evk = reinterpret_cast<float *>(smem_raw);     // [v][KP]
    float *rt = Evk + V * KP;                             // [t][VP]  R
    float *et = rt + TILE_S * VP;                         // [t][KP]  λ, then L = e^{λ - max}
    float *smax = et + TILE_S * KP;                       // [t]
    unsigned char *stg = smem_raw + f32_stage_offset_bytes((size_t)V * KP + (size_t)TILE_S * VP + (size_t)TILE_S * KP + TILE_S);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 16);          // [t][V]
    const int *cnt = p.cnt[m];
    const int v = tid;
    const bool vok = v < V;
    const double *Eg = (unsmoothed ? p.phi : p.Elnphi) + p.goff[m];
    for (int i = tid; i < V * KP; i += blockDim.x) {
        const int vv = i / KP, k = i % KP;
        Evk[i] = k < K ? (float)(unsmoothed ? Eg[k * V + vv] : det_exp(Eg[k * V + vv])) : 0.f;
    }
    const long long ntiles = (p.D + TILE_S - 1) / TILE_S;
    unsigned parity = 0;
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0 && blockIdx.x < ntiles) stage_tile(cnt, blockIdx.x, V, nt, mbar);
    float Ereg[KP];
    double acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        Ereg[k] = vok ? Evk[v * KP + k] : 0.f;
        acc[k] = 0.0;
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * TILE_S;
        // ---- phase 1: L = e^{λ - max_k λ}
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            et[i] = (k < K && d < p.D) ? (float)p.lam_prev[d * p.MK + off + k] : 0.f;
        }
        __syncthreads();
        if (tid < TILE_S) {
            float mx = et[tid * KP];
            for (int k = 1; k < K; ++k) mx = fmaxf(mx, et[tid * KP + k]);
            smax[tid] = mx;
        }
        __syncthreads();
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            et[i] = (k < K && d0 + t < p.D) ? __expf(et[i] - smax[t]) : 0.f;
        }
        __syncthreads();
        mbar_wait(mbar, parity);
        parity ^= 1u;
        // ---- phase 2: Z, R and the tile's statistics, lane <-> term
        if (vok) {
            float tacc[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) tacc[k] = 0.f;
            for (int t = 0; t < TILE_S; ++t) {
                const int ni = nt[t * V + v];
                float R = 0.f;
                if (ni > 0) {
                    const float4 *e4 = reinterpret_cast<const float4 *>(et + t * KP);
                    float ek[KP];
                    float z0 = 0.f, z1 = 0.f;
#pragma unroll
                    for (int k = 0; k < KP; k += 4) {
                        const float4 x = e4[k / 4];
                        ek[k] = x.x; ek[k + 1] = x.y; ek[k + 2] = x.z; ek[k + 3] = x.w;
                        z0 = fmaf(x.x, Ereg[k], z0);
                        z1 = fmaf(x.y, Ereg[k + 1], z1);
                        z0 = fmaf(x.z, Ereg[k + 2], z0);
                        z1 = fmaf(x.w, Ereg[k + 3], z1);
                    }
                    R = __fdividef((float)ni, z0 + z1);
                    if (want_stats) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) tacc[k] = fmaf(ek[k], R, tacc[k]);
                    }
                }
                rt[t * VP + v] = R;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] += (double)tacc[k];
        }
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < ntiles) stage_tile(cnt, tile + gridDim.x, V, nt, mbar);
        // ---- phase 3: sumθ, thread <-> (sample, four consecutive k)
        {
            const int t = lane;
            const long long d = d0 + t;
            const float *row = rt + t * VP;
            for (int kb = 4 * warp; kb < KP; kb += 4 * NW) {
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll 4
                for (int vv = 0; vv < V; ++vv) {
                    const float R = row[vv];
                    const float4 a = *reinterpret_cast<const float4 *>(Evk + vv * KP + kb);
                    g0 = fmaf(a.x, R, g0);
                    g1 = fmaf(a.y, R, g1);
                    g2 = fmaf(a.z, R, g2);
                    g3 = fmaf(a.w, R, g3);
                }
                if (d < p.D) {
                    const float *L = et + t * KP + kb;
                    double *st = p.sumtheta + d * p.MK + off + kb;
                    if (kb + 0 < K) st[0] = (double)(L[0] * g0);
                    if (kb + 1 < K) st[1] = (double)(L[1] * g1);
                    if (kb + 2 < K) st[2] = (double)(L[2] * g2);
                    if (kb + 3 < K) st[3] = (double)(L[3] * g3);
                }
            }
        }
        __syncthreads();
    }
    if (vok && want_stats) {
        double2 *out = partial + (size_t)blockIdx.x * K * V;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) put_partial(out + k * V + v, acc[k], 0.0, p.accum);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// log-likelihood pass of one modality (k_loglik_tile) in float; partial[blockIdx.x * pstride] receives the block's sum
// ------------------------------------------------------------------------------------------------------------------
template <int KP, int NWT>
__global__ void __launch_bounds__(32 * NWT) k_loglik_tile_f32(MmctmDev p, int m, double2 *partial, int pstride) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K[m], V = p.V[m], off = p.koff[m], VP = V | 1, M = p.M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    float *xt = reinterpret_cast<float *>(smem_raw);      // [t][VP]  n log pw
    float *pt = xt + TILE_S * VP;                         // [t][KP]  λ, then props
    float *smax = pt + TILE_S * KP;                       // [t]
    float *ssum = smax + TILE_S;                          // [t]
    float *bsum = ssum + TILE_S;                          // [NW][32]
    unsigned char *stg = smem_raw + f32_stage_offset_bytes((size_t)TILE_S * VP + (size_t)TILE_S * KP + 2 * TILE_S + (size_t)NW * 32);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 16);
    const int *cnt = p.cnt[m];
    const int v = tid;
    const bool vok = v < V;
    const double *ph = p.phi + p.goff[m];
    float Preg[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) Preg[k] = (vok && k < K) ? (float)ph[k * V + v] : 0.f;
    double ahi = 0.0, alo = 0.0;
    const long long ntiles = (p.D + TILE_S - 1) / TILE_S;
    unsigned parity = 0;
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0 && blockIdx.x < ntiles) stage_tile(cnt, blockIdx.x, V, nt, mbar);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * TILE_S;
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            pt[i] = (k < K && d < p.D) ? (float)p.lam[d * p.MK + off + k] : 0.f;
        }
        __syncthreads();
        if (tid < TILE_S) {
            float mx = pt[tid * KP];
            for (int k = 1; k < K; ++k) mx = fmaxf(mx, pt[tid * KP + k]);
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += __expf(pt[tid * KP + k] - mx);
            smax[tid] = mx;
            ssum[tid] = s;
        }
        __syncthreads();
        for (int i = tid; i < TILE_S * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            pt[i] = (k < K && d0 + t < p.D) ? __fdividef(__expf(pt[i] - smax[t]), ssum[t]) : 0.f;
        }
        __syncthreads();
        mbar_wait(mbar, parity);
        parity ^= 1u;
        if (vok) {
            for (int t = 0; t < TILE_S; ++t) {
                const int ni = nt[t * V + v];
                float x = 0.f;
                if (ni > 0) {
                    const float4 *p4 = reinterpret_cast<const float4 *>(pt + t * KP);
                    float w0 = 0.f, w1 = 0.f;
#pragma unroll
                    for (int k = 0; k < KP; k += 4) {
                        const float4 q = p4[k / 4];
                        w0 = fmaf(q.x, Preg[k], w0);
                        w1 = fmaf(q.y, Preg[k + 1], w1);
                        w0 = fmaf(q.z, Preg[k + 2], w0);
                        w1 = fmaf(q.w, Preg[k + 3], w1);
                    }
                    x = (float)ni * __logf(w0 + w1);
                }
                xt[t * VP + v] = x;
            }
        }
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < ntiles) stage_tile(cnt, tile + gridDim.x, V, nt, mbar);
        {
            const float *row = xt + lane * VP;
            const int vb = 32 * warp, ve = min(V, vb + 32);
            float b = 0.f;
            for (int vv = vb; vv < ve; ++vv) b += row[vv];
            bsum[warp * 32 + lane] = b;
        }
        __syncthreads();
        if (tid < TILE_S) {
            const long long d = d0 + tid;
            if (d < p.D && p.N[d * M + m] > 0) {
                float rs = bsum[tid];
                for (int j = 1; j < NW; ++j) rs += bsum[j * 32 + tid];
                dd_add(ahi, alo, (double)rs);
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
        warp_dd_allreduce(ahi, alo);
        if (lane == 0) partial[(size_t)blockIdx.x * pstride] = make_double2(ahi, alo);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// LDA: the fused E pass (k_lda_estep_t32) in float
// ------------------------------------------------------------------------------------------------------------------
template <int KP, int NWT>
__global__ void __launch_bounds__(32 * NWT) k_lda_estep_f32(LdaDev p, double2 *partial, const double *Etab, int want_stats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, V = p.V, VP = V | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    float *Evk = reinterpret_cast<float *>(smem_raw);     // [v][KP]
    float *rt = Evk + V * KP;                             // [t][VP]
    float *et = rt + LDA_TS * VP;                         // [t][KP]  γ, then e^{Elnθ}
    float *ssum = et + LDA_TS * KP;                       // [t]      ψ(Σ_k γ)
    unsigned char *stg = smem_raw + f32_stage_offset_bytes((size_t)V * KP + (size_t)LDA_TS * VP + (size_t)LDA_TS * KP + LDA_TS);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 16);
    const int v = tid;
    const bool vok = v < V;
    for (int i = tid; i < V * KP; i += blockDim.x) {
        const int vv = i / KP, k = i % KP;
        Evk[i] = k < K ? (float)Etab[k * V + vv] : 0.f;
    }
    const long long ntiles = (p.D + LDA_TS - 1) / LDA_TS;
    unsigned parity = 0;
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0 && blockIdx.x < ntiles) stage_tile(p.cnt, blockIdx.x, V, nt, mbar);
    float Ereg[KP];
    double acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        Ereg[k] = vok ? Evk[v * KP + k] : 0.f;
        acc[k] = 0.0;
    }
    const float alpha = (float)p.alpha;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * LDA_TS;
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            et[i] = (k < K && d < p.D) ? (float)p.gamma[d * K + k] : 0.f;
        }
        __syncthreads();
        if (tid < LDA_TS) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += et[tid * KP + k];
            ssum[tid] = (d0 + tid < p.D) ? digamma_f32(s) : 0.f;
        }
        __syncthreads();
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            et[i] = (k < K && d0 + t < p.D) ? __expf(digamma_f32(et[i]) - ssum[t]) : 0.f;
        }
        __syncthreads();
        mbar_wait(mbar, parity);
        parity ^= 1u;
        if (vok) {
            float tacc[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) tacc[k] = 0.f;
            for (int t = 0; t < LDA_TS; ++t) {
                const int ni = nt[t * V + v];
                float R = 0.f;
                if (ni > 0) {
                    const float4 *e4 = reinterpret_cast<const float4 *>(et + t * KP);
                    float ek[KP];
                    float z0 = 0.f, z1 = 0.f;
#pragma unroll
                    for (int k = 0; k < KP; k += 4) {
                        const float4 x = e4[k / 4];
                        ek[k] = x.x; ek[k + 1] = x.y; ek[k + 2] = x.z; ek[k + 3] = x.w;
                        z0 = fmaf(x.x, Ereg[k], z0);
                        z1 = fmaf(x.y, Ereg[k + 1], z1);
                        z0 = fmaf(x.z, Ereg[k + 2], z0);
                        z1 = fmaf(x.w, Ereg[k + 3], z1);
                    }
                    R = __fdividef((float)ni, z0 + z1);
                    if (want_stats) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) tacc[k] = fmaf(ek[k], R, tacc[k]);
                    }
                }
                rt[t * VP + v] = R;
            }
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] += (double)tacc[k];
        }
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < ntiles) stage_tile(p.cnt, tile + gridDim.x, V, nt, mbar);
        {
            const int t = lane;
            const long long d = d0 + t;
            const float *row = rt + t * VP;
            for (int kb = 4 * warp; kb < KP; kb += 4 * NW) {
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
#pragma unroll 4
                for (int vv = 0; vv < V; ++vv) {
                    const float R = row[vv];
                    const float4 a = *reinterpret_cast<const float4 *>(Evk + vv * KP + kb);
                    g0 = fmaf(a.x, R, g0);
                    g1 = fmaf(a.y, R, g1);
                    g2 = fmaf(a.z, R, g2);
                    g3 = fmaf(a.w, R, g3);
                }
                if (d < p.D) {
                    const float *L = et + t * KP + kb;
                    double *gn = p.gamma_next + d * K + kb;
                    if (kb + 0 < K) gn[0] = (double)fmaf(L[0], g0, alpha);
                    if (kb + 1 < K) gn[1] = (double)fmaf(L[1], g1, alpha);
                    if (kb + 2 < K) gn[2] = (double)fmaf(L[2], g2, alpha);
                    if (kb + 3 < K) gn[3] = (double)fmaf(L[3], g3, alpha);
                }
            }
        }
        __syncthreads();
    }
    if (vok && want_stats) {
        double2 *out = partial + (size_t)blockIdx.x * K * V;
#pragma unroll
        for (int k = 0; k < KP; ++k)
            if (k < K) out[k * V + v] = make_double2((double)Ereg[k] * acc[k], 0.0);
    }
}

// LDA log-likelihood pass (k_lda_ll_tile) in float.  partial: [gridDim.x] double2.
template <int KP, int NWT>
__global__ void __launch_bounds__(32 * NWT) k_lda_ll_f32(LdaDev p, double2 *partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int K = p.K, V = p.V, VP = V | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    float *xt = reinterpret_cast<float *>(smem_raw);      // [t][VP]
    float *pt = xt + LDA_TS * VP;                         // [t][KP]  γ, then θ
    float *smax = pt + LDA_TS * KP;                       // [t]  (unused: layout shared with k_loglik_tile_f32)
    float *ssum = smax + LDA_TS;                          // [t]
    float *bsum = ssum + LDA_TS;                          // [NW][32]
    unsigned char *stg = smem_raw + f32_stage_offset_bytes((size_t)LDA_TS * VP + (size_t)LDA_TS * KP + 2 * LDA_TS + (size_t)NW * 32);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stg);
    int *nt = reinterpret_cast<int *>(stg + 16);
    const int v = tid;
    const bool vok = v < V;
    float Breg[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) Breg[k] = (vok && k < K) ? (float)p.beta[k * V + v] : 0.f;
    double ahi = 0.0, alo = 0.0;
    const long long ntiles = (p.D + LDA_TS - 1) / LDA_TS;
    unsigned parity = 0;
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0 && blockIdx.x < ntiles) stage_tile(p.cnt, blockIdx.x, V, nt, mbar);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long d0 = tile * LDA_TS;
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            const long long d = d0 + t;
            pt[i] = (k < K && d < p.D) ? (float)p.gamma[d * K + k] : 0.f;
        }
        __syncthreads();
        if (tid < LDA_TS) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += pt[tid * KP + k];
            ssum[tid] = s;
        }
        __syncthreads();
        for (int i = tid; i < LDA_TS * KP; i += blockDim.x) {
            const int t = i / KP, k = i % KP;
            if (k < K && d0 + t < p.D) pt[i] = __fdividef(pt[i], ssum[t]);
        }
        __syncthreads();
        mbar_wait(mbar, parity);
        parity ^= 1u;
        if (vok) {
            for (int t = 0; t < LDA_TS; ++t) {
                const int ni = nt[t * V + v];
                float x = 0.f;
                if (ni > 0) {
                    const float4 *p4 = reinterpret_cast<const float4 *>(pt + t * KP);
                    float w0 = 0.f, w1 = 0.f;
#pragma unroll
                    for (int k = 0; k < KP; k += 4) {
                        const float4 q = p4[k / 4];
                        w0 = fmaf(q.x, Breg[k], w0);
                        w1 = fmaf(q.y, Breg[k + 1], w1);
                        w0 = fmaf(q.z, Breg[k + 2], w0);
                        w1 = fmaf(q.w, Breg[k + 3], w1);
                    }
                    x = (float)ni * __logf(w0 + w1);
                }
                xt[t * VP + v] = x;
            }
        }
        __syncthreads();
        if (tid == 0 && tile + gridDim.x < ntiles) stage_tile(p.cnt, tile + gridDim.x, V, nt, mbar);
        {
            const float *row = xt + lane * VP;
            const int vb = 32 * warp, ve = min(V, vb + 32);
            float b = 0.f;
            for (int vv = vb; vv < ve; ++vv) b += row[vv];
            bsum[warp * 32 + lane] = b;
        }
        __syncthreads();
        if (tid < LDA_TS && d0 + tid < p.D) {
            float rs = bsum[tid];
            for (int j = 1; j < NW; ++j) rs += bsum[j * 32 + tid];
            dd_add(ahi, alo, (double)rs);
        }
        __syncthreads();
    }
    if (warp == 0) {
        warp_dd_allreduce(ahi, alo);
        if (lane == 0) partial[blockIdx.x] = make_double2(ahi, alo);
    }
}

}  // namespace mmsig
