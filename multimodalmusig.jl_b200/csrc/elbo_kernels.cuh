// elbo_kernels.cuh -- calculate_elbo of the MMCTM (reference src/MMCTM.jl:271-382), evaluated
// once per fit with the staleness of src/MMCTM.jl:490: θ, ζ, sumθ are those of the last E-step
// (old λ / Elnϕ), λ, ν, μ, invΣ, γ, Elnϕ are current.  θ is never stored: the θ-dependent terms
// are rewritten over tables (ElnPX) or recomputed from (lam_prev, Elnphi_prev) (ElnQZ).
#pragma once
#include "mmctm_kernels.cuh"

namespace mmsig {

// block-level dd reduction of NV values per thread-warp-lane -> out[NV] (thread 0 writes)
template <int NV>
__device__ __forceinline__ void block_reduce_dd_write(double (&hi)[NV], double (&lo)[NV], double2 *red /*[8][NV]*/,
                                                      double2 *out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) warp_dd_allreduce(hi[i], lo[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) red[warp * NV + i] = make_double2(hi[i], lo[i]);
    __syncthreads();
    if (threadIdx.x < NV) {
        double h = 0.0, l = 0.0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) dd_merge(h, l, red[wv * NV + threadIdx.x].x, red[wv * NV + threadIdx.x].y);
        out[threadIdx.x] = make_double2(h, l);
    }
}

// Table-only terms (single block): out[0] = ElnPϕ (:271-284), out[1] = ElnQϕ (:338-350),
// out[2] = ElnPX (:318-336) = Σ_kv (Σ n θ)_kv Elnϕ_kv, out[3] = logdet(invΣ) (:292).
__global__ void __launch_bounds__(256) k_elbo_tables(MmctmDev p, double *out) {
    extern __shared__ double lu_smem[];                  // A, B: MK x MK each; piv: MK ints
    __shared__ double2 red[8 * 3];
    const int G = p.goff[p.M], MK = p.MK;
    double *A = lu_smem, *B = lu_smem + MK * MK;
    int *piv = reinterpret_cast<int *>(B + MK * MK);
    double hi[3] = {0, 0, 0}, lo[3] = {0, 0, 0};
    if (p.factored) {
        // IMMCTM: ElnPϕ (src/IMMCTM.jl:244-261) and ElnQϕ (:311-325) over the feature tables
        for (int r = threadIdx.x; r < p.R; r += blockDim.x) {
            const double a = p.alphaf[p.row_alpha[r]];
            const double *g = p.gammaf + p.row_off[r];
            double r0 = 0.0, s0 = 0.0, r1 = 0.0, s1 = 0.0;
            for (int j = 0; j < p.row_len[r]; ++j) { r0 += lgamma(a); s0 += a; r1 += lgamma(g[j]); s1 += g[j]; }
            r0 -= lgamma(s0);
            r1 -= lgamma(s1);
            dd_add(hi[0], lo[0], -r0);
            dd_add(hi[1], lo[1], -r1);
        }
        for (int t = threadIdx.x; t < p.T; t += blockDim.x) {
            const double E = p.Elnphif[t];
            dd_add(hi[0], lo[0], (p.alphaf[p.row_alpha[p.ent_row[t]]] - 1) * E);
            dd_add(hi[1], lo[1], (p.gammaf[t] - 1) * E);
        }
        for (int i = threadIdx.x; i < G; i += blockDim.x) dd_add(hi[2], lo[2], p.stats[i] * p.Elnphi[i]);
    } else {
    // per (m,k) row pieces handled by one thread each; element-wise pieces strided
    if ((int)threadIdx.x < MK) {
        int m = 0;
        while ((int)threadIdx.x >= p.koff[m + 1]) ++m;
        const int k = threadIdx.x - p.koff[m], V = p.V[m];
        const double a = p.alpha[m];
        const double *g = p.gamma + p.goff[m] + k * V;
        // logmvbeta(fill(α, V)) and logmvbeta(γ_mk), src/common.jl:1-9
        double r0 = 0.0, s0 = 0.0, r1 = 0.0, s1 = 0.0;
        for (int v = 0; v < V; ++v) { r0 += lgamma(a); s0 += a; r1 += lgamma(g[v]); s1 += g[v]; }
        r0 -= lgamma(s0);
        r1 -= lgamma(s1);
        dd_add(hi[0], lo[0], -r0);
        dd_add(hi[1], lo[1], -r1);
    }
    for (int i = threadIdx.x; i < G; i += blockDim.x) {
        int m = 0;
        while (i >= p.goff[m + 1]) ++m;
        const double E = p.Elnphi[i];
        dd_add(hi[0], lo[0], (p.alpha[m] - 1) * E);
        dd_add(hi[1], lo[1], (p.gamma[i] - 1) * E);
        dd_add(hi[2], lo[2], p.stats[i] * E);
    }
    }
    double2 res[3];
    __shared__ double2 outsh[3];
    block_reduce_dd_write<3>(hi, lo, red, outsh);
    __syncthreads();
    if (threadIdx.x < 3) { res[0] = outsh[threadIdx.x]; out[threadIdx.x] = dd_round(res[0].x, res[0].y); }
    if (threadIdx.x < 32) {
        for (int i = threadIdx.x; i < MK * MK; i += 32) A[i] = p.invSigma[i];
        __syncwarp();
        double ld = 0.0;
        const bool ok = warp_lu_inverse(MK, A, B, piv, nullptr, &ld);
        if (threadIdx.x == 0) out[3] = ok ? ld : -__longlong_as_double(0x7ff0000000000000LL);
    }
}

// Per-sample terms, warp per sample, lane j = coordinate j.  partial[block][4] dd:
//  [0] Σ_d (Σ_j ν_j S_jj + Δᵀ S Δ)          -> ElnPη (:286-300)
//  [1] Σ_d ElnPZ_d (:302-316), stale sumθ, ζ
//  [2] Σ_d Σ_j log ν_j                        -> ElnQη (:352-358)
__global__ void __launch_bounds__(256) k_elbo_samples(MmctmDev p, double2 *partial) {
    extern __shared__ double es_smem[];                 // S: MK x MK ; per warp: diff[MK]
    __shared__ double2 red[8 * 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = p.MK, M = p.M;
    double *S = es_smem, *dsh = es_smem + MK * MK + warp * MK;
    for (int i = threadIdx.x; i < MK * MK; i += blockDim.x) S[i] = p.invSigma[i];
    __syncthreads();
    double hi[3] = {0, 0, 0}, lo[3] = {0, 0, 0};
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        __syncwarp();
        for (int j = lane; j < MK; j += 32) dsh[j] = p.lam[d * MK + j] - p.mu[j];
        __syncwarp();
        for (int j = lane; j < MK; j += 32) {
            int mod = 0;
            for (int m = 0; m < M; ++m)
                if (j >= p.koff[m]) mod = m;
            const double lam = p.lam[d * MK + j], nu = p.nu[d * MK + j], diff = dsh[j];
            double q = 0.0;
            for (int i = 0; i < MK; ++i) q += S[i * MK + j] * dsh[i];      // (Δᵀ S)_j as in diff' * invΣ * diff
            dd_add(hi[0], lo[0], nu * S[j * MK + j]);
            dd_add(hi[0], lo[0], q * diff);
            const double zeta = p.zeta[d * M + mod], Ndm = p.N[d * M + mod];
            const double c = Ndm / zeta;
            dd_add(hi[1], lo[1], lam * p.sumtheta[d * MK + j]);
            dd_add(hi[1], lo[1], -(c * det_exp(lam + 0.5 * nu)));
            if (j == p.koff[mod]) {
                dd_add(hi[1], lo[1], Ndm);
                dd_add(hi[1], lo[1], -(Ndm * det_log(zeta)));   // 0 * log ζ for an empty row, as the reference evaluates it
            }
            dd_add(hi[2], lo[2], det_log(nu));
        }
    }
    block_reduce_dd_write<3>(hi, lo, red, partial + (size_t)blockIdx.x * 4);
}

// ElnQZ (:360-370) of one modality: Σ n θ log θ with the θ of the last E-step.
// partial[block][4] dd, slot [3] accumulated across modality launches by the caller's combine.
__global__ void __launch_bounds__(256) k_elbo_qz(MmctmDev p, int m, double2 *partial) {
    extern __shared__ double smem[];
    __shared__ double2 red[8];
    const int K = p.K[m], V = p.V[m], KV = K * V, off = p.koff[m];
    double *Eln = smem;
    const double *Eg = p.Elnphi_prev + p.goff[m];
    for (int i = threadIdx.x; i < KV; i += blockDim.x) Eln[i] = det_exp(Eg[i]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double hi[1] = {0.0}, lo[1] = {0.0};
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        const long long beg = p.rowptr[m][d], end = p.rowptr[m][d + 1];
        for (long long w = beg + lane; w < end; w += 32) {
            const int2 r = p.rec[m][w];
            double Z = 0.0;
            for (int k = 0; k < K; ++k) Z += det_exp(p.lam_prev[d * p.MK + off + k]) * Eln[k * V + (r.x & 0xffff)];
            double s = 0.0;
            const double rz = 1.0 / Z;
            for (int k = 0; k < K; ++k) {
                const double th = (det_exp(p.lam_prev[d * p.MK + off + k]) * Eln[k * V + (r.x & 0xffff)]) * rz;
                if (th > 0.0) s += th * det_log(th);
            }
            dd_add(hi[0], lo[0], (double)r.y * s);
        }
    }
    __syncwarp();
    block_reduce_dd_write<1>(hi, lo, red, partial + (size_t)blockIdx.x * 4 + 3);
}

}  // namespace mmsig
