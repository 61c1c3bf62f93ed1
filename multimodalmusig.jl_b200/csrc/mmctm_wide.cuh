// mmctm_wide.cuh -- the per-sample kernels for 32 < ΣK_m <= 64: two coordinates per lane
// (coordinate j lives on lane j % 32, slot j / 32).  Same arithmetic as mmctm_kernels.cuh: the
// fixed 32-leaf tree first adds a lane's slots in slot order (leaf i = v[i] + v[i+32]), then runs
// the butterfly -- exactly the oracle's tree_sum32 for n > 32.  Not tuned like the one-
// coordinate-per-lane path; it exists so that large signature sets are served, not refused.
#pragma once
#include "mmctm_kernels.cuh"

namespace mmsig {

constexpr int WCPL = 2;                 // coordinates per lane
constexpr int WMK = 64;                 // padded ΣK_m
constexpr int WSTRIDE = 66;             // doubles per padded invΣ row in shared memory

struct WideCtx {
    double Sjj[WCPL], muj[WCPL], c[WCPL], s[WCPL], other[WCPL];
    bool active[WCPL];
};

template <bool IS_NU>
__device__ __forceinline__ void wide_eval_local(const double (&x)[WCPL], const WideCtx &c, const double *__restrict__ ST,
                                                double *dsh, int lane, double &tsum, double (&g)[WCPL]) {
    double t[WCPL];
    if (IS_NU) {
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            const double e = det_exp(c.other[s] + 0.5 * x[s]);
            const double grad = (-0.5 * c.Sjj[s] - (c.c[s] / 2) * e) + fast_rcp(2 * x[s]);
            t[s] = (-0.5 * (x[s] * c.Sjj[s]) - c.c[s] * e) + det_log(x[s]) / 2;
            g[s] = -grad;
        }
    } else {
        double diff[WCPL];
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            diff[s] = x[s] - c.muj[s];
            dsh[lane + 32 * s] = c.active[s] ? diff[s] : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            const double e = det_exp(x[s] + c.other[s]);
            double q = 0.0, qo = 0.0;
            const double2 *srow = reinterpret_cast<const double2 *>(ST + (lane + 32 * s) * WSTRIDE);
            const double2 *dv2 = reinterpret_cast<const double2 *>(dsh);
#pragma unroll 4
            for (int i = 0; i < WMK / 2; ++i) {
                const double2 sv = srow[i], dv = dv2[i];
                q = fma(sv.x, dv.x, q);                 // DET: even / odd index chains, then one add
                qo = fma(sv.y, dv.y, qo);
            }
            q = q + qo;
            const double ce = c.c[s] * e;
            const double grad = (-q + c.s[s]) - ce;
            const double a = q * diff[s], b = x[s] * c.s[s];
            t[s] = (b - 0.5 * a) - ce;
            g[s] = -grad;
        }
        __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < WCPL; ++s)
        if (!c.active[s]) { t[s] = 0.0; g[s] = 1.0; }
    tsum = t[0] + t[1];                  // leaf = slot 0 + slot 1, then the butterfly
}

template <bool IS_NU>
__device__ __forceinline__ int wide_mma_solve(double (&x)[WCPL], const WideCtx &c, const double *__restrict__ ST,
                                              double *dsh, int lane, int stop_rule) {
    const double lb = IS_NU ? 1e-7 : -__longlong_as_double(0x7ff0000000000000LL);
    const double xtol_rel = 1e-4, xtol_abs = 1e-4;
    double sigma[WCPL], g[WCPL], gcur[WCPL], xcur[WCPL], xprev[WCPL], xprevprev[WCPL];
    double rho = 1.0, fmin, fcur;
    {
        double t;
        wide_eval_local<IS_NU>(x, c, ST, dsh, lane, t, g);
        fmin = -warp_tree_sum(t);
    }
#pragma unroll
    for (int s = 0; s < WCPL; ++s) { sigma[s] = 1.0; xcur[s] = xprev[s] = xprevprev[s] = x[s]; }
    int nev = 1, k = 0;
    while (true) {
        ++k;
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            if (k > 1) xprevprev[s] = xprev[s];
            xprev[s] = xcur[s];
        }
        while (true) {
            double gl[WCPL], wl[WCPL], xc[WCPL];
#pragma unroll
            for (int s = 0; s < WCPL; ++s) {
                double u = g[s];
                const double v = fabs(g[s]) * sigma[s] + 0.5 * rho;
                const double sigma2 = sigma[s] * sigma[s];
                u *= sigma2;
                const double qv = fast_div(u, v);
                const double r = qv * fast_rcp(sigma[s]);      // DET: (u / v)(1 / sigma)
                const double om = fabs(1 - r * r);
                const double sq = fast_sqrt(om < 0x1p-200 ? 0x1p-200 : om);   // om is 0 or >= 2^-53: sqrt(0) -> 2^-100, and -1 - 2^-100 == -1
                double dx = fast_div(qv, -1 - sq);
                double xn = x[s] + dx;
                if (xn > x[s] + 0.9 * sigma[s]) xn = x[s] + 0.9 * sigma[s];
                else if (xn < x[s] - 0.9 * sigma[s]) xn = x[s] - 0.9 * sigma[s];
                if (xn < lb) xn = lb;
                if (!c.active[s]) xn = x[s];
                dx = xn - x[s];
                const double dx2 = dx * dx;
                const double denominv = fast_rcp(sigma2 - dx2);       // |dx| <= 0.9 sigma
                const double cc = sigma2 * dx;
                gl[s] = (g[s] * cc + (fabs(g[s]) * sigma[s] + 0.5 * rho) * dx2) * denominv;
                wl[s] = 0.5 * dx2 * denominv;
                xc[s] = xn;
            }
            double gterm = gl[0] + gl[1], wterm = wl[0] + wl[1], t;
#pragma unroll
            for (int s = 0; s < WCPL; ++s) xcur[s] = xc[s];
            wide_eval_local<IS_NU>(xcur, c, ST, dsh, lane, t, gcur);
            warp_tree_sum3h(gterm, wterm, t, lane);
            const double gval = fmin + gterm, wval = wterm;
            fcur = -t;
            ++nev;
            const bool inner_done = __all_sync(FULLMASK, gval >= fcur);
            if (fcur < fmin) {
                fmin = fcur;
#pragma unroll
                for (int s = 0; s < WCPL; ++s) { x[s] = xcur[s]; g[s] = gcur[s]; }
            }
            if (nev >= MMA_MAXEVAL) return nev;
            if (inner_done) break;
            if (__all_sync(FULLMASK, fcur > gval)) {
                const double r1 = 10 * rho, r2 = 1.1 * (rho + guarded_div(fcur - gval, wval));
                rho = r1 < r2 ? r1 : r2;
            }
        }
        double ad[WCPL];
#pragma unroll
        for (int s = 0; s < WCPL; ++s) ad[s] = c.active[s] ? fabs(xcur[s] - xprev[s]) : 0.0;
        bool stop;
        if (stop_rule == 1) {
            bool ok = true;
#pragma unroll
            for (int s = 0; s < WCPL; ++s)
                ok = ok && (!c.active[s] || ad[s] < xtol_abs ||
                            ad[s] < xtol_rel * (fabs(xcur[s]) + fabs(xprev[s])) * 0.5 || xcur[s] == xprev[s]);
            stop = __all_sync(FULLMASK, ok);
        } else {
            double dn = ad[0] + ad[1];
            double xn = (c.active[0] ? fabs(xcur[0]) : 0.0) + (c.active[1] ? fabs(xcur[1]) : 0.0);
            warp_tree_sum2h(dn, xn, lane);
            stop = __all_sync(FULLMASK, dn <= xtol_rel * xn) ||
                   __all_sync(FULLMASK, !(ad[0] > xtol_abs) && !(ad[1] > xtol_abs));
        }
        if (stop) break;
        rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;
        if (k > 1) {
#pragma unroll
            for (int s = 0; s < WCPL; ++s) {
                const double s2 = (xcur[s] - xprev[s]) * (xprev[s] - xprevprev[s]);
                sigma[s] *= s2 < 0 ? 0.7 : (s2 > 0 ? 1.2 : 1.0);
            }
        }
    }
    return nev;
}

// sequential block sum from a per-warp shared array v[0..MK): Σ_{i in [lo, hi)} v[i], index order
__device__ __forceinline__ double smem_block_sum(const double *v, int lo, int hi) {
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += v[i];
    return s;
}

__global__ void __launch_bounds__(256) k_solve_wide(MmctmDev p, double2 *partial) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    extern __shared__ double wsm[];
    double *ST = wsm;                                    // WMK x WSTRIDE
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *dsh = wsm + WMK * WSTRIDE + warp * WMK;      // 8 x WMK
    double2 *red = reinterpret_cast<double2 *>(wsm + WMK * WSTRIDE + 8 * WMK);   // 8 x 4 x 32
    const int MK = p.MK, M = p.M;
    for (int t = threadIdx.x; t < WMK * WSTRIDE; t += blockDim.x) {
        const int j = t / WSTRIDE, i = t % WSTRIDE;
        ST[t] = (i < MK && j < MK) ? p.invSigma[j * MK + i] : 0.0;
    }
    for (int i = lane; i < WMK; i += 32) dsh[i] = 0.0;
    __syncthreads();
    WideCtx c;
    int mod[WCPL], blo[WCPL], bhi[WCPL];
#pragma unroll
    for (int s = 0; s < WCPL; ++s) {
        const int j = lane + 32 * s;
        c.active[s] = j < MK;
        mod[s] = 0;
        for (int m = 0; m < M; ++m)
            if (j >= p.koff[m]) mod[s] = m;
        blo[s] = p.koff[mod[s]];
        bhi[s] = p.koff[mod[s] + 1];
        c.Sjj[s] = c.active[s] ? p.invSigma[j * MK + j] : 0.0;
        c.muj[s] = c.active[s] ? p.mu[j] : 0.0;
    }
    double acc[4][WCPL];                                 // Σλ hi/lo, Σν hi/lo
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int s = 0; s < WCPL; ++s) acc[a][s] = 0.0;
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = next_sample(p.work, FULLMASK, 0, lane == 0); d < p.D; d = next_sample(p.work, FULLMASK, 0, lane == 0)) {
        double lam[WCPL], nu[WCPL];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            const long long base = d * MK + lane + 32 * s;
            lam[s] = c.active[s] ? p.lam_prev[base] : 0.0;
            nu[s] = c.active[s] ? p.nu[base] : 1.5;
            c.s[s] = c.active[s] ? p.sumtheta[base] : 0.0;
            dsh[lane + 32 * s] = c.active[s] ? det_exp(lam[s] + 0.5 * nu[s]) : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            const double zeta = c.active[s] ? smem_block_sum(dsh, blo[s], bhi[s]) : 1.0;
            const double Ndm = c.active[s] ? p.N[d * M + mod[s]] : 0.0;
            c.c[s] = c.active[s] ? Ndm / zeta : 0.0;
            if (c.active[s] && lane + 32 * s == blo[s]) p.zeta[d * M + mod[s]] = zeta;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < WCPL; ++s) c.other[s] = lam[s];
        const int nev_nu = wide_mma_solve<true>(nu, c, ST, dsh, lane, p.stop_rule);
#pragma unroll
        for (int s = 0; s < WCPL; ++s) c.other[s] = 0.5 * nu[s];
        const int nev_lam = wide_mma_solve<false>(lam, c, ST, dsh, lane, p.stop_rule);
#pragma unroll
        for (int s = 0; s < WCPL; ++s)
            if (c.active[s]) {
                const long long base = d * MK + lane + 32 * s;
                p.lam[base] = lam[s];
                p.nu[base] = nu[s];
                dd_add(acc[0][s], acc[1][s], lam[s]);
                dd_add(acc[2][s], acc[3][s], nu[s]);
            }
        if (lane == 0) { p.nev_nu[d] = nev_nu; p.nev_lam[d] = nev_lam; }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < WCPL; ++s) {
        red[(warp * 4 + 2 * s) * 32 + lane] = make_double2(acc[0][s], acc[1][s]);
        red[(warp * 4 + 2 * s + 1) * 32 + lane] = make_double2(acc[2][s], acc[3][s]);
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int q = threadIdx.x >> 5;                  // 0: λ slot 0, 1: ν slot 0, 2: λ slot 1, 3: ν slot 1
        const int s = q >> 1, which = q & 1, j = lane + 32 * s;
        double hi = 0.0, lo = 0.0;
        for (int wv = 0; wv < 8; ++wv) dd_merge(hi, lo, red[(wv * 4 + q) * 32 + lane].x, red[(wv * 4 + q) * 32 + lane].y);
        if (j < MK) put_partial(partial + (size_t)blockIdx.x * 2 * MK + which * MK + j, hi, lo, p.accum);
    }
}

// ζ, props for the wide layout
__global__ void __launch_bounds__(256) k_zeta_props_wide(MmctmDev p, double *props_out, int want_zeta) {
    __shared__ double sh[8][WMK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = p.MK, M = p.M;
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        __syncwarp();
        for (int j = lane; j < MK; j += 32)
            sh[warp][j] = want_zeta ? det_exp(p.lam[d * MK + j] + 0.5 * p.nu[d * MK + j]) : det_exp(p.lam[d * MK + j]);
        __syncwarp();
        for (int j = lane; j < MK; j += 32) {
            int mod = 0;
            for (int m = 0; m < M; ++m)
                if (j >= p.koff[m]) mod = m;
            const double s = smem_block_sum(sh[warp], p.koff[mod], p.koff[mod + 1]);
            if (want_zeta) {
                if (j == p.koff[mod]) p.zeta[d * M + mod] = s;
            } else
                props_out[d * MK + j] = sh[warp][j] / s;
        }
    }
}

// ΣΔΔᵀ for rows [r0, r0 + 16): lane owns columns lane, lane + 32.  partial: [gridDim.x][MK*MK + M]
__global__ void __launch_bounds__(256) k_moments_wide(MmctmDev p, double2 *partial, int r0) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    __shared__ double dsh_all[8][WMK];
    __shared__ double2 red[8 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MK = p.MK;
    double *dsh = dsh_all[warp];
    double mhi[WCPL][16], mlo[WCPL][16];
#pragma unroll
    for (int s = 0; s < WCPL; ++s)
#pragma unroll
        for (int i = 0; i < 16; ++i) { mhi[s][i] = 0.0; mlo[s][i] = 0.0; }
    const long long nw = (long long)gridDim.x * 8;
    for (long long d = (long long)blockIdx.x * 8 + warp; d < p.D; d += nw) {
        double diff[WCPL];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < WCPL; ++s) {
            const int j = lane + 32 * s;
            diff[s] = j < MK ? p.lam[d * MK + j] - p.mu[j] : 0.0;
            dsh[j] = diff[s];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const double di = (r0 + i < MK) ? dsh[r0 + i] : 0.0;
#pragma unroll
            for (int s = 0; s < WCPL; ++s) dd_add(mhi[s][i], mlo[s][i], diff[s] * di);
        }
    }
    __syncthreads();
    double2 *out = partial + (size_t)blockIdx.x * (MK * MK + p.M);
#pragma unroll
    for (int s = 0; s < WCPL; ++s)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            red[warp * 32 + lane] = make_double2(mhi[s][i], mlo[s][i]);
            __syncthreads();
            const int j = lane + 32 * s;
            if (warp == 0 && j < MK && r0 + i < MK) {
                double hi = 0.0, lo = 0.0;
                for (int wv = 0; wv < 8; ++wv) dd_merge(hi, lo, red[wv * 32 + lane].x, red[wv * 32 + lane].y);
                out[j * MK + (r0 + i)] = make_double2(hi, lo);       // Σ_d diff_j * diff_{r0+i}
            }
            __syncthreads();
        }
}

}  // namespace mmsig
