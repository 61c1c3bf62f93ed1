// mmctm_split.cuh -- EXPERIMENTAL (compiled only with make EXP=1 -> libmmsig_exp.so; MMSIG_SOLVE=split | split16; not yet measured): the E-step's two
// LD_MMA solves as two kernels, update_ν! (src/MMCTM.jl:156-170) for every sample, then update_λ!
// (:127-143), each with four samples per warp and CPL = 3 or 4 coordinates per lane as k_solve_multi
// (mmctm_pack.cuh).
//
// Why: k_solve_multi fills every lane and shares each reduction between four samples, but its four groups
// sit in different phases most of the time, so a warp pays for the ν evaluation AND the λ evaluation on
// nearly every trip (negative results).  With one phase per kernel the only
// divergence left is the predicated bookkeeping of LD_MMA's inner / outer loop; the ν kernel also drops the
// invΣ rows, μ and sumθ from its working set.  Between the kernels ν and ζ travel through memory (they are
// outputs of the E-step anyway): 8 (ΣK + M) bytes per sample.
//
// Arithmetic is k_solve's (DET specification) operation for operation, so results are bit-identical.
#pragma once
#include "mmctm_pack.cuh"

namespace mmsig {

// slots of one lane: coordinates gl, gl + G, gl + 2G, ...; their sum in the order of the 32-leaf tree's top levels
template <int CPL>
__device__ __forceinline__ double slot_sum_g(const double (&v)[CPL]) {
    if (CPL == 2) return v[0] + v[1];                    // G = 16: level 16
    if (CPL == 3) return (v[0] + v[2]) + v[1];           // G = 8: levels 16, 8 (slot 3 is absent: + 0 exactly)
    return (v[0] + v[2]) + (v[1] + v[CPL - 1]);          // CPL == 4
}

// G lanes per sample (32 / G samples per warp), CPL coordinates per lane (coordinate j on lane j % G, slot j / G);
// (G, CPL) = (8, 3), (8, 4): four samples per warp; (16, 2): two samples per warp, fewer registers.
template <int G, int CPL, int PH>
__global__ void __launch_bounds__(128, (G == 16 ? 4 : MULTI_MIN_BLOCKS)) k_solve_phase(MmctmDev p, double2 *partial) {
    constexpr bool NU = PH == PH_NU;
    constexpr int NG = 32 / G, MKP = G * CPL, STRIDE = 34, NW = 4;       // NW warps per block
    static_assert(MKP <= 32 && (G == 8 || G == 16), "one 32-leaf tree per sample");
    __shared__ __align__(16) double ST[NU ? 2 : 32 * STRIDE];            // invΣ rows: λ phase only
    __shared__ __align__(16) double dsh_all[NW][NG][STRIDE];
    __shared__ double2 red[NW][CPL][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;
    const unsigned gmask = ((1u << G) - 1u) << (grp * G);
    const int MK = p.MK, M = p.M;
    double *dsh = dsh_all[warp][grp];
    if (!NU)
        for (int t = threadIdx.x; t < 32 * STRIDE; t += blockDim.x) {
            const int j = t / STRIDE, i = t % STRIDE;
            ST[t] = (i < MK && j < MK) ? p.invSigma[j * MK + i] : 0.0;
        }
    for (int i = gl; i < STRIDE; i += G) dsh[i] = 0.0;
    __syncthreads();
    bool active[CPL];
    int mod[CPL], blo[CPL], bhi[CPL];
    double Sjj[CPL], muj[CPL];
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
        const int j = gl + G * s;
        active[s] = j < MK;
        mod[s] = 0;
        for (int m = 0; m < M; ++m)
            if (j >= p.koff[m]) mod[s] = m;
        blo[s] = p.koff[mod[s]];
        bhi[s] = p.koff[mod[s] + 1];
        Sjj[s] = active[s] ? p.invSigma[j * MK + j] : 0.0;
        muj[s] = active[s] ? p.mu[j] : 0.0;
    }
    bool busy = false, init = false;
    int k = 0, nev = 0;
    double x[CPL], g[CPL], xcur[CPL], xprev[CPL], xprevprev[CPL], sigma[CPL];
    double cN[CPL], sth[CPL], other[CPL];
    double acch[CPL], accl[CPL];                     // Σ of this phase's result (dd)
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
        x[s] = xcur[s] = xprev[s] = xprevprev[s] = NU ? 1.5 : 0.0;     // a group that never gets a sample evaluates a harmless point
        g[s] = 0.0;
        sigma[s] = 1.0;
        cN[s] = sth[s] = other[s] = 0.0;
        acch[s] = accl[s] = 0.0;
    }
    double fmin = 0.0, rho = 1.0;
    long long dcur = -1;
    bool exhausted = false;
    const double lb = NU ? 1e-7 : -__longlong_as_double(0x7ff0000000000000LL);

    while (true) {
        if (!busy && !exhausted) {
            dcur = next_sample(p.work, gmask, grp * G, gl == 0);
            if (dcur >= p.D) exhausted = true;
        }
        if (!busy && !exhausted) {
            if (NU) {
                // ζ from the old λ, ν (src/MMCTM.jl:450-452), then ν's problem
                double nu0[CPL], lam0[CPL];
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const long long base = dcur * MK + gl + G * s;
                    lam0[s] = active[s] ? p.lam_prev[base] : 0.0;
                    nu0[s] = active[s] ? p.nu[base] : 1.5;
                    dsh[gl + G * s] = active[s] ? det_exp(lam0[s] + 0.5 * nu0[s]) : 0.0;
                }
                __syncwarp(gmask);
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    double zeta = 0.0;
                    for (int i = blo[s]; i < bhi[s]; ++i) zeta += dsh[i];
                    const double Ndm = active[s] ? p.N[dcur * M + mod[s]] : 0.0;
                    cN[s] = active[s] ? Ndm / zeta : 0.0;
                    if (active[s] && gl + G * s == blo[s]) p.zeta[dcur * M + mod[s]] = zeta;
                    x[s] = nu0[s];
                    other[s] = lam0[s];
                }
                __syncwarp(gmask);
            } else {
                // λ's problem: the new ν, the old ζ (both written by the ν kernel), the old sumθ (:454)
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const long long base = dcur * MK + gl + G * s;
                    x[s] = active[s] ? p.lam_prev[base] : 0.0;
                    other[s] = active[s] ? 0.5 * p.nu[base] : 0.75;
                    sth[s] = active[s] ? p.sumtheta[base] : 0.0;
                    const double Ndm = active[s] ? p.N[dcur * M + mod[s]] : 0.0;
                    const double zeta = active[s] ? p.zeta[dcur * M + mod[s]] : 1.0;
                    cN[s] = active[s] ? Ndm / zeta : 0.0;
                }
            }
            busy = true;
            init = true;
        }
        if (!__any_sync(FULLMASK, busy)) break;

        // ---- propose the next point (NLopt mma.c inner iteration, m = 0 constraints)
        double xe[CPL], gl_[CPL], wl_[CPL];
#pragma unroll
        for (int s = 0; s < CPL; ++s) { xe[s] = x[s]; gl_[s] = 0.0; wl_[s] = 0.0; }
        if (busy && !init) {
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                double u = g[s];
                const double v = fabs(g[s]) * sigma[s] + 0.5 * rho;
                const double sigma2 = sigma[s] * sigma[s];
                u *= sigma2;
                const double qv = fast_div(u, v);
                const double r = qv * fast_rcp(sigma[s]);      // DET: (u / v)(1 / sigma)
                const double om = fabs(1 - r * r);
                const double sq = fast_sqrt(om < 0x1p-200 ? 0x1p-200 : om);   // om is 0 or >= 2^-53: sqrt(0) -> 2^-100, and -1 - 2^-100 == -1
                double dx = fast_div(qv, -1 - sq);
                double xc = x[s] + dx;
                const double mv = 0.9 * sigma[s], xhi = x[s] + mv, xlo = x[s] - mv;
                xc = xc > xhi ? xhi : (xc < xlo ? xlo : xc);
                if (xc < lb) xc = lb;
                if (!active[s]) xc = x[s];
                dx = xc - x[s];
                const double dx2 = dx * dx;
                const double denominv = fast_rcp(sigma2 - dx2);       // |dx| <= 0.9 sigma
                const double cc = sigma2 * dx;
                gl_[s] = (g[s] * cc + (fabs(g[s]) * sigma[s] + 0.5 * rho) * dx2) * denominv;
                wl_[s] = 0.5 * dx2 * denominv;
                xe[s] = xc;
            }
        }
        // ---- lane-local part of the objective at xe (src/common.jl:11-36)
        double tl[CPL], gcur[CPL];
        if (NU) {
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const double e = det_exp(other[s] + 0.5 * xe[s]);
                const double grad = (-0.5 * Sjj[s] - (cN[s] / 2) * e) + fast_rcp(2 * xe[s]);
                tl[s] = (-0.5 * (xe[s] * Sjj[s]) - cN[s] * e) + det_log(xe[s]) / 2;
                gcur[s] = -grad;
            }
        } else {
            double diff[CPL];
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                diff[s] = xe[s] - muj[s];
                dsh[gl + G * s] = active[s] ? diff[s] : 0.0;
            }
            __syncwarp(gmask);
            const double2 *dv2 = reinterpret_cast<const double2 *>(dsh);
            double q[CPL], qo[CPL];                     // DET: even / odd index chains, then one add
#pragma unroll
            for (int s = 0; s < CPL; ++s) { q[s] = 0.0; qo[s] = 0.0; }
#pragma unroll 4
            for (int i = 0; i < MKP / 2; ++i) {
                const double2 dv = dv2[i];
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const double2 sv = reinterpret_cast<const double2 *>(ST + (gl + G * s) * STRIDE)[i];
                    q[s] = fma(sv.x, dv.x, q[s]);
                    qo[s] = fma(sv.y, dv.y, qo[s]);
                }
            }
#pragma unroll
            for (int s = 0; s < CPL; ++s) q[s] = q[s] + qo[s];
            __syncwarp(gmask);
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const double e = det_exp(xe[s] + other[s]);
                const double ce = cN[s] * e;
                const double grad = (-q[s] + sth[s]) - ce;
                const double a = q[s] * diff[s], b = xe[s] * sth[s];
                tl[s] = (b - 0.5 * a) - ce;
                gcur[s] = -grad;
            }
        }
        double adl[CPL], xnl[CPL];
        bool ok26 = true, okabs = true;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            if (!active[s] || !busy) { tl[s] = 0.0; gcur[s] = 1.0; }
            adl[s] = (active[s] && busy && !init) ? fabs(xe[s] - xprev[s]) : 0.0;
            xnl[s] = (active[s] && busy) ? fabs(xe[s]) : 0.0;
            ok26 = ok26 && (!active[s] || adl[s] < 1e-4 || adl[s] < 1e-4 * (fabs(xe[s]) + fabs(xprev[s])) * 0.5 ||
                            xe[s] == xprev[s]);
            okabs = okabs && !(adl[s] > 1e-4);
        }
        __syncwarp();
        double gterm = slot_sum_g<CPL>(gl_), wterm = slot_sum_g<CPL>(wl_), t = slot_sum_g<CPL>(tl);
        group_tree_sum3<G>(gterm, wterm, t, lane);
        const double f = -t;
        const double gval = fmin + gterm;
        const bool inner_done = !init && (gval >= f);
        double dn = slot_sum_g<CPL>(adl), xn = slot_sum_g<CPL>(xnl);
        if (__any_sync(FULLMASK, busy && inner_done)) {
            dn = group_tree_sum<G>(dn);
            xn = group_tree_sum<G>(xn);
        }
        const bool all26 = group_all<G>(ok26, lane);
        const bool allabs = group_all<G>(okabs, lane);
        // ---- state transitions (group-uniform predicates)
        bool finish = false;
        if (busy) {
            if (init) {
                fmin = f;
                nev = 1;
                k = 1;
                rho = 1.0;
                init = false;
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    g[s] = gcur[s];
                    xcur[s] = xprev[s] = xprevprev[s] = x[s];
                    sigma[s] = 1.0;
                }
            } else {
                ++nev;
                const bool better = f < fmin;
                if (better) fmin = f;
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    xcur[s] = xe[s];
                    if (better) { x[s] = xe[s]; g[s] = gcur[s]; }
                }
                if (nev >= MMA_MAXEVAL) finish = true;
                else if (inner_done) {
                    const bool stop = (p.stop_rule == 1) ? all26 : ((dn <= 1e-4 * xn) || allabs);
                    if (stop) finish = true;
                    else {
                        rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            if (k > 1) {
                                const double s2 = (xcur[s] - xprev[s]) * (xprev[s] - xprevprev[s]);
                                sigma[s] *= s2 < 0 ? 0.7 : (s2 > 0 ? 1.2 : 1.0);
                            }
                            xprevprev[s] = xprev[s];
                            xprev[s] = xcur[s];
                        }
                        ++k;
                    }
                } else if (f > gval) {
                    const double r1 = 10 * rho, r2 = 1.1 * (rho + guarded_div(f - gval, wterm));
                    rho = r1 < r2 ? r1 : r2;
                }
            }
        }
        if (finish) {
            double *dst = NU ? p.nu : p.lam;
#pragma unroll
            for (int s = 0; s < CPL; ++s)
                if (active[s]) { dst[dcur * MK + gl + G * s] = x[s]; dd_add(acch[s], accl[s], x[s]); }
            if (gl == 0) (NU ? p.nev_nu : p.nev_lam)[dcur] = nev;
            busy = false;
        }
    }
#pragma unroll
    for (int s = 0; s < CPL; ++s) red[warp][s][lane] = make_double2(acch[s], accl[s]);
    __syncthreads();
    // coordinate j = gl + G s: sum over warps and over the four groups; partial: [grid][2 MK], Σλ then Σν
    for (int j = threadIdx.x; j < MK; j += blockDim.x) {
        const int s = j / G, l = j % G;
        double hi = 0.0, lo = 0.0;
        for (int wv = 0; wv < NW; ++wv)
            for (int gg = 0; gg < NG; ++gg) {
                const double2 v = red[wv][s][gg * G + l];
                dd_merge(hi, lo, v.x, v.y);
            }
        put_partial(partial + (size_t)blockIdx.x * 2 * MK + (NU ? MK : 0) + j, hi, lo, p.accum);
    }
}

}  // namespace mmsig
