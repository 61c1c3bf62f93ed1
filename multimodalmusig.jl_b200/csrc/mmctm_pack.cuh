// mmctm_pack.cuh -- k_solve for small models (ΣK_m <= 16): several samples per warp.
//
// With one coordinate per lane a model with ΣK_m = 10 (CTM, K = 10) or 14 (MMCTM([7,7])) leaves
// most of a warp idle.  Here a warp is split into NG = 32 / G groups of G lanes (G = 16, or 8 when
// ΣK_m <= 8) and every group runs its own sample.  The groups must stay in lock step to share
// instruction issue, but LD_MMA's inner / outer loops are data dependent, so the recurrence is
// flattened into a state machine: one objective evaluation per trip and per group, followed by
// predicated (group-uniform) state transitions -- accept / reject, rho growth, x-tolerance test,
// sigma update, nu -> lambda phase change, next sample.  A group that finishes early simply takes
// its next sample; nobody waits.
//
// Arithmetic is that of k_solve (DET specification): the 32-leaf tree with leaves >= G equal to
// zero is the G-leaf butterfly, so restricting the shuffles to a group changes no bit.
#pragma once
#include "mmctm_kernels.cuh"

namespace mmsig {

template <int G>
__device__ __forceinline__ double group_tree_sum(double v) {
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) v = v + shfl_xor_d(v, off);
    return v;
}
// three trees in one pass (recursive halving on the two top levels, then broadcast); see warp_tree_sum3h
template <int G>
__device__ __forceinline__ void group_tree_sum3(double &a, double &b, double &c, int lane) {
    constexpr int TOP = G / 2, SEC = G / 4;
    const bool ut = (lane & TOP) != 0, us = (lane & SEC) != 0;
    double s0 = ut ? a : c, s1 = ut ? b : 0.0;
    double k0 = ut ? c : a, k1 = ut ? 0.0 : b;
    k0 = k0 + shfl_xor_d(s0, TOP);
    k1 = k1 + shfl_xor_d(s1, TOP);
    const double s = us ? k0 : k1;
    double k = us ? k1 : k0;
    k = k + shfl_xor_d(s, SEC);
#pragma unroll
    for (int off = SEC / 2; off >= 1; off >>= 1) k = k + shfl_xor_d(k, off);
    const int gb = lane & ~(G - 1);
    a = shfl_d(k, gb);
    b = shfl_d(k, gb + SEC);
    c = shfl_d(k, gb + TOP);
}
template <int G>
__device__ __forceinline__ bool group_all(bool pred, int lane) {
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
    return (__ballot_sync(FULLMASK, pred) & gmask) == gmask;
}

constexpr int PH_IDLE = 0, PH_NU = 1, PH_LAM = 2;

template <int G>
__global__ void __launch_bounds__(256, 3) k_solve_pack(MmctmDev p, double2 *partial) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    constexpr int NG = 32 / G;                 // samples per warp
    constexpr int STRIDE = G + 2;              // padded invΣ row (doubles); (G+2)/2 odd -> conflict-free LDS.128
    __shared__ __align__(16) double ST[G * STRIDE];
    __shared__ __align__(16) double dsh_all[8][NG][G];
    __shared__ double2 red[8][2][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, gl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    const int MK = p.MK, M = p.M;
    double *dsh = dsh_all[warp][grp];
    for (int t = threadIdx.x; t < G * STRIDE; t += blockDim.x) {
        const int j = t / STRIDE, i = t % STRIDE;
        ST[t] = (i < MK && j < MK) ? p.invSigma[j * MK + i] : 0.0;
    }
    dsh[gl] = 0.0;
    __syncthreads();
    const bool active = gl < MK;
    int mod = 0;
    for (int m = 0; m < M; ++m)
        if (gl >= p.koff[m]) mod = m;
    const int blo = p.koff[mod], bhi = p.koff[mod + 1];
    const double Sjj = active ? p.invSigma[gl * MK + gl] : 0.0;
    const double muj = active ? p.mu[gl] : 0.0;
    const double2 *srow = reinterpret_cast<const double2 *>(ST + gl * STRIDE);

    // per-group state (identical in all lanes of a group unless noted "per lane")
    int phase = PH_IDLE, k = 0, nev = 0, nev_nu = 0;
    bool init = false;
    double x = 0.0, g = 0.0, xcur = 0.0, xprev = 0.0, xprevprev = 0.0, sigma = 1.0, isig = 1.0;      // per lane (isig = 1 / sigma)
    double fmin = 0.0, rho = 1.0;
    double cN = 0.0, sth = 0.0, other = 0.0, lam0 = 0.0;                                  // per lane context
    double lsh = 0.0, lsl = 0.0, nsh = 0.0, nsl = 0.0;
    long long dcur = -1;
    bool exhausted = false;

    while (true) {
        // ---- groups without work take their next sample (src/MMCTM.jl:450-453: ζ from the old λ, ν)
        if (phase == PH_IDLE && !exhausted) {
            dcur = next_sample(p.work, gmask, grp * G, gl == 0);
            if (dcur >= p.D) exhausted = true;
        }
        if (phase == PH_IDLE && !exhausted) {
            const long long base = dcur * MK + gl;
            lam0 = active ? p.lam_prev[base] : 0.0;
            const double nu0 = active ? p.nu[base] : 1.5;
            sth = active ? p.sumtheta[base] : 0.0;
            const double e0 = active ? det_exp(lam0 + 0.5 * nu0) : 0.0;
            double zeta = 0.0;
            for (int i = 0; i < MK; ++i) {
                const int hi = __shfl_sync(gmask, __double2hiint(e0), grp * G + i);
                const int lo = __shfl_sync(gmask, __double2loint(e0), grp * G + i);
                const double ei = __hiloint2double(hi, lo);
                if (i >= blo && i < bhi) zeta += ei;
            }
            const double Ndm = active ? p.N[dcur * M + mod] : 0.0;
            cN = active ? Ndm / zeta : 0.0;
            if (active && gl == blo) p.zeta[dcur * M + mod] = zeta;
            phase = PH_NU;
            init = true;
            x = nu0;
            other = lam0;
        }
        if (!__any_sync(FULLMASK, phase != PH_IDLE)) break;
        const bool busy = phase != PH_IDLE;

        // ---- propose the next point (NLopt mma.c inner iteration, m = 0 constraints)
        double xe = x, gterm = 0.0, wterm = 0.0;
        if (busy && !init) {
            const double lb = (phase == PH_NU) ? 1e-7 : -__longlong_as_double(0x7ff0000000000000LL);
            double u = g;
            const double v = fabs(g) * sigma + 0.5 * rho;
            const double sigma2 = sigma * sigma;
            u *= sigma2;
            const double qv = fast_div(u, v);
                const double r = qv * isig;                 // DET: (u / v)(1 / sigma)
            const double om = fabs(1 - r * r);
            const double sq = fast_sqrt(om < 0x1p-200 ? 0x1p-200 : om);   // om is 0 or >= 2^-53: sqrt(0) -> 2^-100, and -1 - 2^-100 == -1
            double dx = fast_div(qv, -1 - sq);
            double xc = x + dx;
            const double mv = 0.9 * sigma, xhi = x + mv, xlo = x - mv;          // move limits: selects, no branches
            xc = xc > xhi ? xhi : (xc < xlo ? xlo : xc);
            if (xc < lb) xc = lb;
            if (!active) xc = x;
            dx = xc - x;
            const double dx2 = dx * dx;
            const double denominv = fast_rcp(sigma2 - dx2);       // |dx| <= 0.9 sigma
            const double cc = sigma2 * dx;
            gterm = (g * cc + (fabs(g) * sigma + 0.5 * rho) * dx2) * denominv;
            wterm = 0.5 * dx2 * denominv;
            xe = xc;
        }
        // ---- lane-local part of the objective at xe (src/common.jl:11-36)
        double t = 0.0, gcur = 1.0;
        if (phase == PH_NU) {
            const double e = det_exp(other + 0.5 * xe);
            const double grad = (-0.5 * Sjj - (cN / 2) * e) + fast_rcp(2 * xe);
            t = (-0.5 * (xe * Sjj) - cN * e) + det_log(xe) / 2;
            gcur = -grad;
        } else if (phase == PH_LAM) {
            const double diff = xe - muj;
            const double e = det_exp(xe + other);
            dsh[gl] = active ? diff : 0.0;
            __syncwarp(gmask);
            double q = 0.0, qo = 0.0;
            const double2 *dv2 = reinterpret_cast<const double2 *>(dsh);
#pragma unroll
            for (int i = 0; i < G / 2; ++i) {
                const double2 sv = srow[i], dv = dv2[i];
                q = fma(sv.x, dv.x, q);                 // DET: even / odd index chains, then one add
                qo = fma(sv.y, dv.y, qo);
            }
            q = q + qo;
            __syncwarp(gmask);
            const double ce = cN * e;
            const double grad = (-q + sth) - ce;
            const double a = q * diff, b = xe * sth;
            t = (b - 0.5 * a) - ce;
            gcur = -grad;
        }
        if (!active || !busy) { t = 0.0; gcur = 1.0; }
        __syncwarp();
        // ---- the three group sums in one pass
        group_tree_sum3<G>(gterm, wterm, t, lane);
        const double f = -t;
        // x-tolerance quantities of (xe, xprev), needed only when an inner loop ends
        const double ad = (active && busy && !init) ? fabs(xe - xprev) : 0.0;
        double dn = ad, xn = (active && busy) ? fabs(xe) : 0.0;
        const double gval = fmin + gterm;
        const bool inner_done = !init && (gval >= f);
        if (__any_sync(FULLMASK, busy && inner_done)) {
            dn = group_tree_sum<G>(dn);
            xn = group_tree_sum<G>(xn);
        }
        // group votes for the x-tolerance rules, taken by the whole warp outside divergent code
        const bool ok26 = ad < 1e-4 || ad < 1e-4 * (fabs(xe) + fabs(xprev)) * 0.5 || xe == xprev;
        const bool all26 = group_all<G>(ok26 || !active, lane);
        const bool allabs = group_all<G>(!(ad > 1e-4), lane);
        // ---- state transitions (group-uniform predicates)
        bool finish = false;
        if (busy) {
            if (init) {
                fmin = f;
                g = gcur;
                nev = 1;
                xcur = xprev = xprevprev = x;
                k = 1;
                rho = 1.0;
                sigma = 1.0;
                isig = 1.0;
                init = false;
            } else {
                xcur = xe;
                ++nev;
                if (f < fmin) { fmin = f; x = xcur; g = gcur; }
                if (nev >= MMA_MAXEVAL) finish = true;
                else if (inner_done) {
                    const bool stop = (p.stop_rule == 1) ? all26 : ((dn <= 1e-4 * xn) || allabs);
                    if (stop) finish = true;
                    else {
                        rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;
                        if (k > 1) {
                            const double s2 = (xcur - xprev) * (xprev - xprevprev);
                            sigma *= s2 < 0 ? 0.7 : (s2 > 0 ? 1.2 : 1.0);
                            isig = fast_rcp(sigma);
                        }
                        ++k;
                        xprevprev = xprev;
                        xprev = xcur;
                    }
                } else if (f > gval) {
                    const double r1 = 10 * rho, r2 = 1.1 * (rho + guarded_div(f - gval, wterm));
                    rho = r1 < r2 ? r1 : r2;
                }
            }
        }
        if (finish) {
            if (phase == PH_NU) {                 // ν done: λ with the new ν, old ζ (src/MMCTM.jl:454)
                nev_nu = nev;
                const double nu_new = x;
                if (active) { p.nu[dcur * MK + gl] = nu_new; dd_add(nsh, nsl, nu_new); }
                other = 0.5 * nu_new;
                x = lam0;
                phase = PH_LAM;
                init = true;
            } else {
                if (active) { p.lam[dcur * MK + gl] = x; dd_add(lsh, lsl, x); }
                if (gl == 0) { p.nev_nu[dcur] = nev_nu; p.nev_lam[dcur] = nev; }
                phase = PH_IDLE;
            }
        }
    }
    red[warp][0][lane] = make_double2(lsh, lsl);
    red[warp][1][lane] = make_double2(nsh, nsl);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int which = threadIdx.x >> 5;
        if (lane < MK) {
            double hi = 0.0, lo = 0.0;
            for (int wv = 0; wv < 8; ++wv)
                for (int gg = 0; gg < NG; ++gg) dd_merge(hi, lo, red[wv][which][gg * G + lane].x, red[wv][which][gg * G + lane].y);
            put_partial(partial + (size_t)blockIdx.x * 2 * MK + which * MK + lane, hi, lo, p.accum);
        }
    }
}

}  // namespace mmsig
