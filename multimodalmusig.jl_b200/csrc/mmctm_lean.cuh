// mmctm_lean.cuh -- the E-step's two LD_MMA solves for sum(K) <= 32: update_ν! (src/MMCTM.jl:156-170)
// for every sample, then update_λ! (:127-143), one kernel per phase, G lanes per sample and CPL
// coordinates per lane (coordinate j on lane j % G of its group, slot j / G), 32 / G samples per warp.
//
// Why this shape (measured, profiles/r02_solve_ab.md): k_solve (one sample per warp, lane = coordinate)
// spends 390 warp-instructions per objective evaluation, 36 % of them FP64, on 24 of 32 lanes at
// sum(K) = 24; it is bound by instruction issue, not by the FP64 pipe.  Packing a sample into 8 (or 4)
// lanes fills every lane, shares each reduction and each piece of scalar LD_MMA bookkeeping between 4
// (or 8) samples and gives every lane CPL independent dependency chains.  One phase per kernel keeps
// the samples of a warp in the same objective.  LD_MMA's data-dependent inner / outer loops are
// flattened into ONE objective evaluation per trip with select-only state transitions: there is no
// separate "first evaluation" path (the first trip proposes from a zero gradient, i.e. stays at x0),
// no per-trip copies of the iterate (x, g change by selects; xprev / xprevprev / sigma only on an outer
// step) and the only branches are warp-uniform (any group needs a sample / ends an inner loop).
//
// Arithmetic is k_solve's (DET specification) operation for operation; the 32-leaf
// tree sum becomes: slots of a lane in the order of the tree's top levels, then the G-lane butterfly.
// Results are bit-identical to k_solve, k_solve_pack and the oracle, evaluation counts included.
#pragma once
#include "mmctm_pack.cuh"

namespace mmsig {

#ifndef LEAN_MIN_BLOCKS
#define LEAN_MIN_BLOCKS 4          // 126 registers, no spills; measured 20.8 ms against 23.4 ms at 3 (160 registers) and 20.4 ms at 5 (96, spills)
#endif

// sum of a lane's CPL slot values in the order of the 32-leaf tree's top levels: slot s holds leaf
// gl + G s, so tree level 16 / G pairs slot s with s ^ (NS / 2), the next level with s ^ (NS / 4), ...
// (an absent slot is an exact + 0)
template <int G, int CPL>
__device__ __forceinline__ double lean_slot_sum(const double (&v)[CPL]) {
    constexpr int NS = 32 / G;
    double w[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) w[s] = s < CPL ? v[s] : 0.0;
#pragma unroll
    for (int off = NS / 2; off >= 1; off >>= 1)
#pragma unroll
        for (int s = 0; s < off; ++s)
            if (s + off < CPL) w[s] = w[s] + w[s + off];
    return w[0];
}

template <int G>
__device__ __forceinline__ double lean_shfl_xor(double v, int off) {
    return shfl_xor_d(v, off);
}

// the three group sums of one trip; every lane of a group ends with the same bits
template <int G>
__device__ __forceinline__ void lean_sum3(double &a, double &b, double &c, int lane) {
    if (G >= 8) {
        group_tree_sum3<G>(a, b, c, lane);
    } else {
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            const double ta = shfl_xor_d(a, off), tb = shfl_xor_d(b, off), tc = shfl_xor_d(c, off);
            a = a + ta;
            b = b + tb;
            c = c + tc;
        }
    }
}

#ifndef LEAN_MV_UNROLL
#define LEAN_MV_UNROLL 4
#endif
constexpr int kLeanMvUnroll = LEAN_MV_UNROLL;

// asynchronous 8-byte copy global -> shared (LDGSTS): the next sample's inputs travel while the current one is solved
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gmem_src) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// dynamic shared memory of k_solve_lean<G, CPL, PH> (doubles): invΣ rows (λ phase) | per-group Δ vectors | dd sums of
// the results | per-sample context (c, other, [sumθ], xprevprev) | staging of the next sample's inputs
template <int G, int CPL, int PH>
constexpr size_t lean_smem_doubles() {
    constexpr bool NU = PH == PH_NU;
    return (NU ? 0 : G * CPL * 34) + 4 * (32 / G) * 34 + 2 * 4 * CPL * 32 + (NU ? 3 : 4) * CPL * 128 + (NU ? 3 : 5) * CPL * 128;
}

// FULL: sum(K) == G * CPL, no padding coordinate -- every "active" mask is compile-time true
template <int G, int CPL, int PH, bool FULL>
__global__ void __launch_bounds__(128, LEAN_MIN_BLOCKS) k_solve_lean(MmctmDev p, double2 *partial) {
    if (p.ctl && p.ctl[0]) return;        // an earlier iteration of this batch met the convergence rule (mmctm_run_iterations)
    constexpr bool NU = PH == PH_NU;
    constexpr int NG = 32 / G, MKP = G * CPL, STRIDE = 34, NW = 4;       // NW warps per block
    constexpr int NSTG = NU ? 3 : 5;                                     // staged inputs per coordinate: ν: λ, ν, N ; λ: λ, ν, sumθ, N, ζ
    static_assert(MKP <= 32 && (G == 4 || G == 8 || G == 16), "one 32-leaf tree per sample");
    extern __shared__ __align__(16) double lean_smem[];
    double *ST = lean_smem;                                              // invΣ rows: λ phase only
    double *dsh_base = ST + (NU ? 0 : MKP * STRIDE);                     // [NW][NG][STRIDE]
    double2 *red = reinterpret_cast<double2 *>(dsh_base + NW * NG * STRIDE);   // [NW][CPL][32]: Σ of this phase's result per lane and slot (dd)
    // per-sample values read once per trip (or less) live in shared memory, not in registers: the kernel is bound by
    // fixed-latency stalls, and <= 128 registers admit four to five blocks per SM (ncu: profiles/r02_solve_summary.md)
    double *ctx_c = reinterpret_cast<double *>(red + NW * CPL * 32);     // [CPL][128] N_dm / ζ_dm
    double *ctx_o = ctx_c + CPL * 128;                                   // ν: λ_j ; λ: ν_j / 2
    double *ctx_xpp = ctx_o + CPL * 128;                                 // xprevprev
    double *ctx_s = ctx_xpp + CPL * 128;                                 // λ: sumθ_j
    double *stage = ctx_s + (NU ? 0 : CPL * 128);                        // [NSTG][CPL][128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int grp = lane / G, gl = lane % G;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (grp * G));
    const int MK = p.MK, M = p.M;
    double *dsh = dsh_base + (warp * NG + grp) * STRIDE;
    if (!NU)
        for (int t = threadIdx.x; t < MKP * STRIDE; t += blockDim.x) {
            const int j = t / STRIDE, i = t % STRIDE;
            ST[t] = (i < MK && j < MK) ? p.invSigma[j * MK + i] : 0.0;
        }
    for (int i = gl; i < STRIDE; i += G) dsh[i] = 0.0;
    __syncthreads();
    constexpr bool full = FULL;
    bool active[CPL];
    int mod[CPL];
    double cst[CPL];                                  // ν: -0.5 invΣ_jj ; λ: μ_j
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
        const int j = gl + G * s;
        active[s] = FULL || j < MK;
        mod[s] = 0;
        for (int m = 0; m < M; ++m)
            if (j >= p.koff[m]) mod[s] = m;
        const double Sjj = active[s] ? p.invSigma[j * MK + j] : 0.0;
        cst[s] = NU ? -0.5 * Sjj : (active[s] ? p.mu[j] : 0.0);
        red[(warp * CPL + s) * 32 + lane] = make_double2(0.0, 0.0);
    }
    // iterate (per lane and slot) and LD_MMA scalars (identical in the lanes of a group)
    double x[CPL], g[CPL], xp[CPL], sig[CPL], isig[CPL];
#pragma unroll
    for (int s = 0; s < CPL; ++s) {
        x[s] = xp[s] = NU ? 1.5 : cst[s];
        g[s] = 0.0;
        sig[s] = isig[s] = 1.0;
    }
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double fmin = INF, rho = 1.0;
    int k = 1, nev = 0;
    bool first = false, need = true, alive = true;
    const double lb = NU ? 1e-7 : -INF;
    const int stop_rule = p.stop_rule;

    // Sample pipeline of a group: dcur is being solved, dstage has its inputs in flight to (or already in) the
    // staging slots, `ticket` is the atomic counter value this group's leader drew at the previous refill (read at
    // the next one, so neither the atomic nor the loads are ever waited for while a solve could run).
    auto stage_issue = [&](long long d) {
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            if (!active[s]) continue;
            const long long base = d * MK + gl + G * s;
            cp_async8(stage + (0 * CPL + s) * 128 + tid, p.lam_prev + base);
            cp_async8(stage + (1 * CPL + s) * 128 + tid, p.nu + base);
            cp_async8(stage + (2 * CPL + s) * 128 + tid, p.N + d * M + mod[s]);
            if (!NU) {
                cp_async8(stage + (3 * CPL + s) * 128 + tid, p.sumtheta + base);
                cp_async8(stage + (4 * CPL + s) * 128 + tid, p.zeta + d * M + mod[s]);
            }
        }
        cp_async_commit();
    };
    long long dcur = -1;
    long long dstage = next_sample(p.work, gmask, grp * G, gl == 0);
    if (dstage < p.D) stage_issue(dstage);
    unsigned long long ticket = 0;
    if (gl == 0) ticket = atomicAdd(p.work, 1ULL);

    while (true) {
        // ---- groups whose solve ended take their next sample (warp-uniform branch)
        if (__any_sync(FULLMASK, need)) {
            if (need) {
                if (alive) {
                    dcur = dstage;
                    if (dcur >= p.D) alive = false;
                }
                if (alive) {
                    cp_async_wait_all();                       // this lane's staged inputs of dcur have landed
                    if (NU) {
                        // ζ from the old λ, ν (src/MMCTM.jl:450-452), then ν's problem
                        double nu0[CPL];
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const double lam0 = active[s] ? stage[(0 * CPL + s) * 128 + tid] : 0.0;
                            nu0[s] = active[s] ? stage[(1 * CPL + s) * 128 + tid] : 1.5;
                            ctx_o[s * 128 + tid] = lam0;
                            dsh[gl + G * s] = active[s] ? det_exp(lam0 + 0.5 * nu0[s]) : 0.0;
                        }
                        __syncwarp(gmask);
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            const int blo = p.koff[mod[s]], bhi = p.koff[mod[s] + 1];
                            double zeta = 0.0;
                            for (int i = blo; i < bhi; ++i) zeta += dsh[i];
                            const double Ndm = active[s] ? stage[(2 * CPL + s) * 128 + tid] : 0.0;
                            ctx_c[s * 128 + tid] = active[s] ? Ndm / zeta : 0.0;
                            if (active[s] && gl + G * s == blo) p.zeta[dcur * M + mod[s]] = zeta;
                            x[s] = nu0[s];
                        }
                        __syncwarp(gmask);
                    } else {
                        // λ's problem: the new ν, the old ζ (both written by the ν kernel), the old sumθ (:454)
#pragma unroll
                        for (int s = 0; s < CPL; ++s) {
                            x[s] = active[s] ? stage[(0 * CPL + s) * 128 + tid] : 0.0;
                            ctx_o[s * 128 + tid] = active[s] ? 0.5 * stage[(1 * CPL + s) * 128 + tid] : 0.0;
                            ctx_s[s * 128 + tid] = active[s] ? stage[(3 * CPL + s) * 128 + tid] : 0.0;
                            const double Ndm = active[s] ? stage[(2 * CPL + s) * 128 + tid] : 0.0;
                            const double zeta = active[s] ? stage[(4 * CPL + s) * 128 + tid] : 1.0;
                            ctx_c[s * 128 + tid] = active[s] ? Ndm / zeta : 0.0;
                        }
                    }
                    first = true;
                    fmin = 0.0;
                } else {
                    // no sample left: a fixed point of the state machine (zero gradient: the proposal never moves;
                    // fmin = +inf: every trip "ends an inner loop" and "stops", and a dead group's stop is ignored)
#pragma unroll
                    for (int s = 0; s < CPL; ++s) {
                        x[s] = NU ? 1.5 : cst[s];
                        ctx_c[s * 128 + tid] = 0.0;
                        ctx_o[s * 128 + tid] = 0.0;
                        if (!NU) ctx_s[s * 128 + tid] = 0.0;
                    }
                    first = false;
                    fmin = INF;
                }
                // refill the pipeline: the ticket drawn at the previous refill becomes the staged sample, a new ticket is drawn
                {
                    const unsigned hi = (unsigned)__shfl_sync(gmask, (int)(ticket >> 32), grp * G);
                    const unsigned lo = (unsigned)__shfl_sync(gmask, (int)(ticket & 0xffffffffu), grp * G);
                    dstage = (long long)(((unsigned long long)hi << 32) | lo);
                    if (alive && dstage < p.D) stage_issue(dstage);
                    if (alive && gl == 0) ticket = atomicAdd(p.work, 1ULL);
                }
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    g[s] = 0.0;
                    xp[s] = x[s];
                    ctx_xpp[s * 128 + tid] = x[s];
                    sig[s] = isig[s] = 1.0;
                }
                rho = 1.0;
                k = 1;
                nev = 0;
                need = false;
            }
            if (!__any_sync(FULLMASK, alive)) break;
        }

        // ---- propose the next point (NLopt mma.c inner iteration, m = 0 constraints).  On a solve's first
        // trip g = 0, so the proposal is x itself and the trip evaluates the starting point.
        double xe[CPL], gt[CPL], wt[CPL];
        const double hr = 0.5 * rho;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            double u = g[s];
            const double v = fabs(g[s]) * sig[s] + hr;
            const double sigma2 = sig[s] * sig[s];
            u *= sigma2;
            const double qv = fast_div(u, v);
            const double r = qv * isig[s];                  // DET: (u / v)(1 / sigma)
            const double om = fabs(1 - r * r);
            const double sq = fast_sqrt(om < 0x1p-200 ? 0x1p-200 : om);   // om is 0 or >= 2^-53: sqrt(0) -> 2^-100, and -1 - 2^-100 == -1
            double dx = fast_div(qv, -1 - sq);
            double xc = x[s] + dx;
            const double mv = 0.9 * sig[s], xhi = x[s] + mv, xlo = x[s] - mv;
            xc = xc > xhi ? xhi : (xc < xlo ? xlo : xc);
            if (NU) xc = (!first && xc < lb) ? lb : xc;          // the starting point is evaluated as given
            dx = xc - x[s];
            const double dx2 = dx * dx;
            const double denominv = fast_rcp(sigma2 - dx2);       // |dx| <= 0.9 sigma
            const double cc = sigma2 * dx;
            gt[s] = (g[s] * cc + v * dx2) * denominv;
            wt[s] = 0.5 * dx2 * denominv;
            xe[s] = xc;
        }
        // ---- lane-local part of the objective at xe (src/common.jl:11-36)
        double tl[CPL], gcur[CPL];
        if (NU) {
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                // src/common.jl:25-36 with the exact power-of-two scalings folded: (c / 2) e == (c e) / 2 and
                // -0.5 (x S_jj) == x (-0.5 S_jj) bit for bit, so one product serves value and gradient
                const double e = det_exp(ctx_o[s * 128 + tid] + 0.5 * xe[s]);
                const double ce = ctx_c[s * 128 + tid] * e;
                const double grad = fma(-0.5, ce, cst[s]) + fast_rcp(2 * xe[s]);
                tl[s] = (xe[s] * cst[s] - ce) + det_log(xe[s]) / 2;
                gcur[s] = -grad;
            }
        } else {
            double diff[CPL];
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                diff[s] = xe[s] - cst[s];
                dsh[gl + G * s] = (full || active[s]) ? diff[s] : 0.0;
            }
            __syncwarp();
            const double2 *dv2 = reinterpret_cast<const double2 *>(dsh);
            double q[CPL], qo[CPL];                     // DET: even / odd index chains, then one add
#pragma unroll
            for (int s = 0; s < CPL; ++s) { q[s] = 0.0; qo[s] = 0.0; }
#pragma unroll kLeanMvUnroll
            for (int i = 0; i < MKP / 2; ++i) {
                const double2 dv = dv2[i];
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    const double2 sv = reinterpret_cast<const double2 *>(ST + (gl + G * s) * STRIDE)[i];
                    q[s] = fma(sv.x, dv.x, q[s]);
                    qo[s] = fma(sv.y, dv.y, qo[s]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int s = 0; s < CPL; ++s) {
                const double qq = q[s] + qo[s];
                const double e = det_exp(xe[s] + ctx_o[s * 128 + tid]);
                const double ce = ctx_c[s * 128 + tid] * e;
                const double sth = ctx_s[s * 128 + tid];
                const double grad = (-qq + sth) - ce;
                const double a = qq * diff[s], b = xe[s] * sth;
                tl[s] = (b - 0.5 * a) - ce;
                gcur[s] = -grad;
            }
        }
        double adl[CPL], xnl[CPL];
        bool ok26 = true, okabs = true;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            if (!full && !active[s]) { tl[s] = 0.0; gcur[s] = 0.0; }
            const double ax = fabs(xe[s]);
            adl[s] = (full || active[s]) ? fabs(xe[s] - xp[s]) : 0.0;
            xnl[s] = (full || active[s]) ? ax : 0.0;
            okabs = okabs && !(adl[s] > 1e-4);
            if (stop_rule == 1)
                ok26 = ok26 && (!active[s] || adl[s] < 1e-4 || adl[s] < 1e-4 * (ax + fabs(xp[s])) * 0.5 || xe[s] == xp[s]);
        }
        // ---- the three group sums in one pass
        double gterm = lean_slot_sum<G, CPL>(gt), wterm = lean_slot_sum<G, CPL>(wt), t = lean_slot_sum<G, CPL>(tl);
        lean_sum3<G>(gterm, wterm, t, lane);
        const double f = -t;
        const double gval = fmin + gterm;
        const bool inner_done = !first && (gval >= f);
        // x-tolerance (NLopt stop.c) on (xcur, xprev), needed only when an inner loop ends
        bool stop = false;
        if (__any_sync(FULLMASK, inner_done)) {
            if (stop_rule == 1) {
                stop = group_all<G>(ok26, lane);
            } else {
                double dn = lean_slot_sum<G, CPL>(adl), xn = lean_slot_sum<G, CPL>(xnl);
#pragma unroll
                for (int off = G / 2; off >= 1; off >>= 1) {
                    const double ta = shfl_xor_d(dn, off), tb = shfl_xor_d(xn, off);
                    dn = dn + ta;
                    xn = xn + tb;
                }
                const bool allabs = group_all<G>(okabs, lane);      // a full-mask vote: every lane takes it (no short circuit)
                stop = (dn <= 1e-4 * xn) || allabs;
            }
        }
        // ---- state transitions: selects on group-uniform predicates
        const bool better = alive && (first || f < fmin);
        ++nev;
        fmin = better ? f : fmin;
#pragma unroll
        for (int s = 0; s < CPL; ++s) {
            x[s] = better ? xe[s] : x[s];
            g[s] = better ? gcur[s] : g[s];
        }
        const bool finish = alive && ((inner_done && stop) || nev >= MMA_MAXEVAL);
        const bool outer = alive && inner_done && !finish;
        if (!first && !inner_done && f > gval) {                 // rho grows inside an inner loop
            const double r1 = 10 * rho, r2 = 1.1 * (rho + guarded_div(f - gval, wterm));
            rho = r1 < r2 ? r1 : r2;
        }
        first = false;
        if (__any_sync(FULLMASK, outer)) {
            if (outer) {
                rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;
#pragma unroll
                for (int s = 0; s < CPL; ++s) {
                    if (k > 1) {
                        const double s2 = (xe[s] - xp[s]) * (xp[s] - ctx_xpp[s * 128 + tid]);
                        const double gam = s2 < 0 ? 0.7 : (s2 > 0 ? 1.2 : 1.0);
                        sig[s] = sig[s] * gam;
                        isig[s] = fast_rcp(sig[s]);
                    }
                    ctx_xpp[s * 128 + tid] = xp[s];
                    xp[s] = xe[s];
                }
                ++k;
            }
        }
        if (finish) {
            double *dst = NU ? p.nu : p.lam;
#pragma unroll
            for (int s = 0; s < CPL; ++s)
                if (active[s]) {
                    dst[dcur * MK + gl + G * s] = x[s];
                    double2 a = red[(warp * CPL + s) * 32 + lane];
                    dd_add(a.x, a.y, x[s]);
                    red[(warp * CPL + s) * 32 + lane] = a;
                }
            if (gl == 0) (NU ? p.nev_nu : p.nev_lam)[dcur] = nev;
            need = true;
        }
    }
    __syncthreads();
    // coordinate j = gl + G s: sum over warps and over the groups; partial: [grid][2 MK], Σλ then Σν
    for (int j = threadIdx.x; j < MK; j += blockDim.x) {
        const int s = j / G, l = j % G;
        double hi = 0.0, lo = 0.0;
        for (int wv = 0; wv < NW; ++wv)
            for (int gg = 0; gg < NG; ++gg) {
                const double2 v = red[(wv * CPL + s) * 32 + gg * G + l];
                dd_merge(hi, lo, v.x, v.y);
            }
        put_partial(partial + (size_t)blockIdx.x * 2 * MK + (NU ? MK : 0) + j, hi, lo, p.accum);
    }
}

}  // namespace mmsig
