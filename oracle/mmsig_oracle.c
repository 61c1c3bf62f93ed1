/*
 * mmsig_oracle.c -- CPU ORACLE for the MMCTM / CTM / LDA variational-EM loop.
 * TEST INFRASTRUCTURE ONLY (see mmsig_oracle.h).  "parity unpinned" at the
 * NLopt boundary: Julia/NLopt cannot run here; the closed-form pieces are
 * pinned by the reference's own known-answer tests (tests/golden/).
 *
 * Every function cites the reference file:line (under /root/reference) that it
 * restates.  Operation order and staleness follow the reference literally.
 */
#include "mmsig_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ========================================================================
 * Arithmetic modes.
 *
 * ORC_ARITH_LITERAL (default): the reference's operation order, glibc
 *   exp/log, sequential sums -- the literal restatement.
 *
 * ORC_ARITH_DET: same formulas and same addends, but every rounding is pinned
 *   so that an independent implementation (the CUDA product) can reproduce the
 *   result BIT FOR BIT.  Why this exists: the reference's lambda/nu are where
 *   NLopt's MMA happens to stop, and MMA's conservative test `gval >= fcur`
 *   is decided by the last bits of f near convergence (|gval-fcur| ~
 *   rho*dx^2 ~ 1e-12 against |f| ~ 1e4).  One ulp of difference in exp() or in
 *   a summation order flips that branch for a few percent of the samples and
 *   moves their lambda by ~1e-5 -- the reference itself is not reproducible
 *   across libm / BLAS builds at the 1e-12 level north_star asks for.  A
 *   1e-12 parity claim is therefore only meaningful against pinned arithmetic:
 *     (1) exp, log: the fixed algorithms det_exp / det_log below (< 1 ulp,
 *         only + * fma / sqrt and bit moves, all IEEE-754 exact-rounded);
 *     (2) sums over the MK coordinates of one sample (objective values, MMA's
 *         gval / wval, the x-tolerance norms): a fixed 32-leaf binary tree
 *         (tree_sum32); mat-vec rows: an even-index and an odd-index fma chain, added;
 *     (3) the row's log-likelihood (a sum over the nonzeros w of ONE row): blocks
 *         of 32 terms summed in term order, the blocks added in order;  sum-theta_k = exp(lambda_k) * (fma chain
 *         over the row's nonzeros, in term order, of E_kv * R_w), R_w = n_w (1/Z_w);
 *         sums over SAMPLES d (sum lambda, sum nu, the covariance moments, the
 *         LL totals, and the topic-term statistics in their product form
 *         sum n theta_kv = E_kv * sum_d fl(exp(lambda_dk) R_dv)): the EXACTLY ROUNDED sum
 *         of the addends -- order independent by definition, so independent
 *         of how samples are sharded over warps, blocks or GPUs;
 *     (4) theta_k = e_k * (1/Z) (one division per nonzero) instead of e_k / Z,
 *         with e_k = exp(lambda_k) * exp(Elnphi_kv) (the product form of the
 *         reference's own unsmoothed_update_theta!, src/MMCTM.jl:503).
 *   Nothing else changes.  tests/ check LITERAL against the reference's
 *   known-answer tests, DET against LITERAL (agreement to ~1e-13 wherever no
 *   MMA branch flips), and the CUDA path against DET.
 * ===================================================================== */

static inline double u64_as_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t double_as_u64(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
static inline double pow2i(int k) { return u64_as_double((uint64_t)(k + 1023) << 52); }

/* exp: k = rint(x*log2e); r = x - k*ln2 (Cody-Waite, 2 fma); degree-13 Taylor
 * by Horner with fma; two-step scaling by 2^k (exact unless the result is
 * subnormal / overflows). */
static double det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return HUGE_VAL;
    if (x < -745.2) return 0.0;
    double kd = rint(x * 0x1.71547652b82fep+0);
    int k = (int)kd;
    double r = fma(kd, -0x1.62e42fee00000p-1, x);
    r = fma(kd, -0x1.a39ef35793c76p-33, r);
    double p = 0x1.6124613a86d09p-33;          /* 1/13! */
    p = fma(p, r, 0x1.1eed8eff8d898p-29);      /* 1/12! */
    p = fma(p, r, 0x1.ae64567f544e4p-26);
    p = fma(p, r, 0x1.27e4fb7789f5cp-22);
    p = fma(p, r, 0x1.71de3a556c734p-19);
    p = fma(p, r, 0x1.a01a01a01a01ap-16);
    p = fma(p, r, 0x1.a01a01a01a01ap-13);
    p = fma(p, r, 0x1.6c16c16c16c17p-10);
    p = fma(p, r, 0x1.1111111111111p-7);
    p = fma(p, r, 0x1.5555555555555p-5);
    p = fma(p, r, 0x1.5555555555555p-3);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int k1 = k / 2, k2 = k - k1;
    return (p * pow2i(k1)) * pow2i(k2);
}

/* log: fdlibm-style.  x = 2^k * m, m in [sqrt(2)/2, sqrt(2)); f = m-1;
 * s = f/(2+f); log(m) = f - (f^2/2 - s*(f^2/2 + R(s^2))). */
static double det_log(double x)
{
    int k = 0;
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -HUGE_VAL;
    if (x == HUGE_VAL) return x;
    if (x < 0x1p-1022) { x *= 0x1p54; k = -54; }
    uint64_t bits = double_as_u64(x);
    int e = (int)(bits >> 52) - 1023;
    uint64_t mant = bits & 0x000fffffffffffffULL;
    double m;
    if (mant >= 0x6a09e667f3bcdULL) { m = u64_as_double(mant | 0x3fe0000000000000ULL); e += 1; }
    else m = u64_as_double(mant | 0x3ff0000000000000ULL);
    k += e;
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double R = 0x1.2f112df3e5244p-3;
    R = fma(R, z, 0x1.39a09d078c69fp-3);
    R = fma(R, z, 0x1.7466496cb03dep-3);
    R = fma(R, z, 0x1.c71c51d8e78afp-3);
    R = fma(R, z, 0x1.2492494229359p-2);
    R = fma(R, z, 0x1.999999997fa04p-2);
    R = fma(R, z, 0x1.5555555555593p-1);
    R = R * z;
    double hfsq = 0.5 * f * f;
    double dk = (double)k;
    return dk * 0x1.62e42fee00000p-1
           - ((hfsq - (s * (hfsq + R) + dk * 0x1.a39ef35793c76p-33)) - f);
}

double orc_exp(double x) { return det_exp(x); }
double orc_log(double x) { return det_log(x); }

static inline double xexp(int arith, double x) { return arith ? det_exp(x) : exp(x); }
static inline double xlog(int arith, double x) { return arith ? det_log(x) : log(x); }

/* fixed 32-leaf butterfly: a[i] += a[i^16], ^8, ^4, ^2, ^1 (what a warp's
 * xor-shuffle reduction computes).  n > 32: leaf i first adds v[i], v[i+32],
 * ... in that order. */
static double tree_sum32(const double *v, int n)
{
    double a[32], b[32];
    for (int i = 0; i < 32; ++i) a[i] = 0.0;
    for (int i = 0; i < n; ++i) {
        if (i < 32) a[i] = v[i];
        else a[i & 31] += v[i];
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int i = 0; i < 32; ++i) b[i] = a[i] + a[i ^ off];
        memcpy(a, b, sizeof(a));
    }
    return a[0];
}
static double seq_sum(const double *v, int n)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
    return s;
}
static inline double xsum(int arith, const double *v, int n)
{
    return arith ? tree_sum32(v, n) : seq_sum(v, n);
}

/* exactly rounded running sum: double-double accumulator (Knuth TwoSum).  The
 * value hi+lo carries ~106 bits, so the final rounding is the correctly
 * rounded sum of the addends except when the exact sum lies within ~2^-100
 * (relative) of a rounding boundary. */
typedef struct { double hi, lo; } dd_t;
static inline void dd_add(dd_t *a, double x)
{
    double s = a->hi + x;
    double bb = s - a->hi;
    double e = (a->hi - (s - bb)) + (x - bb);
    a->hi = s;
    a->lo += e;
}
static inline double dd_round(dd_t a) { return a.hi + a.lo; }

/* ========================================================================
 * Special functions
 * ===================================================================== */

/* SpecialFunctions.jl src/gamma.jl `digamma(x::Float64)` (third-party, not in
 * /root/reference; compat "0.8.0, ~0.10, ~1, ~2", reference Project.toml:16).
 * Reflection for x <= 0, upward recurrence until x >= 7, then the asymptotic
 * series with coefficients B_2k/(2k), k = 1..8, evaluated by Horner with fma
 * (Julia's @evalpoly uses muladd). Call sites: src/MMCTM.jl:218,
 * src/common.jl:42, src/LDA.jl:79,97. */
static double digamma_a(int arith, double x)
{
    double psi = 0.0;
    if (x <= 0.0) {
        psi = -M_PI / tan(M_PI * x);
        x = 1.0 - x;
    }
    if (x < 7.0) {
        int n = 7 - (int)floor(x);
        for (int v = 1; v <= n - 1; ++v)
            psi -= 1.0 / (x + (double)v);
        psi -= 1.0 / x;
        x += (double)n;
    }
    double t = 1.0 / x;
    psi += xlog(arith, x) - 0.5 * t;
    t *= t;
    static const double c[8] = {
        0.08333333333333333, -0.008333333333333333, 0.003968253968253968,
        -0.004166666666666667, 0.007575757575757576, -0.021092796092796094,
        0.08333333333333333, -0.4432598039215686 };
    double p = c[7];
    for (int i = 6; i >= 0; --i)
        p = fma(p, t, c[i]);
    psi -= t * p;
    return psi;
}
double orc_digamma(double x) { return digamma_a(ORC_ARITH_LITERAL, x); }
double orc_digamma_det(double x) { return digamma_a(ORC_ARITH_DET, x); }

/* `logabsgamma(x)[1]` / deprecated `lgamma(x)` (SpecialFunctions -> openlibm
 * lgamma_r, an fdlibm port; glibc's lgamma is the same family). Call sites:
 * src/common.jl:4,6,45; src/LDA.jl:115,121,143,149. */
double orc_lgamma(double x)
{
    int sign;
    return lgamma_r(x, &sign);
}

/* src/common.jl:1-9 */
double orc_logmvbeta(const double *vals, int n)
{
    double r = 0.0, s = 0.0;
    for (int i = 0; i < n; ++i)
        r += orc_lgamma(vals[i]);
    for (int i = 0; i < n; ++i)
        s += vals[i];
    r -= orc_lgamma(s);
    return r;
}

/* ========================================================================
 * NLopt LD_MMA with zero constraints (third-party; NLopt src/algs/mma/mma.c
 * `mma_minimize` + `dual_func`, src/util/stop.c `nlopt_stop_x`).  Reference
 * call sites: src/MMCTM.jl:128-141 (lambda), :157-168 (nu), :253-266 (alpha).
 * Settings used there: xtol_rel = xtol_abs = 1e-4 (1e-5 for alpha), no ftol,
 * no maxeval, no maxtime; nu and alpha have lb = 1e-7.
 * ===================================================================== */

#define ORC_MMA_RHOMIN 1e-5
#define ORC_MMA_MAXEVAL 10000    /* guard only; NLopt has no limit here */

static int orc_isinf(double x) { return fabs(x) >= HUGE_VAL * 0.99 || isinf(x); }

/* NLopt >= 2.7 stop.c: L1-norm relative test, else all |dx_j| <= xtol_abs */
static int stop_x_27(unsigned n, const double *x, const double *oldx,
                     double xtol_rel, double xtol_abs, int arith)
{
    double dn = 0.0, xn = 0.0;
    if (arith) {
        double t[n ? n : 1];
        for (unsigned i = 0; i < n; ++i) t[i] = fabs(x[i] - oldx[i]);
        dn = tree_sum32(t, (int)n);
        for (unsigned i = 0; i < n; ++i) t[i] = fabs(x[i]);
        xn = tree_sum32(t, (int)n);
    } else {
        for (unsigned i = 0; i < n; ++i) dn += fabs(x[i] - oldx[i]);
        for (unsigned i = 0; i < n; ++i) xn += fabs(x[i]);
    }
    if (dn <= xtol_rel * xn) return 1;
    for (unsigned i = 0; i < n; ++i)
        if (fabs(x[i] - oldx[i]) > xtol_abs) return 0;
    return 1;
}

/* NLopt <= 2.6 stop.c: per coordinate relstop(old, new, reltol, abstol) */
static int relstop_26(double vold, double vnew, double reltol, double abstol)
{
    if (orc_isinf(vold)) return 0;
    return (fabs(vnew - vold) < abstol
            || fabs(vnew - vold) < reltol * (fabs(vnew) + fabs(vold)) * 0.5
            || (reltol > 0 && vnew == vold));
}
static int stop_x_26(unsigned n, const double *x, const double *oldx,
                     double xtol_rel, double xtol_abs)
{
    for (unsigned i = 0; i < n; ++i)
        if (!relstop_26(oldx[i], x[i], xtol_rel, xtol_abs)) return 0;
    return 1;
}

int orc_mma_minimize(unsigned n, orc_func f, void *fdata,
                     const double *lb, const double *ub,
                     double *x, double *minf,
                     double xtol_rel, double xtol_abs,
                     int stop_rule, int arith, int *nouter)
{
    double *sigma = (double *)malloc(sizeof(double) * 8 * (n ? n : 1));
    double *dfdx = sigma + n, *dfdx_cur = dfdx + n, *xcur = dfdx_cur + n;
    double *xprev = xcur + n, *xprevprev = xprev + n;
    double *gterm = xprevprev + n, *wterm = gterm + n;
    double rho = 1.0, fcur;
    unsigned j, k = 0;
    int nevals = 0;

    for (j = 0; j < n; ++j) {
        if (orc_isinf(ub[j]) || orc_isinf(lb[j]))
            sigma[j] = 1.0;
        else
            sigma[j] = 0.5 * (ub[j] - lb[j]);
    }
    fcur = *minf = f(n, x, dfdx, fdata);
    ++nevals;
    memcpy(xcur, x, sizeof(double) * n);

    while (1) { /* outer iterations */
        if (++k > 1) memcpy(xprevprev, xprev, sizeof(double) * n);
        memcpy(xprev, xcur, sizeof(double) * n);

        while (1) { /* inner iterations */
            /* dual_func with m = 0: separable closed-form minimiser of the
               MMA approximant around the best point x (gradient dfdx). */
            double gval = *minf, wval = 0.0;
            for (j = 0; j < n; ++j) {
                double u, v, dx, denominv, c, sigma2, dx2;
                gterm[j] = wterm[j] = 0.0;
                if (sigma[j] == 0) { xcur[j] = x[j]; continue; }
                u = dfdx[j];
                v = fabs(dfdx[j]) * sigma[j] + 0.5 * rho;
                sigma2 = sigma[j] * sigma[j];
                u *= sigma2;
                {
                    /* DET: u / (v sigma) as (u / v) * (1 / sigma) -- the device keeps 1 / sigma
                       per coordinate and refreshes it when sigma moves, one division less per
                       evaluation; differs from the literal by two roundings */
                    double r = arith ? (u / v) * (1.0 / sigma[j]) : u / (v * sigma[j]);
                    dx = (u / v) / (-1 - sqrt(fabs(1 - r * r)));
                }
                xcur[j] = x[j] + dx;
                if (xcur[j] > x[j] + 0.9 * sigma[j]) xcur[j] = x[j] + 0.9 * sigma[j];
                else if (xcur[j] < x[j] - 0.9 * sigma[j]) xcur[j] = x[j] - 0.9 * sigma[j];
                if (xcur[j] > ub[j]) xcur[j] = ub[j];
                else if (xcur[j] < lb[j]) xcur[j] = lb[j];
                dx = xcur[j] - x[j];
                dx2 = dx * dx;
                denominv = 1.0 / (sigma2 - dx2);
                c = sigma2 * dx;
                gterm[j] = (dfdx[j] * c + (fabs(dfdx[j]) * sigma[j] + 0.5 * rho) * dx2)
                           * denominv;
                wterm[j] = 0.5 * dx2 * denominv;
                if (!arith) { gval += gterm[j]; wval += wterm[j]; }
            }
            if (arith) {    /* DET: gval = minf + tree(gterm), wval = tree(wterm) */
                gval = *minf + tree_sum32(gterm, (int)n);
                wval = tree_sum32(wterm, (int)n);
            }

            fcur = f(n, xcur, dfdx_cur, fdata);
            ++nevals;
            int inner_done = gval >= fcur;

            if (fcur < *minf) { /* m = 0: always "feasible" */
                *minf = fcur;
                memcpy(x, xcur, sizeof(double) * n);
                memcpy(dfdx, dfdx_cur, sizeof(double) * n);
            }
            if (nevals >= ORC_MMA_MAXEVAL) goto done;
            if (inner_done) break;
            if (fcur > gval) {
                double r1 = 10 * rho, r2 = 1.1 * (rho + (fcur - gval) / wval);
                rho = r1 < r2 ? r1 : r2;
            }
        }

        if (stop_rule == ORC_STOP_NLOPT26
                ? stop_x_26(n, xcur, xprev, xtol_rel, xtol_abs)
                : stop_x_27(n, xcur, xprev, xtol_rel, xtol_abs, arith))
            goto done;

        /* update rho and sigma for iteration k+1 */
        rho = 0.1 * rho > ORC_MMA_RHOMIN ? 0.1 * rho : ORC_MMA_RHOMIN;
        if (k > 1) {
            for (j = 0; j < n; ++j) {
                double dx2 = (xcur[j] - xprev[j]) * (xprev[j] - xprevprev[j]);
                double gam = dx2 < 0 ? 0.7 : (dx2 > 0 ? 1.2 : 1);
                sigma[j] *= gam;
                if (!orc_isinf(ub[j]) && !orc_isinf(lb[j])) {
                    double w = ub[j] - lb[j];
                    if (sigma[j] > 10 * w) sigma[j] = 10 * w;
                    if (sigma[j] < 0.01 * w) sigma[j] = 0.01 * w;
                }
            }
        }
    }
done:
    if (nouter) *nouter = (int)k;
    free(sigma);
    return nevals;
}

/* ========================================================================
 * Objectives, src/common.jl
 * ===================================================================== */

/* src/common.jl:11-23 */
double orc_lambda_objective(int MK, const double *lam, double *grad,
                            const double *nu, const double *Ndivzeta,
                            const double *sumtheta, const double *mu,
                            const double *invSigma, int arith)
{
    double diff[MK], Eeeta[MK], q[MK];
    for (int j = 0; j < MK; ++j) diff[j] = lam[j] - mu[j];
    for (int j = 0; j < MK; ++j) Eeeta[j] = xexp(arith, lam[j] + 0.5 * nu[j]);
    if (arith) {
        /* DET: row_j(invSigma).diff as two fma chains in index order, one
           over the even and one over the odd indices, then one add (two
           independent dependency chains on the device); the quadratic form
           reuses it (q.diff); per-coordinate terms are fused and reduced by
           the fixed tree. */
        double t[MK];
        for (int j = 0; j < MK; ++j) {
            double s0 = 0.0, s1 = 0.0;
            for (int i = 0; i < MK; i += 2) s0 = fma(invSigma[(size_t)j * MK + i], diff[i], s0);
            for (int i = 1; i < MK; i += 2) s1 = fma(invSigma[(size_t)j * MK + i], diff[i], s1);
            q[j] = s0 + s1;
        }
        for (int j = 0; j < MK; ++j) {
            double ce = Ndivzeta[j] * Eeeta[j];
            if (grad) grad[j] = (-q[j] + sumtheta[j]) - ce;
            double a = q[j] * diff[j], b = lam[j] * sumtheta[j];
            t[j] = (b - 0.5 * a) - ce;
        }
        return tree_sum32(t, MK);
    }
    for (int j = 0; j < MK; ++j) {           /* invSigma * diff (row j) */
        double s = 0.0;
        for (int i = 0; i < MK; ++i) s += invSigma[(size_t)j * MK + i] * diff[i];
        q[j] = s;
    }
    if (grad)
        for (int j = 0; j < MK; ++j)
            grad[j] = -q[j] + sumtheta[j] - Ndivzeta[j] * Eeeta[j];
    double quad = 0.0, lin = 0.0, ee = 0.0;
    /* diff' * invSigma * diff : (invSigma' diff) . diff */
    for (int j = 0; j < MK; ++j) {
        double s = 0.0;
        for (int i = 0; i < MK; ++i) s += invSigma[(size_t)i * MK + j] * diff[i];
        quad += s * diff[j];
    }
    for (int j = 0; j < MK; ++j) lin += lam[j] * sumtheta[j];
    for (int j = 0; j < MK; ++j) ee += Ndivzeta[j] * Eeeta[j];
    return -0.5 * quad + lin - ee;
}

/* src/common.jl:25-36 ; tr(diagm(nu) * invSigma) = sum_j nu_j invSigma_jj */
double orc_nu_objective(int MK, const double *nu, double *grad,
                        const double *lam, const double *Ndivzeta,
                        const double *mu, const double *invSigma, int arith)
{
    (void)mu;
    double Eeeta[MK];
    for (int j = 0; j < MK; ++j) Eeeta[j] = xexp(arith, lam[j] + 0.5 * nu[j]);
    if (arith) {    /* DET: fused per-coordinate terms, fixed tree */
        double t[MK];
        for (int j = 0; j < MK; ++j) {
            double sjj = invSigma[(size_t)j * MK + j];
            if (grad)
                grad[j] = (-0.5 * sjj - (Ndivzeta[j] / 2) * Eeeta[j]) + (1.0 / (2 * nu[j]));
            t[j] = (-0.5 * (nu[j] * sjj) - Ndivzeta[j] * Eeeta[j]) + det_log(nu[j]) / 2;
        }
        return tree_sum32(t, MK);
    }
    if (grad)
        for (int j = 0; j < MK; ++j)
            grad[j] = -0.5 * invSigma[(size_t)j * MK + j]
                      - (Ndivzeta[j] / 2) * Eeeta[j] + (1.0 / (2 * nu[j]));
    double tr = 0.0, ee = 0.0, sl = 0.0;
    for (int j = 0; j < MK; ++j) tr += nu[j] * invSigma[(size_t)j * MK + j];
    for (int j = 0; j < MK; ++j) ee += Ndivzeta[j] * Eeeta[j];
    for (int j = 0; j < MK; ++j) sl += log(nu[j]);
    return -0.5 * tr - ee + sl / 2;
}

/* src/common.jl:38-46 */
double orc_alpha_objective(double alpha, double *grad, double sum_Elnphi,
                           int K, int V)
{
    if (grad)
        *grad = K * V * (orc_digamma(V * alpha) - orc_digamma(alpha)) + sum_Elnphi;
    return K * (orc_lgamma(V * alpha) - V * orc_lgamma(alpha)) + alpha * sum_Elnphi;
}

/* ========================================================================
 * Dense helpers: LinearAlgebra.inv / logdet (LAPACK getrf + getri restated
 * as LU with partial pivoting + triangular solves).  src/MMCTM.jl:211,292.
 * ===================================================================== */
static int lu_factor(int n, double *A, int *piv, int *sign)
{
    *sign = 1;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double a = fabs(A[(size_t)i * n + k]);
            if (a > best) { best = a; p = i; }
        }
        piv[k] = p;
        if (best == 0.0) return -1;
        if (p != k) {
            for (int j = 0; j < n; ++j) {
                double t = A[(size_t)k * n + j];
                A[(size_t)k * n + j] = A[(size_t)p * n + j];
                A[(size_t)p * n + j] = t;
            }
            *sign = -*sign;
        }
        double inv = 1.0 / A[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double l = A[(size_t)i * n + k] * inv;
            A[(size_t)i * n + k] = l;
            for (int j = k + 1; j < n; ++j)
                A[(size_t)i * n + j] -= l * A[(size_t)k * n + j];
        }
    }
    return 0;
}

int orc_inv(int n, const double *A, double *Ainv)
{
    double *LU = (double *)malloc(sizeof(double) * n * n);
    int *piv = (int *)malloc(sizeof(int) * n);
    int sign;
    memcpy(LU, A, sizeof(double) * n * n);
    if (lu_factor(n, LU, piv, &sign)) { free(LU); free(piv); return -1; }
    double *b = (double *)malloc(sizeof(double) * n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) b[i] = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < n; ++k)
            if (piv[k] != k) { double t = b[k]; b[k] = b[piv[k]]; b[piv[k]] = t; }
        for (int i = 0; i < n; ++i) {          /* L y = P b */
            double s = b[i];
            for (int j = 0; j < i; ++j) s -= LU[(size_t)i * n + j] * b[j];
            b[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {     /* U x = y */
            double s = b[i];
            for (int j = i + 1; j < n; ++j) s -= LU[(size_t)i * n + j] * b[j];
            b[i] = s / LU[(size_t)i * n + i];
        }
        for (int i = 0; i < n; ++i) Ainv[(size_t)i * n + c] = b[i];
    }
    free(b); free(LU); free(piv);
    return 0;
}

double orc_logabsdet(int n, const double *A)
{
    double *LU = (double *)malloc(sizeof(double) * n * n);
    int *piv = (int *)malloc(sizeof(int) * n);
    int sign;
    memcpy(LU, A, sizeof(double) * n * n);
    double r = 0.0;
    if (lu_factor(n, LU, piv, &sign)) r = -HUGE_VAL;
    else
        for (int i = 0; i < n; ++i) r += log(fabs(LU[(size_t)i * n + i]));
    free(LU); free(piv);
    return r;
}

/* Julia Base.mapreduce_impl pairwise sum (block 1024) of D vectors of length n
 * stored row-major with stride `stride`; used by mean(model.lambda)
 * (src/MMCTM.jl:201) and sum(diagm.(0 .=> model.nu)) (:205). */
static void pairwise_vecsum(const double *A, size_t stride, int n,
                            int64_t ifirst, int64_t ilast, double *out)
{
    if (ifirst == ilast) {
        for (int j = 0; j < n; ++j) out[j] = A[(size_t)ifirst * stride + j];
    } else if (ilast - ifirst < 1024) {
        for (int j = 0; j < n; ++j)
            out[j] = A[(size_t)ifirst * stride + j] + A[(size_t)(ifirst + 1) * stride + j];
        for (int64_t i = ifirst + 2; i <= ilast; ++i)
            for (int j = 0; j < n; ++j) out[j] += A[(size_t)i * stride + j];
    } else {
        int64_t imid = ifirst + ((ilast - ifirst) >> 1);
        double *v2 = (double *)malloc(sizeof(double) * n);
        pairwise_vecsum(A, stride, n, ifirst, imid, out);
        pairwise_vecsum(A, stride, n, imid + 1, ilast, v2);
        for (int j = 0; j < n; ++j) out[j] += v2[j];
        free(v2);
    }
}

/* ========================================================================
 * MMCTM
 * ===================================================================== */

/* ctor, src/MMCTM.jl:29-91 (init=:random gamma is drawn by the CALLER: the
 * reference uses Julia's global RNG, :61; gamma0 is passed in). */
orc_mmctm *orc_mmctm_new(int M, const int *K, const int *V, int64_t D,
                         const int64_t *const *rowptr, const int32_t *const *term,
                         const int32_t *const *cnt, const double *alpha,
                         const double *gamma0)
{
    orc_mmctm *m = (orc_mmctm *)calloc(1, sizeof(orc_mmctm));
    m->M = M; m->D = D;
    m->K = (int *)malloc(sizeof(int) * M);
    m->V = (int *)malloc(sizeof(int) * M);
    m->koff = (int *)malloc(sizeof(int) * (M + 1));
    m->goff = (int64_t *)malloc(sizeof(int64_t) * (M + 1));
    m->alpha = (double *)malloc(sizeof(double) * M);
    m->koff[0] = 0; m->goff[0] = 0;
    for (int i = 0; i < M; ++i) {
        m->K[i] = K[i]; m->V[i] = V[i]; m->alpha[i] = alpha[i];
        m->koff[i + 1] = m->koff[i] + K[i];
        m->goff[i + 1] = m->goff[i] + (int64_t)K[i] * V[i];
    }
    int MK = m->MK = m->koff[M];
    m->rowptr = (int64_t **)malloc(sizeof(void *) * M);
    m->term = (int32_t **)malloc(sizeof(void *) * M);
    m->cnt = (int32_t **)malloc(sizeof(void *) * M);
    m->theta = (double **)malloc(sizeof(void *) * M);
    m->rz = (double **)malloc(sizeof(void *) * M);
    m->N = (int64_t *)calloc((size_t)D * M, sizeof(int64_t));
    for (int i = 0; i < M; ++i) {
        int64_t nnz = rowptr[i][D];
        m->rowptr[i] = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
        memcpy(m->rowptr[i], rowptr[i], sizeof(int64_t) * (D + 1));
        m->term[i] = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
        m->cnt[i] = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
        memcpy(m->term[i], term[i], sizeof(int32_t) * nnz);
        memcpy(m->cnt[i], cnt[i], sizeof(int32_t) * nnz);
        /* theta = fill(1/K) (:52-57) */
        m->theta[i] = (double *)malloc(sizeof(double) * (nnz ? nnz : 1) * K[i]);
        for (int64_t t = 0; t < nnz * K[i]; ++t) m->theta[i][t] = 1.0 / K[i];
        m->rz[i] = (double *)calloc((size_t)(nnz ? nnz : 1), sizeof(double));
        for (int64_t d = 0; d < D; ++d) {         /* N (:38) */
            int64_t s = 0;
            for (int64_t w = rowptr[i][d]; w < rowptr[i][d + 1]; ++w) s += cnt[i][w];
            m->N[(size_t)d * M + i] = s;
        }
    }
    m->mu = (double *)calloc(MK, sizeof(double));                  /* :44 */
    m->Sigma = (double *)calloc((size_t)MK * MK, sizeof(double));    /* :45 */
    m->invSigma = (double *)calloc((size_t)MK * MK, sizeof(double)); /* :46 */
    for (int j = 0; j < MK; ++j) m->Sigma[(size_t)j * MK + j] = m->invSigma[(size_t)j * MK + j] = 1.0;
    m->lambda = (double *)calloc((size_t)D * MK, sizeof(double));    /* :82 */
    m->nu = (double *)malloc(sizeof(double) * D * MK);               /* :83 */
    for (int64_t t = 0; t < D * MK; ++t) m->nu[t] = 1.0;
    m->zeta = (double *)calloc((size_t)D * M, sizeof(double));
    m->props = (double *)calloc((size_t)D * MK, sizeof(double));
    int64_t G = m->goff[M];
    m->gamma = (double *)malloc(sizeof(double) * G);
    m->Elnphi = (double *)malloc(sizeof(double) * G);
    m->phi = (double *)malloc(sizeof(double) * G);
    memcpy(m->gamma, gamma0, sizeof(double) * G);                    /* :60-63 */
    orc_mmctm_update_Elnphi(m);                                      /* :78-79 */
    memcpy(m->phi, m->gamma, sizeof(double) * G);                    /* :80 */
    for (int64_t d = 0; d < D; ++d) orc_mmctm_update_zeta(m, d);     /* :85-86 */
    m->ll = (double *)calloc(M, sizeof(double));
    m->nev_nu = (int32_t *)calloc(D ? D : 1, sizeof(int32_t));
    m->nev_lambda = (int32_t *)calloc(D ? D : 1, sizeof(int32_t));
    m->expl = (double *)calloc((size_t)(D ? D : 1) * MK, sizeof(double));
    m->sumtheta_e = (double *)calloc((size_t)(D ? D : 1) * MK, sizeof(double));
    m->theta_unsm = 0;
    m->factored = 0;
    m->nfeat = NULL; m->feat = NULL; m->J = NULL; m->foff = NULL; m->aoff = NULL;
    m->alphaf = m->gammaf = m->Elnphif = NULL;
    m->stop_rule = ORC_STOP_NLOPT27;
    m->nthreads = 1;
    m->converged = 0;                                                /* :88 */
    return m;
}

void orc_mmctm_free(orc_mmctm *m)
{
    if (!m) return;
    for (int i = 0; i < m->M; ++i) {
        free(m->rowptr[i]); free(m->term[i]); free(m->cnt[i]); free(m->theta[i]); free(m->rz[i]);
    }
    free(m->rowptr); free(m->term); free(m->cnt); free(m->theta); free(m->rz);
    free(m->expl); free(m->sumtheta_e);
    if (m->factored) {
        for (int i = 0; i < m->M; ++i) { free(m->feat[i]); free(m->J[i]); }
        free(m->nfeat); free(m->feat); free(m->J); free(m->foff); free(m->aoff);
        free(m->alphaf); free(m->gammaf); free(m->Elnphif);
    }
    free(m->K); free(m->V); free(m->koff); free(m->goff); free(m->alpha);
    free(m->N); free(m->mu); free(m->Sigma); free(m->invSigma);
    free(m->lambda); free(m->nu); free(m->zeta); free(m->props);
    free(m->gamma); free(m->Elnphi); free(m->phi); free(m->ll);
    free(m->nev_nu); free(m->nev_lambda);
    free(m);
}

/* src/MMCTM.jl:172-181 */
void orc_mmctm_update_zeta(orc_mmctm *m, int64_t d)
{
    const double *lam = m->lambda + (size_t)d * m->MK, *nu = m->nu + (size_t)d * m->MK;
    for (int i = 0; i < m->M; ++i) {
        double s = 0.0;
        for (int j = m->koff[i]; j < m->koff[i + 1]; ++j)
            s += xexp(m->arith, lam[j] + 0.5 * nu[j]);
        m->zeta[(size_t)d * m->M + i] = s;
    }
}

/* src/MMCTM.jl:183-198 (no max-subtraction, as in the reference) */
void orc_mmctm_update_theta(orc_mmctm *m, int64_t d)
{
    const double *lam = m->lambda + (size_t)d * m->MK;
    for (int i = 0; i < m->M; ++i) {
        int K = m->K[i], V = m->V[i], off = m->koff[i];
        const double *Eln = m->Elnphi + m->goff[i];
        for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
            int v = m->term[i][w];
            double *th = m->theta[i] + (size_t)w * K;
            double s = 0.0;
            for (int k = 0; k < K; ++k) {
                /* DET: product form exp(lambda_k) * exp(Elnphi_kv), as unsmoothed_update_theta! (:503) */
                if (!m->arith && m->factored) {      /* src/IMMCTM.jl:158-166: exp(lambda) times one exp per feature */
                    int nf = m->nfeat[i];
                    int64_t rowlen = (m->foff[i + 1] - m->foff[i]) / K, o = m->foff[i] + (int64_t)k * rowlen;
                    th[k] = exp(lam[off + k]);
                    for (int f = 0; f < nf; ++f) {
                        th[k] *= exp(m->Elnphif[o + m->feat[i][(size_t)v * nf + f]]);
                        o += m->J[i][f];
                    }
                } else
                th[k] = m->arith ? det_exp(lam[off + k]) * det_exp(Eln[(size_t)k * V + v])
                                 : exp(lam[off + k] + Eln[(size_t)k * V + v]);
                s += th[k];
            }
            if (m->arith) {     /* DET: theta_k = e_k * (1/Z), one division per nonzero */
                double rz = 1.0 / s;
                m->rz[i][w] = rz;
                for (int k = 0; k < K; ++k) th[k] = th[k] * rz;
            } else
                for (int k = 0; k < K; ++k) th[k] /= s;
        }
        if (m->arith)
            for (int k = 0; k < K; ++k) m->expl[(size_t)d * m->MK + off + k] = det_exp(lam[off + k]);
    }
}

/* src/MMCTM.jl:496-509 : theta ∝ exp(lambda) * phi (phi, not exp(Elnphi)) */
void orc_mmctm_unsmoothed_update_theta(orc_mmctm *m, int64_t d)
{
    const double *lam = m->lambda + (size_t)d * m->MK;
    for (int i = 0; i < m->M; ++i) {
        int K = m->K[i], V = m->V[i], off = m->koff[i];
        const double *phi = m->phi + m->goff[i];
        for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
            int v = m->term[i][w];
            double *th = m->theta[i] + (size_t)w * K;
            double s = 0.0;
            for (int k = 0; k < K; ++k) {
                th[k] = xexp(m->arith, lam[off + k]) * phi[(size_t)k * V + v];
                s += th[k];
            }
            if (m->arith) {
                double rz = 1.0 / s;
                m->rz[i][w] = rz;
                for (int k = 0; k < K; ++k) th[k] = th[k] * rz;
            } else
                for (int k = 0; k < K; ++k) th[k] /= s;
        }
        if (m->arith)
            for (int k = 0; k < K; ++k) m->expl[(size_t)d * m->MK + off + k] = det_exp(lam[off + k]);
    }
}

/* src/MMCTM.jl:110-117 */
void orc_mmctm_calc_sumtheta(const orc_mmctm *m, int64_t d, double *out)
{
    for (int i = 0; i < m->M; ++i) {
        int K = m->K[i], off = m->koff[i];
        for (int k = 0; k < K; ++k) {
            if (m->arith) {
                /* DET: sum-theta_k = exp(lambda_k) * sum_w E_kv R_w with R_w = n_w * (1/Z_w), the
                   sum as one fma chain over the row's nonzeros in term order (on the device: a
                   thread per sample walking the dense R row; absent terms add exactly 0).  Valid
                   right after update_theta(d): uses the table that call used. */
                int V = m->V[i];
                double s = 0.0;
                for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                    int v = m->term[i][w];
                    double E = m->theta_unsm ? m->phi[m->goff[i] + (size_t)k * V + v]
                                             : det_exp(m->Elnphi[m->goff[i] + (size_t)k * V + v]);
                    s = fma(E, (double)m->cnt[i][w] * m->rz[i][w], s);
                }
                out[off + k] = m->expl[(size_t)d * m->MK + off + k] * s;
                continue;
            }
            double s = 0.0;
            for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w)
                s += m->theta[i][(size_t)w * K + k] * (double)m->cnt[i][w];
            out[off + k] = s;
        }
    }
}

/* src/MMCTM.jl:119-125 */
void orc_mmctm_calc_Ndivzeta(const orc_mmctm *m, int64_t d, double *out)
{
    for (int i = 0; i < m->M; ++i) {
        double c = (double)m->N[(size_t)d * m->M + i] / m->zeta[(size_t)d * m->M + i];
        for (int j = m->koff[i]; j < m->koff[i + 1]; ++j) out[j] = c;
    }
}

typedef struct {
    int MK;
    const double *a, *Ndivzeta, *sumtheta, *mu, *invSigma;
    int arith;
} obj_data;

/* NLopt f_max wrapper: max_objective! negates value and gradient */
static double neg_lambda_obj(unsigned n, const double *x, double *grad, void *p)
{
    obj_data *o = (obj_data *)p;
    double v = orc_lambda_objective((int)n, x, grad, o->a, o->Ndivzeta, o->sumtheta,
                                    o->mu, o->invSigma, o->arith);
    if (grad) for (unsigned j = 0; j < n; ++j) grad[j] = -grad[j];
    return -v;
}
static double neg_nu_obj(unsigned n, const double *x, double *grad, void *p)
{
    obj_data *o = (obj_data *)p;
    double v = orc_nu_objective((int)n, x, grad, o->a, o->Ndivzeta, o->mu, o->invSigma, o->arith);
    if (grad) for (unsigned j = 0; j < n; ++j) grad[j] = -grad[j];
    return -v;
}

/* src/MMCTM.jl:156-170 : LD_MMA, lb = 1e-7, xtol_rel = xtol_abs = 1e-4,
 * start = current nu; result copied regardless of the return code. */
void orc_mmctm_update_nu(orc_mmctm *m, int64_t d)
{
    int MK = m->MK;
    double Ndz[MK], lb[MK], ub[MK], x[MK], minf;
    orc_mmctm_calc_Ndivzeta(m, d, Ndz);
    for (int j = 0; j < MK; ++j) { lb[j] = 1e-7; ub[j] = HUGE_VAL; }
    memcpy(x, m->nu + (size_t)d * MK, sizeof(double) * MK);
    obj_data o = { MK, m->lambda + (size_t)d * MK, Ndz, NULL, m->mu, m->invSigma, m->arith };
    int nev = orc_mma_minimize(MK, neg_nu_obj, &o, lb, ub, x, &minf, 1e-4, 1e-4,
                               m->stop_rule, m->arith, NULL);
    m->nev_nu[d] = nev;
    memcpy(m->nu + (size_t)d * MK, x, sizeof(double) * MK);
}

/* src/MMCTM.jl:127-143 : LD_MMA, unbounded, xtol 1e-4, start = current lambda,
 * uses the NEW nu and the OLD zeta/theta. */
void orc_mmctm_update_lambda(orc_mmctm *m, int64_t d)
{
    int MK = m->MK;
    double Ndz[MK], st[MK], lb[MK], ub[MK], x[MK], minf;
    orc_mmctm_calc_Ndivzeta(m, d, Ndz);
    orc_mmctm_calc_sumtheta(m, d, st);
    memcpy(m->sumtheta_e + (size_t)d * MK, st, sizeof(double) * MK);     /* the ELBO reuses it (stale theta, :490) */
    for (int j = 0; j < MK; ++j) { lb[j] = -HUGE_VAL; ub[j] = HUGE_VAL; }
    memcpy(x, m->lambda + (size_t)d * MK, sizeof(double) * MK);
    obj_data o = { MK, m->nu + (size_t)d * MK, Ndz, st, m->mu, m->invSigma, m->arith };
    int nev = orc_mma_minimize(MK, neg_lambda_obj, &o, lb, ub, x, &minf, 1e-4, 1e-4,
                               m->stop_rule, m->arith, NULL);
    m->nev_lambda[d] = nev;
    memcpy(m->lambda + (size_t)d * MK, x, sizeof(double) * MK);
}

/* src/MMCTM.jl:450-455 */
void orc_mmctm_fitdoc(orc_mmctm *m, int64_t d)
{
    orc_mmctm_update_zeta(m, d);
    orc_mmctm_update_theta(m, d);
    orc_mmctm_update_nu(m, d);
    orc_mmctm_update_lambda(m, d);
}

/* src/MMCTM.jl:200-202 : mean(model.lambda) = pairwise sum / D */
void orc_mmctm_update_mu(orc_mmctm *m)
{
    if (m->arith) {     /* DET: exactly rounded column sums / D */
        for (int j = 0; j < m->MK; ++j) {
            dd_t a = {0.0, 0.0};
            for (int64_t d = 0; d < m->D; ++d) dd_add(&a, m->lambda[(size_t)d * m->MK + j]);
            m->mu[j] = dd_round(a) / (double)m->D;
        }
        return;
    }
    pairwise_vecsum(m->lambda, m->MK, m->MK, 0, m->D - 1, m->mu);
    for (int j = 0; j < m->MK; ++j) m->mu[j] /= (double)m->D;
}

/* src/MMCTM.jl:204-212 */
void orc_mmctm_update_Sigma(orc_mmctm *m)
{
    int MK = m->MK;
    double dg[MK], diff[MK];
    if (m->arith) {
        /* DET: Sigma_ij = exact_round( [i==j] sum_d nu_dj + sum_d diff_i*diff_j ) / D,
           each product rounded once (as in the literal code), the sum exact. */
        dd_t *acc = (dd_t *)calloc((size_t)MK * MK, sizeof(dd_t));
        for (int64_t d = 0; d < m->D; ++d) {
            for (int j = 0; j < MK; ++j) diff[j] = m->lambda[(size_t)d * MK + j] - m->mu[j];
            for (int i = 0; i < MK; ++i) {
                dd_add(&acc[(size_t)i * MK + i], m->nu[(size_t)d * MK + i]);
                for (int j = 0; j < MK; ++j)
                    dd_add(&acc[(size_t)i * MK + j], diff[i] * diff[j]);
            }
        }
        for (int t = 0; t < MK * MK; ++t) m->Sigma[t] = dd_round(acc[t]) / (double)m->D;
        free(acc);
        orc_inv(MK, m->Sigma, m->invSigma);
        return;
    }
    pairwise_vecsum(m->nu, MK, MK, 0, m->D - 1, dg);
    memset(m->Sigma, 0, sizeof(double) * MK * MK);
    for (int j = 0; j < MK; ++j) m->Sigma[(size_t)j * MK + j] = dg[j];
    for (int64_t d = 0; d < m->D; ++d) {
        for (int j = 0; j < MK; ++j) diff[j] = m->lambda[(size_t)d * MK + j] - m->mu[j];
        for (int i = 0; i < MK; ++i)
            for (int j = 0; j < MK; ++j)
                m->Sigma[(size_t)i * MK + j] += diff[i] * diff[j];
    }
    for (int t = 0; t < MK * MK; ++t) m->Sigma[t] /= (double)m->D;
    orc_inv(MK, m->Sigma, m->invSigma);
}

/* src/MMCTM.jl:214-222 */
void orc_mmctm_update_Elnphi(orc_mmctm *m)
{
    for (int i = 0; i < m->M; ++i)
        for (int k = 0; k < m->K[i]; ++k) {
            const double *g = m->gamma + m->goff[i] + (size_t)k * m->V[i];
            double *e = m->Elnphi + m->goff[i] + (size_t)k * m->V[i];
            double s = 0.0;
            for (int v = 0; v < m->V[i]; ++v) s += g[v];
            double ds = digamma_a(m->arith, s);
            for (int v = 0; v < m->V[i]; ++v) e[v] = digamma_a(m->arith, g[v]) - ds;
        }
}

/* src/MMCTM.jl:224-242 */
void orc_mmctm_update_gamma(orc_mmctm *m)
{
    if (m->factored) { orc_immctm_update_gamma(m); return; }
    if (m->arith) {
        /* DET: S_kv = exactly rounded sum over samples of fl(exp(lambda_dk) * R_dv), R = n * (1/Z)
           (the (D x K)^T (D x V) product form of the statistics); sum n theta = E_kv * S_kv and
           gamma_kv = fma(E_kv, S_kv, alpha), E the table of the E-step (still current here). */
        int64_t G = m->goff[m->M];
        dd_t *acc = (dd_t *)calloc((size_t)G, sizeof(dd_t));
        for (int64_t d = 0; d < m->D; ++d)
            for (int i = 0; i < m->M; ++i) {
                int K = m->K[i], V = m->V[i];
                dd_t *g = acc + m->goff[i];
                const double *L = m->expl + (size_t)d * m->MK + m->koff[i];
                for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                    int v = m->term[i][w];
                    double R = (double)m->cnt[i][w] * m->rz[i][w];
                    for (int k = 0; k < K; ++k) dd_add(&g[(size_t)k * V + v], L[k] * R);
                }
            }
        for (int i = 0; i < m->M; ++i)
            for (int64_t t = 0; t < (int64_t)m->K[i] * m->V[i]; ++t) {
                int64_t x = m->goff[i] + t;
                double E = m->theta_unsm ? m->phi[x] : det_exp(m->Elnphi[x]);
                m->gamma[x] = fma(E, dd_round(acc[x]), m->alpha[i]);
            }
        free(acc);
        orc_mmctm_update_Elnphi(m);
        return;
    }
    for (int i = 0; i < m->M; ++i)
        for (int64_t t = 0; t < (int64_t)m->K[i] * m->V[i]; ++t)
            m->gamma[m->goff[i] + t] = m->alpha[i];
    for (int64_t d = 0; d < m->D; ++d)
        for (int i = 0; i < m->M; ++i) {
            int K = m->K[i], V = m->V[i];
            double *g = m->gamma + m->goff[i];
            for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                int v = m->term[i][w];
                double n = (double)m->cnt[i][w];
                for (int k = 0; k < K; ++k)
                    g[(size_t)k * V + v] += m->theta[i][(size_t)w * K + k] * n;
            }
        }
    orc_mmctm_update_Elnphi(m);
}

/* src/MMCTM.jl:145-154 (no max-subtraction) */
void orc_mmctm_update_props(orc_mmctm *m)
{
    for (int64_t d = 0; d < m->D; ++d)
        for (int i = 0; i < m->M; ++i) {
            const double *eta = m->lambda + (size_t)d * m->MK + m->koff[i];
            double *p = m->props + (size_t)d * m->MK + m->koff[i];
            double s = 0.0;
            for (int k = 0; k < m->K[i]; ++k) s += xexp(m->arith, eta[k]);
            for (int k = 0; k < m->K[i]; ++k) p[k] = xexp(m->arith, eta[k]) / s;
        }
}

/* src/MMCTM.jl:244-250 */
void orc_mmctm_update_phi(orc_mmctm *m)
{
    if (m->factored) return;      /* the composite phi follows the feature tables (orc_immctm_update_Elnphi) */
    for (int i = 0; i < m->M; ++i)
        for (int k = 0; k < m->K[i]; ++k) {
            const double *g = m->gamma + m->goff[i] + (size_t)k * m->V[i];
            double *p = m->phi + m->goff[i] + (size_t)k * m->V[i];
            double s = 0.0;
            for (int v = 0; v < m->V[i]; ++v) s += g[v];
            for (int v = 0; v < m->V[i]; ++v) p[v] = g[v] / s;
        }
}

static double neg_alpha_obj(unsigned n, const double *x, double *grad, void *p)
{
    (void)n;
    double *q = (double *)p; /* sum_Elnphi, K, V */
    double g, v = orc_alpha_objective(x[0], &g, q[0], (int)q[1], (int)q[2]);
    if (grad) grad[0] = -g;
    return -v;
}

/* src/MMCTM.jl:252-269 : 1-D LD_MMA, lb 1e-7, xtol 1e-5 */
void orc_mmctm_update_alpha(orc_mmctm *m)
{
    if (m->factored) {          /* src/IMMCTM.jl:223-241: one alpha per (modality, feature) */
        for (int i = 0; i < m->M; ++i) {
            int K = m->K[i], nf = m->nfeat[i];
            int64_t rowlen = (m->foff[i + 1] - m->foff[i]) / K, fo = 0;
            for (int f = 0; f < nf; ++f) {
                double s = 0.0;
                for (int k = 0; k < K; ++k)
                    for (int j = 0; j < m->J[i][f]; ++j) s += m->Elnphif[m->foff[i] + (int64_t)k * rowlen + fo + j];
                double q[3] = { s, (double)K, (double)m->J[i][f] };
                double lb = 1e-7, ub = HUGE_VAL, x = m->alphaf[m->aoff[i] + f], minf;
                orc_mma_minimize(1, neg_alpha_obj, q, &lb, &ub, &x, &minf, 1e-5, 1e-5,
                                 m->stop_rule, ORC_ARITH_LITERAL, NULL);
                m->alphaf[m->aoff[i] + f] = x;
                fo += m->J[i][f];
            }
        }
        return;
    }
    for (int i = 0; i < m->M; ++i) {
        double s = 0.0;
        for (int64_t t = 0; t < (int64_t)m->K[i] * m->V[i]; ++t)
            s += m->Elnphi[m->goff[i] + t];
        double q[3] = { s, (double)m->K[i], (double)m->V[i] };
        double lb = 1e-7, ub = HUGE_VAL, x = m->alpha[i], minf;
        orc_mma_minimize(1, neg_alpha_obj, q, &lb, &ub, &x, &minf, 1e-5, 1e-5,
                         m->stop_rule, ORC_ARITH_LITERAL, NULL);
        m->alpha[i] = x;
    }
}

/* src/MMCTM.jl:384-448 */
void orc_mmctm_loglikelihoods(const orc_mmctm *m, double *ll)
{
    for (int i = 0; i < m->M; ++i) {
        int K = m->K[i], V = m->V[i];
        const double *phi = m->phi + m->goff[i];
        double tot = 0.0;
        dd_t tot_dd = {0.0, 0.0};
        int64_t Ntot = 0;
        for (int64_t d = 0; d < m->D; ++d) {
            int64_t docN = m->N[(size_t)d * m->M + i];       /* :409 */
            if (docN > 0) {
                const double *p = m->props + (size_t)d * m->MK + m->koff[i];
                double dl = 0.0;
                int64_t rb = m->rowptr[i][d], rn = m->rowptr[i][d + 1] - rb;
                double tl[rn ? rn : 1];
                for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                    int v = m->term[i][w];
                    double pw = 0.0;
                    for (int k = 0; k < K; ++k) pw += p[k] * phi[(size_t)k * V + v];
                    if (m->arith) tl[w - rb] = (double)m->cnt[i][w] * det_log(pw);
                    else dl += (double)m->cnt[i][w] * log(pw);
                }
                if (m->arith) {
                    /* DET: blocks of 32 TERMS summed in term order, the blocks added in order
                       (on the device: a thread per (sample, block) over the dense tile row) */
                    int nb = (V + 31) / 32;
                    double bs[nb];
                    for (int j = 0; j < nb; ++j) bs[j] = 0.0;
                    for (int64_t w = 0; w < rn; ++w) bs[m->term[i][rb + w] / 32] += tl[w];
                    dl = bs[0];
                    for (int j = 1; j < nb; ++j) dl += bs[j];
                }
                dl = dl / (double)docN;                       /* :399 */
                if (m->arith) dd_add(&tot_dd, dl * (double)docN);
                else tot += dl * (double)docN;                /* :412 */
                Ntot += docN;
            }
        }
        if (m->arith) tot = dd_round(tot_dd);                 /* DET: exact sum over samples */
        ll[i] = tot / (double)Ntot;
    }
}

/* src/MMCTM.jl:271-382 ; uses whatever theta/zeta are stored (stale at fit! exit) */
double orc_mmctm_elbo(const orc_mmctm *m, double *terms)
{
    int MK = m->MK;
    double t[7] = {0, 0, 0, 0, 0, 0, 0};
    /* IMMCTM: ElnPphi src/IMMCTM.jl:244-261, ElnQphi :311-325 over the feature tables */
    if (m->factored)
        for (int i = 0; i < m->M; ++i)
            for (int k = 0; k < m->K[i]; ++k) {
                int64_t o = m->foff[i] + (int64_t)k * ((m->foff[i + 1] - m->foff[i]) / m->K[i]);
                for (int f = 0; f < m->nfeat[i]; ++f) {
                    int Jf = m->J[i][f];
                    double a = m->alphaf[m->aoff[i] + f];
                    double fillv[Jf];
                    for (int j = 0; j < Jf; ++j) fillv[j] = a;
                    t[0] -= orc_logmvbeta(fillv, Jf);
                    t[4] += -orc_logmvbeta(m->gammaf + o, Jf);
                    for (int j = 0; j < Jf; ++j) {
                        t[0] += (a - 1) * m->Elnphif[o + j];
                        t[4] += (m->gammaf[o + j] - 1) * m->Elnphif[o + j];
                    }
                    o += Jf;
                }
            }
    /* ElnPphi :271-284 */
    for (int i = 0; i < m->M && !m->factored; ++i) {
        double *fillv = (double *)malloc(sizeof(double) * m->V[i]);
        for (int v = 0; v < m->V[i]; ++v) fillv[v] = m->alpha[i];
        for (int k = 0; k < m->K[i]; ++k) {
            t[0] -= orc_logmvbeta(fillv, m->V[i]);
            for (int v = 0; v < m->V[i]; ++v)
                t[0] += (m->alpha[i] - 1) * m->Elnphi[m->goff[i] + (size_t)k * m->V[i] + v];
        }
        free(fillv);
    }
    /* ElnPeta :286-300 */
    {
        double ld = orc_logabsdet(MK, m->invSigma);
        double diff[MK];
        for (int64_t d = 0; d < m->D; ++d) {
            double tr = 0.0, quad = 0.0;
            for (int j = 0; j < MK; ++j) diff[j] = m->lambda[(size_t)d * MK + j] - m->mu[j];
            for (int j = 0; j < MK; ++j) tr += m->nu[(size_t)d * MK + j] * m->invSigma[(size_t)j * MK + j];
            for (int j = 0; j < MK; ++j) {
                double s = 0.0;
                for (int i = 0; i < MK; ++i) s += m->invSigma[(size_t)i * MK + j] * diff[i];
                quad += s * diff[j];
            }
            t[1] += 0.5 * (ld - MK * log(2 * M_PI) - tr - quad);
        }
    }
    /* ElnPZ :302-316 */
    {
        double st[MK], Ndz[MK];
        for (int64_t d = 0; d < m->D; ++d) {
            const double *lam = m->lambda + (size_t)d * MK, *nu = m->nu + (size_t)d * MK;
            if (m->arith) memcpy(st, m->sumtheta_e + (size_t)d * MK, sizeof(double) * MK);   /* tables have moved on */
            else orc_mmctm_calc_sumtheta(m, d, st);
            orc_mmctm_calc_Ndivzeta(m, d, Ndz);
            double a = 0.0, b = 0.0, sN = 0.0, c = 0.0;
            for (int j = 0; j < MK; ++j) a += lam[j] * st[j];
            for (int j = 0; j < MK; ++j) b += Ndz[j] * xexp(m->arith, lam[j] + 0.5 * nu[j]);
            for (int i = 0; i < m->M; ++i) sN += (double)m->N[(size_t)d * m->M + i];
            for (int i = 0; i < m->M; ++i)
                c += (double)m->N[(size_t)d * m->M + i] * xlog(m->arith, m->zeta[(size_t)d * m->M + i]);
            t[2] += a;
            t[2] -= b - sN;
            t[2] -= c;
        }
    }
    /* ElnPX :318-336 */
    for (int64_t d = 0; d < m->D; ++d)
        for (int i = 0; i < m->M; ++i) {
            int K = m->K[i], V = m->V[i];
            for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                int v = m->term[i][w];
                for (int k = 0; k < K; ++k)
                    t[3] += (double)m->cnt[i][w] * m->theta[i][(size_t)w * K + k]
                            * m->Elnphi[m->goff[i] + (size_t)k * V + v];
            }
        }
    /* ElnQphi :338-350 */
    for (int i = 0; i < m->M && !m->factored; ++i)
        for (int k = 0; k < m->K[i]; ++k) {
            const double *g = m->gamma + m->goff[i] + (size_t)k * m->V[i];
            t[4] += -orc_logmvbeta(g, m->V[i]);
            for (int v = 0; v < m->V[i]; ++v)
                t[4] += (g[v] - 1) * m->Elnphi[m->goff[i] + (size_t)k * m->V[i] + v];
        }
    /* ElnQeta :352-358 */
    for (int64_t d = 0; d < m->D; ++d) {
        double sl = 0.0;
        for (int j = 0; j < MK; ++j) sl += xlog(m->arith, m->nu[(size_t)d * MK + j]);
        t[5] += -0.5 * (sl + MK * (log(2 * M_PI) + 1));
    }
    /* ElnQZ :360-370 : n * log(theta^theta), 0^0 = 1 */
    for (int64_t d = 0; d < m->D; ++d)
        for (int i = 0; i < m->M; ++i) {
            int K = m->K[i];
            double s = 0.0;
            for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w)
                for (int k = 0; k < K; ++k) {
                    double th = m->theta[i][(size_t)w * K + k];
                    s += (double)m->cnt[i][w] * log(pow(th, th));
                }
            t[6] += s;
        }
    if (terms) memcpy(terms, t, sizeof(t));
    return t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6];   /* :372-382 */
}

/* body of the fit! loop, src/MMCTM.jl:463-479 */
void orc_mmctm_iterate(orc_mmctm *m, int updateSigma, int autoalpha, double *ll)
{
    m->theta_unsm = 0;
#ifdef _OPENMP
    #pragma omp parallel for schedule(dynamic, 16) num_threads(m->nthreads > 0 ? m->nthreads : 1)
#endif
    for (int64_t d = 0; d < m->D; ++d)
        orc_mmctm_fitdoc(m, d);
    orc_mmctm_update_mu(m);
    if (updateSigma) orc_mmctm_update_Sigma(m);
    orc_mmctm_update_gamma(m);
    if (autoalpha) orc_mmctm_update_alpha(m);
    orc_mmctm_update_props(m);
    orc_mmctm_update_phi(m);
    orc_mmctm_loglikelihoods(m, ll);
}

/* The loop bodies of fit_heldout (src/MMCTM.jl:566-573: E-step, props, LL), transform
 * (:523-538: E-step with unsmoothed theta, [mu, Sigma], props, LL) and predict_modality_eta
 * (:604-609) as flag combinations of one iteration. */
void orc_mmctm_iterate_flags(orc_mmctm *m, unsigned flags, double *ll)
{
    const int unsm = (flags & ORC_FLAG_UNSMOOTHED) != 0;
    m->theta_unsm = unsm;
#ifdef _OPENMP
    #pragma omp parallel for schedule(dynamic, 16) num_threads(m->nthreads > 0 ? m->nthreads : 1)
#endif
    for (int64_t d = 0; d < m->D; ++d) {
        orc_mmctm_update_zeta(m, d);
        if (unsm) orc_mmctm_unsmoothed_update_theta(m, d);
        else orc_mmctm_update_theta(m, d);
        orc_mmctm_update_nu(m, d);
        orc_mmctm_update_lambda(m, d);
    }
    if (!(flags & ORC_FLAG_FREEZE_MU)) orc_mmctm_update_mu(m);
    if (flags & ORC_FLAG_UPDATE_SIGMA) orc_mmctm_update_Sigma(m);
    if (!(flags & ORC_FLAG_FREEZE_TOPICS)) { orc_mmctm_update_gamma(m); }
    orc_mmctm_update_props(m);
    if (!(flags & ORC_FLAG_FREEZE_TOPICS)) orc_mmctm_update_phi(m);
    orc_mmctm_loglikelihoods(m, ll);
}

/* src/common.jl:48-51 */
static int check_convergence_vec(const double *prev, const double *cur, int M, double tol)
{
    double r = 0.0;
    for (int i = 0; i < M; ++i) {
        double v = fabs(prev[i] - cur[i]) / fabs(cur[i]);
        if (v > r || v != v) r = v;
    }
    return r < tol;
}

/* src/MMCTM.jl:457-494 */
int orc_mmctm_fit(orc_mmctm *m, int maxiter, double tol, int updateSigma,
                  int autoalpha, double *ll_hist)
{
    int it = 0;
    for (int iter = 1; iter <= maxiter; ++iter) {
        double *ll = ll_hist + (size_t)(iter - 1) * m->M;
        orc_mmctm_iterate(m, updateSigma, autoalpha, ll);
        it = iter;
        if (iter > 10 && check_convergence_vec(ll - m->M, ll, m->M, tol)) {
            m->converged = 1;
            break;
        }
    }
    m->elbo = orc_mmctm_elbo(m, NULL);
    if (it > 0) memcpy(m->ll, ll_hist + (size_t)(it - 1) * m->M, sizeof(double) * m->M);
    return it;
}

/* ========================================================================
 * LDA, src/LDA.jl.  Matrices keep Julia's column-major layout:
 *   lambda/Elnbeta/beta (V x K): [k*V + v] ; gamma/Elntheta/theta (K x D): [d*K + k]
 * ===================================================================== */

/* ctor, src/LDA.jl:24-54 ; lambda0 = rand(1:100, V, K) drawn by the caller */
orc_lda *orc_lda_new(int K, int V, int64_t D, const int64_t *rowptr,
                     const int32_t *term, const int32_t *cnt,
                     double alpha, double eta, const double *lambda0)
{
    orc_lda *m = (orc_lda *)calloc(1, sizeof(orc_lda));
    m->K = K; m->V = V; m->D = D; m->alpha = alpha; m->eta = eta;
    int64_t nnz = rowptr[D];
    m->rowptr = (int64_t *)malloc(sizeof(int64_t) * (D + 1));
    memcpy(m->rowptr, rowptr, sizeof(int64_t) * (D + 1));
    m->term = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    m->cnt = (int32_t *)malloc(sizeof(int32_t) * (nnz ? nnz : 1));
    memcpy(m->term, term, sizeof(int32_t) * nnz);
    memcpy(m->cnt, cnt, sizeof(int32_t) * nnz);
    m->N = (int64_t *)calloc(D ? D : 1, sizeof(int64_t));
    for (int64_t d = 0; d < D; ++d)
        for (int64_t w = rowptr[d]; w < rowptr[d + 1]; ++w) m->N[d] += cnt[w];
    m->lambda = (double *)malloc(sizeof(double) * V * K);
    m->Elnbeta = (double *)malloc(sizeof(double) * V * K);
    m->beta = (double *)calloc((size_t)V * K, sizeof(double));
    memcpy(m->lambda, lambda0, sizeof(double) * V * K);
    orc_lda_update_Elnbeta(m);                                  /* :39 */
    m->gamma = (double *)malloc(sizeof(double) * K * (D ? D : 1));
    for (int64_t t = 0; t < (int64_t)K * D; ++t) m->gamma[t] = 1.0;  /* :41 */
    m->theta = (double *)calloc((size_t)K * (D ? D : 1), sizeof(double));
    m->Elntheta = (double *)malloc(sizeof(double) * K * (D ? D : 1));
    orc_lda_update_Elntheta(m);                                 /* :44 */
    m->phi = (double *)malloc(sizeof(double) * (nnz ? nnz : 1) * K);
    for (int64_t t = 0; t < nnz * K; ++t) m->phi[t] = 1.0 / K;  /* :46-49 */
    m->nthreads = 1;
    m->factored = 0; m->nfeat = 0; m->feat = NULL; m->J = NULL; m->T = 0;
    m->etaf = m->lambdaf = m->Elnbetaf = NULL;
    return m;
}

void orc_lda_free(orc_lda *m)
{
    if (!m) return;
    if (m->factored) { free(m->feat); free(m->J); free(m->etaf); free(m->lambdaf); free(m->Elnbetaf); }
    free(m->rowptr); free(m->term); free(m->cnt); free(m->N);
    free(m->lambda); free(m->Elnbeta); free(m->beta);
    free(m->gamma); free(m->Elntheta); free(m->theta); free(m->phi);
    free(m);
}

/* src/LDA.jl:78-80 */
void orc_lda_update_Elntheta(orc_lda *m)
{
    for (int64_t d = 0; d < m->D; ++d) {
        double s = 0.0;
        for (int k = 0; k < m->K; ++k) s += m->gamma[(size_t)d * m->K + k];
        double ds = digamma_a(m->arith, s);
        for (int k = 0; k < m->K; ++k)
            m->Elntheta[(size_t)d * m->K + k] = digamma_a(m->arith, m->gamma[(size_t)d * m->K + k]) - ds;
    }
}

/* src/LDA.jl:82-90 : gamma[:,d] = alpha + phi[d] * n_d (phi from the PREVIOUS iteration) */
void orc_lda_update_gamma(orc_lda *m)
{
    int K = m->K;
#ifdef _OPENMP
    #pragma omp parallel for schedule(static) num_threads(m->nthreads > 0 ? m->nthreads : 1)
#endif
    for (int64_t d = 0; d < m->D; ++d)
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w)
                s += m->phi[(size_t)w * K + k] * (double)m->cnt[w];
            m->gamma[(size_t)d * K + k] = m->alpha + s;
        }
    orc_lda_update_Elntheta(m);
}

/* src/LDA.jl:69-76 */
void orc_lda_update_phi(orc_lda *m)
{
    int K = m->K, V = m->V;
#ifdef _OPENMP
    #pragma omp parallel for schedule(static) num_threads(m->nthreads > 0 ? m->nthreads : 1)
#endif
    for (int64_t d = 0; d < m->D; ++d)
        for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w) {
            int v = m->term[w];
            double *p = m->phi + (size_t)w * K;
            double s = 0.0;
            for (int k = 0; k < K; ++k) {
                p[k] = xexp(m->arith, m->Elntheta[(size_t)d * K + k] + m->Elnbeta[(size_t)k * V + v]);
                s += p[k];
            }
            for (int k = 0; k < K; ++k) p[k] /= s;
        }
}

/* src/LDA.jl:226-231 : phi ∝ exp(Elntheta) * beta */
void orc_lda_unsmoothed_update_phi(orc_lda *m)
{
    int K = m->K, V = m->V;
    for (int64_t d = 0; d < m->D; ++d)
        for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w) {
            int v = m->term[w];
            double *p = m->phi + (size_t)w * K;
            double s = 0.0;
            for (int k = 0; k < K; ++k) {
                p[k] = xexp(m->arith, m->Elntheta[(size_t)d * K + k]) * m->beta[(size_t)k * V + v];
                s += p[k];
            }
            for (int k = 0; k < K; ++k) p[k] /= s;
        }
}

/* loop bodies of LDA fit_heldout (src/LDA.jl:275-280) and transform (:242-246) */
double orc_lda_iterate_flags(orc_lda *m, unsigned flags)
{
    orc_lda_update_gamma(m);
    if (flags & ORC_FLAG_UNSMOOTHED) orc_lda_unsmoothed_update_phi(m);
    else orc_lda_update_phi(m);
    if (!(flags & ORC_FLAG_FREEZE_TOPICS)) { orc_lda_update_lambda(m); orc_lda_update_beta(m); }
    orc_lda_update_theta(m);
    return orc_lda_loglikelihood(m);
}

/* src/LDA.jl:96-98 */
void orc_lda_update_Elnbeta(orc_lda *m)
{
    if (m->factored) { orc_ilda_update_Elnbeta(m); return; }
    for (int k = 0; k < m->K; ++k) {
        double s = 0.0;
        for (int v = 0; v < m->V; ++v) s += m->lambda[(size_t)k * m->V + v];
        double ds = digamma_a(m->arith, s);
        for (int v = 0; v < m->V; ++v)
            m->Elnbeta[(size_t)k * m->V + v] = digamma_a(m->arith, m->lambda[(size_t)k * m->V + v]) - ds;
    }
}

/* src/LDA.jl:100-108 */
void orc_lda_update_lambda(orc_lda *m)
{
    int K = m->K, V = m->V;
    if (m->factored) {          /* src/ILDA.jl:105-125: lambda_i[j, :] = eta_i + sum over the nonzeros whose term carries j */
        int nf = m->nfeat;
        int64_t rowlen = m->T / K;
        for (int k = 0; k < K; ++k) {
            int64_t o = (int64_t)k * rowlen;
            for (int f = 0; f < nf; ++f) {
                for (int j = 0; j < m->J[f]; ++j) m->lambdaf[o + j] = m->etaf[f];
                o += m->J[f];
            }
        }
        for (int64_t d = 0; d < m->D; ++d)
            for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w) {
                int v = m->term[w];
                for (int k = 0; k < K; ++k) {
                    double np = m->phi[(size_t)w * K + k] * (double)m->cnt[w];
                    int64_t o = (int64_t)k * rowlen;
                    for (int f = 0; f < nf; ++f) {
                        m->lambdaf[o + m->feat[(size_t)v * nf + f]] += np;
                        o += m->J[f];
                    }
                }
            }
        orc_ilda_update_Elnbeta(m);
        return;
    }
    for (int t = 0; t < V * K; ++t) m->lambda[t] = m->eta;
    for (int64_t d = 0; d < m->D; ++d)
        for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w) {
            int v = m->term[w];
            for (int k = 0; k < K; ++k)
                m->lambda[(size_t)k * V + v] += m->phi[(size_t)w * K + k] * (double)m->cnt[w];
        }
    orc_lda_update_Elnbeta(m);
}

/* src/LDA.jl:110-112 */
void orc_lda_update_beta(orc_lda *m)
{
    if (m->factored) { orc_ilda_compose(m); return; }      /* src/ILDA.jl:127-129 */
    for (int k = 0; k < m->K; ++k) {
        double s = 0.0;
        for (int v = 0; v < m->V; ++v) s += m->lambda[(size_t)k * m->V + v];
        for (int v = 0; v < m->V; ++v)
            m->beta[(size_t)k * m->V + v] = m->lambda[(size_t)k * m->V + v] / s;
    }
}

/* src/LDA.jl:92-94 */
void orc_lda_update_theta(orc_lda *m)
{
    for (int64_t d = 0; d < m->D; ++d) {
        double s = 0.0;
        for (int k = 0; k < m->K; ++k) s += m->gamma[(size_t)d * m->K + k];
        for (int k = 0; k < m->K; ++k)
            m->theta[(size_t)d * m->K + k] = m->gamma[(size_t)d * m->K + k] / s;
    }
}

/* src/LDA.jl:174-188 */
double orc_lda_loglikelihood(const orc_lda *m)
{
    double ll = 0.0;
    int64_t N = 0;
    for (int64_t d = 0; d < m->D; ++d) {
        N += m->N[d];
        for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w) {
            int v = m->term[w];
            double dot = 0.0;
            for (int k = 0; k < m->K; ++k)
                dot += m->theta[(size_t)d * m->K + k] * m->beta[(size_t)k * m->V + v];
            ll += (double)m->cnt[w] * xlog(m->arith, dot);
        }
    }
    return ll / (double)N;
}

/* src/LDA.jl:114-172 */
double orc_lda_elbo(const orc_lda *m, double *terms)
{
    int K = m->K, V = m->V;
    double t[7] = {0, 0, 0, 0, 0, 0, 0};
    double s;
    /* ElnPbeta :114-118 */
    s = 0.0; for (int i = 0; i < V * K; ++i) s += m->Elnbeta[i];
    t[0] = K * (orc_lgamma(V * m->eta) - V * orc_lgamma(m->eta)) + (m->eta - 1) * s;
    if (m->factored) {          /* src/ILDA.jl:131-140, per feature */
        int64_t rowlen = m->T / K, fo = 0;
        t[0] = 0.0;
        for (int f = 0; f < m->nfeat; ++f) {
            double se = 0.0;
            for (int k = 0; k < K; ++k)
                for (int j = 0; j < m->J[f]; ++j) se += m->Elnbetaf[(int64_t)k * rowlen + fo + j];
            t[0] += K * (orc_lgamma(m->J[f] * m->etaf[f]) - m->J[f] * orc_lgamma(m->etaf[f]));
            t[0] += (m->etaf[f] - 1) * se;
            fo += m->J[f];
        }
    }
    /* ElnPtheta :120-124 */
    s = 0.0; for (int64_t i = 0; i < (int64_t)K * m->D; ++i) s += m->Elntheta[i];
    t[1] = m->D * (orc_lgamma(K * m->alpha) - K * orc_lgamma(m->alpha)) + (m->alpha - 1) * s;
    /* ElnPZ :126-132 ; ElnPX :134-140 ; ElnQZ :154-160 (NOT count weighted) */
    for (int64_t d = 0; d < m->D; ++d) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int64_t w = m->rowptr[d]; w < m->rowptr[d + 1]; ++w) {
            int v = m->term[w];
            for (int k = 0; k < K; ++k) {
                double p = m->phi[(size_t)w * K + k];
                a += p * m->Elntheta[(size_t)d * K + k] * (double)m->cnt[w];
                b += p * m->Elnbeta[(size_t)k * V + v] * (double)m->cnt[w];
                c += log(pow(p, p));
            }
        }
        t[2] += a; t[3] += b; t[6] += c;
    }
    /* ElnQbeta :142-146 */
    {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int i = 0; i < V * K; ++i) a += orc_lgamma(m->lambda[i]);
        for (int k = 0; k < K; ++k) {
            double cs = 0.0;
            for (int v = 0; v < V; ++v) cs += m->lambda[(size_t)k * V + v];
            b += orc_lgamma(cs);
        }
        for (int i = 0; i < V * K; ++i) c += (m->lambda[i] - 1) * m->Elnbeta[i];
        t[4] = a - b - c;
        if (m->factored) {
            /* src/ILDA.jl:174-181 as written: `lnq = ...` inside the loop over features, not `+=`, so
               only the LAST feature contributes.  Reproduced (the ELBO is only reported, never optimised). */
            int64_t rowlen = m->T / K, fo = 0;
            for (int f = 0; f < m->nfeat; ++f) {
                a = b = c = 0.0;
                for (int k = 0; k < K; ++k) {
                    double cs = 0.0;
                    for (int j = 0; j < m->J[f]; ++j) {
                        double l = m->lambdaf[(int64_t)k * rowlen + fo + j];
                        a += orc_lgamma(l);
                        cs += l;
                        c += (l - 1) * m->Elnbetaf[(int64_t)k * rowlen + fo + j];
                    }
                    b += orc_lgamma(cs);
                }
                t[4] = a - b - c;
                fo += m->J[f];
            }
        }
    }
    /* ElnQtheta :148-152 */
    {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int64_t i = 0; i < (int64_t)K * m->D; ++i) a += orc_lgamma(m->gamma[i]);
        for (int64_t d = 0; d < m->D; ++d) {
            double cs = 0.0;
            for (int k = 0; k < K; ++k) cs += m->gamma[(size_t)d * K + k];
            b += orc_lgamma(cs);
        }
        for (int64_t i = 0; i < (int64_t)K * m->D; ++i) c += (m->gamma[i] - 1) * m->Elntheta[i];
        t[5] = a - b - c;
    }
    if (terms) memcpy(terms, t, sizeof(t));
    return t[0] + t[1] + t[2] + t[3] - t[4] - t[5] - t[6];   /* :162-172 */
}

/* body of the LDA fit! loop, src/LDA.jl:202-209 */
double orc_lda_iterate(orc_lda *m)
{
    orc_lda_update_gamma(m);
    orc_lda_update_phi(m);
    orc_lda_update_lambda(m);
    orc_lda_update_beta(m);
    orc_lda_update_theta(m);
    return orc_lda_loglikelihood(m);
}

/* src/LDA.jl:198-224 */
int orc_lda_fit(orc_lda *m, int maxiter, double tol, double *ll_hist)
{
    int it = 0;
    for (int iter = 1; iter <= maxiter; ++iter) {
        ll_hist[iter - 1] = orc_lda_iterate(m);
        it = iter;
        if (iter > 10) {
            double r = fabs(ll_hist[iter - 2] - ll_hist[iter - 1]) / fabs(ll_hist[iter - 1]);
            if (r < tol) { m->converged = 1; break; }
        }
    }
    m->elbo = orc_lda_elbo(m, NULL);
    if (it > 0) m->ll = ll_hist[it - 1];
    return it;
}


/* =====================================================================
 * ILDA (src/ILDA.jl): LDA whose topics factorise over features.  The per-sample updates
 * (update_phi :64-78, update_gamma :84-92, log-likelihood :203-233) run over the composite
 * tables Elnbeta_kv = sum_i Elnbeta_i[f(v,i), k] and beta_kv = prod_i beta_i[f(v,i), k] exactly as
 * the LDA's (the literal reference adds / multiplies per nonzero: roundings only); the M-step
 * (:105-129) runs over the feature tables, flat [k][i][j].
 * ===================================================================== */
void orc_ilda_compose(orc_lda *m)
{
    int K = m->K, V = m->V, nf = m->nfeat;
    int64_t rowlen = m->T / K;
    for (int k = 0; k < K; ++k)
        for (int v = 0; v < V; ++v) {
            double e = 0.0, b = 1.0;
            int64_t o = (int64_t)k * rowlen;
            for (int f = 0; f < nf; ++f) {
                int j = m->feat[(size_t)v * nf + f];
                double sl = 0.0;
                for (int jj = 0; jj < m->J[f]; ++jj) sl += m->lambdaf[o + jj];
                e += m->Elnbetaf[o + j];
                b *= m->lambdaf[o + j] / sl;              /* update_beta!, :127-129 */
                o += m->J[f];
            }
            m->Elnbeta[(size_t)k * V + v] = e;
            m->beta[(size_t)k * V + v] = b;
        }
}

/* src/ILDA.jl:97-103 */
void orc_ilda_update_Elnbeta(orc_lda *m)
{
    int K = m->K, nf = m->nfeat;
    int64_t rowlen = m->T / K;
    for (int k = 0; k < K; ++k) {
        int64_t o = (int64_t)k * rowlen;
        for (int f = 0; f < nf; ++f) {
            double sl = 0.0;
            for (int j = 0; j < m->J[f]; ++j) sl += m->lambdaf[o + j];
            double ds = digamma_a(m->arith, sl);
            for (int j = 0; j < m->J[f]; ++j) m->Elnbetaf[o + j] = digamma_a(m->arith, m->lambdaf[o + j]) - ds;
            o += m->J[f];
        }
    }
    orc_ilda_compose(m);
}

void orc_ilda_enable(orc_lda *m, int nfeat, const int *feat, const double *etaf, const double *lambdaf0)
{
    int V = m->V;
    m->nfeat = nfeat;
    m->feat = (int *)malloc(sizeof(int) * (size_t)V * nfeat);
    memcpy(m->feat, feat, sizeof(int) * (size_t)V * nfeat);
    m->J = (int *)calloc(nfeat, sizeof(int));
    int64_t rowlen = 0;
    for (int f = 0; f < nfeat; ++f) {                     /* J = maximum(features, dims=1), :35 */
        for (int v = 0; v < V; ++v)
            if (feat[(size_t)v * nfeat + f] + 1 > m->J[f]) m->J[f] = feat[(size_t)v * nfeat + f] + 1;
        rowlen += m->J[f];
    }
    m->T = rowlen * m->K;
    m->etaf = (double *)malloc(sizeof(double) * nfeat);
    memcpy(m->etaf, etaf, sizeof(double) * nfeat);
    m->lambdaf = (double *)malloc(sizeof(double) * m->T);
    m->Elnbetaf = (double *)malloc(sizeof(double) * m->T);
    memcpy(m->lambdaf, lambdaf0, sizeof(double) * m->T);
    m->factored = 1;
    orc_ilda_update_Elnbeta(m);                           /* :42 */
    for (int i = 0; i < m->K * V; ++i) m->lambda[i] = 0.0 / 0.0;     /* no K x V lambda in this model */
}

/* =====================================================================
 * IMMCTM (src/IMMCTM.jl): feature-factorised topics.  Everything per sample (zeta, theta, nu,
 * lambda: src/IMMCTM.jl:105-172 are the MMCTM functions over the composite table) is shared
 * with the MMCTM above; what differs is the M-step over the feature tables and the two
 * table terms of the ELBO.
 * DET specification: the composite log-table Elnphi_kv = sum_i Elnphi_k,i,f(v,i) (index order,
 * from 0) so that theta uses exp(lambda) * exp(Elnphi_kv) exactly as the MMCTM does (the
 * literal reference multiplies one exp per feature: roundings only); phi_kv = prod_i phi_k,i,f(v,i)
 * (index order, from 1); gamma_k,i,j = alpha_i + (sum over v with f(v,i) = j, ascending v, of
 * sum-n-theta_kv), with sum-n-theta_kv = E_kv S_kv as in the MMCTM.
 * ===================================================================== */
int64_t orc_immctm_table_size(const orc_mmctm *m) { return m->factored ? m->foff[m->M] : 0; }

/* src/IMMCTM.jl:186-195, then the composite K x V tables */
void orc_immctm_update_Elnphi(orc_mmctm *m)
{
    for (int i = 0; i < m->M; ++i) {
        int K = m->K[i], V = m->V[i], nf = m->nfeat[i];
        int64_t rowlen = (m->foff[i + 1] - m->foff[i]) / K;
        for (int k = 0; k < K; ++k) {
            int64_t o = m->foff[i] + (int64_t)k * rowlen;
            for (int f = 0; f < nf; ++f) {
                int Jf = m->J[i][f];
                double sg = 0.0;
                for (int j = 0; j < Jf; ++j) sg += m->gammaf[o + j];
                double ds = digamma_a(m->arith, sg);
                for (int j = 0; j < Jf; ++j) m->Elnphif[o + j] = digamma_a(m->arith, m->gammaf[o + j]) - ds;
                o += Jf;
            }
            for (int v = 0; v < V; ++v) {
                double e = 0.0, ph = 1.0;
                int64_t oo = m->foff[i] + (int64_t)k * rowlen;
                for (int f = 0; f < nf; ++f) {
                    int Jf = m->J[i][f], j = m->feat[i][(size_t)v * nf + f];
                    double sg = 0.0;
                    for (int jj = 0; jj < Jf; ++jj) sg += m->gammaf[oo + jj];
                    e += m->Elnphif[oo + j];
                    ph *= m->gammaf[oo + j] / sg;        /* phi as calculate_loglikelihoods builds it, :423-426 */
                    oo += Jf;
                }
                m->Elnphi[m->goff[i] + (size_t)k * V + v] = e;
                m->phi[m->goff[i] + (size_t)k * V + v] = ph;
            }
        }
    }
}

void orc_immctm_enable(orc_mmctm *m, const int *nfeat, const int *const *feat,
                       const double *alphaf, const double *gammaf0)
{
    int M = m->M;
    m->nfeat = (int *)malloc(sizeof(int) * M);
    m->feat = (int **)malloc(sizeof(void *) * M);
    m->J = (int **)malloc(sizeof(void *) * M);
    m->foff = (int64_t *)malloc(sizeof(int64_t) * (M + 1));
    m->aoff = (int *)malloc(sizeof(int) * (M + 1));
    m->foff[0] = 0; m->aoff[0] = 0;
    for (int i = 0; i < M; ++i) {
        int nf = nfeat[i], V = m->V[i];
        m->nfeat[i] = nf;
        m->feat[i] = (int *)malloc(sizeof(int) * (size_t)V * nf);
        memcpy(m->feat[i], feat[i], sizeof(int) * (size_t)V * nf);
        m->J[i] = (int *)calloc(nf, sizeof(int));
        int64_t rowlen = 0;
        for (int f = 0; f < nf; ++f) {                 /* J = maximum(features, dims=1), src/IMMCTM.jl:44 */
            for (int v = 0; v < V; ++v)
                if (feat[i][(size_t)v * nf + f] + 1 > m->J[i][f]) m->J[i][f] = feat[i][(size_t)v * nf + f] + 1;
            rowlen += m->J[i][f];
        }
        m->foff[i + 1] = m->foff[i] + rowlen * m->K[i];
        m->aoff[i + 1] = m->aoff[i] + nf;
    }
    int64_t T = m->foff[M];
    m->alphaf = (double *)malloc(sizeof(double) * m->aoff[M]);
    memcpy(m->alphaf, alphaf, sizeof(double) * m->aoff[M]);
    m->gammaf = (double *)malloc(sizeof(double) * T);
    m->Elnphif = (double *)malloc(sizeof(double) * T);
    memcpy(m->gammaf, gammaf0, sizeof(double) * T);
    m->factored = 1;
    orc_immctm_update_Elnphi(m);                         /* :69-70 */
    for (int64_t t = 0; t < m->goff[M]; ++t) m->gamma[t] = 0.0 / 0.0;   /* no K x V gamma in this model */
}

/* src/IMMCTM.jl:197-221 */
void orc_immctm_update_gamma(orc_mmctm *m)
{
    int64_t T = m->foff[m->M];
    for (int i = 0; i < m->M; ++i) {
        int K = m->K[i], nf = m->nfeat[i];
        int64_t rowlen = (m->foff[i + 1] - m->foff[i]) / K;
        for (int k = 0; k < K; ++k) {
            int64_t o = m->foff[i] + (int64_t)k * rowlen;
            for (int f = 0; f < nf; ++f) {
                for (int j = 0; j < m->J[i][f]; ++j) m->gammaf[o + j] = m->alphaf[m->aoff[i] + f];
                o += m->J[i][f];
            }
        }
    }
    (void)T;
    if (m->arith) {
        /* DET: sum-n-theta_kv = E_kv S_kv exactly as orc_mmctm_update_gamma (S = exactly rounded
           sum over samples of fl(exp(lambda_dk) R_dv)), then per feature value the plain sum over
           the terms that carry it, ascending v */
        int64_t G = m->goff[m->M];
        dd_t *acc = (dd_t *)calloc((size_t)G, sizeof(dd_t));
        for (int64_t d = 0; d < m->D; ++d)
            for (int i = 0; i < m->M; ++i) {
                int K = m->K[i], V = m->V[i];
                dd_t *g = acc + m->goff[i];
                const double *L = m->expl + (size_t)d * m->MK + m->koff[i];
                for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                    int v = m->term[i][w];
                    double R = (double)m->cnt[i][w] * m->rz[i][w];
                    for (int k = 0; k < K; ++k) dd_add(&g[(size_t)k * V + v], L[k] * R);
                }
            }
        for (int i = 0; i < m->M; ++i) {
            int K = m->K[i], V = m->V[i], nf = m->nfeat[i];
            int64_t rowlen = (m->foff[i + 1] - m->foff[i]) / K;
            for (int k = 0; k < K; ++k)
                for (int v = 0; v < V; ++v) {
                    int64_t x = m->goff[i] + (size_t)k * V + v;
                    double st = det_exp(m->Elnphi[x]) * dd_round(acc[x]);
                    int64_t o = m->foff[i] + (int64_t)k * rowlen;
                    for (int f = 0; f < nf; ++f) {
                        m->gammaf[o + m->feat[i][(size_t)v * nf + f]] += st;
                        o += m->J[i][f];
                    }
                }
        }
        free(acc);
    } else {
        for (int64_t d = 0; d < m->D; ++d)
            for (int i = 0; i < m->M; ++i) {
                int K = m->K[i], nf = m->nfeat[i];
                int64_t rowlen = (m->foff[i + 1] - m->foff[i]) / K;
                for (int64_t w = m->rowptr[i][d]; w < m->rowptr[i][d + 1]; ++w) {
                    int v = m->term[i][w];
                    double n = (double)m->cnt[i][w];
                    for (int k = 0; k < K; ++k) {
                        double nt = m->theta[i][(size_t)w * K + k] * n;
                        int64_t o = m->foff[i] + (int64_t)k * rowlen;
                        for (int f = 0; f < nf; ++f) {
                            m->gammaf[o + m->feat[i][(size_t)v * nf + f]] += nt;
                            o += m->J[i][f];
                        }
                    }
                }
            }
    }
    orc_immctm_update_Elnphi(m);
}

/* =====================================================================
 * Count ingest, src/utils.jl
 * ===================================================================== */

/* make_count_matrix (src/utils.jl:1-7) applied to every sample of one modality, as
 * format_counts_lda (:9-18) / format_counts_mmctm (:24-36) do column by column:
 * idx = findall(counts .> 0); rows [idx, counts[idx]] in ascending idx.  Flat result: CSR with
 * 0-BASED terms.  dense: D*V int64; layout 0: dense[v*D + d] (term-major, the TSV), 1: dense[d*V + v].
 * Pass term == NULL to only fill rowptr (sizing pass).  Returns nnz, or -1 if a count exceeds int32. */
int64_t orc_make_count_csr(int64_t D, int V, const int64_t *dense, int layout,
                           int64_t *rowptr, int32_t *term, int32_t *cnt)
{
    int64_t w = 0;
    rowptr[0] = 0;
    for (int64_t d = 0; d < D; ++d) {
        for (int v = 0; v < V; ++v) {
            int64_t x = layout == 0 ? dense[(size_t)v * D + d] : dense[(size_t)d * V + v];
            if (x > 0) {                         /* counts .> 0 */
                if (x > 2147483647LL) return -1;
                if (term) { term[w] = (int32_t)v; cnt[w] = (int32_t)x; }
                ++w;
            }
        }
        rowptr[d + 1] = w;
    }
    return w;
}
