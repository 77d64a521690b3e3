// pgm_density.cu -- elementwise Polya-Gamma pdf / cdf / logpdf / logcdf series on sm_100a.
//
// Replaces src/pgm_density.c:67-331 and the per-element `dispatch` loop of
// polyagamma/_polyagamma.pyx:433-472.  Same series, same term order, same stopping rules
// (1-ulp stagnation for pdf/cdf, all 199 terms for the log versions), same tail
// approximations inside the cdf pieces (A&S 7.1.28 branch of invgamma_logcdf, Bryc branch of
// norm_logcdf) -- those are reproduced branch for branch because the 1e-10 parity target is
// against the reference's numbers, not against the exact cdf.
//
// What is different: the Gamma-function ratio Gamma(n+h) / (Gamma(h) n!) of term n is carried
// by a recurrence instead of two lgamma calls per term.  For pdf/logpdf the whole term obeys
//     term_n / term_{n-1} = (n-1+h)(2n+h) / (n (2n-2+h)) * exp(-(2n+h-1) / (2x))
// and the exponential factor is itself geometric (ratio exp(-1/x)), so the inner loop has no
// transcendental at all.  For cdf/logcdf log c_n = log c_{n-1} + log((n-1+h)/n).
#include "pgm_internal.h"

namespace pgm {

constexpr int kMaxSeriesTerms = 200;  // PGM_MAX_SERIES_TERMS (src/pgm_density.c:44-46)
constexpr double kDblEps = 2.220446049250313e-16;
constexpr double k2Pi = 6.283185307179586;

__device__ __forceinline__ bool
stagnated(double sum, double prev)
{
    // PGM_ISCLOSE(sum, prev, 0, DBL_EPSILON)  (src/pgm_macros.h:47-48, src/pgm_density.c:87)
    return fabs(sum - prev) <= kDblEps * fmax(fabs(sum), fabs(prev));
}

// Gamma-ratio factor of term n over term n - 1.  The quotient is formed with the MUFU-seeded
// reciprocal (<= 1 ulp; an IEEE division costs ~35 issue slots in the innermost loop) whenever the
// denominator is a normal number, i.e. unless h is below 1e-290.
__device__ __forceinline__ double
term_ratio(double dn, double h, bool fast)
{
    const double num = (dn - 1.0 + h) * (2.0 * dn + h), den = dn * (2.0 * dn - 2.0 + h);
    return fast ? fm::div(num, den) : num / den;
}

// pgm_polyagamma_pdf (src/pgm_density.c:67-92)
__device__ __forceinline__ double
pg_pdf(double x, double h, double z)
{
    if (isnan(x)) return x;
    if (x <= 0.0 || isinf(x)) return 0.0;
    double a = (fabs(z) > 0.0 ? h * log(cosh(0.5 * z)) - 0.5 * z * z * x : 0.0) + (h - 1.0) * kLog2;
    double inv_x = 1.0 / x;
    double term = exp(a - 0.125 * h * h * inv_x) * h;
    double sum = term;
    double q = exp(-inv_x);
    double g = exp(-0.5 * (h + 1.0) * inv_x);
    double sign = -1.0;
    const bool fast = h > 1e-290 && h < 1e150;
    for (int n = 1; n < kMaxSeriesTerms; ++n, sign = -sign) {
        double dn = (double)n;
        term *= term_ratio(dn, h, fast) * g;
        g *= q;
        double prev = sum;
        sum += sign * term;
        if (stagnated(sum, prev)) break;
    }
    return sum / sqrt(k2Pi * x * x * x);
}

// pgm_polyagamma_logpdf (src/pgm_density.c:100-121)
__device__ __forceinline__ double
pg_logpdf(double x, double h, double z)
{
    if (isnan(x)) return x;
    if (x <= 0.0 || isinf(x)) return -INFINITY;
    double inv_x = 1.0 / x;
    double a = (fabs(z) > 0.0 ? h * log(cosh(0.5 * z)) - 0.5 * z * z * x : 0.0) + (h - 1.0) * kLog2 - kLs2Pi -
               1.5 * log(x);
    double first = -0.125 * h * h * inv_x;  // the reference's lgamma(h) cancels against `a`
    double term = 1.0, sum = 1.0, sign = -1.0, prev_term = 1.0;
    double q = exp(-inv_x);
    double g = exp(-0.5 * (h + 1.0) * inv_x);
    const bool fast = h > 1e-290 && h < 1e150;
    for (int n = 1; n < kMaxSeriesTerms; ++n, sign = -sign) {
        double dn = (double)n;
        term *= term_ratio(dn, h, fast) * g;
        g *= q;
        sum += sign * term;
        // The reference adds all 199 terms.  Past their peak the terms decrease monotonically
        // (log-concave in n), so once one is below 1e-30 |sum| neither it nor any later term can
        // change the double-precision sum: stopping here is bit-identical to going on.
        if (term < prev_term && term <= 1e-30 * fabs(sum)) break;
        prev_term = term;
    }
    return a + (log(h) + first + log(sum));
}

// invgamma_logcdf (src/pgm_density.c:149-165)
__device__ __forceinline__ double
invgamma_logcdf(double a, double s2x)
{
    double x = 0.5 * a / s2x;
    if (x > 26.55) {
        const double p0 = 0.3275911, p1 = 0.254829592, p2 = -0.284496736, p3 = 1.421413741, p4 = -1.453152027,
                     p5 = 1.061405429;
        double t = 1.0 / (1.0 + p0 * x);
        return log(t * (p1 + t * (p2 + t * (p3 + t * (p4 + t * p5))))) - x * x;
    }
    return log(erfc(x));
}

// norm_logcdf (src/pgm_density.c:181-209)
__device__ __forceinline__ double
norm_logcdf(double x)
{
    if (x < -37.5) {
        const double p0 = 12.77436324, p1 = 5.575192695, q0 = 25.54872648, q1 = 31.53531977, q2 = 14.38718147,
                     s2p = 2.5066282746310002;
        x = -x;
        double rr = (p0 + x * (p1 + x)) / (x * x * x * s2p + q0 + x * (q1 + q2 * x));
        return log(rr) - 0.5 * x * x;
    }
    double y = x * 0.7071067811865475;
    double z = fabs(y);
    if (z < 1.0) {
        return log(0.5 + 0.5 * erf(y));
    }
    double a = 0.5 * erfc(z);
    if (y > 0.0) {
        return log1p(-a);
    }
    return log(a);
}

// invgauss_logcdf (src/pgm_density.c:216-225)
__device__ __forceinline__ double
invgauss_logcdf(double a, double s2x, double x, double z)
{
    double qm = 2.0 * x * z / a;
    double rr = 2.0 * s2x / a;
    double la = norm_logcdf((qm - 1.0) / rr);
    double lb = z * a + norm_logcdf(-(qm + 1.0) / rr);
    return la + log1p(exp(lb - la));
}

__device__ __forceinline__ double
piece_logcdf(bool ig, double a, double s2x, double x, double z)
{
    return ig ? invgauss_logcdf(a, s2x, x, z) : invgamma_logcdf(a, s2x);
}

// The same piece in linear space, for the (overwhelmingly common) arguments where the reference
// takes none of its tail approximations:
//   z > 0:  F = Phi(A) + e^{z a} Phi(-B),  A = z sqrt(x) - a / (2 sqrt(x)),  B = z sqrt(x) + a / (2 sqrt(x))
//             = 1/2 erfc(-A / sqrt 2) + 1/2 erfcx(B / sqrt 2) e^{-A^2 / 2}        (z a - B^2/2 = -A^2/2)
//   z = 0:  F = erfc(a / (2 sqrt(2x)))
// Two single-path erfcx and one exp, no branch, instead of two (erf|erfc) + two log + log1p + exp
// behind four data-dependent branches.  Returns false where the reference switches to the Bryc
// (A < -37.5 or B > 37.5, src/pgm_density.c:184-194) or A&S 7.1.28 (:154-163) forms: the caller
// then takes the log-space path, which reproduces them.
__device__ __forceinline__ bool
piece_cdf_fast(bool ig, double a, double sx, double half_rsx, double zsx, double& F)
{
    if (ig) {
        const double t = a * half_rsx;  // a / (2 sqrt(x))
        const double A = zsx - t, B = zsx + t;
        if (A < -37.5 || B > 37.5) return false;
        // with a' = A / sqrt 2, b' = B / sqrt 2 >= |a'| and E = e^{-a'^2}:
        //   a' <  0:  F = E/2 (erfcx(-a') + erfcx(b'))        (erfc(-a') = E erfcx(-a'))
        //   a' >= 0:  F = 1 - E/2 (erfcx(a') - erfcx(b'))     (erfc(-a') = 2 - E erfcx(a'))
        // two single-path erfcx and ONE exponential
        const double ap = 0.7071067811865475 * A, bp = 0.7071067811865475 * B;
        const double e1 = fm::erfcx_pos(fabs(ap)), e2 = fm::erfcx_pos(bp);
        const double hE = 0.5 * fm::exp_neg_sq(ap);
        F = ap < 0.0 ? hE * (e1 + e2) : 1.0 - hE * (e1 - e2);
        return true;
    }
    const double t = 0.5 * a / sx;  // sx = sqrt(2x)
    if (t > 26.55) return false;
    F = fm::erfcx_pos(t) * fm::exp_neg_sq(t);
    return true;
}

// pgm_polyagamma_cdf (src/pgm_density.c:236-280)
__device__ __forceinline__ double
pg_cdf(double x, double h, double z)
{
    if (isnan(x)) return x;
    if (x <= 0.0) return 0.0;
    if (isinf(x)) return 1.0;
    z = fabs(z);
    const bool ig = z > 0.0;
    double c, s2x;
    if (ig) {
        c = h * log1p(exp(-z));
        s2x = sqrt(x);
    }
    else {
        c = h * kLog2;
        s2x = sqrt(2.0 * x);
    }
    const double half_rsx = 0.5 / s2x, zsx = z * s2x, ez = exp(-z);
    // ck_n = e^c Gamma(n+h) / (Gamma(h) n!) e^{-n z}, by recurrence; for h > 300 it can overflow and
    // the series (hopelessly ill-conditioned there anyway) stays in log space like the reference
    const bool linear = h <= 300.0;
    double ck = exp(c), lc = 0.0;
    double F;
    double sum = (linear && piece_cdf_fast(ig, h, s2x, half_rsx, zsx, F)) ? ck * F
                                                                           : exp(c + piece_logcdf(ig, h, s2x, x, z));
    double sign = -1.0;
    for (int n = 1; n < kMaxSeriesTerms; ++n, sign = -sign) {
        const double dn = (double)n;
        const double a = 2.0 * dn + h;
        double term;
        if (linear) {
            ck *= fm::div(dn - 1.0 + h, dn) * ez;
            term = piece_cdf_fast(ig, a, s2x, half_rsx, zsx, F) ? ck * F
                                                                 : exp(log(ck) + piece_logcdf(ig, a, s2x, x, z));
        }
        else {
            lc += log((dn - 1.0 + h) / dn);
            term = exp(c + lc + piece_logcdf(ig, a, s2x, x, z) - z * dn);
        }
        const double prev = sum;
        sum += sign * term;
        if (stagnated(sum, prev)) break;
    }
    return sum;
}

// pgm_polyagamma_logcdf (src/pgm_density.c:291-331)
__device__ __forceinline__ double
pg_logcdf(double x, double h, double z)
{
    if (isnan(x)) return x;
    if (x <= 0.0) return -INFINITY;
    if (isinf(x)) return 0.0;
    z = fabs(z);
    const bool ig = z > 0.0;
    double c, s2x;
    if (ig) {
        c = h * log1p(exp(-z));
        s2x = sqrt(x);
    }
    else {
        c = h * kLog2;
        s2x = sqrt(2.0 * x);
    }
    const double half_rsx = 0.5 / s2x, zsx = z * s2x, ez = exp(-z);
    const double first = piece_logcdf(ig, h, s2x, x, z);  // the reference's lgamma(h) cancels
    // ratios F_n / F_0 are formed in linear space only while both are comfortably normal numbers
    double F0 = 0.0;
    const bool linear = h <= 300.0 && first > -600.0 && piece_cdf_fast(ig, h, s2x, half_rsx, zsx, F0);
    double ck = linear ? 1.0 / F0 : 1.0;  // Gamma(n+h) / (Gamma(h) n!) e^{-n z} / F_0
    double lc = 0.0;
    double sum = 1.0, sign = -1.0, prev_term = 1.0;
    for (int n = 1; n < kMaxSeriesTerms; ++n, sign = -sign) {
        const double dn = (double)n;
        const double a = 2.0 * dn + h;
        double F;
        double term;
        if (linear) {
            ck *= fm::div(dn - 1.0 + h, dn) * ez;
            term = (piece_cdf_fast(ig, a, s2x, half_rsx, zsx, F) && F > 1e-280)
                       ? ck * F
                       : exp(log(ck) + piece_logcdf(ig, a, s2x, x, z));
        }
        else {
            lc += log((dn - 1.0 + h) / dn);
            term = exp(lc + piece_logcdf(ig, a, s2x, x, z) - z * dn - first);
        }
        sum += sign * term;
        // same argument as in pg_logpdf: later terms cannot change the sum any more
        if (term < prev_term && term <= 1e-30 * fabs(sum)) break;
        prev_term = term;
    }
    return c + (first + log(sum));
}

// ---------------------------------------------------------------------------------------
// Cancellation-free densities for large h: which = 4..7
// ---------------------------------------------------------------------------------------
// The alternating series above (the reference's, src/pgm_density.c:80-86,268-272) loses cond = sum|term| /
// |sum term| digits: everything for h >~ 40 at small |z| (8-DEFECTS g).  The opt-in "stable" forms invert the
// Laplace transform instead,
//     E exp(-t X) = [cosh(z/2) / cosh r(t)]^h,   r(t) = sqrt(z^2/4 + t/2),    X ~ PG(h, z),
//     pdf(x) = 1/(2 pi i) int exp(psi(t)) dt,       psi(t) = t x + h (log cosh(z/2) - log cosh r(t)),
//     cdf(x) = 1/(2 pi i) int exp(psi(t)) / t dt    (Re t = c > 0;  c < 0 gives cdf(x) - 1),
// along the vertical line through the real saddle point t0 of psi (psi'(t0) = 0 is the equation
// K'(u) = 4x/h of the saddle-point sampler, t0 = -4u - z^2/2): there the integrand is largest at y = 0 and falls
// off like a Gaussian of width 1/sigma, sigma^2 = psi''(t0) = h K''(u) / 16, so nothing cancels, and the
// trapezoid rule in y converges geometrically (analytic integrand).  Step 0.2 / sigma: the nearest singularity
// of the cdf integrand (the pole at t = 0) is kept at least 1.1 / sigma = 5.5 steps away by moving the line to
// c = +1.1 / sigma when |t0| is smaller than that (at most a factor e^{2.4} off the saddle), which bounds the
// quadrature error by e^{-2 pi 5.5} ~ 1e-15.  The sum stops when the integrand has fallen below 1e-18 of it.
// Checked against the same series in 300-digit arithmetic (tests/golden/make_stable_golden.py): <= 1e-12
// relative for h >= 8 wherever the result is above 1e-290.  Below h = 8 the series itself is well conditioned
// and is used (its answers and these agree to 1e-10 on the overlap, tests/test_gpu_parity.py).
struct cplx {
    double re, im;
};
// K'(u) and K''(u) of the cumulant of J*(1, 0): tanh(s)/s with s = sqrt(-2u) for u < 0, tan(s)/s with
// s = sqrt(2u) for u > 0 (src/pgm_saddle.c:100-107, exact functions), K'' = K'^2 + (1 - K') / (2u)
__device__ __forceinline__ void
stable_kprime(double u, double& f, double& fp)
{
    const double s2 = 2.0 * u;
    if (fabs(s2) < 1e-4) {
        const double g = 1.0 / 3.0 + s2 * (2.0 / 15.0 + s2 * (17.0 / 315.0));
        f = 1.0 + s2 * g;
        fp = f * f - g;
        return;
    }
    const double s = ::sqrt(fabs(s2));
    f = (s2 > 0.0 ? tan(s) : tanh(s)) / s;
    fp = f * f + (1.0 - f) / s2;
}
// root u of K'(u) = xp (xp > 0): Newton's iteration kept inside a bracket
__device__ double
stable_saddle(double xp, double& fp)
{
    const double kUMax = 1.2337005501361697;  // pi^2 / 8: the pole of tan
    double s_lo = 1.0 / xp + 1.0;
    double lo = -0.5 * s_lo * s_lo, hi = kUMax;  // K'(lo) < xp < K'(hi -> pole)
    double u = (xp < 1.0) ? -0.5 / (xp * xp) : (xp < 1.5 ? 0.3 : kUMax * (1.0 - 0.4 / xp));
    if (!(u > lo)) u = 0.5 * (lo + hi);
    double f = 1.0;
    for (int it = 0; it < 80; ++it) {
        stable_kprime(u, f, fp);
        const double g = f - xp;
        if (fabs(g) <= 1e-14 * xp) break;
        if (g > 0.0) hi = u;
        else lo = u;
        double un = u - g / fp;
        if (!(un > lo && un < hi)) un = 0.5 * (lo + hi);
        if (un == u) break;
        u = un;
    }
    stable_kprime(u, f, fp);
    return u;
}
__device__ __forceinline__ cplx
csqrt_pos(cplx a)  // principal square root (Re >= 0)
{
    const double m = ::sqrt(a.re * a.re + a.im * a.im);
    cplx r;
    if (a.re >= 0.0) {
        r.re = ::sqrt(0.5 * (m + a.re));
        r.im = (r.re > 0.0) ? 0.5 * a.im / r.re : 0.0;
    }
    else {
        const double q = ::sqrt(0.5 * (m - a.re));
        r.re = 0.5 * fabs(a.im) / q;
        r.im = (a.im >= 0.0) ? q : -q;
    }
    return r;
}
// log cosh r for Re r >= 0: r - log 2 + log(1 + e^{-2r})
__device__ __forceinline__ cplx
clogcosh(cplx r)
{
    const double e = ::exp(-2.0 * r.re);
    double sn, cs;
    sincos(-2.0 * r.im, &sn, &cs);
    const double ar = e * cs, ai = e * sn;  // a = e^{-2r}, |a| <= 1
    cplx v;
    v.re = r.re - kLog2 + 0.5 * log1p(2.0 * ar + ar * ar + ai * ai);
    v.im = r.im + atan2(ai, 1.0 + ar);
    return v;
}

// mode 0: log pdf; 1: log of cdf-or-survival, `upper` says which: upper = false -> log cdf(x),
// upper = true -> log(1 - cdf(x)) (taken whenever the saddle lies left of the pole: x well above the mean)
__device__ double
pg_contour(int mode, double x, double h, double z, bool& upper)
{
    upper = false;
    const double az = fabs(z), hz = 0.5 * az;
    const double lcz = hz - kLog2 + log1p(::exp(-2.0 * hz));  // log cosh(z/2)
    // real saddle point: K'(u) = 4x/h
    const double xp = 4.0 * x / h;
    double fp;
    const double u = stable_saddle(xp, fp);
    const double t0 = -4.0 * u - 0.5 * az * az;
    const double sigma = ::sqrt(0.0625 * h * fp);
    const double dy = 0.2 / sigma;
    double c = t0;
    if (mode == 1) {
        const double cmin = 1.1 / sigma;
        if (fabs(t0) < cmin) c = cmin;
        upper = c < 0.0;
    }
    // psi at the real point c (the integrand is normalised by it)
    double psi0;
    {
        const double w = 0.25 * az * az + 0.5 * c;  // r^2, real; negative: r = i rho, cosh r = cos rho
        const double lcr = (w >= 0.0) ? (::sqrt(w) - kLog2 + log1p(::exp(-2.0 * ::sqrt(w)))) : ::log(cos(::sqrt(-w)));
        psi0 = c * x + h * (lcz - lcr);
    }
    double sum = (mode == 0) ? 0.5 : 0.5 / c;  // the y = 0 term, weight 1/2
    int small = 0;
    for (int k = 1; k < 4096; ++k) {
        const double y = k * dy;
        cplx r2{0.25 * az * az + 0.5 * c, 0.5 * y};
        const cplx lcr = clogcosh(csqrt_pos(r2));
        const double pre = c * x + h * (lcz - lcr.re) - psi0, pim = y * x - h * lcr.im;
        const double mag = ::exp(pre);
        double sn, cs;
        sincos(pim, &sn, &cs);
        double term;
        if (mode == 0) {
            term = mag * cs;
        }
        else {  // Re(e^{i pim} / (c + i y)) = (c cos + y sin) / (c^2 + y^2)
            term = mag * (c * cs + y * sn) / (c * c + y * y);
        }
        sum += term;
        const double scale = (mode == 0) ? 1.0 : 1.0 / hypot(c, y);
        if (mag * scale < 1e-18 * fabs(sum)) {
            if (++small >= 3) break;
        }
        else {
            small = 0;
        }
    }
    // pdf = e^{psi0} dy / pi * sum;  cdf (c > 0) or cdf - 1 (c < 0) = the same with 1/t under the integral
    const double lg = psi0 + ::log(dy * 0.3183098861837907 * fabs(sum));
    if (mode == 1 && upper && !(sum < 0.0)) return -INFINITY;  // (cdf - 1 must be negative: rounding at cdf = 1)
    return lg;
}

constexpr double kStableMinH = 8.0;  // below: the series is well conditioned (and the integrand decays slowly)

__device__ double
pg_pdf_stable(double x, double h, double z, bool want_log)
{
    if (!(h >= kStableMinH)) return want_log ? pg_logpdf(x, h, z) : pg_pdf(x, h, z);
    if (!(x > 0.0) || isinf(x)) return want_log ? -INFINITY : 0.0;
    bool upper;
    const double l = pg_contour(0, x, h, z, upper);
    return want_log ? l : ::exp(l);
}

__device__ double
pg_cdf_stable(double x, double h, double z, bool want_log)
{
    if (!(h >= kStableMinH)) return want_log ? pg_logcdf(x, h, z) : pg_cdf(x, h, z);
    if (!(x > 0.0)) return want_log ? -INFINITY : 0.0;
    if (isinf(x)) return want_log ? 0.0 : 1.0;
    bool upper;
    const double l = pg_contour(1, x, h, z, upper);
    if (!upper) return want_log ? fmin(l, 0.0) : fmin(::exp(l), 1.0);
    const double sf = ::exp(l);  // 1 - cdf
    return want_log ? log1p(-sf) : 1.0 - sf;
}

template <int WHICH>
__global__ void __launch_bounds__(256)
density_kernel(DensityArgs a, uint64_t n)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        int64_t ox = i, oh = i, oz = i;
        if (a.bc.ndim != 0) {
            ox = oh = oz = 0;
            uint64_t rem = i;
            for (int d = a.bc.ndim - 1; d >= 0; --d) {
                uint64_t qd = rem / (uint64_t)a.bc.shape[d];
                int64_t k = (int64_t)(rem - qd * (uint64_t)a.bc.shape[d]);
                ox += k * a.bc.xs[d];
                oh += k * a.bc.hs[d];
                oz += k * a.bc.zs[d];
                rem = qd;
            }
        }
        double x = a.x[ox], h = a.h[oh], z = a.z[oz];
        double v;
        if (WHICH == 0) v = pg_pdf(x, h, z);
        else if (WHICH == 1) v = pg_cdf(x, h, z);
        else if (WHICH == 2) v = pg_logpdf(x, h, z);
        else if (WHICH == 3) v = pg_logcdf(x, h, z);
        else if (WHICH == 4) v = pg_pdf_stable(x, h, z, false);
        else if (WHICH == 5) v = pg_cdf_stable(x, h, z, false);
        else if (WHICH == 6) v = pg_pdf_stable(x, h, z, true);
        else v = pg_cdf_stable(x, h, z, true);
        a.out[i] = v;
    }
}

int
launch_density(int which, DensityArgs a, uint64_t n, cudaStream_t s)
{
    if (n == 0) return 0;
    if (which < 0 || which > 7) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t need = (n + 255) / 256;
    uint64_t full = (uint64_t)sms * 8;
    uint32_t grid = (uint32_t)(need < full ? need : full);
    ScopedKernelTimer timer(kKindDensity, s);
    switch (which) {
        case 0: density_kernel<0><<<grid, 256, 0, s>>>(a, n); break;
        case 1: density_kernel<1><<<grid, 256, 0, s>>>(a, n); break;
        case 2: density_kernel<2><<<grid, 256, 0, s>>>(a, n); break;
        case 3: density_kernel<3><<<grid, 256, 0, s>>>(a, n); break;
        case 4: density_kernel<4><<<grid, 256, 0, s>>>(a, n); break;
        case 5: density_kernel<5><<<grid, 256, 0, s>>>(a, n); break;
        case 6: density_kernel<6><<<grid, 256, 0, s>>>(a, n); break;
        default: density_kernel<7><<<grid, 256, 0, s>>>(a, n); break;
    }
    count_launch(1);
    return (int)cudaGetLastError();
}

}  // namespace pgm
