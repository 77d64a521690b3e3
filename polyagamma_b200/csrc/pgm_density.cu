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
        else v = pg_logcdf(x, h, z);
        a.out[i] = v;
    }
}

int
launch_density(int which, DensityArgs a, uint64_t n, cudaStream_t s)
{
    if (n == 0) return 0;
    if (which < 0 || which > 3) return -1;
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
        default: density_kernel<3><<<grid, 256, 0, s>>>(a, n); break;
    }
    count_launch(1);
    return (int)cudaGetLastError();
}

}  // namespace pgm
