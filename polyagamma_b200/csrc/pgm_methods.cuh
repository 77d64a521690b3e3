// pgm_methods.cuh -- the per-element sampling methods as lane-level state machines.
//
// Each method M provides
//     struct Par    per-(h, z) constants          (the reference's parameter_t)
//     struct State  per-element progress          (partial sum, pieces left)
//     setup(par, h, z, compat)                    (the reference's set_sampling_parameters)
//     begin(par, st, h, z, compat)
//     step(par, st, rng, compat) -> bool          ONE proposal attempt; true when the element is
//                                                 complete and st.acc holds PG(h, z)
// so that a persistent warp can interleave 32 elements at proposal granularity: a lane that
// rejects simply takes another step, a lane that finishes is retired and refilled.  Every lane
// consumes its element's Philox stream in exactly the order of the scalar oracle
// (oracle/pgm_oracle.c), which is what makes results independent of launch shape.
//
// Restates: src/pgm_devroye.c, src/pgm_alternate.c (+ _trunc_points.h), src/pgm_saddle.c and
// the normal / gamma-convolution samplers of src/pgm_random.c, all in FP64.
#pragma once

#include "pgm_device.cuh"

namespace pgm {

// =========================================================================================
// Devroye: exact J*(1, z), summed floor(h) times            (src/pgm_devroye.c)
// =========================================================================================
struct Devroye {
    static constexpr double T = 0.64;

    struct Par {
        double w;  // p / (p + q): probability of the truncated inverse-Gaussian piece
        double K;
        double zz;  // |z| / 2
        double z2;
    };
    struct State {
        double acc;
        uint32_t remaining;
    };

    // set_sampling_parameters (src/pgm_devroye.c:61-83)
    static __device__ __forceinline__ void setup(Par& p, double /*h*/, double z, bool /*compat*/)
    {
        p.zz = 0.5 * fabs(z);
        if (!(p.zz > 0.0)) p.zz = 0.0;  // z = NaN is sampled as z = 0
        if (p.zz > 0.0) {
            const double a = 0.8838834764831844;  // 1 / sqrt(2T)
            const double tt = 0.565685424949238;  // sqrt(T / 2)
            double b = p.zz * tt;
            double t1 = erfc(a - b) * exp(-p.zz);
            double t2 = (a + b < 26.5) ? erfc(a + b) * exp(p.zz) : 0.0;
            double pp = t1 + t2;
            p.z2 = p.zz * p.zz;
            p.K = kPi2_8 + 0.5 * p.z2;
            double q = kPi_2 * exp(-p.K * T) / p.K;
            p.w = (pp + q > 0.0) ? pp / (pp + q) : 1.0;
        }
        else {
            p.w = 0.4223027567786595;  // src/pgm_devroye.c:78
            p.K = kPi2_8;
            p.z2 = 0.0;
        }
    }

    static __device__ __forceinline__ void begin(Par&, State& s, double h, double, bool)
    {
        s.acc = 0.0;
        // `size_t hi = h` (src/pgm_devroye.c:221); counts above 2^32-1 are clamped
        s.remaining = (h >= 4294967295.0) ? 4294967295u : (uint32_t)h;
    }

    static __device__ __forceinline__ bool empty(const State& s) { return s.remaining == 0; }

    // random_right_bounded_invgauss (src/pgm_devroye.c:113-142)
    static __device__ __forceinline__ double trunc_invgauss(const Par& p, Rng& r)
    {
        double x;
        if (p.zz < 1.5625) {
            for (;;) {
                double e1, e2;
                do {
                    e1 = r.exponential();
                    e2 = r.exponential();
                } while (e1 * e1 > 3.125 * e2);
                x = 1.0 + T * e1;
                x = T / (x * x);
                if (!(p.zz > 0.0)) break;
                if (r.uniform32() < exp(-0.5 * p.z2 * x)) break;
            }
            return x;
        }
        double mu2 = 1.0 / p.z2;
        do {
            double y = r.normal();
            double w = (p.zz + 0.5 * y * y) * mu2;
            x = mu2 / (w + sqrt(fmax(w * w - mu2, 0.0)));
            if (r.uniform_co() * (1.0 + x * p.zz) > 1.0) {
                x = mu2 / x;
            }
        } while (x >= T);
        return x;
    }

    // piecewise_coef (src/pgm_devroye.c:36-46)
    static __device__ __forceinline__ double coef(int n, double lp, double e)
    {
        double a = n + 0.5;
        return kPi * a * exp(lp - e * a * a);
    }

    // One pass of the loop in random_jacobi_star (src/pgm_devroye.c:160-190)
    static __device__ __forceinline__ bool step(const Par& p, State& s, Rng& r, bool)
    {
        double x, lp, e;
        if (r.uniform_co() < p.w) {
            x = trunc_invgauss(p, r);
        }
        else {
            x = T + r.exponential() / p.K;
        }
        if (x > T) {
            lp = 0.0;
            e = 0.5 * kPi * kPi * x;
        }
        else {
            lp = -1.5 * (kLogPi_2 + log(x));
            e = 2.0 / x;
        }
        double sum = coef(0, lp, e);
        double u = r.uniform32() * sum;
        sum -= coef(1, lp, e);
        bool accept = (u <= sum);
        if (!accept) {
            accept = true;  // if the terms vanish before a decision (never in practice)
            for (int n = 2; n < 400; n += 2) {
                sum += coef(n, lp, e);
                if (u > sum) {
                    accept = false;
                    break;
                }
                sum -= coef(n + 1, lp, e);
                if (u <= sum) break;
            }
        }
        if (accept) {
            s.acc += x;
            if (--s.remaining == 0) {
                s.acc *= 0.25;
                return true;
            }
        }
        return false;
    }
};

// =========================================================================================
// Alternate: J*(h, z) for h <= 4, chunked above             (src/pgm_alternate.c)
// =========================================================================================

// Truncation points t(h), h = 1, 1.125, ..., 4: numerical data of
// src/pgm_alternate_trunc_points.h:5-23.
static __constant__ double c_trunc_f[25] = {
    1.273239366, 1.901515423, 2.281992126, 2.607829551, 2.910421526, 3.200449543, 3.482766779,
    3.759955106, 4.033540671, 4.304486011, 4.573437633, 4.840840644, 5.107017272, 5.372204821,
    5.636581947, 5.900288573, 6.163432428, 6.426094330, 6.688351603, 6.950254767, 7.211854235,
    7.473186206, 7.734284136, 7.995175158, 8.255882407};

struct Alternate {
    struct Par {
        double w_right;  // q / (p + q): probability of the truncated-gamma (right) piece
        double h, zz, z2, t, t_inv, lambda, half_h2, lgammah, hlog2, h_z, h_z2, ca;
    };
    struct State {
        double acc;
        double h_rem;   // shape still to be drawn after the current piece
        double chunk;   // 0 when the current piece is the last one
    };

    // get_truncation_point (src/pgm_alternate.c:38-70) as a true linear interpolation
    static __device__ __forceinline__ double trunc_point(double h)
    {
        if (h <= 1.0) return c_trunc_f[0];
        if (h >= 4.0) return c_trunc_f[24];
        double pos = (h - 1.0) * 8.0;
        int i = (int)pos;
        double frac = pos - i;
        return c_trunc_f[i] + (c_trunc_f[i + 1] - c_trunc_f[i]) * frac;
    }

    // set_sampling_parameters (src/pgm_alternate.c:139-176); mixture weight in log space
    static __device__ inline void setup_piece(Par& p, double h, double zz, bool compat)
    {
        p.h = h;
        p.zz = zz;
        p.t = trunc_point(h);
        p.t_inv = 1.0 / p.t;
        p.half_h2 = 0.5 * h * h;
        p.lgammah = pgm_lgamma(h);
        p.hlog2 = h * kLog2;
        p.ca = compat ? 0.5 * kLogPi_2 : kLogPi_2;  // src/pgm_alternate.c:90
        double lp;
        double a = h / sqrt(2.0 * p.t);
        if (zz > 0.0) {
            p.z2 = zz * zz;
            p.h_z = h / zz;
            p.h_z2 = p.h_z * p.h_z;
            p.lambda = kPi2_8 + 0.5 * p.z2;
            double b = zz * sqrt(0.5 * p.t);
            double hz = h * zz;
            double t1 = 0.5 * erfc(a - b) * exp(-hz);
            double t2 = (a + b < 26.5 && hz < 700.0) ? 0.5 * erfc(a + b) * exp(hz) : 0.0;
            lp = p.hlog2 + log(t1 + t2);
        }
        else {
            p.z2 = 0.0;
            p.h_z = INFINITY;
            p.h_z2 = INFINITY;
            p.lambda = kPi2_8;
            lp = p.hlog2 + log(erfc(a));
        }
        double lq = h * (kLogPi_2 - log(p.lambda)) + log_upper_gamma(h, p.lambda * p.t, true);
        p.w_right = 1.0 / (1.0 + exp(lp - lq));
        if (!(p.w_right >= 0.0)) p.w_right = 0.0;
    }

    static __device__ __forceinline__ void setup(Par&, double, double, bool) {}

    // random_polyagamma_alternate (src/pgm_alternate.c:294-323): piece schedule
    static __device__ inline void begin(Par& p, State& s, double h, double z, bool compat)
    {
        double zz = 0.5 * fabs(z);
        if (!(zz > 0.0)) zz = 0.0;
        s.acc = 0.0;
        if (h > 4.0) {
            s.chunk = (h >= 5.0) ? 4.0 : 3.0;
            s.h_rem = h;
            setup_piece(p, s.chunk, zz, compat);
        }
        else {
            s.chunk = 0.0;
            s.h_rem = 0.0;
            setup_piece(p, h, zz, compat);
        }
    }

    static __device__ __forceinline__ bool empty(const State&) { return false; }

    static __device__ __forceinline__ double inv_left_gamma(const Par& p, Rng& r)
    {
        return 1.0 / left_bounded_gamma(r, 0.5, p.half_h2, p.t_inv);
    }

    // random_right_bounded_invgauss (src/pgm_alternate.c:200-218)
    static __device__ inline double trunc_invgauss(const Par& p, Rng& r)
    {
        double x;
        if (p.t < p.h_z) {
            do {
                x = inv_left_gamma(p, r);
            } while (!(r.uniform32() < exp(-0.5 * p.z2 * x)));
            return x;
        }
        do {
            double y = r.normal();
            double w = p.h_z + 0.5 * y * y / p.z2;
            x = p.h_z2 / (w + sqrt(fmax(w * w - p.h_z2, 0.0)));
            if (r.uniform_co() * (p.h_z + x) > p.h_z) {
                x = p.h_z2 / x;
            }
        } while (x >= p.t);
        return x;
    }

    // One pass of the loop in random_jacobi_star (src/pgm_alternate.c:229-263)
    static __device__ inline bool step(Par& p, State& s, Rng& r, bool compat)
    {
        double x;
        if (r.uniform32() < p.w_right) {
            x = left_bounded_gamma(r, p.h, p.lambda, p.t);
        }
        else if (p.zz > 0.0) {
            x = trunc_invgauss(p, r);
        }
        else {
            x = inv_left_gamma(p, r);
        }
        double logx = log(x);
        double lk;
        if (x > p.t) {  // bounding_kernel (src/pgm_alternate.c:86-99)
            lk = p.h * p.ca + (p.h - 1.0) * logx - kPi2_8 * x - p.lgammah;
        }
        else {
            lk = p.hlog2 - p.half_h2 / x - 1.5 * logx - kLs2Pi + log(p.h);
        }
        double u = r.uniform32() * exp(lk);
        // piecewise_coef (src/pgm_alternate.c:75-83) with c_n = c_{n-1} (n - 1 + h) / n
        double base = p.hlog2 - kLs2Pi - 1.5 * logx;
        double inv2x = 0.5 / x;
        double c = 1.0;
        double a_prev = p.h * exp(base - p.h * p.h * inv2x);
        double sum = a_prev;
        bool accept = false;
        for (int n = 1; n < 2000; ++n) {
            double tn = 2.0 * n + p.h;
            c *= (n - 1.0 + p.h) / n;
            double a = c * tn * exp(base - tn * tn * inv2x);
            if (n & 1) {
                sum -= a;
                if (u <= sum) {
                    accept = true;
                    break;
                }
            }
            else {
                sum += a;
                if (a <= a_prev && u > sum) break;
            }
            a_prev = a;
        }
        if (!accept) return false;
        s.acc += x;
        if (s.chunk > 0.0) {
            s.h_rem -= s.chunk;
            if (!(s.h_rem > 4.0)) {
                double hr = s.h_rem;
                s.chunk = 0.0;
                s.h_rem = 0.0;
                setup_piece(p, hr, p.zz, compat);
            }
            return false;
        }
        s.acc *= 0.25;
        return true;
    }
};

// =========================================================================================
// Saddle-point approximation                                 (src/pgm_saddle.c)
// =========================================================================================
struct Saddle {
    struct Par {
        double w_left;  // p / (p + q): probability of the truncated inverse-Gaussian piece
        double h, half_z2, log_cosh_z, xl, xc, logxc;
        double slope_l, icpt_l, slope_r, icpt_r, log_coef_l, log_coef_r, log_sqrt_h2pi, hrho;
    };
    struct State {
        double acc;
    };

    static __device__ __forceinline__ double log_cosh(double s)
    {
        s = fabs(s);
        return s + log1p(exp(-2.0 * s)) - kLog2;
    }

    // cumulant_prime (src/pgm_saddle.c:100-107) with exact tan / tanh
    static __device__ __forceinline__ void kprime(double u, double& f, double& fp)
    {
        double s2 = 2.0 * u;
        if (fabs(s2) < 1e-4) {
            double g = 1.0 / 3.0 + s2 * (2.0 / 15.0 + s2 * (17.0 / 315.0));
            f = 1.0 + s2 * g;
            fp = f * f - g;
            return;
        }
        if (s2 > 0.0) {
            double s = sqrt(s2);
            f = tan(s) / s;
        }
        else {
            double s = sqrt(-s2);
            f = tanh(s) / s;
        }
        fp = f * f + (1.0 - f) / s2;
    }

    // cumulant (src/pgm_saddle.c:91-94) minus its log cosh z offset
    static __device__ __forceinline__ double cumulant0(double u)
    {
        if (u < 0.0) return -log_cosh(sqrt(-2.0 * u));
        if (u > 0.0) return -log(cos(sqrt(2.0 * u)));
        return 0.0;
    }

    // select_starting_guess (src/pgm_saddle.c:119-125)
    static __device__ __forceinline__ double guess(double x)
    {
        if (x <= 0.25) {
            double s = (1.0 - 2.0 * exp(-2.0 / x)) / x;
            return -0.5 * s * s;
        }
        return x <= 0.5 ? -1.78 : x <= 1.0 ? -0.147 : x <= 1.5 ? 0.345 : x <= 2.5 ? 0.72
               : x <= 4.0 ? 0.95 : 1.15;
    }

    // newton_raphson (src/pgm_saddle.c:140-158)
    static __device__ inline double newton(double x, double& fprime)
    {
        const double umax = 1.2337005501361697;
        double u = guess(x);
        double f, fp;
        for (int n = 0; n < 25; ++n) {
            kprime(u, f, fp);
            double fval = f - x;
            if (fabs(fval) <= 1e-9 * x || !(fp > 1e-300)) break;
            double un = u - fval / fp;
            if (un >= umax) un = 0.5 * (u + umax);
            if (un == u) break;
            u = un;
        }
        kprime(u, f, fp);
        fprime = fp;
        return u;
    }

    // invgauss_logcdf (src/pgm_saddle.c:282-292)
    static __device__ __forceinline__ double invgauss_logcdf(double x, double mu, double lambda)
    {
        double qm = x / mu;
        double tm = mu / lambda;
        double rr = sqrt(x / lambda);
        double a = log_ndtr((qm - 1.0) / rr);
        double b = 2.0 / tm + log_ndtr(-(qm + 1.0) / rr);
        return a + log1p(exp(b - a));
    }

    // set_sampling_parameters (src/pgm_saddle.c:180-230) + mixture weight (:317-328)
    static __device__ inline void setup(Par& p, double h, double z, bool compat)
    {
        double zz = 0.5 * fabs(z);
        if (!(zz > 0.0)) zz = 0.0;
        double xl, logxl;
        if (zz > 0.0) {
            xl = tanh(zz) / zz;
            logxl = log(xl);
            p.half_z2 = 0.5 * zz * zz;
            p.log_cosh_z = log_cosh(zz);
        }
        else {
            xl = 1.0;
            logxl = 0.0;
            p.half_z2 = 0.0;
            p.log_cosh_z = 0.0;
        }
        p.h = h;
        p.xl = xl;
        p.xc = 2.75 * xl;
        double xr = 3.0 * xl;
        double xc_inv = 1.0 / p.xc;
        double xl_inv = 1.0 / xl;
        double fp_r, fp_c;
        double ur = newton(xr, fp_r);
        (void)newton(p.xc, fp_c);
        double tr = ur + p.half_z2;

        p.slope_l = -0.5 * xl_inv * xl_inv;
        p.icpt_l = -0.5 * xc_inv + xl_inv;
        p.logxc = 1.0116009116784799 + logxl;
        p.slope_r = -tr - 1.0 / xr;
        p.icpt_r = (p.log_cosh_z + cumulant0(ur)) + 1.0 - 1.0986122886681098 - logxl + p.logxc;

        double alpha_r = fp_c * xc_inv * xc_inv;
        double alpha_l = xc_inv * alpha_r;
        p.log_sqrt_h2pi = 0.5 * log(h / (2.0 * kPi));
        p.log_coef_l = p.log_sqrt_h2pi - 0.5 * log(alpha_l);
        p.log_coef_r = p.log_sqrt_h2pi - 0.5 * log(alpha_r);

        double lp = h * (0.5 * xc_inv + p.icpt_l - xl_inv) + invgauss_logcdf(p.xc, xl, h) -
                    0.5 * log(alpha_l);
        p.hrho = -h * p.slope_r;
        double lq = log_upper_gamma(h, p.hrho * p.xc, false) + p.log_coef_r +
                    h * (p.icpt_r - (compat ? 0.0 : p.logxc) - log(p.hrho));
        p.w_left = 1.0 / (1.0 + exp(lq - lp));
        if (!(p.w_left >= 0.0)) p.w_left = 1.0;
    }

    static __device__ __forceinline__ void begin(Par&, State& s, double, double, bool) { s.acc = 0.0; }
    static __device__ __forceinline__ bool empty(const State&) { return false; }

    // One pass of the proposal loop of random_polyagamma_saddle (src/pgm_saddle.c:331-347)
    static __device__ inline bool step(const Par& p, State& s, Rng& r, bool compat)
    {
        const double h = p.h;
        double x;
        if (r.uniform32() < p.w_left) {
            double mu = p.xl, mu2 = p.xl * p.xl;
            do {
                double y = r.normal();
                double w = mu + 0.5 * mu2 * y * y / h;
                x = mu2 / (w + sqrt(fmax(w * w - mu2, 0.0)));
                if (r.uniform_co() * (mu + x) > mu) {
                    x = mu2 / x;
                }
            } while (x >= p.xc);
        }
        else {
            x = left_bounded_gamma(r, h, p.hrho, p.xc);
        }
        double logx = log(x);
        double lk;
        if (x > p.xc) {  // bounding_kernel (src/pgm_saddle.c:250-263)
            double sgn = compat ? 1.0 : -1.0;
            lk = h * (sgn * p.logxc + p.slope_r * x + p.icpt_r) + (h - 1.0) * logx + p.log_coef_r;
        }
        else {
            lk = 0.5 * h * (1.0 / p.xc - 1.0 / x) + h * (p.slope_l * x + p.icpt_l) - 1.5 * logx +
                 p.log_coef_l;
        }
        double fp;
        double u = newton(x, fp);  // saddle_point (src/pgm_saddle.c:235-244)
        double t = u + p.half_z2;
        double lphi = h * (p.log_cosh_z + cumulant0(u) - t * x) + p.log_sqrt_h2pi - 0.5 * log(fp);
        if (log(r.uniform32()) + lk <= lphi) {
            s.acc = 0.25 * h * x;
            return true;
        }
        return false;
    }
};

// =========================================================================================
// Normal approximation (h > 50) and truncated gamma convolution   (src/pgm_random.c)
// =========================================================================================

// random_polyagamma_normal_approx (src/pgm_random.c:47-66), cancellation-free moments
__device__ __forceinline__ double
normal_approx_sample(Rng& r, double h, double z)
{
    double a = fabs(z);
    double mean, var;
    if (!(a > 0.0)) {
        mean = 0.25 * h;
        var = h / 24.0;
    }
    else {
        double e = exp(-a);
        double ope = 1.0 + e;
        mean = 0.5 * h * ((1.0 - e) / ope) / a;
        if (a < 1.0) {
            double a2 = a * a;
            double S = 1.0 / 6.0 + a2 * (1.0 / 120.0 + a2 * (1.0 / 5040.0 + a2 * (1.0 / 362880.0 +
                       a2 * (1.0 / 39916800.0 + a2 * (1.0 / 6227020800.0 + a2 * (1.0 / 1307674368000.0))))));
            var = 0.25 * h * S * (4.0 * e / (ope * ope));
        }
        else {
            var = 0.25 * h * 2.0 * (1.0 - e * e - 2.0 * a * e) / (ope * ope * a * a * a);
        }
    }
    return mean + r.normal() * sqrt(var);
}

// random_polyagamma_gamma_conv (src/pgm_random.c:79-97): 200 Marsaglia-Tsang gammas
__device__ inline double
gamma_conv_sample(Rng& r, double h, double z)
{
    double zz = 0.5 * fabs(z);
    if (!(zz > 0.0)) zz = 0.0;
    const double pi2 = 9.869604401089358;
    double z2 = zz * zz;
    bool boost = h < 1.0;
    double d = (boost ? h + 1.0 : h) - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double inv_h = 1.0 / h;
    double acc = 0.0;
    bool has_spare = false;
    double spare = 0.0;
    for (int k = 0; k < 200; ++k) {
        double g;
        for (;;) {
            double x, v;
            do {
                if (has_spare) {
                    x = spare;
                    has_spare = false;
                }
                else {
                    r.normal_pair(x, spare);
                    has_spare = true;
                }
                v = 1.0 + c * x;
            } while (v <= 0.0);
            v = v * v * v;
            double u = r.uniform_oc();
            double x2 = x * x;
            if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) {
                g = d * v;
                break;
            }
        }
        if (boost) {
            g *= exp(log(r.uniform_oc()) * inv_h);
        }
        double m = k + 0.5;
        acc += g / (pi2 * m * m + z2);
    }
    return 0.5 * acc;
}

// random_polyagamma_hybrid (src/pgm_random.c:105-121); SIGNED z as in the reference
enum : int { kGamma = 0, kDevroye = 1, kAlternate = 2, kSaddle = 3, kHybrid = 4, kNormal = 5 };

__host__ __device__ __forceinline__ int
hybrid_select(double h, double z)
{
    if (h > 50.0) return kNormal;
    if (h >= 8.0 || (h > 4.0 && z <= 4.0)) return kSaddle;
    if (h == 1.0 || (h == floor(h) && z <= 1.0)) return kDevroye;
    return kAlternate;
}

}  // namespace pgm
