// pgm_methods.cuh -- the per-element sampling methods, split the way the GPU wants them.
//
// The reference does "set_sampling_parameters(h, z); loop { propose; accept? }" per element
// (fill2 pays the setup for every element).  Here each method M is cut into
//
//   make_record(h, z, compat, rec[])   dense, divergence-free part: everything that depends
//                                      only on (h, z) and is expensive (erfc, lgamma, incomplete
//                                      gamma, Newton solves ...).  Run by setup_kernel<M> over a
//                                      whole bucket, one thread per element, result written as
//                                      M::kFields doubles per element (structure of arrays).
//   load(rec, par, st)                 cheap: rebuild the lane's working set from the record
//   step(par, st, rng) -> bool         ONE proposal attempt; true when the element is complete
//                                      and st.acc holds PG(h, z)
//
// so the persistent sampling kernel only carries the small `step` code (instruction-cache
// resident) and a warp interleaves 32 elements at proposal granularity: a lane that rejects
// just takes another step, a lane that finishes is retired and refilled.  Every lane consumes
// its element's Philox stream in exactly the order of the scalar oracle (oracle/pgm_oracle.c),
// which is what makes results independent of launch shape.
//
// Restates: src/pgm_devroye.c, src/pgm_alternate.c (+ _trunc_points.h), src/pgm_saddle.c and
// the normal / gamma-convolution samplers of src/pgm_random.c, all in FP64.
#pragma once

#include "pgm_device.cuh"

namespace pgm {

// Unqualified log / exp / sqrt in this header are the lean fm:: versions (same results to ~1 ulp
// for finite normal arguments, libdevice fallback otherwise).
using fm::exp;
using fm::log;
using fm::sqrt;

constexpr int kMaxRecordFields = 8;  // saddle-full is the largest record

// Which sampling kernels keep a single out-of-line copy of the Philox block function (bit i =
// method i in launch order: 0 devroye, 1 alternate, 2 saddle-full, 3 saddle-left, 4 saddle-far,
// 5 alternate-left, 6 gamma).  Chosen per kernel by measurement (profiles/README.md).
#ifndef PGM_RNG_OUTLINE_MASK
#define PGM_RNG_OUTLINE_MASK 0x3e
#endif
template <int BIT>
using RngFor = RngT<((PGM_RNG_OUTLINE_MASK >> BIT) & 1) != 0>;

// =========================================================================================
// Devroye: exact J*(1, z), summed floor(h) times            (src/pgm_devroye.c)
// =========================================================================================
// LEFT fixes the proposal of the truncated inverse-Gaussian piece at compile time: the hybrid
// dispatch buckets devroye elements by |z|/2 < 1/T (exponential pairs) or not (Michael-Schucany-Haas)
// so that the lanes of a warp do not serialise the two (cfg2, z ~ N(0, 3): 70 % / 30 %; 13.5 of 32
// lanes active when mixed); the explicit devroye method goes through the same bucketing.
enum : int { kDevroyePair = 1, kDevroyeMSH = 2 };

template <int LEFT>
struct DevroyeT {
    using Rng = RngFor<0>;
    static constexpr double T = 0.64;
    static constexpr int kFields = 4;  // w, K, zz, count

    struct Par {
        double w;   // p / (p + q): probability of the truncated inverse-Gaussian piece
        double inv_K;
        double zz;  // |z| / 2
        double z2;
        double mu2;  // 1 / z2
    };
    struct State {
        double acc;
        uint32_t remaining;
        uint32_t in_left;  // 1: inside the truncated inverse-Gaussian retry loop (no new mixture draw)
    };

    // set_sampling_parameters (src/pgm_devroye.c:61-83)
    static __device__ __forceinline__ void make_record(double h, double z, bool /*compat*/, double* rec)
    {
        double zz = 0.5 * fabs(z);
        if (!(zz > 0.0)) zz = 0.0;  // z = NaN is sampled as z = 0
        double w, K;
        if (zz > 0.0) {
            // p = erfc(a - b) e^{-zz} + erfc(a + b) e^{zz} and q = (pi/2) e^{-K T} / K with a = 1 / sqrt(2T),
            // b = zz sqrt(T/2), 2ab = zz: every term carries E = e^{-(a^2 + b^2)} --
            //   erfc(a -+ b) e^{-+zz} = erfcx(a -+ b) E,      e^{-K T} = E e^{a^2 - pi^2 T / 8}
            // -- so p / (p + q) needs no exponential at all while a >= b, and one (of (b - a)^2, through
            // erfc(a - b) = 2 - erfc(b - a)) otherwise: two single-path erfcx instead of two erfc and
            // three exp.
            const double a = 0.8838834764831844;  // 1 / sqrt(2T)
            const double tt = 0.565685424949238;  // sqrt(T / 2)
            const double b = zz * tt;
            const double d = b - a;
            K = kPi2_8 + 0.5 * zz * zz;
            const double Q = 1.557784085126967 * fm::rcp(K);  // (pi/2) e^{a^2 - pi^2 T / 8} / K
            const double e1 = fm::erfcx_pos(fabs(d)), e2 = fm::erfcx_pos(a + b);
            if (d <= 0.0) {
                const double P = e1 + e2;
                w = fm::div(P, P + Q);
            }
            else if (d * d < 600.0) {
                const double P = 2.0 * exp(d * d) - e1 + e2;
                w = fm::div(P, P + Q);
            }
            else {
                w = 1.0;  // q / p < 1e-260
            }
        }
        else {
            w = 0.4223027567786595;  // src/pgm_devroye.c:78
            K = kPi2_8;
        }
        rec[0] = w;
        rec[1] = K;
        rec[2] = zz;
        // `size_t hi = h` (src/pgm_devroye.c:221); counts above 2^32-1 are clamped
        rec[3] = (h >= 4294967295.0) ? 4294967295.0 : floor(h);
    }

    // false: nothing to draw (floor(h) == 0), the sample is 0
    template <class Rec>
    static __device__ __forceinline__ bool load(const Rec& rec, Par& p, State& s, bool)
    {
        p.w = rec[0];
        p.inv_K = fm::rcp(rec[1]);
        p.zz = rec[2];
        p.z2 = p.zz * p.zz;
        p.mu2 = (p.zz > 0.0) ? fm::rcp(p.z2) : 0.0;
        s.acc = 0.0;
        s.remaining = (uint32_t)rec[3];
        s.in_left = 0;
        return s.remaining != 0;
    }

    // piecewise_coef (src/pgm_devroye.c:36-46)
    static __device__ __forceinline__ double coef(int n, double lp, double e)
    {
        double a = n + 0.5;
        return kPi * a * exp(lp - e * a * a);
    }

    // One trip of random_jacobi_star (src/pgm_devroye.c:160-190) with the retry loops of
    // random_right_bounded_invgauss (:113-142) FLATTENED into the trip: a lane whose exponential
    // pair / Michael-Schucany-Haas draw fails returns with in_left set and simply proposes again
    // next trip, instead of spinning while the other 31 lanes wait.  A trip consumes exactly EIGHT words of
    // the element's stream (two Philox blocks, generated by all lanes together at its top; oracle/pgm_oracle.c
    // dev_jacobi_star states which word is what): lanes that refill a four-word buffer whenever it runs dry do
    // so alone -- 11 of 32 lanes active in the Philox rounds, a quarter of the kernel's instructions.  All three
    // proposal kinds start from the same quantity L = -log(U): the tail exponential, the first exponential of
    // the pair, half the squared Box-Muller radius -- so that logarithm is computed once, convergently.
    static __device__ __forceinline__ bool step(Par& p, State& s, Rng& r, bool)
    {
        uint32_t w[8];
        r.two_blocks(w);
        double x, lp, e;
        const bool left = s.in_left ? true : (Rng::to_co(w[0], w[1]) < p.w);
        const double L = -fm::log_pn(Rng::to_oc(w[2], w[3]));
        if (left) {
            s.in_left = 1;
            constexpr bool pair = LEFT == kDevroyePair;
            if (pair) {  // |z|/2 < 1/T: exponential-pair proposal of Devroye (2009)
                const double e2 = -fm::log_pn(Rng::to_oc(w[4], w[5]));
                if (L * L > 3.125 * e2) return false;
                x = 1.0 + T * L;
                x = T * fm::rcp(x * x);
                if (p.zz > 0.0 && !(Rng::to_u32(w[6]) < exp(-0.5 * p.z2 * x))) return false;
            }
            else {  // IG(1/z, 1) by Michael-Schucany-Haas; only the square of the normal is needed
                const double c = fm::cos2pi(Rng::to_co(w[4], w[5]));
                const double y2 = 2.0 * L * c * c;
                const double mu2 = p.mu2;
                const double ww = (p.zz + 0.5 * y2) * mu2;
                x = mu2 * fm::rcp(ww + fm::sqrt_nn(ww * ww - mu2));
                if (Rng::to_u32(w[6]) * (1.0 + x * p.zz) > 1.0) {
                    x = mu2 * fm::rcp(x);
                }
                if (x >= T) return false;
            }
            s.in_left = 0;
        }
        else {
            x = T + L * p.inv_K;
        }
        if (x > T) {
            lp = 0.0;
            e = 0.5 * kPi * kPi * x;
        }
        else {
            lp = -1.5 * (kLogPi_2 + fm::log_pn(x));
            e = 2.0 * fm::rcp(x);
        }
        double sum = coef(0, lp, e);
        double u = Rng::to_u32(w[7]) * sum;
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
    using Rng = RngFor<1>;
    // h, zz, then (w_right, lgamma(h)) of the first piece and of the remainder piece
    static constexpr int kFields = 6;

    struct Par {
        double w_right;  // q / (p + q): probability of the truncated-gamma (right) piece
        double h, zz, z2, t, t_inv, lambda, half_h2, lgammah, hlog2, h_z, h_z2, ca, inv_z2, logh;
    };
    struct State {
        double acc;
        double h_rem;  // shape still to be drawn, including the current piece (chunked only)
        double chunk;  // 0 when the current piece is the last one
        double w_last, lg_last;  // record of the remainder piece
    };

    // get_truncation_point (src/pgm_alternate.c:38-70) as a true linear interpolation.
    // PGM_TRUNC_SCALE (build-time, default 1) rescales the table: 0.5 gives the acceptance-optimal points (the
    // reference's generator, scripts/generate_alternate_truncation_pts.py:53-83, doubles the argmin).  The A/B
    // is in profiles/r2_trunc_table_ab.txt (tools/ab_trunc_table.sh): fewer proposals per sample, but slower on
    // the B200 -- the smaller t moves proposals to the costlier pieces and lengthens the setup's continued
    // fraction -- so the reference's table stays.
#ifndef PGM_TRUNC_SCALE
#define PGM_TRUNC_SCALE 1.0
#endif
    static __device__ __forceinline__ double trunc_point(double h)
    {
        constexpr double ks = PGM_TRUNC_SCALE;
        if (h <= 1.0) return ks * c_trunc_f[0];
        if (h >= 4.0) return ks * c_trunc_f[24];
        double pos = (h - 1.0) * 8.0;
        int i = (int)pos;
        double frac = pos - i;
        return ks * (c_trunc_f[i] + (c_trunc_f[i + 1] - c_trunc_f[i]) * frac);
    }

    // the expensive half of set_sampling_parameters (src/pgm_alternate.c:139-176): the mixture
    // weight q / (p + q), in log space, and lgamma(h)
    static __device__ inline void piece_weights(double h, double zz, double& w_right, double& lgammah)
    {
        double t = trunc_point(h);
        lgammah = pgm_lgamma(h);
        double hlog2 = h * kLog2;
        double lp, lambda;
        double a = h * fm::rsqrt(2.0 * t);
        if (zz > 0.0) {
            lambda = kPi2_8 + 0.5 * zz * zz;
            // invgauss_cdf (src/pgm_alternate.c:104-114): 1/2 [erfc(a - b) e^{-hz} + erfc(a + b) e^{hz}] with
            // 2ab = hz, so both terms carry e^{-(a^2 + b^2)} (see Devroye::make_record):
            //   a >= b:  log = -(a^2 + b^2) + log(1/2 (erfcx(a - b) + erfcx(a + b)))
            //   a <  b:  log = -hz + log1p(1/2 e^{-(b-a)^2} (erfcx(a + b) - erfcx(b - a)))
            const double b = zz * sqrt(0.5 * t);
            const double d = b - a;
            const double e1 = fm::erfcx_pos(fabs(d)), e2 = fm::erfcx_pos(a + b);
            if (d <= 0.0) lp = hlog2 - (a * a + b * b) + log(0.5 * (e1 + e2));
            else lp = hlog2 - h * zz + log1p(0.5 * fm::exp_neg_sq(d) * (e2 - e1));
        }
        else {
            lambda = kPi2_8;
            lp = hlog2 - a * a + log(fm::erfcx_pos(a));  // log erfc(a)
        }
        double lq = h * (kLogPi_2 - log(lambda)) + log_upper_gamma(h, lambda * t, true, lgammah);
        w_right = 1.0 / (1.0 + exp(lp - lq));
        if (!(w_right >= 0.0)) w_right = 0.0;
    }

    // piece schedule of random_polyagamma_alternate (src/pgm_alternate.c:294-323)
    static __device__ inline void make_record(double h, double z, bool /*compat*/, double* rec)
    {
        double zz = 0.5 * fabs(z);
        if (!(zz > 0.0)) zz = 0.0;
        rec[0] = h;
        rec[1] = zz;
        if (h > 4.0) {
            double chunk = (h >= 5.0) ? 4.0 : 3.0;
            double hr = h;
            while (hr > 4.0) hr -= chunk;
            piece_weights(chunk, zz, rec[2], rec[3]);
            piece_weights(hr, zz, rec[4], rec[5]);
        }
        else {
            piece_weights(h, zz, rec[2], rec[3]);
            rec[4] = 0.0;
            rec[5] = 0.0;
        }
    }

    // the cheap half of set_sampling_parameters
    static __device__ __forceinline__ void set_piece(Par& p, double h, double zz, double w_right, double lgammah,
                                                     bool compat)
    {
        p.w_right = w_right;
        p.lgammah = lgammah;
        p.h = h;
        p.zz = zz;
        p.t = trunc_point(h);
        p.t_inv = fm::rcp(p.t);
        p.half_h2 = 0.5 * h * h;
        p.hlog2 = h * kLog2;
        p.logh = log(h);
        p.ca = compat ? 0.5 * kLogPi_2 : kLogPi_2;  // src/pgm_alternate.c:90
        p.z2 = zz * zz;
        p.lambda = kPi2_8 + 0.5 * p.z2;
        if (zz > 1e-100) {  // (below: h / zz is "infinite" for every test it enters, as in the oracle)
            const double inv_z = fm::rcp(zz);
            p.inv_z2 = inv_z * inv_z;
            p.h_z = h * inv_z;
            p.h_z2 = p.h_z * p.h_z;
        }
        else {
            p.inv_z2 = INFINITY;
            p.h_z = INFINITY;
            p.h_z2 = INFINITY;
        }
    }

    template <class Rec>
    static __device__ __forceinline__ bool load(const Rec& rec, Par& p, State& s, bool compat)
    {
        double h = rec[0], zz = rec[1];
        s.acc = 0.0;
        if (h > 4.0) {
            s.chunk = (h >= 5.0) ? 4.0 : 3.0;
            s.h_rem = h;
            s.w_last = rec[4];
            s.lg_last = rec[5];
            set_piece(p, s.chunk, zz, rec[2], rec[3], compat);
        }
        else {
            s.chunk = 0.0;
            s.h_rem = 0.0;
            s.w_last = 0.0;
            s.lg_last = 0.0;
            set_piece(p, h, zz, rec[2], rec[3], compat);
        }
        return true;
    }

    static __device__ __forceinline__ double inv_left_gamma(const Par& p, Rng& r)
    {
        return fm::rcp(left_bounded_gamma(r, 0.5, p.half_h2, p.t_inv));
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
            double w = p.h_z + 0.5 * y * y * p.inv_z2;
            x = p.h_z2 * fm::rcp(w + fm::sqrt_nn(w * w - p.h_z2));
            if (r.uniform_co() * (p.h_z + x) > p.h_z) {
                x = p.h_z2 * fm::rcp(x);
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
        const double rx = fm::rcp(x);
        double lk;
        if (x > p.t) {  // bounding_kernel (src/pgm_alternate.c:86-99)
            lk = p.h * p.ca + (p.h - 1.0) * logx - kPi2_8 * x - p.lgammah;
        }
        else {
            lk = p.hlog2 - p.half_h2 * rx - 1.5 * logx - kLs2Pi + p.logh;
        }
        double u = r.uniform32() * exp(lk);
        // piecewise_coef (src/pgm_alternate.c:75-83) with c_n = c_{n-1} (n - 1 + h) / n
        double base = p.hlog2 - kLs2Pi - 1.5 * logx;
        double inv2x = 0.5 * rx;
        double c = 1.0;
        double a_prev = p.h * exp(base - p.h * p.h * inv2x);
        double sum = a_prev;
        bool accept = false;
        for (int n = 1; n < 2000; ++n) {
            double tn = 2.0 * n + p.h;
            c *= fm::div(n - 1.0 + p.h, (double)n);
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
                set_piece(p, hr, p.zz, s.w_last, s.lg_last, compat);
            }
            return false;
        }
        s.acc *= 0.25;
        return true;
    }
};

// Alternate elements whose right (truncated gamma) piece is unreachable: for |z| >= 14 its proposal
// mass is below 2^-33 for every h <= 4 (tools/alternate_right_weight.py), so the 32-bit mixture
// uniform always selects the truncated inverse-Gaussian; and since h/zz <= 4/7 < t(h) that draw is
// always the Michael-Schucany-Haas branch.  No erfc, lgamma or incomplete gamma is ever needed:
// the record is just (h, zz).
__host__ __device__ __forceinline__ bool
alternate_left_only(double z)
{
    return fabs(z) >= 14.0;
}

struct AlternateLeft {
    using Rng = RngFor<5>;
    static constexpr int kFields = 2;  // h, zz

    struct Par {
        double h, zz, inv_z2, t, half_h2, hlog2, h_z, h_z2;
    };
    struct State {
        double acc;
        double h_rem;
        double chunk;
    };

    static __device__ __forceinline__ void make_record(double h, double z, bool, double* rec)
    {
        rec[0] = h;
        rec[1] = 0.5 * fabs(z);
    }

    static __device__ __forceinline__ void set_piece(Par& p, double h, double zz)
    {
        p.h = h;
        p.zz = zz;
        p.t = Alternate::trunc_point(h);
        p.half_h2 = 0.5 * h * h;
        p.hlog2 = h * kLog2;
        const double inv_z = fm::rcp(zz);
        p.inv_z2 = inv_z * inv_z;
        p.h_z = h * inv_z;
        p.h_z2 = p.h_z * p.h_z;
    }

    template <class Rec>
    static __device__ __forceinline__ bool load(const Rec& rec, Par& p, State& s, bool)
    {
        double h = rec[0], zz = rec[1];
        s.acc = 0.0;
        if (h > 4.0) {
            s.chunk = (h >= 5.0) ? 4.0 : 3.0;
            s.h_rem = h;
            set_piece(p, s.chunk, zz);
        }
        else {
            s.chunk = 0.0;
            s.h_rem = 0.0;
            set_piece(p, h, zz);
        }
        return true;
    }

    // random_jacobi_star (src/pgm_alternate.c:229-263), left piece only
    static __device__ __forceinline__ bool step(Par& p, State& s, Rng& r, bool)
    {
        (void)r.uniform32();  // mixture draw: always the left piece here
        double x;
        do {
            double y = r.normal();
            double w = p.h_z + 0.5 * y * y * p.inv_z2;
            x = p.h_z2 * fm::rcp(w + fm::sqrt_nn(w * w - p.h_z2));
            if (r.uniform_co() * (p.h_z + x) > p.h_z) {
                x = p.h_z2 * fm::rcp(x);
            }
        } while (x >= p.t);
        const double logx = log(x);
        const double base = p.hlog2 - kLs2Pi - 1.5 * logx;
        const double inv2x = 0.5 * fm::rcp(x);
        // k(x) = a_0(x) on the left of t: u = U * a_0
        double a_prev = p.h * exp(base - p.h * p.h * inv2x);
        const double u = r.uniform32() * a_prev;
        double c = 1.0;
        double sum = a_prev;
        bool accept = false;
        for (int n = 1; n < 2000; ++n) {
            double tn = 2.0 * n + p.h;
            c *= (n - 1.0 + p.h) * fm::rcp((double)n);
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
                set_piece(p, hr, p.zz);
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
//
// FULL = false is the specialisation for elements whose right-hand (truncated gamma) envelope
// piece carries less than 2^-33 of the proposal mass: the 32-bit mixture uniform can then never
// select it, so neither the right tangent nor the incomplete gamma function is needed and the
// per-proposal code has a single branch-free path.  saddle_mode() is the (h, z) rule.
namespace saddle {

constexpr double kUMax = 1.2337005501361697;  // just below pi^2 / 8, the pole of tan

// K'(u), K''(u), K'''(u) of the cumulant generating function (cumulant_prime,
// src/pgm_saddle.c:100-107, with exact tan / tanh instead of the rational tanh_x :56-74)
template <bool NEG = false>
__device__ __forceinline__ void
kprime3(double u, double& f, double& fp, double& fpp)
{
    double s2 = 2.0 * u;
    if (!NEG && fabs(s2) < 1e-4) {
        double g = 1.0 / 3.0 + s2 * (2.0 / 15.0 + s2 * (17.0 / 315.0));
        f = 1.0 + s2 * g;
        fp = f * f - g;
        fpp = 16.0 / 15.0 + s2 * (136.0 / 105.0);
        return;
    }
    double a = fabs(s2);
    double rs = fm::rsqrt(a);
    double s = a * rs;
    double inv = (!NEG && s2 > 0.0) ? rs * rs : -(rs * rs);
    double tt;
    if (!NEG && s2 > 0.0) {
        tt = tan(s);
    }
    else {
        // tanh s = (1 - e^{-2s}) / (1 + e^{-2s}); below s = 0.1 the subtraction would cost digits: Taylor
        // series of tanh(s)/s in s^2 there (first omitted term < 2e-18)
        const double e2 = fm::exp(-2.0 * s);
        tt = (1.0 - e2) * fm::rcp(1.0 + e2);
        if (s < 0.1) {
            double g = -929569.0 / 638512875.0;
            g = fma(g, a, 21844.0 / 6081075.0);
            g = fma(g, a, -1382.0 / 155925.0);
            g = fma(g, a, 62.0 / 2835.0);
            g = fma(g, a, -17.0 / 315.0);
            g = fma(g, a, 2.0 / 15.0);
            g = fma(g, a, -1.0 / 3.0);
            tt = s * fma(g, a, 1.0);
        }
    }
    f = tt * rs;
    double t = (1.0 - f) * inv;
    fp = f * f + t;
    fpp = 2.0 * f * fp - fp * inv - 2.0 * t * inv;
}

// cumulant (src/pgm_saddle.c:91-94) minus its log cosh z offset, from K'(u) = f:
//   u > 0: -log cos(s)  = 1/2 log(1 + tan^2 s)
//   u < 0: -log cosh(s) = 1/2 log(1 - tanh^2 s), or -(s - log 2 + log1p(e^-2s)) once tanh -> 1
__device__ __forceinline__ double
cumulant0(double u, double f)
{
    double s2 = 2.0 * u;
    if (s2 > 0.0) {
        return 0.5 * log(1.0 + f * f * s2);
    }
    if (s2 < 0.0) {
        double s = sqrt(-s2);
        double th = f * s;
        if (s < 3.0) {
            return 0.5 * log((1.0 - th) * (1.0 + th));
        }
        double e = (1.0 - th) * fm::rcp(1.0 + th);  // e^{-2s}
        return -(s - kLog2 + e * (1.0 - e * (0.5 - e * (1.0 / 3.0 - 0.25 * e))));
    }
    return 0.0;
}

// select_starting_guess (src/pgm_saddle.c:119-125); for x <= 0.25 the large-argument solution of
// tanh(s)/s = x replaces the constant -9
// Large-argument solution of tanh(s)/s = x: two fixed-point passes of s = tanh(s)/x from s = 1/x
// written out with tanh s = 1 - 2 e^{-2s} + 2 e^{-4s}; relative residual < 1e-9 for x <= 0.2.
__device__ __forceinline__ double
asymptotic_guess(double x)
{
    const double xi = fm::rcp(x);
    const double e = exp(-2.0 * xi);
    const double s = (1.0 - 2.0 * e + e * e * (2.0 - 8.0 * xi)) * xi;
    return -0.5 * s * s;
}

__device__ __forceinline__ double
table_guess(double x)
{
    if (x <= 0.25) return asymptotic_guess(x);
    return x <= 0.5 ? -1.78 : x <= 1.0 ? -0.147 : x <= 1.5 ? 0.345 : x <= 2.5 ? 0.72 : x <= 4.0 ? 0.95 : 1.15;
}

// newton_raphson (src/pgm_saddle.c:140-158): root u of K'(u) = x, relative residual <= 1e-9.
// Halley's iteration (uses K''', which is free once K' and K'' are known) falls back to a
// Newton step when its denominator is not positive.  On return f = K'(u), fp = K''(u) were
// evaluated AT the returned u.
template <bool NEG = false>
__device__ __forceinline__ double
solve(double x, double u, double& f, double& fp)
{
    double fpp;
    for (int n = 0; n < 30; ++n) {
        kprime3<NEG>(u, f, fp, fpp);
        double g = f - x;
        if (fabs(g) <= 1e-9 * x || !(fp > 1e-300)) return u;
        double den = 2.0 * fp * fp - g * fpp;
        double du = (den > 0.0) ? 2.0 * g * fp * fm::rcp(den) : g * fm::rcp(fp);
        double un = u - du;
        if (un >= kUMax) un = 0.5 * (u + kUMax);  // never step onto / over the pole
        if (NEG && un > -1e-3) un = 0.5 * (u - 1e-3);
        if (un == u) return u;
        u = un;
    }
    kprime3<NEG>(u, f, fp, fpp);
    return u;
}

// Starting points for the two per-(h, z) solves of the FULL setup, K'(u) = 2.75 xl and K'(u) = 3 xl with
// xl = tanh(zz)/zz: both are functions of zz alone, tabulated on zz in [0, kGuessZMax] (step
// 1/16, filled once per device by init_tables_kernel with the solver itself) and
// interpolated linearly -- error ~1e-4, so one Halley step lands below the 1e-9 tolerance and
// every lane of a warp takes the same two evaluations.
constexpr double kGuessZMax = 26.0;
constexpr int kGuessPerUnit = 16;
constexpr int kGuessN = (int)(kGuessZMax * kGuessPerUnit) + 2;
static __device__ double g_guess_ur[kGuessN];

__device__ __forceinline__ double
table_interp(const double* tab, double zz)
{
    double pos = zz * kGuessPerUnit;
    int i = (int)pos;
    double fr = pos - i;
    double a = __ldg(tab + i), b = __ldg(tab + i + 1);
    return a + (b - a) * fr;
}

// ---- the saddle point as a table ---------------------------------------------------------------------
// K'(u) = x defines ONE function u(x), whatever h and z are (the reference solves it by Newton-Raphson for
// every proposal, src/pgm_saddle.c:140-158,235-244).  The acceptance test needs two functions of it,
//     G(x) = K0(u(x)) - u(x) x      (K0 = cumulant without its log cosh z offset; dG/dx = -u)
//     D(x) = K2(u(x))               (K2 = second derivative of the cumulant),
// and only through the combinations
//     Gr(x) = G(x) + 1/(2x) - log 2,      Hr(x) = 1/2 log(D(x) / x^3),
// which are smooth, bounded, and vanish like e^{-2/x} as x -> 0 (G ~ -1/(2x) + log 2, D ~ x^3 there).  Both
// are tabulated per device as cubic Hermite pieces on a grid that follows the floating-point format --
// 2^kTabBits cells per binade of x, so the cell index is a shift of the high word -- over x in
// [2^kTabEmin, 2^kTabEmax).  Interpolation error < 4e-11 absolute (tests/test_gpu_parity.py::
// test_saddle_table against the direct solve); below the range both functions are < 1e-25 (taken as 0),
// above it (x >= 8: a right-piece proposal far in the tail) the direct solve runs.
// Every lane of a warp therefore takes the same dozen instructions per proposal instead of one to four
// tanh / tan evaluations: the far / mid / left split of the saddle buckets (round 1) is gone.
#ifndef PGM_SADDLE_TAB_BITS
#define PGM_SADDLE_TAB_BITS 7
#endif
constexpr int kTabBits = PGM_SADDLE_TAB_BITS;
constexpr int kTabEmin = -5, kTabEmax = 3;
constexpr int kTabCells = (kTabEmax - kTabEmin) << kTabBits;
static __device__ double2 g_saddle_tab[4 * kTabCells];  // per cell: cubic in t of Gr (c0 c1 | c2 c3), then of Hr

// Gr, Hr and their x-derivatives by the direct solve (table generation; x above the table)
static __device__ __forceinline__ void
eval_direct(double x, double& gr, double& dgr, double& hr, double& dhr)
{
    double u = table_guess(x), f = 0.0, fp = 1.0, fpp = 0.0;
    for (int n = 0; n < 80; ++n) {  // Halley, to the last bit
        kprime3(u, f, fp, fpp);
        const double g = f - x;
        const double den = 2.0 * fp * fp - g * fpp;
        const double du = (den > 0.0) ? 2.0 * g * fp / den : g / fp;
        double un = u - du;
        if (un >= kUMax) un = 0.5 * (u + kUMax);
        const bool done = (un == u) || fabs(du) <= 2e-16 * fabs(u);
        u = un;
        if (done) break;
    }
    kprime3(u, f, fp, fpp);
    const double rx = 1.0 / x;
    gr = (cumulant0(u, f) - u * x) + 0.5 * rx - kLog2;
    dgr = -u - 0.5 * rx * rx;
    hr = 0.5 * ::log(fp * rx * rx * rx);
    dhr = 0.5 * (fpp / (fp * fp) - 3.0 * rx);  // dD/dx = K3 du/dx = K3 / K2
}

// The table lookup in two halves, so that the caller can put independent work (the logarithm of the
// acceptance uniform) between the loads and their first use.
struct Lookup {
    double2 g01, g23, h01, h23;
    double t;
    int cell;
};

__device__ __forceinline__ Lookup
lookup_issue(double x)
{
    Lookup q;
    const int hi = __double2hiint(x);
    q.cell = (hi >> (20 - kTabBits)) - ((1023 + kTabEmin) << kTabBits);
    const int c = min(max(q.cell, 0), kTabCells - 1);  // (out of range: any valid cell, the values are not used)
    const double x0 = __hiloint2double(hi & ~((1 << (20 - kTabBits)) - 1), 0);                // left end of the cell
    const double inv_w = __hiloint2double(((2046 + kTabBits) << 20) - (hi & 0x7ff00000), 0);  // 1 / cell width
    q.t = (x - x0) * inv_w;
    const double2* p = g_saddle_tab + 4 * c;
    q.g01 = __ldg(p), q.g23 = __ldg(p + 1), q.h01 = __ldg(p + 2), q.h23 = __ldg(p + 3);
    return q;
}

// ABOVE: x may lie above the table (x >= 8: only a right-piece proposal of the full saddle sampler can); the
// direct solve then runs, inlined (see the note on out-of-line calls in pgm_device.cuh).
template <bool ABOVE>
__device__ __forceinline__ void
lookup_finish(const Lookup& q, double x, double& gr, double& hr)
{
    gr = fma(fma(fma(q.g23.y, q.t, q.g23.x), q.t, q.g01.y), q.t, q.g01.x);
    hr = fma(fma(fma(q.h23.y, q.t, q.h23.x), q.t, q.h01.y), q.t, q.h01.x);
    if ((unsigned)q.cell >= (unsigned)kTabCells) {
        gr = 0.0;
        hr = 0.0;
        if (ABOVE && q.cell > 0) {
            double d0, d1;
            eval_direct(x, gr, d0, hr, d1);
        }
    }
}

template <bool ABOVE>
__device__ __forceinline__ void
lookup(double x, double& gr, double& hr)
{
    lookup_finish<ABOVE>(lookup_issue(x), x, gr, hr);
}

// invgauss_logcdf (src/pgm_saddle.c:282-292)
__device__ __forceinline__ double
invgauss_logcdf(double x, double mu, double lambda)
{
    // (x, mu, lambda are positive normal numbers here: MUFU-seeded reciprocals)
    // cdf = Phi(A) + e^{2 lambda / mu} Phi(-B); with a' = A / sqrt 2, b' = B / sqrt 2 the exponent is
    // b'^2 - a'^2, so both terms carry e^{-a'^2} (the devroye setup's identity):
    //   a' <  0:  log cdf = -a'^2 + log(1/2 (erfcx(-a') + erfcx(b')))
    //   a' >= 0:  log cdf = log1p(-1/2 e^{-a'^2} (erfcx(a') - erfcx(b')))
    const double qm = fm::div(x, mu);
    const double s = 0.7071067811865475 * fm::rsqrt(fm::div(x, lambda));
    const double ap = (qm - 1.0) * s, bp = (qm + 1.0) * s;
    const double e1 = fm::erfcx_pos(fabs(ap)), e2 = fm::erfcx_pos(bp);
    if (ap < 0.0) return log(0.5 * (e1 + e2)) - ap * ap;
    return log1p(-0.5 * fm::exp_neg_sq(ap) * (e1 - e2));
}

}  // namespace saddle

// Elements of the saddle method whose right envelope piece is unreachable (weight < 2^-33, see
// tools/saddle_right_weight.py for the map of log10 w_right over (h, |z|/2)): the 32-bit mixture uniform can
// never select it, so neither the right tangent nor the incomplete gamma function is needed.  Only used with
// the consistent envelope; ref_compat keeps the reference's (much larger) right weight.
enum : int { kSaddleLeft = 1, kSaddleFull = 2 };

__host__ __device__ __forceinline__ int
saddle_mode(double h, double z, bool compat)
{
    return (compat || !(h * (1.0 + 0.125 * fabs(z)) >= 22.0)) ? kSaddleFull : kSaddleLeft;
}

template <int MODE>
struct SaddleT {
    using Rng = RngFor<(MODE == kSaddleFull ? 2 : 3)>;
    static constexpr bool FULL = (MODE == kSaddleFull);
    // h, xl, kappa, A [, w_left, kappa_r, A_r, slope_r]
    static constexpr int kFields = FULL ? 8 : 4;

    struct Par {
        double h, inv_h, xl, xc, kappa, A;
        double w_left, kappa_r, A_r, slope_r;  // FULL only
    };
    struct State {
        double acc;
    };

    // set_sampling_parameters (src/pgm_saddle.c:180-230) + mixture weight (:317-328, log space).
    // With L_l(x) = slope_l x + icpt_l the left tangent (slope_l = -1/(2 xl^2), icpt_l = 1/xl - 1/(2 xc)) and
    // alpha_l = K2(u(xc)) / xc^3, the left acceptance test  U k_l(x) <= phi(x)  (bounding_kernel :250-263,
    // saddle_point :235-244, common sqrt(h / 2 pi) cancelled) reads, in the tabulated functions,
    //     log U + Hr(x) <= h (Gr(x) - kappa x) + A,
    //     kappa = z^2/2 + slope_l,   A = h (log 2 + log cosh z - 1/xl) + Hr(xc)      (z halved)
    // -- the 1/(2x) of the envelope cancels the one in G.
    static __device__ inline void make_record(double h, double z, bool compat, double* rec)
    {
        using namespace saddle;
        double zz = 0.5 * fabs(z);
        if (!(zz > 0.0)) zz = 0.0;
        double xl, half_z2, l2cz;  // l2cz = log 2 + log cosh(zz) = zz + log1p(e^{-2 zz})
        // (below 1e-100 every quantity here equals its z = 0 value in double precision; the cut keeps
        // the MUFU-seeded reciprocals on normal numbers)
        if (zz > 1e-100) {
            const double e = fm::exp_poly(-2.0 * zz);
            xl = fm::div(1.0 - e, (1.0 + e) * zz);  // tanh(zz) / zz
            half_z2 = 0.5 * zz * zz;
            l2cz = zz + log(1.0 + e);  // (absolute accuracy is what counts here: no log1p needed)
        }
        else {
            xl = 1.0;
            half_z2 = 0.0;
            l2cz = kLog2;
        }
        const double xc = 2.75 * xl;
        const double xl_inv = fm::rcp(xl);
        double gr_c, hr_c;
        lookup<false>(xc, gr_c, hr_c);  // (xc <= 2.75)
        const double slope_l = -0.5 * xl_inv * xl_inv;
        rec[0] = h;
        rec[1] = xl;
        rec[2] = half_z2 + slope_l;
        rec[3] = h * (l2cz - xl_inv) + hr_c;
        if (FULL) {
            const double log_cosh_z = l2cz - kLog2;
            const double xc_inv = fm::rcp(xc);
            const double icpt_l = -0.5 * xc_inv + xl_inv;  // cumulant(ul) = 0, src/pgm_saddle.c:216
            const double logxl = log(xl);
            const double xr = 3.0 * xl;
            double f, fp_r;
            const bool tabulated = zz < kGuessZMax;
            double ur = solve(xr, tabulated ? table_interp(g_guess_ur, zz) : table_guess(xr), f, fp_r);
            const double tr = ur + half_z2;
            const double logxc = 1.0116009116784799 + logxl;  // log(2.75)
            const double slope_r = -tr - fm::rcp(xr);
            const double icpt_r = (log_cosh_z + cumulant0(ur, f)) + 1.0 - 1.0986122886681098 - logxl + logxc;
            const double log_sqrt_h2pi = 0.5 * log(h / (2.0 * kPi));
            // 1/2 log alpha_l = Hr(xc);  1/2 log alpha_r = Hr(xc) + 1/2 log xc   (alpha_r = K2(u(xc)) / xc^2)
            const double half_log_alpha_r = hr_c + 0.5 * logxc;
            const double log_coef_r = log_sqrt_h2pi - half_log_alpha_r;
            double lp = h * (0.5 * xc_inv + icpt_l - xl_inv) + invgauss_logcdf(xc, xl, h) - hr_c;
            const double hrho = -h * slope_r;
            double lq = log_upper_gamma(h, hrho * xc, false) + log_coef_r +
                        h * (icpt_r - (compat ? 0.0 : logxc) - log(hrho));
            double w_left = 1.0 / (1.0 + exp(lq - lp));
            if (!(w_left >= 0.0)) w_left = 1.0;
            rec[4] = w_left;
            // right test:  log U + log k_r(x) <= log phi(x)  with  log k_r = h slope_r x + (h - 1) log x + c_r,
            // c_r = h (-+log xc + icpt_r) - 1/2 log alpha_r - h log cosh z  (src/pgm_saddle.c:257 uses +log(xc);
            // the derivation needs -log(xc), SURVEY 8-DEFECTS i); in the tabulated functions
            //     log U + Hr(x) + (h + 1/2) log x + h kappa_r x + h / (2x) <= h Gr(x) + A_r
            const double c_r = h * ((compat ? logxc : -logxc) + icpt_r) - half_log_alpha_r - h * log_cosh_z;
            rec[5] = slope_r + half_z2;
            rec[6] = h * kLog2 - c_r;
            rec[7] = slope_r;
        }
    }

    template <class Rec>
    static __device__ __forceinline__ bool load(const Rec& rec, Par& p, State& s, bool)
    {
        p.h = rec[0];
        p.inv_h = fm::rcp(p.h);
        p.xl = rec[1];
        p.xc = 2.75 * p.xl;
        p.kappa = rec[2];
        p.A = rec[3];
        if (FULL) {
            p.w_left = rec[4];
            p.kappa_r = rec[5];
            p.A_r = rec[6];
            p.slope_r = rec[7];
        }
        s.acc = 0.0;
        return true;
    }

    // One pass of the proposal loop of random_polyagamma_saddle (src/pgm_saddle.c:331-347).
    static __device__ __forceinline__ bool step(Par& p, State& s, Rng& r, bool)
    {
        using namespace saddle;
        const double h = p.h;
        double x;
        bool left = true;
        double um = r.uniform32();  // mixture draw (always consumed, as in the oracle)
        if (FULL) left = um < p.w_left;
        if (left) {
            const double mu = p.xl, mu2 = p.xl * p.xl;
            do {  // IG(mu = xl, lambda = h) | x < xc, Michael-Schucany-Haas; only the normal's square is needed
                const double L = -fm::log_pn(r.uniform_oc());
                const double c = fm::cos2pi_poly(r.uniform_co());
                const double y2 = 2.0 * L * c * c;
                double w = mu + 0.5 * mu2 * y2 * p.inv_h;
                x = mu2 * fm::rcp(w + fm::sqrt_nn(w * w - mu2));
                if (r.uniform_co() * (mu + x) > mu) {
                    x = mu2 * fm::rcp(x);
                }
            } while (x >= p.xc);
        }
        else {
            x = left_bounded_gamma(r, h, -h * p.slope_r, p.xc);
        }
        const Lookup q = lookup_issue(x);  // (the loads fly while the logarithm is computed)
        const double lu = fm::log_pn(r.uniform32());
        double gr, hr;
        lookup_finish<FULL>(q, x, gr, hr);
        bool accept;
        if (FULL && x > p.xc) {
            accept = lu + hr + (h + 0.5) * log(x) + h * (p.kappa_r * x + 0.5 * fm::rcp(x)) <= h * gr + p.A_r;
        }
        else {
            accept = lu + hr <= h * (gr - p.kappa * x) + p.A;
        }
        if (accept) {
            s.acc = 0.25 * h * x;
            return true;
        }
        return false;
    }
};

using SaddleFull = SaddleT<kSaddleFull>;
using SaddleLeft = SaddleT<kSaddleLeft>;


// =========================================================================================
// Normal approximation (h > 50) and truncated gamma convolution   (src/pgm_random.c)
// =========================================================================================

// random_polyagamma_normal_approx (src/pgm_random.c:47-66), cancellation-free moments
template <class R>
__device__ __forceinline__ double
normal_approx_sample(R& r, double h, double z)
{
    double a = fabs(z);
    double mean, var;
    if (!(a > 0.0)) {
        mean = 0.25 * h;
        var = h / 24.0;
    }
    else {
        // tanh(a/2) = (1 - e)/(1 + e) and sech^2(a/2) = 4 e / (1 + e)^2 with e = e^{-a}; one
        // reciprocal rinv = 1 / ((1 + e) a) serves the mean and the variance
        const double e = fm::exp_poly(-a);  // (the tile kernel's two methods stay table-free: see fm::exp)
        const double ope = 1.0 + e;
        const double rinv = fm::rcp(ope * a);
        mean = 0.5 * h * (1.0 - e) * rinv;
        if (a < 1.0) {
            const double a2 = a * a;
            // (sinh a - a) / a^3 = sum_k a^(2k) / (2k+3)!
            const double S = 1.0 / 6.0 + a2 * (1.0 / 120.0 + a2 * (1.0 / 5040.0 + a2 * (1.0 / 362880.0 +
                             a2 * (1.0 / 39916800.0 + a2 * (1.0 / 6227020800.0 + a2 * (1.0 / 1307674368000.0))))));
            const double inv_ope = rinv * a;
            var = h * S * e * inv_ope * inv_ope;
        }
        else {
            // (sinh a - a) sech^2(a/2) / (4 a^3) = (1 - e^2 - 2 a e) / (2 (1 + e)^2 a^3)
            var = 0.5 * h * (1.0 - e * e - 2.0 * a * e) * rinv * rinv * (ope * rinv);
        }
    }
    // mean + N sqrt(var) with the Box-Muller radius folded into one square root
    const double u1 = r.uniform_oc();
    const double u2 = r.uniform_co();
    return mean + fm::sqrt_nn(-2.0 * var * fm::log_pn(u1)) * fm::cos2pi_poly(u2);
}

// random_polyagamma_gamma_conv (src/pgm_random.c:79-97): 200 Marsaglia-Tsang gammas
using GammaRng = RngFor<6>;

// random_polyagamma_gamma_conv (src/pgm_random.c:79-97): 1/2 sum_{k<200} Gamma(h, 1) / (pi^2 (k+1/2)^2 + z^2/4)
// with Marsaglia-Tsang gammas (shape < 1: Gamma(h + 1) U^{1/h}).  One trip of the loop draws ONE
// Box-Muller pair and makes TWO Marsaglia-Tsang attempts with it, so all lanes of a warp stay in
// phase (a "spare normal" flag drifts apart between lanes after the first rejection and halves the
// active lanes); the draw order is exactly that of the oracle's spare-normal loop.
__device__ inline double
gamma_conv_sample(GammaRng& r, double h, double z)
{
    double zz = 0.5 * fabs(z);
    if (!(zz > 0.0)) zz = 0.0;
    const double pi2 = 9.869604401089358;
    const double z2 = zz * zz;
    const bool boost = h < 1.0;
    const double d = (boost ? h + 1.0 : h) - 1.0 / 3.0;
    const double c = fm::rsqrt(9.0 * d);
    const double inv_h = fm::rcp(h);
    double acc = 0.0;
    int k = 0;
    while (k < 200) {
        double nrm[2];
        r.normal_pair(nrm[0], nrm[1]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (k < 200) {
                const double x = nrm[j];
                double v = 1.0 + c * x;
                if (v > 0.0) {
                    v = v * v * v;
                    const double u = r.uniform_oc();
                    const double x2 = x * x;
                    // Marsaglia-Tsang's squeeze u < 1 - 0.0331 x^4 only implies the test below; in a warp
                    // some lane fails it on 93 % of the attempts, so the two logarithms are paid for
                    // anyway -- evaluating them for every lane keeps the warp converged.
                    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) {
                        double g = d * v;
                        if (boost) g *= exp(log(r.uniform_oc()) * inv_h);
                        const double m = k + 0.5;
                        acc += g * fm::rcp(pi2 * m * m + z2);
                        ++k;
                    }
                }
            }
        }
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
