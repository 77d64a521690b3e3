/* oracle/pgm_oracle.c -- TEST INFRASTRUCTURE ONLY (see pgm_oracle.h).
 *
 * Scalar C restatement of the reference's Polya-Gamma path (citations are into the
 * zoj613/polyagamma checkout, file:line) in the FP64 arithmetic and Philox stream discipline
 * of the B200 kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load the library built from this file.
 *
 * Stream discipline (shared with polyagamma_b200/csrc/philox.cuh):
 *   call key  = first two words of Philox4x32-10(key=(seed_lo,seed_hi),
 *               ctr=(offset_lo, offset_hi, 'PGM1', 'key!'))
 *   element i = words Philox(key=call key, ctr=(i_lo, i_hi, block, 0)), block = 0,1,2,...
 *               consumed strictly in order, 4 words per block
 *   uniform32 = (w + 0.5) * 2^-32                    in (0,1)   -- the reference's next_float
 *   uniform53 = (((hi<<32)|lo) >> 11) * 2^-53         in [0,1)   -- the reference's next_double
 *               (lo = first word drawn, hi = second); the (0,1] form adds 1 before scaling
 *   exponential = -log(uniform53 in (0,1])
 *   normal      = sqrt(-2 log U1) * cos(2 pi U2), U1 in (0,1], U2 in [0,1)  (4 words)
 *   gamma(a)    = Marsaglia-Tsang (a >= 1), Gamma(a+1) * U^(1/a) (a < 1); inside the
 *                 gamma-convolution sampler normals are drawn in Box-Muller pairs
 */
#include "pgm_oracle.h"

#include <math.h>
#include <string.h>

#define PGO_PI 3.141592653589793238462643383279503
#define PGO_PI_2 1.570796326794896619231321691639751
#define PGO_PI2_8 1.233700550136169827354311374984519
#define PGO_LOGPI_2 0.4515827052894548647261952298948821 /* log(pi/2) */
#define PGO_LS2PI 0.9189385332046727417803297364056177   /* log(sqrt(2 pi)) */
#define PGO_LOG2 0.6931471805599453094172321214581766

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11).  Replaces numpy's bitgen_t (include/pgm_random.h:4).
 * ---------------------------------------------------------------------------------------- */
void
pgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void
pgo_call_key(uint64_t seed, uint64_t offset, uint32_t key[2])
{
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t c[4] = {(uint32_t)offset, (uint32_t)(offset >> 32), 0x50474D31u, 0x6B657921u};
    uint32_t o[4];
    pgo_philox4x32_10(c, k, o);
    key[0] = o[0];
    key[1] = o[1];
}

typedef struct {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int pos;
    uint64_t words;
} pgo_rng;

static void
rng_init(pgo_rng* r, const uint32_t key[2], uint64_t index)
{
    r->key[0] = key[0];
    r->key[1] = key[1];
    r->ctr[0] = (uint32_t)index;
    r->ctr[1] = (uint32_t)(index >> 32);
    r->ctr[2] = 0;
    r->ctr[3] = 0;
    r->pos = 4;
    r->words = 0;
}

static uint32_t
rng_u32(pgo_rng* r)
{
    if (r->pos == 4) {
        pgo_philox4x32_10(r->ctr, r->key, r->buf);
        if (++r->ctr[2] == 0) {
            ++r->ctr[3];
        }
        r->pos = 0;
    }
    r->words++;
    return r->buf[r->pos++];
}

/* next_float (src/pgm_macros.h:54-55), 32-bit and never 0 or 1 */
static double
rng_uniform32(pgo_rng* r)
{
    return ((double)rng_u32(r) + 0.5) * 2.3283064365386963e-10;
}

static uint64_t
rng_u53(pgo_rng* r)
{
    uint64_t lo = rng_u32(r);
    uint64_t hi = rng_u32(r);
    return ((hi << 32) | lo) >> 11;
}

/* next_double (src/pgm_macros.h:61-62): [0, 1) */
static double
rng_uniform_co(pgo_rng* r)
{
    return (double)rng_u53(r) * 1.1102230246251565e-16;
}

/* (0, 1] */
static double
rng_uniform_oc(pgo_rng* r)
{
    return ((double)rng_u53(r) + 1.0) * 1.1102230246251565e-16;
}

/* replaces numpy random_standard_exponential (call sites src/pgm_devroye.c:122-123,169;
 * src/pgm_common.h:141,147,153) */
static double
rng_exponential(pgo_rng* r)
{
    return -log(rng_uniform_oc(r));
}

/* replaces numpy random_standard_normal (src/pgm_random.c:64, src/pgm_devroye.c:131,
 * src/pgm_alternate.c:211, src/pgm_saddle.c:335) */
static double
rng_normal(pgo_rng* r)
{
    double u1 = rng_uniform_oc(r);
    double u2 = rng_uniform_co(r);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * PGO_PI * u2);
}

typedef struct {
    int has;
    double spare;
} pgo_spare;

static double
rng_normal_paired(pgo_rng* r, pgo_spare* s)
{
    if (s->has) {
        s->has = 0;
        return s->spare;
    }
    double u1 = rng_uniform_oc(r);
    double u2 = rng_uniform_co(r);
    double rad = sqrt(-2.0 * log(u1));
    double ang = 2.0 * PGO_PI * u2;
    s->spare = rad * sin(ang);
    s->has = 1;
    return rad * cos(ang);
}

/* replaces numpy random_standard_gamma (src/pgm_random.c:92).  Marsaglia & Tsang (2000). */
static double
rng_gamma(pgo_rng* r, double a, pgo_spare* sp)
{
    double boost = (a < 1.0) ? 1.0 : 0.0;
    double d = (a + boost) - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    double g;
    for (;;) {
        double x, v;
        do {
            x = rng_normal_paired(r, sp);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        double u = rng_uniform_oc(r);
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) {
            g = d * v;
            break;
        }
    }
    if (boost != 0.0) {
        g *= exp(log(rng_uniform_oc(r)) / a);
    }
    return g;
}

void
pgo_stream_words(uint64_t seed, uint64_t offset, uint64_t index, size_t n, uint32_t* out)
{
    uint32_t key[2];
    pgo_rng r;
    pgo_call_key(seed, offset, key);
    rng_init(&r, key, index);
    for (size_t i = 0; i < n; ++i) {
        out[i] = rng_u32(&r);
    }
}

void
pgo_stream_variates(uint64_t seed, uint64_t offset, uint64_t index, int kind, double param,
                    size_t n, double* out)
{
    uint32_t key[2];
    pgo_rng r;
    pgo_spare sp = {0, 0.0};
    pgo_call_key(seed, offset, key);
    rng_init(&r, key, index);
    for (size_t i = 0; i < n; ++i) {
        switch (kind) {
            case 0: out[i] = rng_uniform32(&r); break;
            case 1: out[i] = rng_uniform_co(&r); break;
            case 2: out[i] = rng_exponential(&r); break;
            case 3: out[i] = rng_normal(&r); break;
            default: out[i] = rng_gamma(&r, param, &sp); break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Shared primitives (src/pgm_common.h)
 * ---------------------------------------------------------------------------------------- */

/* pgm_lgamma (src/pgm_common.h:40-117): log-factorial table for integer z <= 200, else libm.
 * The table is generated (long double running sum) instead of being transcribed. */
static double g_logfact[201];
static int g_logfact_ready = 0;

static void
logfact_init(void)
{
    long double acc = 0.0L;
    g_logfact[0] = 0.0;
    g_logfact[1] = 0.0; /* lgamma(1) = lgamma(2) = 0 */
    for (int k = 2; k <= 200; ++k) {
        /* g_logfact[k] = lgamma(k) = log((k-1)!) */
        acc += logl((long double)(k - 1));
        g_logfact[k] = (double)acc;
    }
    g_logfact_ready = 1;
}

double
pgo_lgamma(double z)
{
    if (!g_logfact_ready) {
        logfact_init();
    }
    if (z < 201.0 && z >= 1.0 && z == floor(z)) {
        return g_logfact[(int)z];
    }
    return lgamma(z);
}

/* scaled complementary error function exp(x^2) erfc(x), x >= 0 (used where the reference's
 * FP32 expf()*erfcf() products overflow/underflow: src/pgm_saddle.c:268,282-292) */
double
pgo_erfcx(double x)
{
    if (x < 10.0) {
        return exp(x * x) * erfc(x);
    }
    /* asymptotic series: 1/(x sqrt(pi)) * sum (-1)^k (2k-1)!! / (2 x^2)^k */
    double r = 0.5 / (x * x);
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 30; ++k) {
        term *= -(2.0 * k - 1.0) * r;
        sum += term;
        if (fabs(term) < 1e-17 * fabs(sum)) {
            break;
        }
    }
    return sum / (x * 1.7724538509055160273);
}

/* log of the standard normal cdf, accurate in both tails (replaces log_norm_cdf,
 * src/pgm_saddle.c:268, whose erfcf underflows) */
static double
log_ndtr(double x)
{
    double y = x * 0.70710678118654752440;
    if (y < -1.0) {
        return log(0.5 * pgo_erfcx(-y)) - y * y;
    }
    if (y < 1.0) {
        return log(0.5 * erfc(-y));
    }
    return log1p(-0.5 * erfc(y));
}

/* Modified Lentz evaluation of the two continued fractions of Abergel & Moisan (2020),
 * restating confluent_x_smaller / confluent_p_smaller (src/pgm_common.h:181-211, 237-265)
 * in FP64. */
#define PGO_LENTZ_TINY 1e-290
#define PGO_LENTZ_EPS 4.0e-16
#define PGO_LENTZ_MAXIT 1000

static double
confluent_x_smaller(double p, double x)
{
    double a = 1.0, b = p;
    double r = -(p - 1.0) * x;
    double s = 0.5 * x;
    double f = a / b, c = a / PGO_LENTZ_TINY, d = 1.0 / b;
    for (int n = 2; n < PGO_LENTZ_MAXIT; ++n) {
        a = (n & 1) ? s * (n - 1) : r - s * n;
        b += 1.0;
        c = b + a / c;
        if (fabs(c) < PGO_LENTZ_TINY) c = PGO_LENTZ_TINY;
        d = a * d + b;
        if (fabs(d) < PGO_LENTZ_TINY) d = PGO_LENTZ_TINY;
        d = 1.0 / d;
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < PGO_LENTZ_EPS) break;
    }
    return f;
}

static double
confluent_p_smaller(double p, double x)
{
    double a = 1.0, b = x - p + 1.0;
    double f = a / b, c = a / PGO_LENTZ_TINY, d = 1.0 / b;
    for (int n = 1; n < PGO_LENTZ_MAXIT; ++n) {
        a = n * (p - n);
        b += 2.0;
        c = b + a / c;
        if (fabs(c) < PGO_LENTZ_TINY) c = PGO_LENTZ_TINY;
        d = a * d + b;
        if (fabs(d) < PGO_LENTZ_TINY) d = PGO_LENTZ_TINY;
        d = 1.0 / d;
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < PGO_LENTZ_EPS) break;
    }
    return f;
}

/* log of the upper incomplete gamma function Gamma(p, x) (normalized: log Q(p, x)).
 * Restates upper_incomplete_gamma (src/pgm_common.h:291-340) in log space so that neither the
 * PGM_MAX_EXP clamps (:325-338) nor the 0*inf of src/pgm_saddle.c:325-326 can occur. */
double
pgo_log_upper_gamma(double p, double x, int normalized)
{
    double lg = pgo_lgamma(p);
    if (x > p) {
        double f = confluent_p_smaller(p, x);
        double v = log(f) - x + p * log(x);
        return normalized ? v - lg : v;
    }
    double f = confluent_x_smaller(p, x);
    double lower = f * exp(-x + p * log(x) - lg); /* P(p, x) */
    double lq = log1p(-lower);
    return normalized ? lq : lq + lg;
}

/* random_left_bounded_gamma (src/pgm_common.h:126-157): X ~ Gamma(a, rate b) | X > t */
static double
left_bounded_gamma(pgo_rng* r, double a, double b, double t)
{
    double x;
    if (a > 1.0) { /* Dagpunar (1978), :131-145 */
        b = t * b;
        double amin1 = a - 1.0;
        double bmina = b - a;
        double c0 = 0.5 * (bmina + sqrt(bmina * bmina + 4.0 * b)) / b;
        double one_minus_c0 = 1.0 - c0;
        double log_m = amin1 * (log(amin1 / one_minus_c0) - 1.0);
        double threshold;
        do {
            x = b + rng_exponential(r) / c0;
            threshold = amin1 * log(x) - x * one_minus_c0 - log_m;
        } while (log(rng_uniform32(r)) > threshold);
        return t * (x / b);
    }
    else if (a == 1.0) { /* :146-148 */
        return t + rng_exponential(r) / b;
    }
    else if (a == 0.5) { /* Philippe (1997) A4, :149-156, with u < x^(-1/2) written as u^2 x < 1 */
        double tb = t * b;
        double u;
        do {
            x = 1.0 + rng_exponential(r) / tb;
            u = rng_uniform32(r);
        } while (u * u * x > 1.0);
        return t * x;
    }
    else {
        double amin1 = a - 1.0;
        double tb = t * b;
        do {
            x = 1.0 + rng_exponential(r) / tb;
        } while (log(rng_uniform32(r)) > amin1 * log(x));
        return t * x;
    }
}

/* ------------------------------------------------------------------------------------------
 * Devroye sampler (src/pgm_devroye.c)
 * ---------------------------------------------------------------------------------------- */
#define PGO_T 0.64

typedef struct {
    double w; /* proposal_probability p/(p+q) */
    double K;
    double zz;
    double z2;
} dev_par;

/* set_sampling_parameters (src/pgm_devroye.c:61-83) in FP64 */
static void
dev_setup(dev_par* p, double z)
{
    p->zz = 0.5 * fabs(z);
    if (!(p->zz > 0.0)) p->zz = 0.0; /* z = NaN is sampled as z = 0 */
    if (p->zz > 0.0) {
        const double a = 0.8838834764831844;  /* 1/sqrt(2T) */
        const double tt = 0.565685424949238;  /* sqrt(T/2) */
        double b = p->zz * tt;
        double t1 = erfc(a - b) * exp(-p->zz);
        double t2 = (a + b < 26.5) ? erfc(a + b) * exp(p->zz) : 0.0;
        double pp = t1 + t2;
        p->z2 = p->zz * p->zz;
        p->K = PGO_PI2_8 + 0.5 * p->z2;
        double q = PGO_PI_2 * exp(-p->K * PGO_T) / p->K;
        p->w = (pp + q > 0.0) ? pp / (pp + q) : 1.0;
    }
    else {
        p->w = 0.4223027567786595; /* :78 */
        p->K = PGO_PI2_8;
        p->z2 = 0.0;
    }
}

/* piecewise_coef (src/pgm_devroye.c:36-46): a_n(x) = pi (n+1/2) exp(lp - e (n+1/2)^2) */
static double
dev_coef(int n, double lp, double e)
{
    double a = n + 0.5;
    return PGO_PI * a * exp(lp - e * a * a);
}

static double
u53_from(uint32_t lo, uint32_t hi)
{
    return (double)(((((uint64_t)hi) << 32) | lo) >> 11);
}

/* random_jacobi_star (src/pgm_devroye.c:160-190) with random_right_bounded_invgauss (:113-142) inside,
 * written as the kernels run it: a sequence of TRIPS, each consuming exactly EIGHT words (two Philox blocks)
 * of the element's stream whatever it does with them --
 *     w0 w1   mixture uniform (53 bits, [0, 1)); ignored while a truncated inverse-Gaussian draw is being retried
 *     w2 w3   L = -log U, U 53 bits in (0, 1]: the tail exponential, the first exponential of Devroye's pair, or
 *             half the squared Box-Muller radius of the Michael-Schucany-Haas normal
 *     w4 w5   the pair's second exponential, or the Box-Muller angle
 *     w6      32-bit uniform: the exponential-tilt test of the pair proposal, or the root choice of M-S-H
 *     w7      32-bit uniform of the alternating-series test
 * -- so that all lanes of a warp generate their two blocks together at the top of a trip (a lane that refills
 * its four-word buffer whenever it runs dry does so alone: 11 of 32 lanes active in the Philox rounds, a
 * quarter of the kernel's instructions).  A trip ends in: a retry of the left piece (the nested loops of the
 * reference, flattened), a rejected proposal, or an accepted J*(1, z) draw. */
static double
dev_jacobi_star(pgo_rng* r, const dev_par* p)
{
    int in_left = 0;
    for (;;) {
        uint32_t w[8];
        for (int k = 0; k < 8; ++k) w[k] = rng_u32(r);
        double x, lp, e;
        const int left = in_left ? 1 : (u53_from(w[0], w[1]) * 1.1102230246251565e-16 < p->w);
        const double L = -log((u53_from(w[2], w[3]) + 1.0) * 1.1102230246251565e-16);
        const double u6 = ((double)w[6] + 0.5) * 2.3283064365386963e-10;
        if (left) {
            in_left = 1;
            if (p->zz < 1.5625) { /* 1/T: exponential-pair proposal of Devroye (2009) */
                const double e2 = -log((u53_from(w[4], w[5]) + 1.0) * 1.1102230246251565e-16);
                if (L * L > 3.125 * e2) continue; /* 2/T */
                x = 1.0 + PGO_T * L;
                x = PGO_T / (x * x);
                if (p->zz > 0.0 && !(u6 < exp(-0.5 * p->z2 * x))) continue;
            }
            else { /* Michael, Schucany & Haas; smaller root written without cancellation */
                const double c = cos(2.0 * PGO_PI * (u53_from(w[4], w[5]) * 1.1102230246251565e-16));
                const double y2 = 2.0 * L * c * c;
                const double mu2 = 1.0 / p->z2;
                const double ww = (p->zz + 0.5 * y2) * mu2;
                x = mu2 / (ww + sqrt(fmax(ww * ww - mu2, 0.0)));
                if (u6 * (1.0 + x * p->zz) > 1.0) {
                    x = mu2 / x;
                }
                if (x >= PGO_T) continue;
            }
            in_left = 0;
        }
        else {
            x = PGO_T + L / p->K;
        }
        if (x > PGO_T) {
            lp = 0.0;
            e = 0.5 * PGO_PI * PGO_PI * x;
        }
        else {
            lp = -1.5 * (PGO_LOGPI_2 + log(x));
            e = 2.0 / x;
        }
        double s = dev_coef(0, lp, e);
        double u = (((double)w[7] + 0.5) * 2.3283064365386963e-10) * s;
        s -= dev_coef(1, lp, e);
        if (u <= s) return x;
        int rejected = 0;
        for (int n = 2; n < 400; n += 2) {
            s += dev_coef(n, lp, e);
            if (u > s) {
                rejected = 1;
                break;
            }
            s -= dev_coef(n + 1, lp, e);
            if (u <= s) return x;
        }
        if (!rejected) return x; /* unreachable in practice: terms vanish super-exponentially */
    }
}

/* random_polyagamma_devroye (src/pgm_devroye.c:211-227), one element */
static double
dev_sample(pgo_rng* r, double h, double z)
{
    dev_par p;
    dev_setup(&p, z);
    double acc = 0.0;
    uint64_t hi = (uint64_t)h; /* truncation as `size_t hi = h` (:221) */
    while (hi--) {
        acc += dev_jacobi_star(r, &p);
    }
    return 0.25 * acc;
}

/* ------------------------------------------------------------------------------------------
 * Alternate sampler (src/pgm_alternate.c)
 * ---------------------------------------------------------------------------------------- */

/* Truncation points t(h), h = 1, 1.125, ..., 4: numerical data of
 * src/pgm_alternate_trunc_points.h:5-23 (output of scripts/generate_alternate_truncation_pts.py) */
static const double g_trunc_f[25] = {
    1.273239366, 1.901515423, 2.281992126, 2.607829551, 2.910421526, 3.200449543, 3.482766779,
    3.759955106, 4.033540671, 4.304486011, 4.573437633, 4.840840644, 5.107017272, 5.372204821,
    5.636581947, 5.900288573, 6.163432428, 6.426094330, 6.688351603, 6.950254767, 7.211854235,
    7.473186206, 7.734284136, 7.995175158, 8.255882407};

/* get_truncation_point (src/pgm_alternate.c:38-70).  The reference's "interpolation" is a step
 * function with an out-of-bounds read for h in (1, 1.125); t only has to be used
 * consistently by the envelope and the mixture weights, so plain linear interpolation is used. */
double
pgo_alternate_trunc_point(double h)
{
    if (h <= 1.0) return g_trunc_f[0];
    if (h >= 4.0) return g_trunc_f[24];
    double pos = (h - 1.0) * 8.0;
    int i = (int)pos;
    double frac = pos - i;
    return g_trunc_f[i] + (g_trunc_f[i + 1] - g_trunc_f[i]) * frac;
}

typedef struct {
    double w_right; /* q/(p+q) */
    double h, zz, z2, t, t_inv, lambda, half_h2, lgammah, hlog2, h_z, h_z2;
    double ca;      /* log(pi/2) (default) or log(sqrt(pi/2)) (ref_compat), :90 */
} alt_par;

/* set_sampling_parameters (src/pgm_alternate.c:139-176), weights in log space */
static void
alt_setup(alt_par* p, double h, double z_half, int compat)
{
    p->h = h;
    p->zz = z_half;
    p->t = pgo_alternate_trunc_point(h);
    p->t_inv = 1.0 / p->t;
    p->half_h2 = 0.5 * h * h;
    p->lgammah = pgo_lgamma(h);
    p->hlog2 = h * PGO_LOG2;
    p->ca = compat ? 0.5 * PGO_LOGPI_2 : PGO_LOGPI_2;
    double lp;
    double a = h / sqrt(2.0 * p->t);
    if (z_half > 0.0) {
        p->z2 = z_half * z_half;
        p->h_z = h / z_half;
        p->h_z2 = p->h_z * p->h_z;
        p->lambda = PGO_PI2_8 + 0.5 * p->z2;
        double b = z_half * sqrt(0.5 * p->t);
        double hz = h * z_half;
        /* invgauss_cdf (:104-114): 0.5 [erfc(a-b) + e^{2hz} erfc(a+b)], times e^{-hz} */
        double t1 = 0.5 * erfc(a - b) * exp(-hz);
        double t2 = (a + b < 26.5 && hz < 700.0) ? 0.5 * erfc(a + b) * exp(hz) : 0.0;
        lp = p->hlog2 + log(t1 + t2);
    }
    else {
        p->z2 = 0.0;
        p->h_z = INFINITY;
        p->h_z2 = INFINITY;
        p->lambda = PGO_PI2_8;
        lp = p->hlog2 + log(erfc(a)); /* :167,170 */
    }
    /* q = (pi/2 / lambda)^h Q(h, lambda t)  (:172-173) */
    double lq = h * (PGO_LOGPI_2 - log(p->lambda)) + pgo_log_upper_gamma(h, p->lambda * p->t, 1);
    double d = lp - lq;
    p->w_right = 1.0 / (1.0 + exp(d));
    if (!(p->w_right >= 0.0)) p->w_right = 0.0;
}

static double
alt_inv_left_gamma(pgo_rng* r, const alt_par* p)
{
    return 1.0 / left_bounded_gamma(r, 0.5, p->half_h2, p->t_inv);
}

/* random_right_bounded_invgauss (src/pgm_alternate.c:200-218): IG(h/z, h^2) | x < t */
static double
alt_trunc_invgauss(pgo_rng* r, const alt_par* p)
{
    double x;
    if (p->t < p->h_z) {
        do {
            x = alt_inv_left_gamma(r, p);
        } while (!(rng_uniform32(r) < exp(-0.5 * p->z2 * x)));
        return x;
    }
    do {
        double y = rng_normal(r);
        double w = p->h_z + 0.5 * y * y / p->z2;
        x = p->h_z2 / (w + sqrt(fmax(w * w - p->h_z2, 0.0)));
        if (rng_uniform_co(r) * (p->h_z + x) > p->h_z) {
            x = p->h_z2 / x;
        }
    } while (x >= p->t);
    return x;
}

/* random_jacobi_star (src/pgm_alternate.c:229-263) with piecewise_coef (:75-83) and
 * bounding_kernel (:86-99).  The Gamma ratio Gamma(n+h)/(Gamma(h) n!) is carried by the
 * recurrence c_n = c_{n-1} (n-1+h)/n.  Decisions: accept on an odd partial sum that covers u
 * (:252-254); reject on an even partial sum below u once the coefficients are decreasing
 * (the reference waits until the FP32 term vanishes, :256-259 -- same decision, later). */
static double
alt_jacobi_star(pgo_rng* r, const alt_par* p)
{
    for (;;) {
        double x;
        if (rng_uniform32(r) < p->w_right) {
            x = left_bounded_gamma(r, p->h, p->lambda, p->t);
        }
        else if (p->zz > 0.0) {
            x = alt_trunc_invgauss(r, p);
        }
        else {
            x = alt_inv_left_gamma(r, p);
        }
        double logx = log(x);
        double lk;
        if (x > p->t) {
            lk = p->h * p->ca + (p->h - 1.0) * logx - PGO_PI2_8 * x - p->lgammah;
        }
        else {
            lk = p->hlog2 - p->half_h2 / x - 1.5 * logx - PGO_LS2PI + log(p->h);
        }
        double u = rng_uniform32(r) * exp(lk);
        double base = p->hlog2 - PGO_LS2PI - 1.5 * logx;
        double inv2x = 0.5 / x;
        double c = 1.0;
        double a_prev = p->h * exp(base - p->h * p->h * inv2x);
        double s = a_prev;
        int decided = 0, accept = 0;
        for (int n = 1; n < 2000; ++n) {
            double tn = 2.0 * n + p->h;
            c *= (n - 1.0 + p->h) / n;
            double a = c * tn * exp(base - tn * tn * inv2x);
            if (n & 1) {
                s -= a;
                if (u <= s) {
                    decided = 1;
                    accept = 1;
                    break;
                }
            }
            else {
                s += a;
                if (a <= a_prev && u > s) {
                    decided = 1;
                    break;
                }
            }
            a_prev = a;
        }
        if (decided && accept) return x;
        /* undecided after 2000 terms cannot happen for finite x; treat as reject */
    }
}

/* random_polyagamma_alternate (src/pgm_alternate.c:294-323), one element */
static double
alt_sample(pgo_rng* r, double h, double z, int compat)
{
    alt_par p;
    double zz = 0.5 * fabs(z);
    double acc = 0.0;
    if (!(zz > 0.0)) zz = 0.0; /* NaN z behaves as z = 0 */
    if (h > 4.0) {
        double chunk = (h >= 5.0) ? 4.0 : 3.0; /* :302 */
        alt_setup(&p, chunk, zz, compat);
        while (h > 4.0) {
            acc += alt_jacobi_star(r, &p);
            h -= chunk;
        }
    }
    alt_setup(&p, h, zz, compat);
    acc += alt_jacobi_star(r, &p);
    return 0.25 * acc;
}

/* ------------------------------------------------------------------------------------------
 * Saddle-point sampler (src/pgm_saddle.c)
 * ---------------------------------------------------------------------------------------- */

static double
log_cosh(double s)
{
    s = fabs(s);
    return s + log1p(exp(-2.0 * s)) - PGO_LOG2;
}

/* cumulant_prime (src/pgm_saddle.c:100-107): K'(u) and K''(u), exact tan/tanh instead of the
 * rational tanh_x (:56-74) / FP32 tanf (:77-81); series near u = 0 where (1-f)/s is 0/0 */
static void
sad_kprime(double u, double* f, double* fp)
{
    double s2 = 2.0 * u;
    if (fabs(s2) < 1e-4) {
        *f = 1.0 + s2 * (1.0 / 3.0 + s2 * (2.0 / 15.0 + s2 * (17.0 / 315.0)));
        *fp = (*f) * (*f) - (1.0 / 3.0 + s2 * (2.0 / 15.0 + s2 * (17.0 / 315.0)));
        return;
    }
    if (s2 > 0.0) {
        double s = sqrt(s2);
        *f = tan(s) / s;
    }
    else {
        double s = sqrt(-s2);
        *f = tanh(s) / s;
    }
    *fp = (*f) * (*f) + (1.0 - *f) / s2;
}

/* cumulant macro (src/pgm_saddle.c:91-94) without the log_cosh_z offset */
static double
sad_cumulant0(double u)
{
    if (u < 0.0) return -log_cosh(sqrt(-2.0 * u));
    if (u > 0.0) return -log(cos(sqrt(2.0 * u)));
    return 0.0;
}

/* select_starting_guess (src/pgm_saddle.c:119-125); for x <= 0.25 the large-argument solution
 * of tanh(s)/s = x is used instead of the constant -9 */
static double
sad_guess(double x)
{
    if (x <= 0.25) {
        double s = (1.0 - 2.0 * exp(-2.0 / x)) / x;
        return -0.5 * s * s;
    }
    return x <= 0.5 ? -1.78 : x <= 1.0 ? -0.147 : x <= 1.5 ? 0.345 : x <= 2.5 ? 0.72
           : x <= 4.0 ? 0.95 : 1.15;
}

#define PGO_SADDLE_MAXIT 25
#define PGO_SADDLE_UMAX 1.2337005501361697 /* just below pi^2/8 */

/* newton_raphson (src/pgm_saddle.c:140-158): root u of K'(u) = x.  Converged when the residual
 * is <= 1e-9 x (the reference: 1e-5 absolute); returns K''(u) through fprime. */
double
pgo_saddle_newton(double x, double* fprime)
{
    double u = sad_guess(x);
    double f, fp;
    for (int n = 0; n < PGO_SADDLE_MAXIT; ++n) {
        sad_kprime(u, &f, &fp);
        double fval = f - x;
        if (fabs(fval) <= 1e-9 * x || !(fp > 1e-300)) {
            break;
        }
        double un = u - fval / fp;
        if (un >= PGO_SADDLE_UMAX) {
            un = 0.5 * (u + PGO_SADDLE_UMAX); /* never step onto/over the pole of tan */
        }
        if (un == u) {
            break;
        }
        u = un;
    }
    sad_kprime(u, &f, &fp);
    *fprime = fp;
    return u;
}

typedef struct {
    double w_left; /* p/(p+q) */
    double h, zz, half_z2, log_cosh_z;
    double xl, xc, logxc;
    double slope_l, icpt_l, slope_r, icpt_r;
    double log_coef_l, log_coef_r, log_sqrt_h2pi;
    double hrho;   /* h * rho_r */
    double sgn;    /* -1 (consistent envelope) or +1 (ref_compat, src/pgm_saddle.c:257) */
} sad_par;

/* invgauss_logcdf (src/pgm_saddle.c:282-292), Giner & Smyth (2016) */
static double
sad_invgauss_logcdf(double x, double mu, double lambda)
{
    double qm = x / mu;
    double tm = mu / lambda;
    double rr = sqrt(x / lambda);
    double a = log_ndtr((qm - 1.0) / rr);
    double b = 2.0 / tm + log_ndtr(-(qm + 1.0) / rr);
    return a + log1p(exp(b - a));
}

/* set_sampling_parameters (src/pgm_saddle.c:180-230) and the mixture weight of
 * random_polyagamma_saddle (:317-328), the latter in log space */
static void
sad_setup(sad_par* p, double h, double zz, int compat)
{
    double xl, logxl;
    if (zz > 0.0) {
        xl = tanh(zz) / zz;
        logxl = log(xl);
        p->half_z2 = 0.5 * zz * zz;
        p->log_cosh_z = log_cosh(zz);
    }
    else {
        xl = 1.0;
        logxl = 0.0;
        p->half_z2 = 0.0;
        p->log_cosh_z = 0.0;
    }
    p->h = h;
    p->zz = zz;
    p->xl = xl;
    p->xc = 2.75 * xl;
    double xr = 3.0 * xl;
    double xc_inv = 1.0 / p->xc;
    double xl_inv = 1.0 / xl;
    double fp_r, fp_c;
    double ur = pgo_saddle_newton(xr, &fp_r);
    (void)pgo_saddle_newton(p->xc, &fp_c);
    double tr = ur + p->half_z2;

    p->slope_l = -0.5 * xl_inv * xl_inv;
    /* cumulant(ul) = log cosh z - log cosh(sqrt(z^2)) = 0  (:216) */
    p->icpt_l = -0.5 * xc_inv + xl_inv;
    p->logxc = 1.0116009116784799 + logxl; /* log(2.75) */
    p->slope_r = -tr - 1.0 / xr;
    p->icpt_r = (p->log_cosh_z + sad_cumulant0(ur)) + 1.0 - 1.0986122886681098 - logxl + p->logxc;

    double alpha_r = fp_c * xc_inv * xc_inv;
    double alpha_l = xc_inv * alpha_r;
    p->log_sqrt_h2pi = 0.5 * log(h / (2.0 * PGO_PI));
    p->log_coef_l = p->log_sqrt_h2pi - 0.5 * log(alpha_l);
    p->log_coef_r = p->log_sqrt_h2pi - 0.5 * log(alpha_r);
    p->sgn = compat ? 1.0 : -1.0;

    double sqrt_rho = xl_inv; /* sqrt(-2 slope_l) */
    double lp = h * (0.5 * xc_inv + p->icpt_l - sqrt_rho) + sad_invgauss_logcdf(p->xc, xl, h) -
                0.5 * log(alpha_l);
    p->hrho = -h * p->slope_r;
    double lq = pgo_log_upper_gamma(h, p->hrho * p->xc, 0) + p->log_coef_r +
                h * (p->icpt_r - (compat ? 0.0 : p->logxc) - log(p->hrho));
    p->w_left = 1.0 / (1.0 + exp(lq - lp));
    if (!(p->w_left >= 0.0)) p->w_left = 1.0;
}

/* random_polyagamma_saddle (src/pgm_saddle.c:309-350), one element.  The acceptance test
 * u k(x) <= phi(x) (saddle_point :235-244, bounding_kernel :250-263) is taken in log space. */
static double
sad_sample(pgo_rng* r, double h, double z, int compat)
{
    sad_par p;
    double zz = 0.5 * fabs(z);
    if (!(zz > 0.0)) zz = 0.0;
    sad_setup(&p, h, zz, compat);
    double mu = p.xl, mu2 = p.xl * p.xl;
    double x;
    for (;;) {
        if (rng_uniform32(r) < p.w_left) {
            do { /* IG(mu = xl, lambda = h) | x < xc */
                double y = rng_normal(r);
                double w = mu + 0.5 * mu2 * y * y / h;
                x = mu2 / (w + sqrt(fmax(w * w - mu2, 0.0)));
                if (rng_uniform_co(r) * (mu + x) > mu) {
                    x = mu2 / x;
                }
            } while (x >= p.xc);
        }
        else {
            x = left_bounded_gamma(r, h, p.hrho, p.xc);
        }
        double logx = log(x);
        double lk;
        if (x > p.xc) {
            lk = h * (p.sgn * p.logxc + p.slope_r * x + p.icpt_r) + (h - 1.0) * logx + p.log_coef_r;
        }
        else {
            lk = 0.5 * h * (1.0 / p.xc - 1.0 / x) + h * (p.slope_l * x + p.icpt_l) - 1.5 * logx +
                 p.log_coef_l;
        }
        double fp;
        double u = pgo_saddle_newton(x, &fp);
        double t = u + p.half_z2;
        double lphi = h * (p.log_cosh_z + sad_cumulant0(u) - t * x) + p.log_sqrt_h2pi - 0.5 * log(fp);
        if (log(rng_uniform32(r)) + lk <= lphi) {
            break;
        }
    }
    return 0.25 * h * x;
}

/* ------------------------------------------------------------------------------------------
 * Normal approximation and gamma convolution (src/pgm_random.c)
 * ---------------------------------------------------------------------------------------- */

/* random_polyagamma_normal_approx (src/pgm_random.c:47-66).  sech^2 and sinh(z) - z are taken
 * in cancellation-free forms (SURVEY 8-DEFECTS d: `1 - tanh^2` is 0 for |z| >= 40). */
static void
normal_moments(double h, double z, double* mean, double* sd)
{
    double a = fabs(z);
    if (!(a > 0.0)) {
        *mean = 0.25 * h;
        *sd = sqrt(h / 24.0);
        return;
    }
    double e = exp(-a);
    double ope = 1.0 + e;
    *mean = 0.5 * h * ((1.0 - e) / ope) / a; /* tanh(a/2) = (1-e^-a)/(1+e^-a) */
    double var;
    if (a < 1.0) {
        double a2 = a * a;
        /* (sinh a - a)/a^3 = sum a^(2k) / (2k+3)! */
        double S = 1.0 / 6.0 + a2 * (1.0 / 120.0 + a2 * (1.0 / 5040.0 + a2 * (1.0 / 362880.0 +
                   a2 * (1.0 / 39916800.0 + a2 * (1.0 / 6227020800.0 + a2 * (1.0 / 1307674368000.0))))));
        var = 0.25 * h * S * (4.0 * e / (ope * ope));
    }
    else {
        /* (sinh a - a) sech^2(a/2) = 2 (1 - e^-2a - 2 a e^-a) / (1 + e^-a)^2 */
        var = 0.25 * h * 2.0 * (1.0 - e * e - 2.0 * a * e) / (ope * ope * a * a * a);
    }
    *sd = sqrt(var);
}

static double
normal_sample(pgo_rng* r, double h, double z)
{
    double mean, sd;
    normal_moments(h, z, &mean, &sd);
    return mean + rng_normal(r) * sd;
}

/* random_polyagamma_gamma_conv (src/pgm_random.c:79-97): 200-term truncated sum */
static double
gamma_sample(pgo_rng* r, double h, double z)
{
    double zz = 0.5 * fabs(z);
    if (!(zz > 0.0)) zz = 0.0;
    const double pi2 = 9.869604401089358;
    double z2 = zz * zz;
    double acc = 0.0;
    pgo_spare sp = {0, 0.0};
    for (int k = 0; k < 200; ++k) {
        double m = k + 0.5;
        acc += rng_gamma(r, h, &sp) / (pi2 * m * m + z2);
    }
    return 0.5 * acc;
}

/* ------------------------------------------------------------------------------------------
 * Dispatch (src/pgm_random.c:105-180)
 * ---------------------------------------------------------------------------------------- */

/* random_polyagamma_hybrid (src/pgm_random.c:105-121); compares the SIGNED z like the reference */
int
pgo_hybrid_select(double h, double z)
{
    if (h > 50.0) return PGO_NORMAL;
    if (h >= 8.0 || (h > 4.0 && z <= 4.0)) return PGO_SADDLE;
    if (h == 1.0 || (h == floor(h) && z <= 1.0)) return PGO_DEVROYE;
    return PGO_ALTERNATE;
}

static _Thread_local uint64_t g_last_words = 0;

uint64_t
pgo_last_words_consumed(void)
{
    return g_last_words;
}

/* pgm_random_polyagamma (src/pgm_random.c:144-155) with nan_or_inf (:137-141) */
static double
sample_one(const uint32_t key[2], uint64_t index, double h, double z, int method)
{
    int compat = (method & PGO_REF_COMPAT) != 0;
    int m = method & 0xff;
    if (!(h > 0.0)) return NAN;          /* h <= 0 or NaN */
    if (isinf(h)) return INFINITY;
    if (m == PGO_HYBRID) m = pgo_hybrid_select(h, z);
    pgo_rng r;
    rng_init(&r, key, index);
    double out;
    switch (m) {
        case PGO_GAMMA: out = gamma_sample(&r, h, z); break;
        case PGO_DEVROYE: out = dev_sample(&r, h, z); break;
        case PGO_ALTERNATE: out = alt_sample(&r, h, z, compat); break;
        case PGO_SADDLE: out = sad_sample(&r, h, z, compat); break;
        default: out = normal_sample(&r, h, z); break;
    }
    g_last_words = r.words;
    return out;
}

/* pgm_random_polyagamma_fill (src/pgm_random.c:158-170) */
void
pgo_random_polyagamma_fill(uint64_t seed, uint64_t offset, double h, double z, int method,
                           size_t n, double* out, uint64_t first_index)
{
    uint32_t key[2];
    pgo_call_key(seed, offset, key);
    for (size_t i = 0; i < n; ++i) {
        out[i] = sample_one(key, first_index + i, h, z, method);
    }
}

/* pgm_random_polyagamma_fill2 (src/pgm_random.c:173-180) */
void
pgo_random_polyagamma_fill2(uint64_t seed, uint64_t offset, const double* h, const double* z,
                            int method, size_t n, double* out, uint64_t first_index)
{
    uint32_t key[2];
    pgo_call_key(seed, offset, key);
    for (size_t i = 0; i < n; ++i) {
        out[i] = sample_one(key, first_index + i, h[i], z[i], method);
    }
}

void
pgo_setup_probe(int method, double h, double z, double* out)
{
    int compat = (method & PGO_REF_COMPAT) != 0;
    double zz = 0.5 * fabs(z);
    switch (method & 0xff) {
        case PGO_DEVROYE: {
            dev_par p;
            dev_setup(&p, z);
            out[0] = p.w;
            out[1] = p.K;
            break;
        }
        case PGO_ALTERNATE: {
            alt_par p;
            alt_setup(&p, h, zz, compat);
            out[0] = p.w_right;
            out[1] = p.t;
            out[2] = p.lambda;
            break;
        }
        case PGO_SADDLE: {
            sad_par p;
            sad_setup(&p, h, zz, compat);
            out[0] = p.w_left;
            out[1] = p.xc;
            out[2] = p.slope_l;
            out[3] = p.icpt_l;
            out[4] = p.slope_r;
            out[5] = p.icpt_r;
            out[6] = p.log_coef_l;
            out[7] = p.log_coef_r;
            break;
        }
        default: {
            normal_moments(h, z, &out[0], &out[1]);
            break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Densities (src/pgm_density.c) -- term-for-term restatement with libm lgamma, so that this
 * file agrees with the compiled reference to rounding; the CUDA kernels (which use Gamma-ratio
 * recurrences) are then compared against it at 1e-10.
 * ---------------------------------------------------------------------------------------- */
#define PGO_MAX_SERIES_TERMS 200
#define PGO_DBL_EPSILON 2.220446049250313e-16
#define PGO_2PI 6.283185307179586

static int
isclose_ulp(double a, double b)
{
    /* PGM_ISCLOSE(a, b, 0, DBL_EPSILON)  (src/pgm_macros.h:47-48) */
    double m = fabs(a) > fabs(b) ? fabs(a) : fabs(b);
    return fabs(a - b) <= PGO_DBL_EPSILON * m;
}

static double g_cond; /* sum |term| / |sum term| of the last series */

/* pgm_polyagamma_pdf (src/pgm_density.c:67-92) */
double
pgo_polyagamma_pdf(double x, double h, double z)
{
    g_cond = 1.0;
    if (x <= 0.0 || isinf(x) || isnan(x)) {
        return isnan(x) ? NAN : 0.0;
    }
    double a = (fabs(z) > 0.0 ? h * log(cosh(0.5 * z)) - 0.5 * z * z * x : 0.0) + (h - 1.0) * PGO_LOG2;
    double sum = exp(a - 0.125 * h * h / x) * h;
    double asum = fabs(sum);
    double sign = -1.0;
    a -= pgo_lgamma(h);
    for (unsigned n = 1; n < PGO_MAX_SERIES_TERMS; ++n, sign = -sign) {
        double twonh = 2.0 * n + h;
        double term = exp(a + pgo_lgamma(n + h) - 0.125 * twonh * twonh / x - pgo_lgamma(n + 1.0)) * twonh;
        double prev = sum;
        sum += sign * term;
        asum += fabs(term);
        if (isclose_ulp(sum, prev)) break;
    }
    g_cond = asum / fabs(sum);
    return sum / sqrt(PGO_2PI * x * x * x);
}

/* pgm_polyagamma_logpdf (src/pgm_density.c:100-121): no early exit */
double
pgo_polyagamma_logpdf(double x, double h, double z)
{
    g_cond = 1.0;
    if (x <= 0.0 || isinf(x) || isnan(x)) {
        return isnan(x) ? NAN : -INFINITY;
    }
    double lg = pgo_lgamma(h);
    double a = (fabs(z) > 0.0 ? h * log(cosh(0.5 * z)) - 0.5 * z * z * x : 0.0) + (h - 1.0) * PGO_LOG2 -
               PGO_LS2PI - 1.5 * log(x) - lg;
    double first = lg - 0.125 * h * h / x;
    double sum = 1.0, asum = 1.0, sign = -1.0;
    for (unsigned n = 1; n < PGO_MAX_SERIES_TERMS; ++n, sign = -sign) {
        double t = 2.0 * n + h;
        double curr = pgo_lgamma(n + h) - 0.125 * t * t / x - pgo_lgamma(n + 1.0);
        double term = exp(curr - first) * t / h;
        sum += sign * term;
        asum += term;
    }
    g_cond = asum / fabs(sum);
    return a + (log(h) + first + log(sum));
}

typedef struct {
    double s2x, a, x, z;
} cdf_args;

/* invgamma_logcdf (src/pgm_density.c:149-165) incl. the A&S 7.1.28-style tail branch */
static double
invgamma_logcdf(const cdf_args* arg)
{
    double x = 0.5 * arg->a / arg->s2x;
    if (x > 26.55) {
        const double p0 = 0.3275911, p1 = 0.254829592, p2 = -0.284496736, p3 = 1.421413741,
                     p4 = -1.453152027, p5 = 1.061405429;
        double t = 1.0 / (1.0 + p0 * x);
        return log(t * (p1 + t * (p2 + t * (p3 + t * (p4 + t * p5))))) - x * x;
    }
    return log(erfc(x));
}

/* norm_logcdf (src/pgm_density.c:181-209) incl. the Bryc (2001) branch for x < -37.5 */
static double
norm_logcdf(double x)
{
    if (x < -37.5) {
        const double p0 = 12.77436324, p1 = 5.575192695, q0 = 25.54872648, q1 = 31.53531977,
                     q2 = 14.38718147, s2p = 2.5066282746310002;
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

/* invgauss_logcdf (src/pgm_density.c:216-225) */
static double
invgauss_logcdf(const cdf_args* arg)
{
    double qm = 2.0 * arg->x * arg->z / arg->a;
    double rr = 2.0 * arg->s2x / arg->a;
    double a = norm_logcdf((qm - 1.0) / rr);
    double b = arg->z * arg->a + norm_logcdf(-(qm + 1.0) / rr);
    return a + log1p(exp(b - a));
}

/* pgm_polyagamma_cdf (src/pgm_density.c:236-280) */
double
pgo_polyagamma_cdf(double x, double h, double z)
{
    g_cond = 1.0;
    if (isnan(x)) return NAN;
    if (x <= 0.0) return 0.0;
    if (isinf(x)) return 1.0;
    z = fabs(z);
    double c, zn;
    int ig = z > 0.0;
    cdf_args arg = {.a = h, .x = x, .z = z};
    if (ig) {
        c = h * log1p(exp(-z));
        arg.s2x = sqrt(x);
        zn = z;
    }
    else {
        c = h * PGO_LOG2;
        arg.s2x = sqrt(2.0 * x);
        zn = 0.0;
    }
    double sum = exp(c + (ig ? invgauss_logcdf(&arg) : invgamma_logcdf(&arg)));
    double asum = fabs(sum), sign = -1.0;
    c -= pgo_lgamma(h);
    for (unsigned n = 1; n < PGO_MAX_SERIES_TERMS; ++n, sign = -sign, zn = z * n) {
        arg.a = 2.0 * n + h;
        double lc = ig ? invgauss_logcdf(&arg) : invgamma_logcdf(&arg);
        double term = exp(c + pgo_lgamma(n + h) + lc - pgo_lgamma(n + 1.0) - zn);
        double prev = sum;
        sum += sign * term;
        asum += fabs(term);
        if (isclose_ulp(sum, prev)) break;
    }
    g_cond = asum / fabs(sum);
    return sum;
}

/* pgm_polyagamma_logcdf (src/pgm_density.c:291-331): no early exit */
double
pgo_polyagamma_logcdf(double x, double h, double z)
{
    g_cond = 1.0;
    if (isnan(x)) return NAN;
    if (x <= 0.0) return -INFINITY;
    if (isinf(x)) return 0.0;
    z = fabs(z);
    double c, zn;
    int ig = z > 0.0;
    double lg = pgo_lgamma(h);
    cdf_args arg = {.a = h, .x = x, .z = z};
    if (ig) {
        c = h * log1p(exp(-z)) - lg;
        arg.s2x = sqrt(x);
        zn = z;
    }
    else {
        c = h * PGO_LOG2 - lg;
        arg.s2x = sqrt(2.0 * x);
        zn = 0.0;
    }
    double first = lg + (ig ? invgauss_logcdf(&arg) : invgamma_logcdf(&arg));
    double sum = 1.0, asum = 1.0, sign = -1.0;
    for (unsigned n = 1; n < PGO_MAX_SERIES_TERMS; ++n, sign = -sign, zn = z * n) {
        arg.a = 2.0 * n + h;
        double lc = ig ? invgauss_logcdf(&arg) : invgamma_logcdf(&arg);
        double curr = pgo_lgamma(n + h) + lc - pgo_lgamma(n + 1.0) - zn;
        double term = exp(curr - first);
        sum += sign * term;
        asum += term;
    }
    g_cond = asum / fabs(sum);
    return c + (first + log(sum));
}

void
pgo_density_array(int which, const double* x, const double* h, const double* z, size_t n,
                  double* out, double* cond)
{
    for (size_t i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = pgo_polyagamma_pdf(x[i], h[i], z[i]); break;
            case 1: out[i] = pgo_polyagamma_cdf(x[i], h[i], z[i]); break;
            case 2: out[i] = pgo_polyagamma_logpdf(x[i], h[i], z[i]); break;
            default: out[i] = pgo_polyagamma_logcdf(x[i], h[i], z[i]); break;
        }
        if (cond) cond[i] = g_cond;
    }
}
