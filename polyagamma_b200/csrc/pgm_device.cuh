// pgm_device.cuh -- device-side building blocks shared by the sampler and density kernels:
// Philox4x32-10 element streams, the variates built on them, and the FP64 special functions
// the samplers' per-element setup needs.  sm_100a only.
//
// Replaces: numpy's bitgen_t + ziggurat/legacy variates (call sites listed in SURVEY 8a C6),
// src/pgm_macros.h:54-62 (next_float / next_double) and src/pgm_common.h (pgm_lgamma,
// random_left_bounded_gamma, upper_incomplete_gamma).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pgm {

constexpr double kPi = 3.141592653589793238462643383279503;
constexpr double kPi_2 = 1.570796326794896619231321691639751;
constexpr double kPi2_8 = 1.233700550136169827354311374984519;
constexpr double kLogPi_2 = 0.4515827052894548647261952298948821;  // log(pi/2)
constexpr double kLs2Pi = 0.9189385332046727417803297364056177;    // log(sqrt(2 pi))
constexpr double kLog2 = 0.6931471805599453094172321214581766;

// ---------------------------------------------------------------------------------------
// Philox4x32-10.  Host+device so the ABI layer derives the call key with the same code.
// ---------------------------------------------------------------------------------------
// rk: the ten round keys (k0 + r W0, k1 + r W1), precomputed on the host.  A kernel that inlines the
// block function reads them straight from its parameter bank as instruction operands, instead of
// issuing 18 uniform adds per block to bump the key.
struct PhiloxKey {
    uint32_t k0, k1;
    uint32_t rk[20];
};

__host__ __device__ __forceinline__ void
philox_round_keys(PhiloxKey& key)
{
    uint32_t k0 = key.k0, k1 = key.k1;
    for (int r = 0; r < 10; ++r) {
        key.rk[2 * r] = k0;
        key.rk[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__host__ __device__ __forceinline__ void
philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one 32x32->64 multiply per half-round (IMAD.WIDE.U32 on the device)
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// the same ten rounds with the round keys given
__device__ __forceinline__ void
philox4x32_10_rk(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, const uint32_t* __restrict__ rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
    }
}

// Call key = first two output words of Philox(key = seed, ctr = (offset, 'PGM1', 'key!')).
__host__ __device__ inline PhiloxKey
derive_call_key(uint64_t seed, uint64_t offset)
{
    uint32_t c0 = (uint32_t)offset, c1 = (uint32_t)(offset >> 32), c2 = 0x50474D31u, c3 = 0x6B657921u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    PhiloxKey key;
    key.k0 = c0;
    key.k1 = c1;
    philox_round_keys(key);
    return key;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------
// fm:: -- lean FP64 elementary functions for the sampling kernels.
//
// ncu showed ~70 % of the instructions of the sampling kernels are not FP64: libdevice's
// log/exp/cospi/division/sqrt spend most of their issue slots on special-case tests, exponent
// bookkeeping and UMOV pairs that materialise every 64-bit polynomial coefficient.  These versions
// assume the arguments the samplers actually produce (finite, normal; anything else falls back to
// libdevice), take their coefficients from __constant__ memory (two doubles per LDCU.128) and
// replace IEEE division / square root by MUFU seeds with two Newton steps.  Accuracy (checked in
// tests/test_gpu_parity.py::test_fast_math): log, exp <= 2 ulp; rcp, sqrt <= 1 ulp; cos(2 pi u)
// <= 1e-15 absolute.  Polynomials: fdlibm e_log.c (Lg1..Lg7), k_sin.c / k_cos.c; exp: degree-13
// Taylor on |r| <= ln2/2.
// ---------------------------------------------------------------------------------------
namespace fm {

static __constant__ double c_log[8] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                       2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                       1.479819860511658591e-01, 0.0};
static __constant__ double c_exp[14] = {1.0,
                                        1.0,
                                        0.5,
                                        1.6666666666666666e-01,
                                        4.1666666666666664e-02,
                                        8.3333333333333332e-03,
                                        1.3888888888888889e-03,
                                        1.9841269841269841e-04,
                                        2.4801587301587302e-05,
                                        2.7557319223985893e-06,
                                        2.7557319223985888e-07,
                                        2.5052108385441720e-08,
                                        2.0876756987868100e-09,
                                        1.6059043836821613e-10};
static __constant__ double c_sin[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03,
                                       -1.98412698298579493134e-04, 2.75573137070700676789e-06,
                                       -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double c_cos[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03,
                                       2.48015872894767294178e-05,  -2.75573143513906633035e-07,
                                       2.08757232129817482790e-09,  -1.13596475577881948265e-11};
constexpr double kLn2Hi = 6.93147180369123816490e-01;
constexpr double kLn2Lo = 1.90821492927058770002e-10;

// PGM_TABLE_MATH (default 1): fm::exp by a table of 2^(j/64) and a degree-5 polynomial, fm::cos2pi by ONE
// polynomial whose coefficients (sine or cosine kernel) are loaded per lane -- both read their tables through
// the read-only path (L1 resident: 512 + 96 bytes).  An FP64 instruction cannot take a constant-bank operand
// here, so every polynomial coefficient costs a load besides its FMA; halving the polynomials pays twice.
// 0 makes them the table-free exp_poly / cos2pi_poly everywhere.
#ifndef PGM_TABLE_MATH
#define PGM_TABLE_MATH 1
#endif
// 2^(j/64), correctly rounded (mpmath)
static __device__ const double g_exp2_64[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
// [0] the cosine kernel's coefficients, [1] the sine kernel's (the values of c_cos / c_sin), in pairs
static __device__ const double2 g_sincos_coef[2][3] = {
    {{4.16666666666666019037e-02, -1.38888888888741095749e-03}, {2.48015872894767294178e-05, -2.75573143513906633035e-07},
     {2.08757232129817482790e-09, -1.13596475577881948265e-11}},
    {{-1.66666666666666324348e-01, 8.33333333332248946124e-03}, {-1.98412698298579493134e-04, 2.75573137070700676789e-06},
     {-2.50507602534068634195e-08, 1.58969099521155010221e-10}}};

// The libdevice fallbacks for arguments outside the lean paths' domain.  They are INLINED on purpose.  Out of
// line (__noinline__: a few KB less never-executed code between the hot instructions of every call site) the
// persistent sampling kernels hung intermittently -- sample_kernel<Alternate> on chunked elements, one warp
// never leaving the kernel, once in 2 to 80 calls depending on unrelated code changes, never with the
// fallbacks inlined (tests/test_gpu_parity.py::test_multi_pass_stress).  Calls into functions with a stack
// frame from deep inside divergent rejection loops are avoided throughout for that reason; the hot paths use
// the guard-free log_pn / sqrt_nn below instead.
static __device__ __forceinline__ double slow_sqrt(double x) { return ::sqrt(x); }
static __device__ __forceinline__ double slow_log(double x) { return ::log(x); }
static __device__ __forceinline__ double slow_exp(double x) { return ::exp(x); }

// 1/x for finite normal x: MUFU.RCP64H seed (~20 bits) + two Newton steps
__device__ __forceinline__ double
rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// a/b with one residual correction (result within 1 ulp)
__device__ __forceinline__ double
div(double a, double b)
{
    const double r = rcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// 1/sqrt(x) and sqrt(x) for finite normal x > 0: MUFU.RSQ64H seed + two Newton steps
__device__ __forceinline__ double
rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}

__device__ __forceinline__ double
sqrt(double x)
{
    if (!(x > 2.3e-308 && x < 1.7e308)) return slow_sqrt(x);  // zero, denormal, negative, NaN, inf
    const double y = rsqrt(x);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);
}

// sqrt(x) for finite x >= 0 where an ABSOLUTE error of 1e-150 is harmless (x below 1e-300, zero included, is
// taken as 1e-300): no special-case branch, so the call site stays one basic block.  NaN in, NaN out.
__device__ __forceinline__ double
sqrt_nn(double x)
{
    x = (x < 1e-300) ? 1e-300 : x;
    const double y = rsqrt(x);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);
}

// natural logarithm of a positive NORMAL finite x -- what the samplers feed it almost everywhere (uniforms in
// (0, 1], proposals, shapes): no special-case test, the call site stays one basic block.
__device__ __forceinline__ double
log_pn(double x)
{
    const long long bits = __double_as_longlong(x);
    int e = (int)(bits >> 52) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);  // [1, 2)
    if (m > 1.4142135623730951) {
        m *= 0.5;
        e += 1;
    }
    const double f = m - 1.0;
    const double s = f * rcp(2.0 + f);
    const double z = s * s;
    double R = c_log[6];
#pragma unroll
    for (int k = 5; k >= 0; --k) R = fma(R, z, c_log[k]);
    R *= z;
    const double hfsq = 0.5 * f * f;
    const double de = (double)e;
    return de * kLn2Hi - ((hfsq - (s * (hfsq + R) + de * kLn2Lo)) - f);
}

// natural logarithm, any argument (zero, denormal, negative, inf, NaN: libdevice, out of line)
__device__ __forceinline__ double
log(double x)
{
    if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return slow_log(x);
    return log_pn(x);
}

// e^x for x <= 700; x < -707 flushes to 0 (the samplers never need denormal results), huge or NaN arguments go
// to libdevice.  Two forms, both within 2 ulp: `exp` reads 2^(j/64) from a table and needs a degree-5
// polynomial; `exp_poly` is table-free (degree 13 on |r| <= ln2/2).  Which one is faster depends on the kernel:
// the list kernels and the densities gain 2-4 % from the table, the tile kernel (64 registers, 4 blocks per SM)
// loses 2.5 % to the extra load, so its two methods (normal approximation, left-only saddle) call exp_poly.
__device__ __forceinline__ double
exp_poly(double x)
{
    if (!(x <= 700.0)) return slow_exp(x);  // huge or NaN
    if (x < -707.0) return 0.0;
    // n = round(x / ln 2) by the 1.5 * 2^52 trick: the integer lands in the low word
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const int n = __double2loint(t);
    const double dn = t - 6755399441055744.0;
    double r = fma(dn, -kLn2Hi, x);
    r = fma(dn, -kLn2Lo, r);
    double p = c_exp[13];
#pragma unroll
    for (int k = 12; k >= 0; --k) p = fma(p, r, c_exp[k]);
    // scale by 2^n (|n| <= 1022 here): add n to the exponent field
    return __longlong_as_double(__double_as_longlong(p) + ((long long)n << 52));
}

__device__ __forceinline__ double
exp(double x)
{
#if PGM_TABLE_MATH
    if (!(x <= 700.0)) return slow_exp(x);  // huge or NaN
    if (x < -707.0) return 0.0;
    // k = round(64 x / ln 2) = 64 n + j;  e^x = 2^n 2^(j/64) e^r,  |r| <= ln2/128: degree 5 is below 4e-17
    const double t = fma(x, 92.332482616893656, 6755399441055744.0);
    const int k = __double2loint(t);
    const double dn = t - 6755399441055744.0;
    double r = fma(dn, -kLn2Hi * 0.015625, x);  // (dn kLn2Hi / 64 is exact: 33 + 17 significant bits)
    r = fma(dn, -kLn2Lo * 0.015625, r);
    const double T = __ldg(&g_exp2_64[k & 63]);
    double q = fma(r, 8.3333333333333332e-03, 4.1666666666666664e-02);
    q = fma(q, r, 1.6666666666666666e-01);
    q = fma(q, r, 0.5);
    q = fma(q * r, r, r);  // e^r - 1
    const double v = fma(T, q, T);
    return __longlong_as_double(__double_as_longlong(v) + ((long long)(k >> 6) << 52));
#else
    return exp_poly(x);
#endif
}

// cos(2 pi u) for u in [0, 1]: quadrant from round(4u), sine/cosine kernels on [-pi/4, pi/4].  `cos2pi` runs ONE
// Horner chain on the coefficients of the lane's quadrant (loaded per lane), `cos2pi_poly` both fixed
// polynomials and a select; same values to 1e-16, same trade as exp / exp_poly.
__device__ __forceinline__ double
cos2pi_poly(double u)
{
    const double v = 4.0 * u;
    const double t = v + 6755399441055744.0;
    const int q = __double2loint(t);
    const double r = v - (t - 6755399441055744.0);
    const double y = r * 1.5707963267948966;
    const double y2 = y * y;
    double sn = c_sin[5], cs = c_cos[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        sn = fma(sn, y2, c_sin[k]);
        cs = fma(cs, y2, c_cos[k]);
    }
    sn = fma(y * y2, sn, y);
    cs = fma(y2 * y2, cs, fma(-0.5, y2, 1.0));
    const double val = (q & 1) ? sn : cs;
    return ((q + 1) & 2) ? -val : val;  // q: 0 -> cos, 1 -> -sin, 2 -> -cos, 3 -> sin, 4 -> cos
}

__device__ __forceinline__ double
cos2pi(double u)
{
#if PGM_TABLE_MATH
    const double v = 4.0 * u;
    const double t = v + 6755399441055744.0;
    const int q = __double2loint(t);
    const double r = v - (t - 6755399441055744.0);
    const double y = r * 1.5707963267948966;
    const double y2 = y * y;
    // sine kernel  y + y y2 S(y2)  or cosine kernel  1 - y2/2 + y2 y2 C(y2)
    const bool odd = (q & 1) != 0;
    const double2* c = g_sincos_coef[odd ? 1 : 0];
    const double2 c01 = __ldg(c), c23 = __ldg(c + 1), c45 = __ldg(c + 2);
    double p = fma(c45.y, y2, c45.x);
    p = fma(p, y2, c23.y);
    p = fma(p, y2, c23.x);
    p = fma(p, y2, c01.y);
    p = fma(p, y2, c01.x);
    const double A = odd ? y : fma(-0.5, y2, 1.0);
    const double B = (odd ? y : y2) * y2;
    const double val = fma(B, p, A);
    return ((q + 1) & 2) ? -val : val;
#else
    return cos2pi_poly(u);
#endif
}

// sin and cos of 2 pi u together (Box-Muller pairs)
__device__ __forceinline__ void
sincos2pi(double u, double& s_out, double& c_out)
{
    const double v = 4.0 * u;
    const double t = v + 6755399441055744.0;
    const int q = __double2loint(t);
    const double r = v - (t - 6755399441055744.0);
    const double y = r * 1.5707963267948966;
    const double y2 = y * y;
    double sn = c_sin[5], cs = c_cos[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        sn = fma(sn, y2, c_sin[k]);
        cs = fma(cs, y2, c_cos[k]);
    }
    sn = fma(y * y2, sn, y);
    cs = fma(y2 * y2, cs, fma(-0.5, y2, 1.0));
    // angle = (pi/2)(q + r): rotate by q quadrants
    const double c0 = (q & 1) ? sn : cs, s0 = (q & 1) ? cs : sn;
    c_out = ((q + 1) & 2) ? -c0 : c0;
    s_out = (q & 2) ? -s0 : s0;
}

// erfcx(z) = e^{z^2} erfc(z) for finite z >= 0 on ONE code path: (1 + z/4) erfcx(z) is smooth in
// t = 1 / (1 + z/4) on (0, 1] and a 24-term Chebyshev series in 2t - 1 (Clenshaw recurrence) gives
// <= 2e-15 relative (tools/gen_erfcx_table.py: coefficients from mpmath, error check of this very
// recurrence; tests/test_gpu_parity.py::test_fast_math).  Used by the cdf / logcdf kernels; in the
// sampler setups libdevice's erfc measured marginally faster (one call per element, no loop) and stays.
// libdevice's erf/erfc/erfcx switch between range-specific kernels, which the lanes of a warp with
// mixed arguments run one after the other.
static __constant__ double c_erfcx[24] = {
#include "pgm_erfcx_table.inc"
};

__device__ __forceinline__ double
erfcx_pos(double z)
{
    const double t = rcp(fma(0.25, z, 1.0));
    const double u2 = fma(4.0, t, -2.0);  // 2 (2t - 1)
    double b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int k = 23; k >= 1; --k) {
        const double b0 = fma(u2, b1, c_erfcx[k] - b2);
        b2 = b1;
        b1 = b0;
    }
    return t * fma(0.5 * u2, b1, c_erfcx[0] - b2);
}

// e^{-x^2} with the rounding error of the square carried along (x^2 up to ~700 would otherwise cost
// 7e-14 of relative accuracy)
__device__ __forceinline__ double
exp_neg_sq(double x)
{
    const double hi = x * x;
    const double lo = fma(x, x, -hi);
    return exp(-hi) * (1.0 - lo);
}

}  // namespace fm

// One private stream per output element: words of Philox(call key, ctr = (idx, block, 0)),
// block = 0, 1, 2, ..., consumed strictly in order.  The four buffered words live in
// registers and are shifted down on every draw (no local-memory indexing).
// OUTLINE: keep ONE copy of the block function per kernel (persistent samplers: many call sites,
// instruction-cache bound) or inline it (dense kernels with a single draw site).
// The outlined copy takes and returns VALUES: a member function would take the address of the
// generator and force its state into local memory (STL/LDL around every call).
static __device__ __noinline__ uint4
philox_block_outlined(uint32_t i0, uint32_t i1, uint32_t blk, uint32_t k0, uint32_t k1)
{
    uint32_t b0 = i0, b1 = i1, b2 = blk, b3 = 0;
    philox4x32_10(b0, b1, b2, b3, k0, k1);
    return make_uint4(b0, b1, b2, b3);
}

// RK (inline variant only): XOR the precomputed round keys of the kernel's parameter bank instead of
// bumping the key per round -- pays in a dense kernel with one block per element (normal: -2.4 %),
// costs in the persistent kernels, where the pointer has to be kept live (devroye: +2.5 %).
template <bool OUTLINE, bool RK = false>
struct RngT {
    uint32_t k0, k1, i0, i1, blk;
    uint32_t b0, b1, b2, b3;
    int left;
    const uint32_t* rk;  // round keys in the kernel's parameter bank (inline variant)

    __device__ __forceinline__ void init(const PhiloxKey& key, uint64_t index)
    {
        k0 = key.k0;
        k1 = key.k1;
        rk = RK ? key.rk : nullptr;
        i0 = (uint32_t)index;
        i1 = (uint32_t)(index >> 32);
        blk = 0;
        left = 0;
    }

    // The next TWO blocks of the stream at once (the stream must stand at a block boundary: a sampler that
    // uses this draws all its words this way).  Every lane of a warp generates its blocks together, and the
    // two sets of rounds interleave.
    __device__ __forceinline__ void two_blocks(uint32_t (&w)[8])
    {
        w[0] = i0, w[1] = i1, w[2] = blk, w[3] = 0;
        w[4] = i0, w[5] = i1, w[6] = blk + 1, w[7] = 0;
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int rnd = 0; rnd < 10; ++rnd) {
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
                const uint64_t p0 = (uint64_t)0xD2511F53u * w[j], p1 = (uint64_t)0xCD9E8D57u * w[j + 2];
                const uint32_t n0 = (uint32_t)(p1 >> 32) ^ w[j + 1] ^ ka, n2 = (uint32_t)(p0 >> 32) ^ w[j + 3] ^ kb;
                w[j] = n0;
                w[j + 1] = (uint32_t)p1;
                w[j + 2] = n2;
                w[j + 3] = (uint32_t)p0;
            }
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        blk += 2;
    }
    static __device__ __forceinline__ double to_co(uint32_t lo, uint32_t hi)  // 53 bits, [0, 1)
    {
        return (double)((((uint64_t)hi << 32) | lo) >> 11) * 1.1102230246251565e-16;
    }
    static __device__ __forceinline__ double to_oc(uint32_t lo, uint32_t hi)  // 53 bits, (0, 1]
    {
        return ((double)((((uint64_t)hi << 32) | lo) >> 11) + 1.0) * 1.1102230246251565e-16;
    }
    static __device__ __forceinline__ double to_u32(uint32_t w)  // 32 bits, (0, 1)
    {
        return ((double)w + 0.5) * 2.3283064365386963e-10;
    }

    // Continue the element's stream at word `wpos` (the saddle queue of the tile kernel re-enters a stream
    // where a rejected proposal left it) / the position reached.
    __device__ __forceinline__ void seek(uint32_t wpos)
    {
        blk = wpos >> 2;
        left = 0;
        const uint32_t skip = wpos & 3u;
        if (skip != 0) {
            refill_body();
            for (uint32_t i = 0; i < skip; ++i) {
                b0 = b1;
                b1 = b2;
                b2 = b3;
            }
            left = 4 - (int)skip;
        }
    }
    __device__ __forceinline__ uint32_t word_position() const { return 4u * blk - (uint32_t)left; }

    // One shared copy of the block function per kernel: u32() has many call sites in the
    // samplers and the persistent kernels live or die by their instruction-cache footprint.
    __device__ __forceinline__ void refill_body()
    {
        if (OUTLINE) {
            const uint4 w = philox_block_outlined(i0, i1, blk, k0, k1);
            b0 = w.x;
            b1 = w.y;
            b2 = w.z;
            b3 = w.w;
        }
        else {
            b0 = i0;
            b1 = i1;
            b2 = blk;
            b3 = 0;
            if (RK) philox4x32_10_rk(b0, b1, b2, b3, rk);
            else philox4x32_10(b0, b1, b2, b3, k0, k1);
        }
        ++blk;
        left = 4;
    }

    __device__ __forceinline__ uint32_t u32()
    {
        if (left == 0) refill_body();
        uint32_t w = b0;
        b0 = b1;
        b1 = b2;
        b2 = b3;
        --left;
        return w;
    }

    // next_float (src/pgm_macros.h:54-55): 32-bit uniform in (0, 1)
    __device__ __forceinline__ double uniform32() { return ((double)u32() + 0.5) * 2.3283064365386963e-10; }

    __device__ __forceinline__ uint64_t u53()
    {
        uint64_t lo = u32();
        uint64_t hi = u32();
        return ((hi << 32) | lo) >> 11;
    }
    // next_double (src/pgm_macros.h:61-62): [0, 1)
    __device__ __forceinline__ double uniform_co() { return (double)u53() * 1.1102230246251565e-16; }
    // (0, 1]
    __device__ __forceinline__ double uniform_oc() { return ((double)u53() + 1.0) * 1.1102230246251565e-16; }
    // replaces numpy random_standard_exponential
    __device__ __forceinline__ double exponential() { return -fm::log_pn(uniform_oc()); }
    // replaces numpy random_standard_normal: one Box-Muller draw (4 words)
    __device__ __forceinline__ double normal()
    {
        double u1 = uniform_oc();
        double u2 = uniform_co();
        return fm::sqrt_nn(-2.0 * fm::log_pn(u1)) * fm::cos2pi(u2);
    }
    // Box-Muller pair (used by the gamma-convolution sampler)
    __device__ __forceinline__ void normal_pair(double& n0, double& n1)
    {
        double u1 = uniform_oc();
        double u2 = uniform_co();
        double rad = fm::sqrt_nn(-2.0 * fm::log_pn(u1));
        double s, c;
        fm::sincos2pi(u2, s, c);
        n0 = rad * c;
        n1 = rad * s;
    }
};

// ---------------------------------------------------------------------------------------
// Special functions
// ---------------------------------------------------------------------------------------

// pgm_lgamma (src/pgm_common.h:40-117): lgamma(k) for integer k in [1, 200] from a table
// (generated by tools/gen_logfact_table.py), CUDA lgamma otherwise.
static __constant__ double c_logfact[201] = {
#include "pgm_logfact_table.inc"
};

// log Gamma(x), x > 0, by Stirling's series at y = x + 7 (x < 7) or y = x with the shift product
// divided out: two logarithms and ~20 FMAs on ONE code path.  libdevice's lgamma switches between
// range-specific kernels (three of them inside (1, 4)), which the lanes of a warp with mixed h run one
// after the other.  Absolute error < 2e-14 (tests/test_gpu_parity.py::test_fast_math).
__device__ __forceinline__ double
lgamma_stirling(double x)
{
    const bool shift = x < 7.0;
    const double y = shift ? x + 7.0 : x;
    const double iy = fm::rcp(y), w = iy * iy;
    // B_2n / (2n (2n - 1)), n = 1..8; the first omitted term is 43867/244188 y^-17 < 5e-16 at y = 7
    double s = -3617.0 / 122400.0;
    s = fma(s, w, 1.0 / 156.0);
    s = fma(s, w, -691.0 / 360360.0);
    s = fma(s, w, 1.0 / 1188.0);
    s = fma(s, w, -1.0 / 1680.0);
    s = fma(s, w, 1.0 / 1260.0);
    s = fma(s, w, -1.0 / 360.0);
    s = fma(s, w, 1.0 / 12.0);
    double lg = fma(y - 0.5, fm::log(y), -y) + (kLs2Pi + s * iy);
    if (shift) {
        const double prod = x * (x + 1.0) * (x + 2.0) * (x + 3.0) * (x + 4.0) * (x + 5.0) * (x + 6.0);
        lg -= fm::log(prod);
    }
    return lg;
}

__device__ __forceinline__ double
pgm_lgamma(double z)
{
    if (z < 201.0 && z >= 1.0 && z == floor(z)) {
        return c_logfact[(int)z];
    }
    if (z > 0.0 && z < 1e300) return lgamma_stirling(z);
    return lgamma(z);
}

// log of the standard normal cdf, both tails (replaces log_norm_cdf, src/pgm_saddle.c:268)
__device__ __forceinline__ double
log_ndtr(double x)
{
    double y = x * 0.70710678118654752440;
    if (y < -1.0) {
        return log(0.5 * erfcx(-y)) - y * y;
    }
    if (y < 1.0) {
        return log(0.5 * erfc(-y));
    }
    return log1p(-0.5 * erfc(y));
}

constexpr double kLentzTiny = 1e-290;
constexpr double kLentzEps = 4.0e-16;
constexpr int kLentzMaxIt = 1000;

// confluent_x_smaller (src/pgm_common.h:181-211), FP64 modified Lentz
__device__ inline double
confluent_x_smaller(double p, double x)
{
    double a = 1.0, b = p;
    double r = -(p - 1.0) * x;
    double s = 0.5 * x;
    double f = a / b, c = a / kLentzTiny, d = 1.0 / b;
    for (int n = 2; n < kLentzMaxIt; ++n) {
        a = (n & 1) ? s * (n - 1) : r - s * n;
        b += 1.0;
        c = b + a * fm::rcp(c);  // |c|, |d| >= 1e-290: normal numbers, the MUFU-seeded reciprocal applies
        if (fabs(c) < kLentzTiny) c = kLentzTiny;
        d = a * d + b;
        if (fabs(d) < kLentzTiny) d = kLentzTiny;
        d = fm::rcp(d);
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < kLentzEps) break;
    }
    return f;
}

// confluent_p_smaller (src/pgm_common.h:237-265), FP64 modified Lentz
__device__ inline double
confluent_p_smaller(double p, double x)
{
    double a = 1.0, b = x - p + 1.0;
    double f = a / b, c = a / kLentzTiny, d = 1.0 / b;
    for (int n = 1; n < kLentzMaxIt; ++n) {
        a = n * (p - n);
        b += 2.0;
        c = b + a * fm::rcp(c);
        if (fabs(c) < kLentzTiny) c = kLentzTiny;
        d = a * d + b;
        if (fabs(d) < kLentzTiny) d = kLentzTiny;
        d = fm::rcp(d);
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < kLentzEps) break;
    }
    return f;
}

// log Gamma(p, x) / log Q(p, x): upper_incomplete_gamma (src/pgm_common.h:291-340) in log space
__device__ inline double
log_upper_gamma(double p, double x, bool normalized, double lg)  // lg = pgm_lgamma(p)
{
    if (x > p) {
        double f = confluent_p_smaller(p, x);
        double v = log(f) - x + p * log(x);
        return normalized ? v - lg : v;
    }
    double f = confluent_x_smaller(p, x);
    double lower = f * exp(-x + p * log(x) - lg);
    double lq = log1p(-lower);
    return normalized ? lq : lq + lg;
}

__device__ inline double
log_upper_gamma(double p, double x, bool normalized)
{
    return log_upper_gamma(p, x, normalized, pgm_lgamma(p));
}

using RngOutlined = RngT<true>;
using RngInline = RngT<false>;
using RngInlineRK = RngT<false, true>;

// random_left_bounded_gamma (src/pgm_common.h:126-157): X ~ Gamma(a, rate b) | X > t
template <class R>
__device__ inline double
left_bounded_gamma(R& r, double a, double b, double t)
{
    double x;
    if (a > 1.0) {  // Dagpunar (1978)
        b = t * b;
        double amin1 = a - 1.0;
        double bmina = b - a;
        // (a, b, t are positive normal numbers: divisions by MUFU-seeded reciprocals, hoisted out of
        // the rejection loops)
        const double inv_b = fm::rcp(b);
        double c0 = 0.5 * (bmina + sqrt(bmina * bmina + 4.0 * b)) * inv_b;
        double one_minus_c0 = 1.0 - c0;
        double log_m = amin1 * (log(fm::div(amin1, one_minus_c0)) - 1.0);
        const double inv_c0 = fm::rcp(c0);
        double threshold;
        do {
            x = b + r.exponential() * inv_c0;
            threshold = amin1 * log(x) - x * one_minus_c0 - log_m;
        } while (log(r.uniform32()) > threshold);
        return t * (x * inv_b);
    }
    else if (a == 1.0) {
        return t + r.exponential() * fm::rcp(b);
    }
    else if (a == 0.5) {  // Philippe (1997) A4; u < x^(-1/2) written as u^2 x < 1
        const double inv_tb = fm::rcp(t * b);
        double u;
        do {
            x = 1.0 + r.exponential() * inv_tb;
            u = r.uniform32();
        } while (u * u * x > 1.0);
        return t * x;
    }
    else {
        double amin1 = a - 1.0;
        const double inv_tb = fm::rcp(t * b);
        do {
            x = 1.0 + r.exponential() * inv_tb;
        } while (log(r.uniform32()) > amin1 * log(x));
        return t * x;
    }
}

#endif  // __CUDACC__

}  // namespace pgm
