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
struct PhiloxKey {
    uint32_t k0, k1;
};

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

// Call key = first two output words of Philox(key = seed, ctr = (offset, 'PGM1', 'key!')).
__host__ __device__ inline PhiloxKey
derive_call_key(uint64_t seed, uint64_t offset)
{
    uint32_t c0 = (uint32_t)offset, c1 = (uint32_t)(offset >> 32), c2 = 0x50474D31u, c3 = 0x6B657921u;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    return PhiloxKey{c0, c1};
}

#ifdef __CUDACC__

// One private stream per output element: words of Philox(call key, ctr = (idx, block, 0)),
// block = 0, 1, 2, ..., consumed strictly in order.  The four buffered words live in
// registers and are shifted down on every draw (no local-memory indexing).
// OUTLINE: keep ONE copy of the block function per kernel (persistent samplers: many call sites,
// instruction-cache bound) or inline it (dense kernels with a single draw site).
template <bool OUTLINE>
struct RngT {
    uint32_t k0, k1, i0, i1, blk;
    uint32_t b0, b1, b2, b3;
    int left;

    __device__ __forceinline__ void init(PhiloxKey key, uint64_t index)
    {
        k0 = key.k0;
        k1 = key.k1;
        i0 = (uint32_t)index;
        i1 = (uint32_t)(index >> 32);
        blk = 0;
        left = 0;
    }

    // One shared copy of the block function per kernel: u32() has many call sites in the
    // samplers and the persistent kernels live or die by their instruction-cache footprint.
    __device__ __forceinline__ void refill_body()
    {
        b0 = i0;
        b1 = i1;
        b2 = blk;
        b3 = 0;
        philox4x32_10(b0, b1, b2, b3, k0, k1);
        ++blk;
        left = 4;
    }
    __device__ __noinline__ void refill_outlined() { refill_body(); }

    __device__ __forceinline__ uint32_t u32()
    {
        if (left == 0) {
            if (OUTLINE) refill_outlined();
            else refill_body();
        }
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
    __device__ __forceinline__ double exponential() { return -log(uniform_oc()); }
    // replaces numpy random_standard_normal: one Box-Muller draw (4 words)
    __device__ __forceinline__ double normal()
    {
        double u1 = uniform_oc();
        double u2 = uniform_co();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    // Box-Muller pair (used by the gamma-convolution sampler)
    __device__ __forceinline__ void normal_pair(double& n0, double& n1)
    {
        double u1 = uniform_oc();
        double u2 = uniform_co();
        double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
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

__device__ __forceinline__ double
pgm_lgamma(double z)
{
    if (z < 201.0 && z >= 1.0 && z == floor(z)) {
        return c_logfact[(int)z];
    }
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
        c = b + a / c;
        if (fabs(c) < kLentzTiny) c = kLentzTiny;
        d = a * d + b;
        if (fabs(d) < kLentzTiny) d = kLentzTiny;
        d = 1.0 / d;
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
        c = b + a / c;
        if (fabs(c) < kLentzTiny) c = kLentzTiny;
        d = a * d + b;
        if (fabs(d) < kLentzTiny) d = kLentzTiny;
        d = 1.0 / d;
        double delta = c * d;
        f *= delta;
        if (fabs(delta - 1.0) < kLentzEps) break;
    }
    return f;
}

// log Gamma(p, x) / log Q(p, x): upper_incomplete_gamma (src/pgm_common.h:291-340) in log space
__device__ inline double
log_upper_gamma(double p, double x, bool normalized)
{
    double lg = pgm_lgamma(p);
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

using RngOutlined = RngT<true>;
using RngInline = RngT<false>;

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
        double c0 = 0.5 * (bmina + sqrt(bmina * bmina + 4.0 * b)) / b;
        double one_minus_c0 = 1.0 - c0;
        double log_m = amin1 * (log(amin1 / one_minus_c0) - 1.0);
        double threshold;
        do {
            x = b + r.exponential() / c0;
            threshold = amin1 * log(x) - x * one_minus_c0 - log_m;
        } while (log(r.uniform32()) > threshold);
        return t * (x / b);
    }
    else if (a == 1.0) {
        return t + r.exponential() / b;
    }
    else if (a == 0.5) {  // Philippe (1997) A4; u < x^(-1/2) written as u^2 x < 1
        double tb = t * b;
        double u;
        do {
            x = 1.0 + r.exponential() / tb;
            u = r.uniform32();
        } while (u * u * x > 1.0);
        return t * x;
    }
    else {
        double amin1 = a - 1.0;
        double tb = t * b;
        do {
            x = 1.0 + r.exponential() / tb;
        } while (log(r.uniform32()) > amin1 * log(x));
        return t * x;
    }
}

#endif  // __CUDACC__

}  // namespace pgm
