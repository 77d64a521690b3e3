/* oracle/pgm_oracle.h -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * CPU restatement (plain C, scalar, one element at a time) of the Polya-Gamma hot path of
 * zoj613/polyagamma, in the arithmetic and with the random stream the B200 kernels use:
 *
 *   - everything in FP64 (the reference mixes FP32/FP64);
 *   - randomness = Philox4x32-10, one private stream per output element, keyed by
 *     (seed, offset) and counted by (global element index, block);
 *   - sampling variates: -log(U) exponentials, Box-Muller normals, Marsaglia-Tsang gammas
 *     (the reference uses numpy's ziggurat / legacy gamma through bitgen_t).
 *
 * Each function cites the reference file:line it restates.  Because the stream differs from
 * numpy's PCG64+ziggurat, sampler parity with the reference is STATISTICAL (tests/ pin this
 * oracle against the compiled reference in oracle/_ref and against exact PG moments); density
 * parity is numerical (<= 1e-13 against oracle/_ref).  The CUDA kernels are then checked
 * element-by-element against this oracle on the same (seed, offset, index).
 */
#ifndef PGM_ORACLE_H
#define PGM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* include/pgm_random.h:13 */
enum { PGO_GAMMA = 0, PGO_DEVROYE = 1, PGO_ALTERNATE = 2, PGO_SADDLE = 3, PGO_HYBRID = 4 };
/* OR-ed into `method`: reproduce the reference's envelope constants (SURVEY 8-DEFECTS a, i). */
#define PGO_REF_COMPAT 0x100

/* internal method ids reported by pgo_hybrid_select (5 = normal approximation) */
enum { PGO_NORMAL = 5 };

/* ---- Philox ---- */
void pgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void pgo_call_key(uint64_t seed, uint64_t offset, uint32_t key[2]);
/* first `n` 32-bit words of element `index`'s stream */
void pgo_stream_words(uint64_t seed, uint64_t offset, uint64_t index, size_t n, uint32_t* out);
/* n variates drawn consecutively from element `index`'s stream.
 * kind: 0 uniform32 (0,1), 1 uniform53 [0,1), 2 exponential, 3 normal, 4 gamma(shape=param) */
void pgo_stream_variates(uint64_t seed, uint64_t offset, uint64_t index, int kind, double param,
                         size_t n, double* out);

/* ---- samplers (src/pgm_random.c:144-180) ---- */
int pgo_hybrid_select(double h, double z);
void pgo_random_polyagamma_fill(uint64_t seed, uint64_t offset, double h, double z, int method,
                                size_t n, double* out, uint64_t first_index);
void pgo_random_polyagamma_fill2(uint64_t seed, uint64_t offset, const double* h, const double* z,
                                 int method, size_t n, double* out, uint64_t first_index);
/* words of randomness consumed by the last element drawn on this thread (instrumentation) */
uint64_t pgo_last_words_consumed(void);

/* ---- helper functions exposed for unit tests ---- */
double pgo_lgamma(double z);
double pgo_erfcx(double x);
double pgo_log_upper_gamma(double p, double x, int normalized);
double pgo_alternate_trunc_point(double h);
double pgo_saddle_newton(double x, double* fprime);
/* out[0..] = devroye {w, K}; alternate {w_right, t, lambda}; saddle {w_left, xc, slope_l,
 * icpt_l, slope_r, icpt_r, log_coef_l, log_coef_r} */
void pgo_setup_probe(int method, double h, double z, double* out);

/* ---- densities (src/pgm_density.c) ---- */
double pgo_polyagamma_pdf(double x, double h, double z);
double pgo_polyagamma_cdf(double x, double h, double z);
double pgo_polyagamma_logpdf(double x, double h, double z);
double pgo_polyagamma_logcdf(double x, double h, double z);
/* which: 0 pdf, 1 cdf, 2 logpdf, 3 logcdf; cond (optional) receives sum|term|/|sum term| */
void pgo_density_array(int which, const double* x, const double* h, const double* z, size_t n,
                       double* out, double* cond);

#ifdef __cplusplus
}
#endif
#endif
