/* pgm_cuda_shim.h -- host-interop mode of the C-ABI: libpgm_cuda_shim.so exports the
 * REFERENCE's own symbols, with the reference's own signatures, so that code compiled against
 * zoj613/polyagamma's headers (its Cython module, `cimport polyagamma` users -- reference
 * polyagamma/__init__.pxd:1-15, README.md:117-144 -- or plain C like examples/c_polyagamma.c:104)
 * runs on the GPU by relinking: replace the objects of src/pgm_*.c with -lpgm_cuda_shim.
 *
 *   reference declaration                                   replaced here by
 *   include/pgm_random.h:47-74   pgm_random_polyagamma{,_fill,_fill2}(bitgen_t*, ...)   same name
 *   include/pgm_density.h:4-14   pgm_polyagamma_{pdf,cdf,logpdf,logcdf}(x, h, z)         same name
 *
 * All pointers are HOST pointers, as in the reference.  Randomness: every call takes two
 * next_uint64() values from the caller's numpy BitGenerator, (seed, offset), and draws from the
 * Philox streams they name (include/pgm_cuda.h); so the caller's generator still decides the
 * result, a seeded generator reproduces it, and the output equals
 * pgm_cuda_random_polyagamma_fill2_host(seed, offset, ...) -- but it is not the reference's
 * PCG64 + ziggurat stream (sample-level parity with numpy's generators is unpinned).
 * The reference functions cannot fail; if the GPU call fails the outputs are filled with NaN,
 * a message goes to stderr, and pgm_cuda_shim_last_error() returns the code.
 * The device is the calling thread's current CUDA device (device 0 by default).
 *
 * A translation unit that already included numpy's <numpy/random/bitgen.h> (as the reference's
 * pgm_random.h does) keeps numpy's definition of bitgen_t; the layout below is the same. */
#ifndef PGM_CUDA_SHIM_H
#define PGM_CUDA_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef NUMPY_CORE_INCLUDE_NUMPY_RANDOM_BITGEN_H_
typedef struct bitgen {
    void* state;
    uint64_t (*next_uint64)(void* st);
    uint32_t (*next_uint32)(void* st);
    double (*next_double)(void* st);
    uint64_t (*next_raw)(void* st);
} bitgen_t;
#endif

#ifndef PGM_RANDOM_H
typedef enum { GAMMA, DEVROYE, ALTERNATE, SADDLE, HYBRID } sampler_t; /* include/pgm_random.h:13 */
#endif

double pgm_random_polyagamma(bitgen_t* bitgen_state, double h, double z, sampler_t method);
void pgm_random_polyagamma_fill(bitgen_t* bitgen_state, double h, double z, sampler_t method, size_t n,
                                double* out);
void pgm_random_polyagamma_fill2(bitgen_t* bitgen_state, const double* h, const double* z, sampler_t method,
                                 size_t n, double* out);

double pgm_polyagamma_pdf(double x, double h, double z);
double pgm_polyagamma_cdf(double x, double h, double z);
double pgm_polyagamma_logpdf(double x, double h, double z);
double pgm_polyagamma_logcdf(double x, double h, double z);

/* 0, or the code of the last failed GPU call made through this shim by the calling thread. */
int pgm_cuda_shim_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* PGM_CUDA_SHIM_H */
