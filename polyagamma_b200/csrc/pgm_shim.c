/* pgm_shim.c -- libpgm_cuda_shim.so: the reference's C symbols on top of libpgm_cuda.so
 * (include/pgm_cuda_shim.h).  Plain C, no CUDA calls of its own. */
#include "../../include/pgm_cuda_shim.h"

#include <math.h>
#include <stdio.h>

#include "../../include/pgm_cuda.h"

static _Thread_local int g_last_error = 0;

static void
failed(const char* what, int rc, double* out, size_t n)
{
    g_last_error = rc;
    fprintf(stderr, "libpgm_cuda_shim: %s failed with code %d; output set to NaN\n", what, rc);
    for (size_t i = 0; i < n; ++i) out[i] = NAN;
}

static void
draw_into(bitgen_t* bg, const double* h, int h_stride, const double* z, int z_stride, sampler_t method, size_t n,
          double* out)
{
    /* always consume the two words, n == 0 included: the generator advances by the same amount
     * whatever the request size */
    const uint64_t seed = bg->next_uint64(bg->state);
    const uint64_t offset = bg->next_uint64(bg->state);
    if (n == 0) return;
    const int rc =
        pgm_cuda_random_polyagamma_fill2_host(seed, offset, h, h_stride, z, z_stride, (int)method, n, out, 0, -1);
    if (rc != 0) failed("pgm_cuda_random_polyagamma_fill2_host", rc, out, n);
    else g_last_error = 0;
}

double
pgm_random_polyagamma(bitgen_t* bitgen_state, double h, double z, sampler_t method)
{
    double out;
    draw_into(bitgen_state, &h, 0, &z, 0, method, 1, &out);
    return out;
}

void
pgm_random_polyagamma_fill(bitgen_t* bitgen_state, double h, double z, sampler_t method, size_t n, double* out)
{
    draw_into(bitgen_state, &h, 0, &z, 0, method, n, out);
}

void
pgm_random_polyagamma_fill2(bitgen_t* bitgen_state, const double* h, const double* z, sampler_t method, size_t n,
                            double* out)
{
    draw_into(bitgen_state, h, 1, z, 1, method, n, out);
}

static double
density(int which, double x, double h, double z)
{
    double out;
    const int rc = pgm_cuda_polyagamma_density_host(which, &x, 0, &h, 0, &z, 0, 1, &out, -1);
    if (rc != 0) failed("pgm_cuda_polyagamma_density_host", rc, &out, 1);
    else g_last_error = 0;
    return out;
}

double
pgm_polyagamma_pdf(double x, double h, double z)
{
    return density(0, x, h, z);
}
double
pgm_polyagamma_cdf(double x, double h, double z)
{
    return density(1, x, h, z);
}
double
pgm_polyagamma_logpdf(double x, double h, double z)
{
    return density(2, x, h, z);
}
double
pgm_polyagamma_logcdf(double x, double h, double z)
{
    return density(3, x, h, z);
}

int
pgm_cuda_shim_last_error(void)
{
    return g_last_error;
}
