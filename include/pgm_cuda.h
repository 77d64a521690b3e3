/* include/pgm_cuda.h -- C-ABI of the B200-native Polya-Gamma hot path.
 *
 * Drop-in boundary for zoj613/polyagamma's C API:
 *     include/pgm_random.h:47-74   pgm_random_polyagamma / _fill / _fill2
 *     include/pgm_density.h:4-14   pgm_polyagamma_{pdf,cdf,logpdf,logcdf}
 *
 * What changes with respect to the reference signatures, and nothing else:
 *   - `bitgen_t*` is replaced by a counter-based Philox4x32-10 stream named by
 *     `(uint64_t seed, uint64_t offset)`.  Element i of a call draws only from the private
 *     stream (seed, offset, first_index + i), so results do not depend on launch shape,
 *     chunking or on how an index range is sharded over GPUs.
 *   - data pointers are DEVICE pointers (FP64), unless the function name ends in `_host`;
 *   - a trailing `void* stream` (a `cudaStream_t`; NULL = default stream).  Device-pointer
 *     entry points are asynchronous on that stream and never synchronise the host;
 *   - functions return `int`: 0 on success, a positive `cudaError_t` for CUDA failures,
 *     PGM_CUDA_EINVAL for bad arguments.  As in the reference, an invalid shape parameter is
 *     NOT an error: it is encoded in the data (h <= 0 or NaN -> NaN, h = +inf -> +inf;
 *     src/pgm_random.c:137-155,162-168).
 *
 * No torch types, no C++ types.  Every in/out buffer is caller-owned.  What the library allocates:
 *   - per call, stream-ordered scratch from a PRIVATE memory pool of the device (the device's default
 *     pool is left alone): the element lists and records of the divergent method buckets, 232 B per
 *     element of a pass (<= 2^25 elements: <= 7.8 GB; the pass is halved while the allocation fails).
 *     Freed blocks stay cached in that pool; pgm_cuda_release_scratch returns them to the driver;
 *   - once per device, 80 KB of tables (saddle-point functions) filled by a kernel on a private stream;
 *     the first call on a device waits for it -- the only host synchronisation on a device-pointer path
 *     (call pgm_cuda_init before a stream capture to have it out of the way);
 *   - `_host` entry points only: pinned staging rings, sized by the largest call so far (see there).
 *
 * Threads: every entry point may be called from several host threads at once, each on its own
 * stream (the reference serialises callers with the BitGenerator's lock,
 * polyagamma/_polyagamma.pyx:172-176; here a call owns no generator state -- its draws are named by
 * (seed, offset, index)).  Concurrent calls produce exactly what the same calls produce one after
 * the other (tests/test_gpu_threads.py).
 */
#ifndef PGM_CUDA_H
#define PGM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same values as the reference's `sampler_t` (include/pgm_random.h:13). */
typedef enum {
    PGM_CUDA_GAMMA = 0,
    PGM_CUDA_DEVROYE = 1,
    PGM_CUDA_ALTERNATE = 2,
    PGM_CUDA_SADDLE = 3,
    PGM_CUDA_HYBRID = 4
} pgm_cuda_sampler_t;

/* OR into `method` to reproduce the reference's envelope constants bit-for-constant
 * (src/pgm_alternate.c:90 uses log(sqrt(pi/2)); src/pgm_saddle.c:257,325-326 use +log(xc)).
 * Default (flag clear) uses the mathematically consistent envelopes. */
#define PGM_CUDA_REF_COMPAT 0x100

#define PGM_CUDA_EINVAL (-1)
#define PGM_CUDA_MAX_NDIM 8

/* ---- sampling: replaces include/pgm_random.h ------------------------------------------- */

/* pgm_random_polyagamma (include/pgm_random.h:47-48): one draw, returned on the host.  One kernel launch
 * that writes into mapped pinned memory of a slot the calling thread keeps per device (no allocation, no
 * copy call after the first draw), then a stream synchronisation.  Returns NaN if a CUDA call fails. */
double pgm_cuda_random_polyagamma(uint64_t seed, uint64_t offset, double h, double z, int method);

/* pgm_random_polyagamma_fill (include/pgm_random.h:61-63): n iid draws of PG(h, z). */
int pgm_cuda_random_polyagamma_fill(uint64_t seed, uint64_t offset, double h, double z,
                                    int method, size_t n, double* d_out, uint64_t first_index,
                                    void* stream);

/* pgm_random_polyagamma_fill2 (include/pgm_random.h:72-74): out[i] ~ PG(h[i], z[i]).
 * d_h, d_z, d_out hold n contiguous doubles. */
int pgm_cuda_random_polyagamma_fill2(uint64_t seed, uint64_t offset, const double* d_h,
                                     const double* d_z, int method, size_t n, double* d_out,
                                     uint64_t first_index, void* stream);

/* fill2 over a broadcast index space (what the reference's NpyIter loop feeds to fill2,
 * polyagamma/_polyagamma.pyx:213-251), without materialising the broadcast operands:
 * out is C-contiguous with `shape[0..ndim)`; element strides (in doubles, 0 = broadcast) are
 * given for h and z.  ndim <= PGM_CUDA_MAX_NDIM.  `shape`/strides are HOST arrays. */
int pgm_cuda_random_polyagamma_fill2_strided(uint64_t seed, uint64_t offset, const double* d_h,
                                             const double* d_z, int method, int ndim,
                                             const int64_t* shape, const int64_t* h_strides,
                                             const int64_t* z_strides, double* d_out,
                                             uint64_t first_index, void* stream);

/* Host-buffer form of fill2 (the reference's calling convention: all pointers are host memory,
 * include/pgm_random.h:72-74).  h_stride / z_stride are 1 (array of n) or 0 (scalar broadcast).  Copies are
 * chunked and overlapped with the kernels on a ring of three internal streams of `device` (< 0: the current
 * one); pinned (page-locked) buffers give the full speed of the host link, pageable ones (an ordinary malloc or
 * numpy array) of 2^16 elements or more are bounced through pinned buffers of the ring (3 x 3 x 16 MiB, kept until
 * pgm_cuda_release_scratch) by a small pool of host threads.  Blocks until `out` is complete; on failure it
 * still does not return before every copy it queued has finished.  The ring's device staging (3 streams x 4
 * operands x the chunk size, at most 64 MiB each, sized by the largest call so far) stays allocated until
 * pgm_cuda_release_scratch.  Calls on different devices run concurrently; calls on one device take turns. */
int pgm_cuda_random_polyagamma_fill2_host(uint64_t seed, uint64_t offset, const double* h,
                                          int h_stride, const double* z, int z_stride, int method,
                                          size_t n, double* out, uint64_t first_index, int device);

/* The same over SEVERAL devices from one host thread: the index range is cut into `ndevices` contiguous
 * shards (shard i = [n i / ndevices, n (i + 1) / ndevices), on devices[i]), every shard runs through its own
 * device's ring, and the chunks are queued round-robin so that all devices copy and compute at once.
 * Elements keep their global Philox indices: the result is bit-identical for every ndevices. */
int pgm_cuda_random_polyagamma_fill2_host_multi(uint64_t seed, uint64_t offset, const double* h, int h_stride,
                                                const double* z, int z_stride, int method, size_t n,
                                                double* out, uint64_t first_index, const int* devices,
                                                int ndevices);

/* Device time (CUDA events, first copy queued -> last copy done; the slowest device of a multi-device
 * call) of the calling THREAD's most recent host-buffer fill, in milliseconds. */
double pgm_cuda_last_host_fill_ms(void);

/* Parameter check of polyagamma/_polyagamma.pyx:411-430 (any h <= 1e-4, or non-integer h when
 * method is DEVROYE) over n contiguous device doubles.  Synchronises; *result = 1 if any
 * element is invalid. */
int pgm_cuda_any_invalid_h(const double* d_h, size_t n, int method, int* result, void* stream);

/* ---- densities: replaces include/pgm_density.h ----------------------------------------- */
/* Elementwise over n contiguous doubles (the reference's functions are scalar and the Cython
 * `dispatch` loop calls them once per element, polyagamma/_polyagamma.pyx:433-472). */
int pgm_cuda_polyagamma_pdf(const double* d_x, const double* d_h, const double* d_z, size_t n,
                            double* d_out, void* stream);
int pgm_cuda_polyagamma_cdf(const double* d_x, const double* d_h, const double* d_z, size_t n,
                            double* d_out, void* stream);
int pgm_cuda_polyagamma_logpdf(const double* d_x, const double* d_h, const double* d_z, size_t n,
                               double* d_out, void* stream);
int pgm_cuda_polyagamma_logcdf(const double* d_x, const double* d_h, const double* d_z, size_t n,
                               double* d_out, void* stream);

/* Broadcast form: which = 0 pdf, 1 cdf, 2 logpdf, 3 logcdf; out is C-contiguous with
 * `shape`; strides in doubles (0 = broadcast); shape/strides are HOST arrays.
 * which = 4 .. 7 (not in the reference): the same four functions computed WITHOUT the
 * alternating series' cancellation for h >= 8 -- by the trapezoid rule along the vertical line through the
 * saddle point of the Laplace inversion integral (csrc/pgm_density.cu) -- where the reference's series loses
 * sum|term| / |sum term| digits: all of them for h >~ 40 at small |z| (src/pgm_density.c:80-86,268-272).
 * Agrees with the series in 300-digit arithmetic to 1e-12 (tests/golden/make_stable_golden.py). */
int pgm_cuda_polyagamma_density_strided(int which, const double* d_x, const double* d_h,
                                        const double* d_z, int ndim, const int64_t* shape,
                                        const int64_t* x_strides, const int64_t* h_strides,
                                        const int64_t* z_strides, double* d_out, void* stream);

/* Host-buffer form of the four density functions (which: 0 pdf, 1 cdf, 2 logpdf, 3 logcdf): x, h, z, out
 * are HOST arrays; a stride of 0 broadcasts a single value, 1 walks a contiguous array of n.  Chunked
 * H2D -> kernel -> D2H on a private stream; returns after `out` is complete.  device < 0: current. */
int pgm_cuda_polyagamma_density_host(int which, const double* x, int x_stride, const double* h, int h_stride,
                                     const double* z, int z_stride, size_t n, double* out, int device);

/* A device result on its way to PAGEABLE host memory -- what the reference's front-end returns is a freshly
 * allocated numpy array (polyagamma/_polyagamma.pyx:452-472).  d_src[0..n) -> dst[0..n): chunks go through
 * pinned bounce buffers on the device's ring of streams and are moved into `dst` by several host threads (the
 * first touch of a new array is page-fault bound: one thread manages ~2 GB/s).  Waits for the work queued on
 * `stream` before it reads d_src; returns when dst is complete. */
int pgm_cuda_download(double* dst, const double* d_src, size_t n, void* stream);

/* ---- next row of the path: the logistic-regression Gibbs omega-step ---------- */
/* The caller pattern of the batched fill (README.md:146-148; Polson, Scott & Windle 2013):
 *     psi = X beta;   omega_i ~ PG(n_i, psi_i);   G = X^T diag(omega) X
 * X: n x p row-major FP64 with leading dimension ldx (in doubles), 1 <= p <= 64; beta: p;
 * d_ntrials: n shape parameters, or NULL to use `ntrials_scalar` for every row; d_gram: p x p
 * row-major, overwritten (deterministic reduction order); d_omega / d_psi: n doubles each, or NULL
 * if the caller does not need them.  omega is exactly what pgm_cuda_random_polyagamma_fill2 draws
 * for (n_trials, psi) with the same (seed, offset, first_index). */
int pgm_cuda_logistic_omega_step(uint64_t seed, uint64_t offset, const double* d_X, size_t n, int p, size_t ldx,
                                 const double* d_beta, const double* d_ntrials, double ntrials_scalar,
                                 int method, double* d_gram, double* d_omega, double* d_psi,
                                 uint64_t first_index, void* stream);

/* ---- set-up / tear-down ------------------------------------------------------------------ */
/* The one-time, per-device initialisation (tables, scratch pool, occupancy of the persistent kernels),
 * which the first call on a device would otherwise do.  Synchronises a private stream. */
int pgm_cuda_init(void);

/* ---- introspection ---------------------------------------------------------------------- */
/* Number of kernels this library has launched in this process (all threads). */
uint64_t pgm_cuda_launch_count(void);
/* Per-kernel device time, for roofline accounting.  Between _begin and _end every kernel this
 * library launches is bracketed by CUDA events on its own stream; _end synchronises those
 * events and returns, per kernel kind, the summed milliseconds and the launch count.
 * Kinds: 0 the drain kernel (every divergent bucket of a mid-size fill in one launch), 1 the one-launch
 * kernel of small fills, 2 the tile
 * kernel (classification, the in-tile normal / left-saddle draws, list compaction);
 * sampling kernels over a list or an identity range: 3 devroye (both proposal variants), 4 alternate,
 * 5 saddle (both envelope pieces), 6 saddle-left (left piece only), 7 unused, 8 alternate-left (left
 * piece only, |z| >= 14), 9 unused; 10 normal, 11 gamma, 12 density; per-element setup kernels 13..19 in
 * the order of 3..9 (8 computes its records inside the sampling kernel: no setup launch);
 * 20 psi = X beta, 21 weighted Gram matrix.
 * `cap` = length of both output arrays (>= 22). */
#define PGM_CUDA_NUM_KERNEL_KINDS 22
void pgm_cuda_profile_begin(void);
int pgm_cuda_profile_end(double* ms, uint64_t* launches, int cap);
/* FP64 (DFMA) peak of the current device in TFLOP/s, measured by a register-resident FMA
 * microbenchmark: the denominator of the FP64 roofline.  Synchronises `stream`. */
int pgm_cuda_fp64_peak_tflops(double* tflops, void* stream);
/* Returns the cached scratch of the current device (the private pool's free blocks and the `_host`
 * staging rings) to the driver.  Synchronises the device. */
int pgm_cuda_release_scratch(void);
/* Diagnostic: with PGM_CUDA_EVLOG=1 in the environment the fill pipeline records an event after every
 * operation it enqueues; this prints to stderr which of them the device has not completed (a watchdog can
 * call it from another thread when a call does not return; tools/stress_probe.py). */
void pgm_cuda_debug_report(void);
/* Unit-test hook: applies one of the kernels' own elementary functions (csrc/pgm_device.cuh, fm::)
 * elementwise.  which: 0 rcp, 1 sqrt, 2 log, 3 exp, 4 cos(2 pi x), 5 rsqrt, 9 log Gamma, 10 erfcx (x >= 0), 11 erfc (x >= 0), 6 sin(2 pi x) and
 * 7 cos(2 pi x) of the paired routine, 8 1/x by fm::div. */
int pgm_cuda_test_math(int which, const double* d_in, size_t n, double* d_out, void* stream);
/* "polyagamma_b200 <version> sm_100a" */
const char* pgm_cuda_version(void);
/* Method the hybrid sampler picks for (h, z): 0 gamma, 1 devroye, 2 alternate, 3 saddle,
 * 5 normal approximation (src/pgm_random.c:105-121).  Host-side, no GPU needed. */
int pgm_cuda_hybrid_select(double h, double z);
/* The Philox call key derived from (seed, offset); host-side, no GPU needed. */
void pgm_cuda_call_key(uint64_t seed, uint64_t offset, uint32_t key[2]);

#ifdef __cplusplus
}
#endif
#endif /* PGM_CUDA_H */
