// pgm_internal.h -- host-side glue shared by the translation units of libpgm_cuda.so.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "pgm_device.cuh"

namespace pgm {

constexpr int kMaxNdim = 8;

// Broadcast index space: ndim == 0 means "linear" (operands are scalar or contiguous).
struct Bcast {
    int ndim;
    int64_t shape[kMaxNdim];
    int64_t hs[kMaxNdim];  // element strides of the 2nd operand (h)
    int64_t zs[kMaxNdim];  // element strides of the 3rd operand (z)
    int64_t xs[kMaxNdim];  // element strides of the 1st operand (x; densities only)
};

struct FillArgs {
    PhiloxKey key;
    const double* h;   // nullptr: use h_scalar
    const double* z;   // nullptr: use z_scalar
    double h_scalar, z_scalar;
    int64_t h_stride, z_stride;  // linear mode: 1 (array) or 0 (one-element array)
    double* out;
    uint64_t first_index;  // Philox index of out[0]
    uint64_t base;         // first element of the current chunk
    int compat;
    Bcast bc;
};

struct DensityArgs {
    const double* x;
    const double* h;
    const double* z;
    double* out;
    Bcast bc;
};

int launch_fill(FillArgs a, uint64_t n, int method, cudaStream_t s);
int launch_check_h(const double* d_h, size_t n, int devroye, int* d_flag, cudaStream_t s);
int launch_density(int which, DensityArgs a, uint64_t n, cudaStream_t s);
int launch_logistic_omega_step(uint64_t seed, uint64_t offset, const double* d_X, size_t n, int p, size_t ldx,
                               const double* d_beta, const double* d_ntrials, double ntrials_scalar, int method,
                               double* d_gram, double* d_omega, double* d_psi, uint64_t first_index, cudaStream_t s);
int device_init();                                          // per-device one-time setup (tables, scratch pool)
int scratch_alloc(void** p, size_t bytes, cudaStream_t s);  // stream-ordered, from the library's private pool
int scratch_release();
void host_rings_release();  // pgm_host.cu: the staging buffers of the current device's host-buffer ring
void evlog_report();  // PGM_CUDA_EVLOG diagnostic: which enqueued operations has the device completed?
uint64_t launch_count();
void count_launch(uint64_t n);
int measure_fp64_peak(double* tflops, cudaStream_t s);
int launch_test_math(int which, const double* d_in, size_t n, double* d_out, cudaStream_t s);

// kernel kinds reported by pgm_cuda_profile_end
enum : int {
    kKindDrain = 0,     // every divergent bucket of a mid-size fill in one launch
    kKindSmall = 1,     // the one-launch kernel for small fills
    kKindTile = 2,      // tile kernel: classify + in-tile normal / left-only saddle + list compaction
    // sampling kernels: + 0 devroye, 1 alternate, 2 saddle (full), 3 saddle-left (identity range only), 4 unused,
    // 5 alternate-left, 6 unused
    kKindDevroye = 3,
    kKindNormal = 10,
    kKindGamma = 11,
    kKindDensity = 12,
    kKindSetup = 13,  // setup kernels, same order as the sampling kernels
    kKindGibbs = 20,  // + 0 psi = X beta, 1 weighted Gram matrix
    kNumKinds = 22
};

class ScopedKernelTimer {
public:
    ScopedKernelTimer(int kind, cudaStream_t s);
    ~ScopedKernelTimer();

private:
    int kind_;
    cudaStream_t stream_;
    bool on_;
    cudaEvent_t e0_, e1_;
};
void profile_begin();
int profile_end(double* ms, uint64_t* launches, int cap);

}  // namespace pgm
