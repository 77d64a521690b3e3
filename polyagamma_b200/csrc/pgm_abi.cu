// pgm_abi.cu -- extern "C" entry points declared in include/pgm_cuda.h.
//
// Thin argument marshalling over launch_fill / launch_density: no torch types, no C++ types
// in the signatures.  See the header for which reference symbol each function replaces.
#include <math.h>
#include <string.h>

#include <mutex>

#include "../../include/pgm_cuda.h"
#include "pgm_internal.h"
#include "pgm_methods.cuh"

// (no `using namespace pgm`: that namespace re-declares log/exp/sqrt for the kernels)
using pgm::DensityArgs;
using pgm::FillArgs;
using pgm::PhiloxKey;
using pgm::derive_call_key;
using pgm::hybrid_select;
using pgm::kDevroye;
using pgm::kMaxNdim;
using pgm::launch_check_h;
using pgm::launch_count;
using pgm::launch_density;
using pgm::launch_fill;
using pgm::launch_logistic_omega_step;
using pgm::launch_test_math;
using pgm::measure_fp64_peak;
using pgm::profile_begin;
using pgm::profile_end;

namespace {

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (dev >= 0 && dev != prev) {
            cudaSetDevice(dev);
            switched = true;
        }
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

FillArgs
make_fill_args(uint64_t seed, uint64_t offset, double* d_out, uint64_t first_index)
{
    FillArgs a;
    memset(&a, 0, sizeof(a));
    a.key = derive_call_key(seed, offset);
    a.out = d_out;
    a.first_index = first_index;
    a.h_stride = 1;
    a.z_stride = 1;
    return a;
}

// product of the extents; false for a negative extent or a product beyond 2^62
bool
element_count(int ndim, const int64_t* shape, uint64_t& n)
{
    n = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return false;
        if (shape[d] != 0 && n > (uint64_t(1) << 62) / (uint64_t)shape[d]) return false;
        n *= (uint64_t)shape[d];
    }
    return true;
}

// Drop size-1 dimensions and merge neighbours that are contiguous for every operand, so the
// per-element index decomposition in the kernels is as short as possible.
// Returns the collapsed ndim; strides arrays are rewritten in place.
int
collapse_dims(int ndim, const int64_t* shape, int nops, const int64_t* const* strides_in, int64_t* shape_out,
              int64_t (*strides_out)[kMaxNdim])
{
    int m = 0;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] == 1) continue;
        bool merged = false;
        if (m > 0) {
            merged = true;
            for (int o = 0; o < nops; ++o) {
                // outer stride must equal inner stride * inner extent
                if (strides_out[o][m - 1] != strides_in[o][d] * shape[d]) merged = false;
            }
            if (merged) {
                shape_out[m - 1] *= shape[d];
                for (int o = 0; o < nops; ++o) strides_out[o][m - 1] = strides_in[o][d];
            }
        }
        if (!merged) {
            shape_out[m] = shape[d];
            for (int o = 0; o < nops; ++o) strides_out[o][m] = strides_in[o][d];
            ++m;
        }
    }
    return m;
}

}  // namespace

extern "C" {

const char*
pgm_cuda_version(void)
{
    return "polyagamma_b200 0.1.0 sm_100a";
}

int
pgm_cuda_init(void)
{
    return pgm::device_init();
}

void
pgm_cuda_debug_report(void)
{
    pgm::evlog_report();
}

uint64_t
pgm_cuda_launch_count(void)
{
    return launch_count();
}

void
pgm_cuda_profile_begin(void)
{
    profile_begin();
}

int
pgm_cuda_profile_end(double* ms, uint64_t* launches, int cap)
{
    if (ms == nullptr || launches == nullptr || cap < 0) return PGM_CUDA_EINVAL;
    return profile_end(ms, launches, cap);
}

int
pgm_cuda_fp64_peak_tflops(double* tflops, void* stream)
{
    if (tflops == nullptr) return PGM_CUDA_EINVAL;
    return measure_fp64_peak(tflops, (cudaStream_t)stream);
}

int
pgm_cuda_logistic_omega_step(uint64_t seed, uint64_t offset, const double* d_X, size_t n, int p, size_t ldx,
                             const double* d_beta, const double* d_ntrials, double ntrials_scalar, int method,
                             double* d_gram, double* d_omega, double* d_psi, uint64_t first_index, void* stream)
{
    if (d_gram == nullptr || d_beta == nullptr || (n != 0 && d_X == nullptr)) return PGM_CUDA_EINVAL;
    return launch_logistic_omega_step(seed, offset, d_X, n, p, ldx, d_beta, d_ntrials, ntrials_scalar, method, d_gram,
                                      d_omega, d_psi, first_index, (cudaStream_t)stream);
}

int
pgm_cuda_release_scratch(void)
{
    pgm::host_rings_release();
    return pgm::scratch_release();
}

int
pgm_cuda_test_math(int which, const double* d_in, size_t n, double* d_out, void* stream)
{
    if (n != 0 && (d_in == nullptr || d_out == nullptr)) return PGM_CUDA_EINVAL;
    return launch_test_math(which, d_in, n, d_out, (cudaStream_t)stream);
}

int
pgm_cuda_hybrid_select(double h, double z)
{
    return hybrid_select(h, z);
}

void
pgm_cuda_call_key(uint64_t seed, uint64_t offset, uint32_t key[2])
{
    PhiloxKey k = derive_call_key(seed, offset);
    key[0] = k.k0;
    key[1] = k.k1;
}

int
pgm_cuda_random_polyagamma_fill(uint64_t seed, uint64_t offset, double h, double z, int method, size_t n,
                                double* d_out, uint64_t first_index, void* stream)
{
    if (n != 0 && d_out == nullptr) return PGM_CUDA_EINVAL;
    FillArgs a = make_fill_args(seed, offset, d_out, first_index);
    a.h_scalar = h;
    a.z_scalar = z;
    return launch_fill(a, n, method, (cudaStream_t)stream);
}

int
pgm_cuda_random_polyagamma_fill2(uint64_t seed, uint64_t offset, const double* d_h, const double* d_z, int method,
                                 size_t n, double* d_out, uint64_t first_index, void* stream)
{
    if (n != 0 && (d_out == nullptr || d_h == nullptr || d_z == nullptr)) return PGM_CUDA_EINVAL;
    FillArgs a = make_fill_args(seed, offset, d_out, first_index);
    a.h = d_h;
    a.z = d_z;
    return launch_fill(a, n, method, (cudaStream_t)stream);
}

int
pgm_cuda_random_polyagamma_fill2_strided(uint64_t seed, uint64_t offset, const double* d_h, const double* d_z,
                                         int method, int ndim, const int64_t* shape, const int64_t* h_strides,
                                         const int64_t* z_strides, double* d_out, uint64_t first_index,
                                         void* stream)
{
    if (ndim < 0 || ndim > PGM_CUDA_MAX_NDIM) return PGM_CUDA_EINVAL;
    if (ndim > 0 && (shape == nullptr || h_strides == nullptr || z_strides == nullptr)) return PGM_CUDA_EINVAL;
    uint64_t n = 1;
    if (!element_count(ndim, shape, n)) return PGM_CUDA_EINVAL;
    if (n == 0) return 0;
    if (d_out == nullptr || d_h == nullptr || d_z == nullptr) return PGM_CUDA_EINVAL;
    FillArgs a = make_fill_args(seed, offset, d_out, first_index);
    a.h = d_h;
    a.z = d_z;
    const int64_t* sin[2] = {h_strides, z_strides};
    int64_t sout[2][kMaxNdim];
    int64_t shp[kMaxNdim];
    int m = collapse_dims(ndim, shape, 2, sin, shp, sout);
    if (m <= 1) {
        // one effective dimension: plain linear addressing with stride 0 or 1 ... or k
        a.h_stride = m ? sout[0][0] : 0;
        a.z_stride = m ? sout[1][0] : 0;
    }
    else {
        a.bc.ndim = m;
        for (int d = 0; d < m; ++d) {
            a.bc.shape[d] = shp[d];
            a.bc.hs[d] = sout[0][d];
            a.bc.zs[d] = sout[1][d];
        }
    }
    return launch_fill(a, n, method, (cudaStream_t)stream);
}

int
pgm_cuda_any_invalid_h(const double* d_h, size_t n, int method, int* result, void* stream)
{
    if (result == nullptr) return PGM_CUDA_EINVAL;
    *result = 0;
    if (n == 0) return 0;
    if (d_h == nullptr) return PGM_CUDA_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    int* d_flag = nullptr;
    cudaError_t e;
    if ((e = (cudaError_t)pgm::scratch_alloc((void**)&d_flag, sizeof(int), s)) != cudaSuccess) return (int)e;
    cudaMemsetAsync(d_flag, 0, sizeof(int), s);
    int rc = launch_check_h(d_h, n, (method & 0xff) == kDevroye, d_flag, s);
    if (rc == 0) {
        e = cudaMemcpyAsync(result, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        rc = (int)e;
    }
    cudaFreeAsync(d_flag, s);
    return rc;
}

// ---- densities --------------------------------------------------------------------------
static int
density_linear(int which, const double* d_x, const double* d_h, const double* d_z, size_t n, double* d_out,
               void* stream)
{
    if (n == 0) return 0;
    if (d_x == nullptr || d_h == nullptr || d_z == nullptr || d_out == nullptr) return PGM_CUDA_EINVAL;
    DensityArgs a;
    memset(&a, 0, sizeof(a));
    a.x = d_x;
    a.h = d_h;
    a.z = d_z;
    a.out = d_out;
    return launch_density(which, a, n, (cudaStream_t)stream);
}

int
pgm_cuda_polyagamma_pdf(const double* d_x, const double* d_h, const double* d_z, size_t n, double* d_out,
                        void* stream)
{
    return density_linear(0, d_x, d_h, d_z, n, d_out, stream);
}

int
pgm_cuda_polyagamma_cdf(const double* d_x, const double* d_h, const double* d_z, size_t n, double* d_out,
                        void* stream)
{
    return density_linear(1, d_x, d_h, d_z, n, d_out, stream);
}

int
pgm_cuda_polyagamma_logpdf(const double* d_x, const double* d_h, const double* d_z, size_t n, double* d_out,
                           void* stream)
{
    return density_linear(2, d_x, d_h, d_z, n, d_out, stream);
}

int
pgm_cuda_polyagamma_logcdf(const double* d_x, const double* d_h, const double* d_z, size_t n, double* d_out,
                           void* stream)
{
    return density_linear(3, d_x, d_h, d_z, n, d_out, stream);
}

int
pgm_cuda_polyagamma_density_strided(int which, const double* d_x, const double* d_h, const double* d_z, int ndim,
                                    const int64_t* shape, const int64_t* x_strides, const int64_t* h_strides,
                                    const int64_t* z_strides, double* d_out, void* stream)
{
    if (ndim < 0 || ndim > PGM_CUDA_MAX_NDIM || which < 0 || which > 7) return PGM_CUDA_EINVAL;
    if (ndim > 0 && (shape == nullptr || x_strides == nullptr || h_strides == nullptr || z_strides == nullptr))
        return PGM_CUDA_EINVAL;
    uint64_t n = 1;
    if (!element_count(ndim, shape, n)) return PGM_CUDA_EINVAL;
    if (n == 0) return 0;
    if (d_x == nullptr || d_h == nullptr || d_z == nullptr || d_out == nullptr) return PGM_CUDA_EINVAL;
    DensityArgs a;
    memset(&a, 0, sizeof(a));
    a.x = d_x;
    a.h = d_h;
    a.z = d_z;
    a.out = d_out;
    const int64_t* sin[3] = {x_strides, h_strides, z_strides};
    int64_t sout[3][kMaxNdim];
    int64_t shp[kMaxNdim];
    int m = collapse_dims(ndim, shape, 3, sin, shp, sout);
    // ndim == 0 in the kernel means "all three operands contiguous"; anything else goes
    // through the strided path (a single collapsed dimension with strides 0/1 included)
    const bool linear = (m == 1 && sout[0][0] == 1 && sout[1][0] == 1 && sout[2][0] == 1);
    if (!linear) {
        if (m == 0) {  // every dimension has extent 1
            m = 1;
            shp[0] = 1;
            sout[0][0] = sout[1][0] = sout[2][0] = 0;
        }
        a.bc.ndim = m;
        for (int d = 0; d < m; ++d) {
            a.bc.shape[d] = shp[d];
            a.bc.xs[d] = sout[0][d];
            a.bc.hs[d] = sout[1][d];
            a.bc.zs[d] = sout[2][d];
        }
    }
    return launch_density(which, a, n, (cudaStream_t)stream);
}

}  // extern "C"
