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
    uint64_t n = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return PGM_CUDA_EINVAL;
        n *= (uint64_t)shape[d];
    }
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

double
pgm_cuda_random_polyagamma(uint64_t seed, uint64_t offset, double h, double z, int method)
{
    double* d = nullptr;
    double out = nan("");
    if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) return out;
    int rc = pgm_cuda_random_polyagamma_fill(seed, offset, h, z, method, 1, d, 0, nullptr);
    if (rc == 0 && cudaMemcpy(&out, d, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        out = nan("");
    }
    cudaFree(d);
    return out;
}

// ---- host-buffer fill2: chunked H2D -> kernels -> D2H on a ring of streams ---------------
namespace {
constexpr int kHostStreams = 3;
constexpr size_t kHostChunk = size_t(1) << 23;  // elements per ring slot (64 MiB per buffer)

struct HostRing {
    bool ready = false;
    cudaEvent_t ev_start, ev_end, ev_join[kHostStreams];
    cudaStream_t streams[kHostStreams];
    double* d_h[kHostStreams];
    double* d_z[kHostStreams];
    double* d_out[kHostStreams];
};
HostRing g_rings[64];
std::mutex g_ring_mutex;
double g_last_host_fill_ms = 0.0;

int
ring_init(HostRing& r)
{
    if (r.ready) return 0;
    for (int i = 0; i < kHostStreams; ++i) {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&r.streams[i], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
        if ((e = cudaMalloc(&r.d_h[i], kHostChunk * sizeof(double))) != cudaSuccess) return (int)e;
        if ((e = cudaMalloc(&r.d_z[i], kHostChunk * sizeof(double))) != cudaSuccess) return (int)e;
        if ((e = cudaMalloc(&r.d_out[i], kHostChunk * sizeof(double))) != cudaSuccess) return (int)e;
        cudaEventCreateWithFlags(&r.ev_join[i], cudaEventDisableTiming);
    }
    cudaEventCreate(&r.ev_start);
    cudaEventCreate(&r.ev_end);
    r.ready = true;
    return 0;
}
}  // namespace

int
pgm_cuda_random_polyagamma_fill2_host(uint64_t seed, uint64_t offset, const double* h, int h_stride,
                                      const double* z, int z_stride, int method, size_t n, double* out,
                                      uint64_t first_index, int device)
{
    if (n == 0) return 0;
    if (h == nullptr || z == nullptr || out == nullptr) return PGM_CUDA_EINVAL;
    if ((h_stride != 0 && h_stride != 1) || (z_stride != 0 && z_stride != 1)) return PGM_CUDA_EINVAL;
    DeviceGuard guard(device);
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_ring_mutex);
    HostRing& r = g_rings[dev & 63];
    int rc = ring_init(r);
    if (rc != 0) return rc;

    // device-side clock of the whole call: start on stream 0 before any copy is queued, end on
    // stream 0 after it has joined the other ring streams
    cudaEventRecord(r.ev_start, r.streams[0]);
    for (int i = 1; i < kHostStreams; ++i) cudaStreamWaitEvent(r.streams[i], r.ev_start, 0);
    size_t c = 0;
    for (size_t off = 0; off < n; off += kHostChunk, ++c) {
        const int slot = (int)(c % kHostStreams);
        cudaStream_t s = r.streams[slot];
        const size_t cn = (n - off) < kHostChunk ? (n - off) : kHostChunk;
        FillArgs a = make_fill_args(seed, offset, r.d_out[slot], first_index + off);
        cudaError_t e;
        if (h_stride) {
            if ((e = cudaMemcpyAsync(r.d_h[slot], h + off, cn * sizeof(double), cudaMemcpyHostToDevice, s)) !=
                cudaSuccess)
                return (int)e;
            a.h = r.d_h[slot];
        }
        else {
            a.h_scalar = h[0];
        }
        if (z_stride) {
            if ((e = cudaMemcpyAsync(r.d_z[slot], z + off, cn * sizeof(double), cudaMemcpyHostToDevice, s)) !=
                cudaSuccess)
                return (int)e;
            a.z = r.d_z[slot];
        }
        else {
            a.z_scalar = z[0];
        }
        if ((rc = launch_fill(a, cn, method, s)) != 0) return rc;
        if ((e = cudaMemcpyAsync(out + off, r.d_out[slot], cn * sizeof(double), cudaMemcpyDeviceToHost, s)) !=
            cudaSuccess)
            return (int)e;
    }
    for (int i = 1; i < kHostStreams; ++i) {
        cudaEventRecord(r.ev_join[i], r.streams[i]);
        cudaStreamWaitEvent(r.streams[0], r.ev_join[i], 0);
    }
    cudaEventRecord(r.ev_end, r.streams[0]);
    for (int i = 0; i < kHostStreams; ++i) {
        cudaError_t e = cudaStreamSynchronize(r.streams[i]);
        if (e != cudaSuccess) return (int)e;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.ev_start, r.ev_end) == cudaSuccess) g_last_host_fill_ms = ms;
    return 0;
}

double
pgm_cuda_last_host_fill_ms(void)
{
    return g_last_host_fill_ms;
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
    if (ndim < 0 || ndim > PGM_CUDA_MAX_NDIM || which < 0 || which > 3) return PGM_CUDA_EINVAL;
    uint64_t n = 1;
    for (int d = 0; d < ndim; ++d) {
        if (shape[d] < 0) return PGM_CUDA_EINVAL;
        n *= (uint64_t)shape[d];
    }
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

int
pgm_cuda_polyagamma_density_host(int which, const double* x, int x_stride, const double* h, int h_stride,
                                 const double* z, int z_stride, size_t n, double* out, int device)
{
    if (which < 0 || which > 3) return PGM_CUDA_EINVAL;
    if (n == 0) return 0;
    if (x == nullptr || h == nullptr || z == nullptr || out == nullptr) return PGM_CUDA_EINVAL;
    const int strides[3] = {x_stride, h_stride, z_stride};
    for (int k = 0; k < 3; ++k)
        if (strides[k] != 0 && strides[k] != 1) return PGM_CUDA_EINVAL;
    DeviceGuard guard(device);
    cudaStream_t s = nullptr;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
    const size_t chunk = n < kHostChunk ? n : kHostChunk;
    const double* src[3] = {x, h, z};
    double* d[4] = {nullptr, nullptr, nullptr, nullptr};
    int rc = 0;
    for (int k = 0; k < 4 && rc == 0; ++k) {
        const size_t len = (k < 3 && strides[k] == 0) ? 1 : chunk;
        rc = pgm::scratch_alloc((void**)&d[k], len * sizeof(double), s);
    }
    for (int k = 0; k < 3 && rc == 0; ++k)
        if (strides[k] == 0) rc = (int)cudaMemcpyAsync(d[k], src[k], sizeof(double), cudaMemcpyHostToDevice, s);
    for (size_t off = 0; off < n && rc == 0; off += chunk) {
        const size_t cn = (n - off) < chunk ? (n - off) : chunk;
        for (int k = 0; k < 3 && rc == 0; ++k)
            if (strides[k] == 1)
                rc = (int)cudaMemcpyAsync(d[k], src[k] + off, cn * sizeof(double), cudaMemcpyHostToDevice, s);
        if (rc != 0) break;
        DensityArgs a;
        memset(&a, 0, sizeof(a));
        a.x = d[0];
        a.h = d[1];
        a.z = d[2];
        a.out = d[3];
        if (!(x_stride == 1 && h_stride == 1 && z_stride == 1)) {
            a.bc.ndim = 1;
            a.bc.shape[0] = (int64_t)cn;
            a.bc.xs[0] = x_stride;
            a.bc.hs[0] = h_stride;
            a.bc.zs[0] = z_stride;
        }
        rc = launch_density(which, a, cn, s);
        if (rc == 0) rc = (int)cudaMemcpyAsync(out + off, d[3], cn * sizeof(double), cudaMemcpyDeviceToHost, s);
    }
    for (int k = 0; k < 4; ++k)
        if (d[k]) cudaFreeAsync(d[k], s);
    e = cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return rc != 0 ? rc : (int)e;
}

}  // extern "C"
