// pgm_host.cu -- the HOST-buffer entry points of include/pgm_cuda.h: the reference's calling convention (all
// pointers are host memory, include/pgm_random.h:47-74, include/pgm_density.h:4-14) on top of the device
// pipeline.
//
//   fill2_host / fill2_host_multi / density_host   chunked H2D -> kernels -> D2H on a ring of streams PER DEVICE
//   random_polyagamma (one draw) / tiny density calls   one launch that writes straight into mapped pinned
//                                                       memory of a per-thread slot: no allocation, no copy call
//
// A ring (three streams, each with device staging for one chunk of every operand) belongs to one device and is
// guarded by that device's mutex: calls on different devices run side by side, calls on one device take turns.
// The staging is sized by the largest chunk asked for so far (a one-element call allocates three doubles per
// stream, not three 64 MiB buffers) and is returned to the driver by pgm_cuda_release_scratch.  Whatever
// happens, no entry point returns while a copy it queued may still touch the caller's memory.
#include <math.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/pgm_cuda.h"
#include "pgm_internal.h"

using pgm::DensityArgs;
using pgm::FillArgs;

namespace {

constexpr int kStreams = 3;
constexpr size_t kChunkMax = size_t(1) << 23;  // elements per ring slot (64 MiB per operand)
constexpr size_t kChunkMin = size_t(1) << 20;
constexpr int kOperands = 4;                   // x / h / z staging + result

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

struct Ring {
    std::mutex mu;
    bool ready = false;
    size_t cap = 0;  // elements each staging buffer holds
    cudaStream_t streams[kStreams] = {};
    cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_join[kStreams] = {};
    double* buf[kStreams][kOperands] = {};
    // pinned bounce buffers: [slot][0] results on their way to pageable memory (pgm_cuda_download and fills into
    // pageable arrays), [slot][1..2] pageable h / z operands on their way to the device
    double* pin[kStreams][3] = {};
    size_t pin_cap = 0;  // elements each holds
    int pin_count = 0;   // buffers per slot that exist (1 after a download, 3 after a bounced fill)
    std::atomic<double> last_ms{0.0};
};
Ring g_rings[64];
thread_local double t_last_ms = 0.0;

void
ring_free_pinned(Ring& r)
{
    for (auto& slot : r.pin)
        for (auto& p : slot) {
            if (p) cudaFreeHost(p);
            p = nullptr;
        }
    r.pin_cap = 0;
    r.pin_count = 0;
}

// `count` pinned buffers of `need` elements per slot (kept, and grown, across calls)
int
ring_prepare_pinned(Ring& r, size_t need, int count)
{
    if (r.pin_cap >= need && r.pin_count >= count) return 0;
    const size_t cap = need > r.pin_cap ? need : r.pin_cap;
    const int cnt = count > r.pin_count ? count : r.pin_count;
    ring_free_pinned(r);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < kStreams && e == cudaSuccess; ++i)
        for (int k = 0; k < cnt && e == cudaSuccess; ++k) e = cudaHostAlloc((void**)&r.pin[i][k], cap * sizeof(double), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        ring_free_pinned(r);
        (void)cudaGetLastError();
        return (int)e;
    }
    r.pin_cap = cap;
    r.pin_count = cnt;
    return 0;
}

void
ring_free_buffers(Ring& r)
{
    for (auto& slot : r.buf)
        for (auto& p : slot) {
            if (p) cudaFree(p);
            p = nullptr;
        }
    r.cap = 0;
    ring_free_pinned(r);
}

void
ring_destroy(Ring& r)
{
    ring_free_buffers(r);
    for (int i = 0; i < kStreams; ++i) {
        if (r.streams[i]) cudaStreamDestroy(r.streams[i]);
        if (r.ev_join[i]) cudaEventDestroy(r.ev_join[i]);
        r.streams[i] = nullptr;
        r.ev_join[i] = nullptr;
    }
    if (r.ev_start) cudaEventDestroy(r.ev_start);
    if (r.ev_end) cudaEventDestroy(r.ev_end);
    r.ev_start = r.ev_end = nullptr;
    r.ready = false;
}

// streams and events once, staging for `need` elements per operand (grown, never shrunk, by later calls);
// on failure everything this call created is released again, so a retry starts clean
int
ring_prepare(Ring& r, size_t need, int noperands)
{
    cudaError_t e = cudaSuccess;
    if (!r.ready) {
        for (int i = 0; i < kStreams && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&r.streams[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.ev_join[i], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreate(&r.ev_start);
        if (e == cudaSuccess) e = cudaEventCreate(&r.ev_end);
        if (e != cudaSuccess) {
            ring_destroy(r);
            (void)cudaGetLastError();
            return (int)e;
        }
        r.ready = true;
    }
    bool have = need <= r.cap;
    for (int i = 0; i < kStreams && have; ++i)
        for (int k = 0; k < noperands; ++k) have = have && r.buf[i][k] != nullptr;
    if (have) return 0;
    // (the ring is idle: every call drains it before it returns)
    const size_t cap = need > r.cap ? need : r.cap;
    ring_free_buffers(r);
    for (int i = 0; i < kStreams && e == cudaSuccess; ++i)
        for (int k = 0; k < kOperands && e == cudaSuccess; ++k) e = cudaMalloc(&r.buf[i][k], cap * sizeof(double));
    if (e != cudaSuccess) {
        ring_free_buffers(r);
        (void)cudaGetLastError();
        return (int)e;
    }
    r.cap = cap;
    return 0;
}

int
ring_drain(Ring& r)
{
    int rc = 0;
    for (int i = 0; i < kStreams; ++i) {
        if (r.streams[i] == nullptr) continue;
        cudaError_t e = cudaStreamSynchronize(r.streams[i]);
        if (e != cudaSuccess && rc == 0) rc = (int)e;
    }
    return rc;
}

// elements per chunk: enough chunks that the copies of one overlap the kernels of another, none so small that
// launch overhead shows
size_t
chunk_for(size_t n)
{
    size_t c = n / (4 * kStreams);
    if (c > kChunkMax) c = kChunkMax;
    if (c < kChunkMin) c = kChunkMin;
    if (c > n) c = n;
    return c;
}

struct HostFill {
    uint64_t seed, offset;
    const double *h, *z;
    int h_stride, z_stride, method;
    double* out;
};

// queue chunk [off, off + cn) of a fill on ring slot `slot` (no waiting); hsrc / zsrc / odst are the chunk's own
// host arrays (the caller's, or their pinned copies)
int
queue_fill_chunk(Ring& r, int slot, const HostFill& f, const double* hsrc, const double* zsrc, double* odst, size_t off,
                 size_t cn, uint64_t first_index)
{
    cudaStream_t s = r.streams[slot];
    FillArgs a;
    memset(&a, 0, sizeof(a));
    a.key = pgm::derive_call_key(f.seed, f.offset);
    a.out = r.buf[slot][3];
    a.first_index = first_index + off;
    a.h_stride = 1;
    a.z_stride = 1;
    cudaError_t e;
    if (f.h_stride) {
        if ((e = cudaMemcpyAsync(r.buf[slot][1], hsrc, cn * sizeof(double), cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return (int)e;
        a.h = r.buf[slot][1];
    }
    else {
        a.h_scalar = f.h[0];
    }
    if (f.z_stride) {
        if ((e = cudaMemcpyAsync(r.buf[slot][2], zsrc, cn * sizeof(double), cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return (int)e;
        a.z = r.buf[slot][2];
    }
    else {
        a.z_scalar = f.z[0];
    }
    if (int rc = pgm::launch_fill(a, cn, f.method, s)) return rc;
    return (int)cudaMemcpyAsync(odst, r.buf[slot][3], cn * sizeof(double), cudaMemcpyDeviceToHost, s);
}

// memcpy by several host threads.  The destination of a result is usually a freshly allocated array whose first
// touch (page faults) one thread cannot do faster than ~2 GB/s, and a single thread's memcpy between pageable
// and pinned memory is well below what the host link carries.  The workers are started once per process and
// sleep between calls; callers (one per device ring) take turns.
class CopyPool {
  public:
    void copy(double* dst, const double* src, size_t n)
    {
        constexpr size_t kMinPerThread = size_t(1) << 17;  // 1 MiB
        std::lock_guard<std::mutex> turn(turn_);
        start_workers();
        size_t parts = workers_.size() + 1;
        if (n / kMinPerThread < parts) parts = n / kMinPerThread;
        if (parts <= 1) {
            memcpy(dst, src, n * sizeof(double));
            return;
        }
        const size_t per = ((n + parts - 1) / parts + 511) & ~size_t(511);  // (page-sized pieces)
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst, src_ = src, n_ = n, per_ = per;
            next_ = 1;  // piece 0 is the caller's
            pending_ = 0;
            for (size_t lo = per; lo < n; lo += per) ++pending_;
            ++generation_;
        }
        cv_.notify_all();
        memcpy(dst, src, (per < n ? per : n) * sizeof(double));
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }

  private:
    void start_workers()
    {
        if (!workers_.empty() || started_) return;
        started_ = true;
        unsigned hw = std::thread::hardware_concurrency();
        const unsigned nt = hw >= 16 ? 7 : (hw >= 4 ? hw / 2 - 1 : 0);
        for (unsigned i = 0; i < nt; ++i) workers_.emplace_back([this] { run(); });
    }
    void run()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            for (;;) {
                const size_t lo = next_ * per_;
                if (lo >= n_) break;
                ++next_;
                const size_t cnt = (n_ - lo) < per_ ? (n_ - lo) : per_;
                double* d = dst_ + lo;
                const double* sp = src_ + lo;
                lk.unlock();
                memcpy(d, sp, cnt * sizeof(double));
                lk.lock();
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::mutex turn_, mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> workers_;
    bool started_ = false, stop_ = false;
    uint64_t generation_ = 0;
    double* dst_ = nullptr;
    const double* src_ = nullptr;
    size_t n_ = 0, per_ = 0, next_ = 0, pending_ = 0;
};
CopyPool g_copy_pool;

void
parallel_memcpy(double* dst, const double* src, size_t n)
{
    g_copy_pool.copy(dst, src, n);
}

// is p ordinary (pageable) host memory?  Pinned / registered / managed memory can be the end point of an
// asynchronous copy at link speed; pageable memory is staged by the driver, synchronously, by one thread.
bool
is_pageable(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// A fill whose operands or result live in PAGEABLE host memory (plain numpy arrays: what a caller of the
// reference has).  Handing such pointers to cudaMemcpyAsync makes the driver stage them itself, synchronously
// and with one thread (~13 GB/s over all three arrays on the bench box).  Here the chunks go through the ring's
// pinned buffers instead: the copy pool moves chunk c of h and z in, the device works on it asynchronously, and
// the results of chunk c - 3 are moved out while it does.
int
bounced_fill(Ring& r, const HostFill& f, size_t n, uint64_t first_index)
{
    constexpr size_t kBounceChunk = size_t(1) << 21;  // 16 MiB per buffer, nine buffers
    const size_t chunk = n < kBounceChunk ? n : kBounceChunk;
    int rc = ring_prepare(r, chunk, kOperands);
    if (rc == 0) rc = ring_prepare_pinned(r, chunk, 3);
    if (rc != 0) return rc;
    cudaEventRecord(r.ev_start, r.streams[0]);
    for (int k = 1; k < kStreams; ++k) cudaStreamWaitEvent(r.streams[k], r.ev_start, 0);
    const size_t nchunks = (n + chunk - 1) / chunk;
    auto span = [&](size_t c, size_t& off, size_t& cn) {
        off = c * chunk;
        cn = (n - off) < chunk ? (n - off) : chunk;
    };
    auto land = [&](size_t c) {
        const int slot = (int)(c % kStreams);
        const cudaError_t es = cudaStreamSynchronize(r.streams[slot]);
        if (es != cudaSuccess) return (int)es;
        size_t off, cn;
        span(c, off, cn);
        parallel_memcpy(f.out + off, r.pin[slot][0], cn);
        return 0;
    };
    for (size_t c = 0; c < nchunks && rc == 0; ++c) {
        const int slot = (int)(c % kStreams);
        if (c >= (size_t)kStreams && (rc = land(c - kStreams)) != 0) break;
        size_t off, cn;
        span(c, off, cn);
        if (f.h_stride) parallel_memcpy(r.pin[slot][1], f.h + off, cn);
        if (f.z_stride) parallel_memcpy(r.pin[slot][2], f.z + off, cn);
        rc = queue_fill_chunk(r, slot, f, r.pin[slot][1], r.pin[slot][2], r.pin[slot][0], off, cn, first_index);
    }
    for (size_t c = nchunks > (size_t)kStreams ? nchunks - kStreams : 0; c < nchunks && rc == 0; ++c) rc = land(c);
    if (rc == 0) {
        for (int k = 1; k < kStreams; ++k) {
            cudaEventRecord(r.ev_join[k], r.streams[k]);
            cudaStreamWaitEvent(r.streams[0], r.ev_join[k], 0);
        }
        cudaEventRecord(r.ev_end, r.streams[0]);
    }
    const int drc = ring_drain(r);
    if (rc == 0) rc = drc;
    float ms = 0.f;
    if (rc == 0 && cudaEventElapsedTime(&ms, r.ev_start, r.ev_end) == cudaSuccess) {
        r.last_ms.store(ms);
        t_last_ms = ms;
    }
    if (rc != 0) (void)cudaGetLastError();
    return rc;
}

// ---- one draw / a handful of points: mapped pinned memory, one launch --------------------------------------
struct TinySlot {
    double* host = nullptr;  // 8 doubles of mapped pinned memory: [0] result, [1..3] x, h, z of a density point
    double* dev = nullptr;
    cudaStream_t stream = nullptr;
    int device = -1;
};
thread_local TinySlot t_tiny[64];

TinySlot*
tiny_slot()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    TinySlot& s = t_tiny[dev & 63];
    if (s.host == nullptr) {
        if (cudaHostAlloc((void**)&s.host, 8 * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&s.dev, s.host, 0) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) {
            if (s.host) cudaFreeHost(s.host);
            s = TinySlot{};
            (void)cudaGetLastError();
            return nullptr;
        }
        s.device = dev;
    }
    return &s;
}

}  // namespace

namespace pgm {
// called by pgm_cuda_release_scratch: the staging of the current device's ring
void
host_rings_release()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    Ring& r = g_rings[dev & 63];
    std::lock_guard<std::mutex> lock(r.mu);
    ring_drain(r);
    ring_free_buffers(r);
}
}  // namespace pgm

extern "C" {

int
pgm_cuda_random_polyagamma_fill2_host(uint64_t seed, uint64_t offset, const double* h, int h_stride, const double* z,
                                      int z_stride, int method, size_t n, double* out, uint64_t first_index, int device)
{
    int dev = device;
    if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    return pgm_cuda_random_polyagamma_fill2_host_multi(seed, offset, h, h_stride, z, z_stride, method, n, out, first_index,
                                                       &dev, 1);
}

int
pgm_cuda_random_polyagamma_fill2_host_multi(uint64_t seed, uint64_t offset, const double* h, int h_stride,
                                            const double* z, int z_stride, int method, size_t n, double* out,
                                            uint64_t first_index, const int* devices, int ndevices)
{
    if (n == 0) return 0;
    if (h == nullptr || z == nullptr || out == nullptr || devices == nullptr) return PGM_CUDA_EINVAL;
    if ((h_stride != 0 && h_stride != 1) || (z_stride != 0 && z_stride != 1)) return PGM_CUDA_EINVAL;
    if (ndevices < 1 || ndevices > 64) return PGM_CUDA_EINVAL;
    for (int i = 0; i < ndevices; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j] || devices[i] < 0) return PGM_CUDA_EINVAL;
    if ((size_t)ndevices > n) ndevices = (int)n;
    DeviceGuard guard(-1);
    // pageable arrays (one device, enough elements that the copies matter): bounce them through pinned memory
    if (ndevices == 1 && n >= (size_t(1) << 16) &&
        ((h_stride && is_pageable(h)) || (z_stride && is_pageable(z)) || is_pageable(out))) {
        if (cudaSetDevice(devices[0]) != cudaSuccess) return (int)cudaGetLastError();
        Ring& r = g_rings[devices[0] & 63];
        std::lock_guard<std::mutex> lock(r.mu);
        const HostFill fb{seed, offset, h, z, h_stride, z_stride, method, out};
        return bounced_fill(r, fb, n, first_index);
    }

    // contiguous shards of the index range, one per device (the partition of polyagamma_b200/sharding.py);
    // a shard's elements keep their global Philox indices, so the result does not depend on ndevices
    struct Shard {
        Ring* ring;
        size_t lo, hi, chunk, next;
        int chunks_queued;
    } sh[64];
    int locked = 0, rc = 0;
    // (rings are locked in ascending device order: two multi-device calls cannot deadlock)
    int order[64];
    for (int i = 0; i < ndevices; ++i) order[i] = i;
    for (int i = 0; i < ndevices; ++i)
        for (int j = i + 1; j < ndevices; ++j)
            if (devices[order[j]] < devices[order[i]]) {
                int t = order[i];
                order[i] = order[j];
                order[j] = t;
            }
    for (; locked < ndevices && rc == 0; ++locked) {
        const int i = order[locked];
        Shard& s = sh[i];
        s.ring = &g_rings[devices[i] & 63];
        s.ring->mu.lock();
        s.lo = n * (size_t)i / ndevices;
        s.hi = n * (size_t)(i + 1) / ndevices;
        s.chunk = chunk_for(s.hi - s.lo);
        s.next = s.lo;
        s.chunks_queued = 0;
        if (cudaSetDevice(devices[i]) != cudaSuccess) rc = (int)cudaGetLastError();
        else rc = ring_prepare(*s.ring, s.chunk, kOperands);
        if (rc == 0) {
            // device-side clock of the shard: start before any copy is queued, end after the ring has joined
            cudaEventRecord(s.ring->ev_start, s.ring->streams[0]);
            for (int k = 1; k < kStreams; ++k) cudaStreamWaitEvent(s.ring->streams[k], s.ring->ev_start, 0);
        }
    }
    const HostFill f{seed, offset, h, z, h_stride, z_stride, method, out};
    // queue the chunks round-robin over the devices: one host thread keeps every device's copy engines and SMs
    // busy (the copies from pinned memory and the launches are asynchronous)
    bool more = (rc == 0);
    while (more && rc == 0) {
        more = false;
        for (int i = 0; i < ndevices && rc == 0; ++i) {
            Shard& s = sh[i];
            if (s.next >= s.hi) continue;
            const size_t cn = (s.hi - s.next) < s.chunk ? (s.hi - s.next) : s.chunk;
            if (cudaSetDevice(devices[i]) != cudaSuccess) {
                rc = (int)cudaGetLastError();
                break;
            }
            rc = queue_fill_chunk(*s.ring, s.chunks_queued % kStreams, f, f.h + (f.h_stride ? s.next : 0),
                                  f.z + (f.z_stride ? s.next : 0), f.out + s.next, s.next, cn, first_index);
            s.next += cn;
            s.chunks_queued += 1;
            more = more || s.next < s.hi;
        }
    }
    // drain every ring that was locked -- also on failure: queued copies may still be touching the caller's memory
    double ms_max = 0.0;
    for (int j = 0; j < locked; ++j) {
        const int i = order[j];
        Shard& s = sh[i];
        cudaSetDevice(devices[i]);
        if (rc == 0 && s.ring->ready) {
            for (int k = 1; k < kStreams; ++k) {
                cudaEventRecord(s.ring->ev_join[k], s.ring->streams[k]);
                cudaStreamWaitEvent(s.ring->streams[0], s.ring->ev_join[k], 0);
            }
            cudaEventRecord(s.ring->ev_end, s.ring->streams[0]);
        }
        const int drc = ring_drain(*s.ring);
        if (rc == 0) rc = drc;
        float ms = 0.f;
        if (rc == 0 && cudaEventElapsedTime(&ms, s.ring->ev_start, s.ring->ev_end) == cudaSuccess) {
            s.ring->last_ms.store(ms);
            if (ms > ms_max) ms_max = ms;
        }
        s.ring->mu.unlock();
    }
    if (rc == 0) t_last_ms = ms_max;
    else (void)cudaGetLastError();
    return rc;
}

double
pgm_cuda_last_host_fill_ms(void)
{
    return t_last_ms;
}

double
pgm_cuda_random_polyagamma(uint64_t seed, uint64_t offset, double h, double z, int method)
{
    TinySlot* s = tiny_slot();
    if (s == nullptr) return nan("");
    FillArgs a;
    memset(&a, 0, sizeof(a));
    a.key = pgm::derive_call_key(seed, offset);
    a.out = s->dev;  // the kernel writes the draw straight into mapped host memory
    a.h_stride = 1;
    a.z_stride = 1;
    a.h_scalar = h;
    a.z_scalar = z;
    s->host[0] = nan("");
    if (pgm::launch_fill(a, 1, method, s->stream) != 0) return nan("");
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return nan("");
    return s->host[0];
}

int
pgm_cuda_polyagamma_density_host(int which, const double* x, int x_stride, const double* h, int h_stride,
                                 const double* z, int z_stride, size_t n, double* out, int device)
{
    if (which < 0 || which > 7) return PGM_CUDA_EINVAL;
    if (n == 0) return 0;
    if (x == nullptr || h == nullptr || z == nullptr || out == nullptr) return PGM_CUDA_EINVAL;
    const int strides[3] = {x_stride, h_stride, z_stride};
    for (int k = 0; k < 3; ++k)
        if (strides[k] != 0 && strides[k] != 1) return PGM_CUDA_EINVAL;
    DeviceGuard guard(device);
    const double* src[3] = {x, h, z};

    if (n == 1) {
        // the scalar functions of include/pgm_density.h, as the reference's dispatch loop calls them once per
        // point (polyagamma/_polyagamma.pyx:458-469): operands and result live in mapped pinned memory
        TinySlot* s = tiny_slot();
        if (s == nullptr) return (int)cudaErrorMemoryAllocation;
        for (int k = 0; k < 3; ++k) s->host[1 + k] = src[k][0];
        DensityArgs a;
        memset(&a, 0, sizeof(a));
        a.x = s->dev + 1;
        a.h = s->dev + 2;
        a.z = s->dev + 3;
        a.out = s->dev;
        int rc = pgm::launch_density(which, a, 1, s->stream);
        if (rc == 0) rc = (int)cudaStreamSynchronize(s->stream);
        if (rc == 0) out[0] = s->host[0];
        return rc;
    }

    int dev = 0;
    cudaGetDevice(&dev);
    Ring& r = g_rings[dev & 63];
    std::lock_guard<std::mutex> lock(r.mu);
    const size_t chunk = chunk_for(n);
    int rc = ring_prepare(r, chunk, kOperands);
    if (rc != 0) return rc;
    // broadcast operands: one element, uploaded once per ring slot
    size_t c = 0;
    for (size_t off = 0; off < n && rc == 0; off += chunk, ++c) {
        const int slot = (int)(c % kStreams);
        cudaStream_t s = r.streams[slot];
        const size_t cn = (n - off) < chunk ? (n - off) : chunk;
        for (int k = 0; k < 3 && rc == 0; ++k) {
            if (strides[k] == 1)
                rc = (int)cudaMemcpyAsync(r.buf[slot][k], src[k] + off, cn * sizeof(double), cudaMemcpyHostToDevice, s);
            else if (c < (size_t)kStreams)
                rc = (int)cudaMemcpyAsync(r.buf[slot][k], src[k], sizeof(double), cudaMemcpyHostToDevice, s);
        }
        if (rc != 0) break;
        DensityArgs a;
        memset(&a, 0, sizeof(a));
        a.x = r.buf[slot][0];
        a.h = r.buf[slot][1];
        a.z = r.buf[slot][2];
        a.out = r.buf[slot][3];
        if (!(x_stride == 1 && h_stride == 1 && z_stride == 1)) {
            a.bc.ndim = 1;
            a.bc.shape[0] = (int64_t)cn;
            a.bc.xs[0] = x_stride;
            a.bc.hs[0] = h_stride;
            a.bc.zs[0] = z_stride;
        }
        rc = pgm::launch_density(which, a, cn, s);
        if (rc == 0) rc = (int)cudaMemcpyAsync(out + off, r.buf[slot][3], cn * sizeof(double), cudaMemcpyDeviceToHost, s);
    }
    const int drc = ring_drain(r);
    if (rc == 0) rc = drc;
    if (rc != 0) (void)cudaGetLastError();
    return rc;
}

int
pgm_cuda_download(double* dst, const double* d_src, size_t n, void* stream)
{
    if (n == 0) return 0;
    if (dst == nullptr || d_src == nullptr) return PGM_CUDA_EINVAL;
    int dev = 0;
    cudaGetDevice(&dev);
    Ring& r = g_rings[dev & 63];
    std::lock_guard<std::mutex> lock(r.mu);
    int rc = ring_prepare(r, 0, 0);
    if (rc != 0) return rc;
    constexpr size_t kPinChunk = size_t(1) << 22;  // 32 MiB per bounce buffer
    const size_t chunk = n < kPinChunk ? n : kPinChunk;
    if ((rc = ring_prepare_pinned(r, chunk, 1)) != 0) return rc;
    cudaError_t e = cudaSuccess;
    // the producer's stream has to be done with d_src before the ring's streams read it
    if ((e = cudaEventRecord(r.ev_join[0], (cudaStream_t)stream)) != cudaSuccess) return (int)e;
    for (int i = 0; i < kStreams; ++i) cudaStreamWaitEvent(r.streams[i], r.ev_join[0], 0);
    const size_t nchunks = (n + chunk - 1) / chunk;
    auto land = [&](size_t c) {  // chunk c has arrived in its bounce buffer: move it to the caller's array
        const int slot = (int)(c % kStreams);
        const cudaError_t es = cudaStreamSynchronize(r.streams[slot]);
        if (es != cudaSuccess) return (int)es;
        const size_t off = c * chunk, cn = (n - off) < chunk ? (n - off) : chunk;
        parallel_memcpy(dst + off, r.pin[slot][0], cn);
        return 0;
    };
    for (size_t c = 0; c < nchunks && rc == 0; ++c) {
        const int slot = (int)(c % kStreams);
        if (c >= (size_t)kStreams) rc = land(c - kStreams);
        if (rc != 0) break;
        const size_t off = c * chunk, cn = (n - off) < chunk ? (n - off) : chunk;
        rc = (int)cudaMemcpyAsync(r.pin[slot][0], d_src + off, cn * sizeof(double), cudaMemcpyDeviceToHost, r.streams[slot]);
    }
    for (size_t c = nchunks > (size_t)kStreams ? nchunks - kStreams : 0; c < nchunks && rc == 0; ++c) rc = land(c);
    const int drc = ring_drain(r);
    if (rc == 0) rc = drc;
    if (rc != 0) (void)cudaGetLastError();
    return rc;
}

}  // extern "C"
