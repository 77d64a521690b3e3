// pgm_sampler.cu -- the batched Polya-Gamma fill on sm_100a.
//
// Replaces src/pgm_random.c:144-180 (pgm_random_polyagamma / _fill / _fill2 and the hybrid
// dispatch).  Pipeline for one chunk of the output index range:
//
//   hybrid:   bucket_kernel     one pass: bucket id per element (+ NaN/inf for invalid h), one list
//                               reservation per bucket and block, per-bucket element lists
//             per bucket        setup_kernel<M> (per-element records) + sample_kernel<M>, or
//                               dense_kernel, over that bucket's list
//   explicit: the same kernels over the identity list
//
// sample_kernel: every lane owns one element at a time and advances it by ONE proposal
// attempt per loop trip (pgm_methods.cuh); lanes whose element completed are retired and
// refilled from a per-bucket work counter, found with __ballot_sync/__popc, so a warp never
// idles on another lane's rejection loop.  Nothing here synchronises the host.
#include <stdlib.h>

#include "pgm_internal.h"
#include "pgm_methods.cuh"

namespace pgm {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void count_launch(uint64_t n) { g_launches.fetch_add(n); }

// Optional per-kernel timing (bench.py's roofline leg): CUDA events recorded on the launching
// stream around every kernel while profiling is on; read back by profile_end().
struct TimedLaunch {
    int kind;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mutex;
static bool g_prof_on = false;
static std::vector<TimedLaunch> g_prof;

ScopedKernelTimer::ScopedKernelTimer(int kind, cudaStream_t s) : kind_(kind), stream_(s), on_(false)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (!g_prof_on) return;
    on_ = true;
    cudaEventCreate(&e0_);
    cudaEventCreate(&e1_);
    cudaEventRecord(e0_, stream_);
}

ScopedKernelTimer::~ScopedKernelTimer()
{
    if (!on_) return;
    cudaEventRecord(e1_, stream_);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof.push_back(TimedLaunch{kind_, e0_, e1_});
}

void
profile_begin()
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (auto& t : g_prof) {
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    g_prof.clear();
    g_prof_on = true;
}

int
profile_end(double* ms, uint64_t* launches, int cap)
{
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_on = false;
    for (int i = 0; i < cap; ++i) {
        ms[i] = 0.0;
        launches[i] = 0;
    }
    int rc = 0;
    for (auto& t : g_prof) {
        float f = 0.f;
        cudaError_t e = cudaEventSynchronize(t.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&f, t.e0, t.e1);
        if (e != cudaSuccess) rc = (int)e;
        if (t.kind >= 0 && t.kind < cap) {
            ms[t.kind] += f;
            launches[t.kind] += 1;
        }
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    g_prof.clear();
    return rc;
}

// ---------------------------------------------------------------------------------------
// operand addressing: scalar / contiguous / broadcast-strided
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void
load_hz_strided(const FillArgs& a, uint64_t g, double& h, double& z)
{
    int64_t oh = 0, oz = 0;
    uint64_t rem = g;
#pragma unroll 1
    for (int d = a.bc.ndim - 1; d >= 0; --d) {
        uint64_t q = rem / (uint64_t)a.bc.shape[d];
        int64_t i = (int64_t)(rem - q * (uint64_t)a.bc.shape[d]);
        oh += i * a.bc.hs[d];
        oz += i * a.bc.zs[d];
        rem = q;
    }
    h = a.h[oh];
    z = a.z[oz];
}

__device__ __forceinline__ void
load_hz(const FillArgs& a, uint64_t g, double& h, double& z)
{
    if (a.bc.ndim == 0) {
        h = a.h ? a.h[g * a.h_stride] : a.h_scalar;
        z = a.z ? a.z[g * a.z_stride] : a.z_scalar;
        return;
    }
    load_hz_strided(a, g, h, z);
}

// nan_or_inf (src/pgm_random.c:137-141); true if h is not a valid shape
__host__ __device__ __forceinline__ bool
invalid_h(double h, double& value)
{
    if (!(h > 0.0)) {
        value = NAN;
        return true;
    }
    if (isinf(h)) {
        value = INFINITY;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// bucketing pre-pass
// ---------------------------------------------------------------------------------------
constexpr int kBucketThreads = 256;
constexpr int kBucketItems = 8;
constexpr int kBucketTile = kBucketThreads * kBucketItems;  // elements per block
constexpr uint8_t kBucketNone = 255;

// bucket ids (also the order of the per-method lists inside the chunk's index list)
// (the gamma method never needs the pre-pass: it runs on the identity range)
enum : int { kBDevroye = 0, kBAlternate = 1, kBSaddleFull = 2, kBSaddleLeft = 3, kBSaddleFar = 4, kBAlternateLeft = 5,
             kBSaddleMid = 6, kBNormal = 7, kBDevroyeMSH = 8, kNumBuckets = 9, kBGamma = 9 };

// ctl layout (uint32): [0..9] bucket counts, [10..19] record bases, [20..29] work counters, [30] ticket
constexpr int kCtlCount = 0, kCtlBase = 10, kCtlWork = 20, kCtlTicket = 30, kCtlWords = 32;
static_assert(kNumBuckets <= 10, "ctl layout");

__host__ __device__ __forceinline__ int
bucket_of(int method, double h, double z, bool compat)
{
    switch (method) {
        // |z|/2 < 1/T: exponential-pair proposal, else Michael-Schucany-Haas (z = NaN is sampled as z = 0)
        case kDevroye: return (0.5 * fabs(z) >= 1.5625) ? kBDevroyeMSH : kBDevroye;
        case kAlternate: return alternate_left_only(z) ? kBAlternateLeft : kBAlternate;
        case kSaddle: {
            // the far kernel's buckets: |z| >= 14 (proposals below x = 0.2 almost surely: closed-form
            // solve for every lane) and 8 <= |z| < 14 (a mix of closed-form and iterative solves)
            int sm = saddle_mode(h, z, compat);
            return sm == kSaddleFull ? kBSaddleFull : sm == kSaddleLeft ? kBSaddleLeft
                   : fabs(z) >= 14.0 ? kBSaddleFar : kBSaddleMid;
        }
        case kNormal: return kBNormal;
        default: return kBGamma;
    }
}

// Layout of the pre-pass: a block owns kTilesPerBlock consecutive tiles of kBucketTile elements;
// thread t of a tile owns the 8 CONSECUTIVE elements tile*2048 + t*8 .. +7, so h and z are read
// with 16-byte loads and the 8 bucket ids are stored as one 8-byte word.  Per-bucket counts travel
// packed two to a 32-bit word (16 bits each; a block holds kTilesPerBlock * 2048 <= 8192 elements).
// Measured on cfg4 (ms per 1e9 elements): 1 tile 5.13, 2 tiles 4.64, 4 tiles 4.87, 8 tiles 8.9.
constexpr int kTilesPerBlock = 2;
constexpr int kPacked = (kNumBuckets + 1) / 2;

__device__ __forceinline__ void
packed_add(uint32_t (&pc)[kPacked], int bucket)
{
#pragma unroll
    for (int w = 0; w < kPacked; ++w) {
        pc[w] += (bucket == 2 * w) ? 1u : (bucket == 2 * w + 1) ? 0x10000u : 0u;
    }
}

__device__ __forceinline__ uint32_t
packed_get(const uint32_t (&pc)[kPacked], int q)
{
    return (q & 1) ? (pc[q >> 1] >> 16) : (pc[q >> 1] & 0xffffu);
}

// One pass builds the per-method element lists.  Phase 1: every thread classifies its 8 elements of
// each of the block's tiles (all loads of the block are in flight together) and stores the 8 bucket
// ids as one 64-bit word in shared memory.  The block's per-bucket totals then reserve a
// range of each bucket's list with ONE atomicAdd per bucket and block (ctl[kCtlCount + q] ends up as
// the bucket's length).  Phase 2 re-reads the ids from shared memory in striped order and writes the
// element indices.  Bucket q's list is list + q * list_stride; the order of the entries inside a
// block's range and of the blocks' ranges depends on scheduling, which changes which lane draws an
// element, never what is drawn (private Philox stream per element).
// The last block to finish (ticket counter, no spinning) turns the totals into the record bases.
// `method` is the requested sampler (kHybrid: src/pgm_random.c:105-121 decides per element).
// VEC: h and z are contiguous, 16-byte aligned arrays (or scalars).
template <bool VEC>
__global__ void __launch_bounds__(kBucketThreads)
bucket_kernel(FillArgs a, uint32_t n, int method, uint32_t* __restrict__ ctl, uint32_t* __restrict__ list,
              uint32_t list_stride)
{
    __shared__ uint32_t s_cnt[kPacked];
    __shared__ uint32_t s_run[kNumBuckets];  // next free list slot of each bucket for this block
    __shared__ __align__(8) uint8_t s_meth[kTilesPerBlock][kBucketTile];
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x < kPacked) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const bool compat = a.compat != 0;
    uint32_t pc[kPacked];
#pragma unroll
    for (int w = 0; w < kPacked; ++w) pc[w] = 0;
#pragma unroll
    for (int tile = 0; tile < kTilesPerBlock; ++tile) {
        const uint32_t e0 = (blockIdx.x * kTilesPerBlock + tile) * kBucketTile + threadIdx.x * kBucketItems;
        uint64_t word = ~0ull;  // kBucketNone in every slot
        if (e0 < n) {
            const bool full = (e0 + kBucketItems <= n);
            auto classify_one = [&](int k, double h, double z) {
                uint32_t b = kBucketNone;
                double bad;
                if (invalid_h(h, bad)) {
                    a.out[a.base + e0 + k] = bad;
                }
                else {
                    int m = (method == kHybrid) ? hybrid_select(h, z) : method;
                    b = (uint32_t)bucket_of(m, h, z, compat);
                    packed_add(pc, (int)b);
                }
                word = (word & ~(0xffull << (8 * k))) | ((uint64_t)b << (8 * k));
            };
            if (VEC && full) {
                double hv[kBucketItems], zv[kBucketItems];
                const uint64_t g0 = a.base + e0;
                if (a.h) {
                    const double2* p = reinterpret_cast<const double2*>(a.h + g0);
#pragma unroll
                    for (int k = 0; k < kBucketItems / 2; ++k) {
                        double2 v = __ldg(p + k);
                        hv[2 * k] = v.x;
                        hv[2 * k + 1] = v.y;
                    }
                }
                else {
#pragma unroll
                    for (int k = 0; k < kBucketItems; ++k) hv[k] = a.h_scalar;
                }
                if (a.z) {
                    const double2* p = reinterpret_cast<const double2*>(a.z + g0);
#pragma unroll
                    for (int k = 0; k < kBucketItems / 2; ++k) {
                        double2 v = __ldg(p + k);
                        zv[2 * k] = v.x;
                        zv[2 * k + 1] = v.y;
                    }
                }
                else {
#pragma unroll
                    for (int k = 0; k < kBucketItems; ++k) zv[k] = a.z_scalar;
                }
#pragma unroll
                for (int k = 0; k < kBucketItems; ++k) classify_one(k, hv[k], zv[k]);
            }
            else {
#pragma unroll 1
                for (int k = 0; k < kBucketItems; ++k) {
                    if (e0 + k < n) {
                        double h, z;
                        load_hz(a, a.base + e0 + k, h, z);
                        classify_one(k, h, z);
                    }
                }
            }
        }
        *reinterpret_cast<uint64_t*>(&s_meth[tile][threadIdx.x * kBucketItems]) = word;
    }
    // block totals -> one reservation per bucket
#pragma unroll
    for (int w = 0; w < kPacked; ++w) {
        // a thread holds at most 8 * kTilesPerBlock <= 32 elements, a warp 1024: the packed 16-bit sums
        // cannot carry
        for (int off = 16; off > 0; off >>= 1) pc[w] += __shfl_xor_sync(0xffffffffu, pc[w], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int w = 0; w < kPacked; ++w) atomicAdd(&s_cnt[w], pc[w]);
    }
    __syncthreads();
    if (threadIdx.x < kNumBuckets) {
        uint32_t v = s_cnt[threadIdx.x >> 1];
        v = (threadIdx.x & 1) ? (v >> 16) : (v & 0xffffu);
        const uint32_t base = v ? atomicAdd(&ctl[kCtlCount + threadIdx.x], v) : 0u;
        s_run[threadIdx.x] = threadIdx.x * list_stride + base;
    }
    __syncthreads();  // s_run and s_meth are visible
    // Phase 2, striped: lane t of stripe j takes element t0 + 256 j + t, so the lanes of a warp that
    // share a bucket are consecutive elements and get consecutive list slots (coalesced stores).
    // The lanes of each bucket are found with one MATCH.ANY on the bucket id; the lowest lane of a
    // group reserves the group's slots from the block's running counter.
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int tile = 0; tile < kTilesPerBlock; ++tile) {
        const uint32_t t0 = (blockIdx.x * kTilesPerBlock + tile) * kBucketTile;
        if (t0 >= n) break;  // block-uniform
#pragma unroll
        for (int j = 0; j < kBucketItems; ++j) {
            const uint32_t i = j * kBucketThreads + threadIdx.x;
            const uint32_t q = s_meth[tile][i];
            const bool valid = q < (uint32_t)kNumBuckets;
            const uint32_t same = __match_any_sync(0xffffffffu, q);
            const int leader = __ffs(same) - 1;
            uint32_t slot = 0;
            if (valid && (int)lane == leader) slot = atomicAdd(&s_run[q], (uint32_t)__popc(same));
            slot = __shfl_sync(0xffffffffu, slot, leader < 0 ? 0 : leader);
            if (valid) list[slot + __popc(same & lt)] = t0 + i;
        }
    }
    // the last block turns the bucket totals into the bases of the packed record regions
    if (threadIdx.x == 0) {
        __threadfence();
        const uint32_t ticket = atomicAdd(&ctl[kCtlTicket], 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            uint32_t base = 0;
            for (int q = 0; q < kNumBuckets; ++q) {
                ctl[kCtlBase + q] = base;
                base += atomicAdd(&ctl[kCtlCount + q], 0u);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// work description shared by the per-method kernels
// ---------------------------------------------------------------------------------------
// A bucket is either the identity range [0, count_imm) of the chunk (ctl == nullptr) or the
// bucket-th list (list + bucket * list_stride), whose length and record base live in ctl on the
// device.
struct Bucket {
    const uint32_t* list;
    const uint32_t* ctl;
    int bucket;
    uint32_t count_imm;
    uint32_t list_stride;  // bucket q's list starts at list + q * list_stride
    double* params;  // record storage for this chunk (kMaxRecordFields doubles per element)
    int uniform;     // 1: one record (position 0) serves every element (scalar h and z)
};

struct BucketView {
    const uint32_t* list;
    uint32_t count;
    double* recs;     // field f of position p at recs[f * stride + p]
    uint32_t stride;
};

__device__ __forceinline__ BucketView
view_of(const Bucket& b)
{
    BucketView v;
    v.list = b.list;
    v.count = b.count_imm;
    v.recs = b.params;
    if (b.ctl) {
        uint32_t base = b.ctl[kCtlBase + b.bucket];
        v.count = b.ctl[kCtlCount + b.bucket];
        v.list = b.list + (size_t)b.bucket * b.list_stride;
        v.recs = b.params + (size_t)base * kMaxRecordFields;
    }
    v.stride = b.uniform ? 1u : v.count;
    return v;
}

struct RecView {
    const double* base;
    uint32_t stride, pos;
    __device__ __forceinline__ double operator[](int f) const { return base[(size_t)f * stride + pos]; }
};

// ---------------------------------------------------------------------------------------
// setup kernel: one thread per element, no rejection loops
// ---------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(256)
setup_kernel(FillArgs a, Bucket b)
{
    const BucketView v = view_of(b);
    const uint32_t nrec = b.uniform ? 1u : v.count;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < nrec; pos += gridDim.x * blockDim.x) {
        uint32_t e = v.list ? v.list[pos] : pos;
        uint64_t g = a.base + e;
        double h, z, bad;
        load_hz(a, g, h, z);
        double rec[M::kFields];
        if (invalid_h(h, bad)) {
            // only reachable on the identity range (bucket_kernel filters invalid h itself)
            if (!b.uniform) a.out[g] = bad;
#pragma unroll
            for (int f = 0; f < M::kFields; ++f) rec[f] = NAN;
        }
        else {
            M::make_record(h, z, a.compat != 0, rec);
        }
#pragma unroll
        for (int f = 0; f < M::kFields; ++f) v.recs[(size_t)f * v.stride + pos] = rec[f];
    }
}

// ---------------------------------------------------------------------------------------
// persistent rejection-sampling kernel
// ---------------------------------------------------------------------------------------
constexpr int kPersistThreads = 256;
constexpr uint32_t kWorkBatch = 64;  // elements a warp reserves per atomic

// FUSED: the lane computes its element's record itself at refill (registers only) instead of reading
// what setup_kernel stored -- for a method whose setup is short and whose lanes all refill on nearly
// every trip (saddle-far: 1.004 proposals per sample), this saves the record's round trip through
// HBM and the second gather of (h, z).
struct LocalRec {
    const double* v;
    __device__ __forceinline__ double operator[](int f) const { return v[f]; }
};

template <class M, bool FUSED = false>
__global__ void __launch_bounds__(kPersistThreads)
sample_kernel(FillArgs a, Bucket b, uint32_t* __restrict__ counter)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const BucketView v = view_of(b);
    const uint32_t count = v.count;
    const bool compat = a.compat != 0;

    typename M::Par par;
    typename M::State st;
    typename M::Rng rng;
    uint64_t g = 0;
    bool active = false;
    uint32_t wnext = 0, wend = 0;  // warp-uniform reserved range of list positions
    bool exhausted = (count == 0);

    for (;;) {
        uint32_t need = __ballot_sync(0xffffffffu, !active);
        if (need != 0 && !(exhausted && wnext == wend)) {
            if (wnext == wend) {
                uint32_t w = 0;
                if (lane == 0) w = atomicAdd(counter, kWorkBatch);
                w = __shfl_sync(0xffffffffu, w, 0);
                if (w >= count) {
                    exhausted = true;
                }
                else {
                    wnext = w;
                    wend = min(w + kWorkBatch, count);
                }
            }
            if (wnext < wend) {
                uint32_t pos = wnext + __popc(need & lt);
                if (!active && pos < wend) {
                    uint32_t e = v.list ? v.list[pos] : pos;
                    g = a.base + e;
                    if (FUSED) {  // (bucketed lists only: invalid h never reaches a list)
                        double h, z, loc[M::kFields];
                        load_hz(a, g, h, z);
                        M::make_record(h, z, compat, loc);
                        if (M::load(LocalRec{loc}, par, st, compat)) {
                            rng.init(a.key, a.first_index + g);
                            active = true;
                        }
                        else {
                            a.out[g] = 0.0;
                        }
                    }
                    else {
                        RecView rec{v.recs, v.stride, b.uniform ? 0u : pos};
                        double r0 = rec[0];
                        if (r0 == r0) {  // NaN record: invalid h, already written by setup_kernel
                            if (M::load(rec, par, st, compat)) {
                                rng.init(a.key, a.first_index + g);
                                active = true;
                            }
                            else {
                                a.out[g] = 0.0;  // nothing to draw (devroye with floor(h) == 0)
                            }
                        }
                    }
                }
                wnext = min(wnext + (uint32_t)__popc(need), wend);
            }
        }
        if (active) {
            if (M::step(par, st, rng, compat)) {
                a.out[g] = st.acc;
                active = false;
            }
        }
        if (exhausted && wnext == wend && !__any_sync(0xffffffffu, active)) break;
    }
}

// ---------------------------------------------------------------------------------------
// dense kernels: one pass, no outer rejection loop (normal approximation, gamma convolution)
// ---------------------------------------------------------------------------------------
#ifndef PGM_NORMAL_PER_TRIP
#define PGM_NORMAL_PER_TRIP 2
#endif
constexpr int kNormalPerTrip = PGM_NORMAL_PER_TRIP;

template <int KIND>
__global__ void __launch_bounds__(256)
dense_kernel(FillArgs a, Bucket b)
{
    const BucketView v = view_of(b);
    const uint32_t stride = gridDim.x * blockDim.x;
    if (KIND == kNormal) {
        // kNormalPerTrip elements per trip (their dependency chains interleave and the constants of
        // the polynomial kernels are fetched once per trip), software-pipelined two deep: the list
        // indices of trip t + 2 and the (h, z) gathers of trip t + 1 are issued before trip t's
        // arithmetic, so the two dependent memory round trips never sit on the critical path.
        // Positions past the end are clamped (valid loads, results unused).
        if (v.count == 0) return;
        constexpr int E = kNormalPerTrip;
        const uint32_t last = v.count - 1;
        auto entry = [&](uint32_t p) -> uint32_t {
            p = min(p, last);
            return v.list ? v.list[p] : p;
        };
        uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t e[E], nx[E];
        double h[E], z[E];
#pragma unroll
        for (int i = 0; i < E; ++i) {
            e[i] = entry(pos + i * stride);
            nx[i] = entry(pos + (E + i) * stride);
        }
#pragma unroll
        for (int i = 0; i < E; ++i) load_hz(a, a.base + e[i], h[i], z[i]);
        for (; pos < v.count; pos += E * stride) {
            uint32_t nn[E];
            double hn[E], zn[E];
#pragma unroll
            for (int i = 0; i < E; ++i) nn[i] = entry(pos + (2 * E + i) * stride);
#pragma unroll
            for (int i = 0; i < E; ++i) load_hz(a, a.base + nx[i], hn[i], zn[i]);
            double out[E];
#pragma unroll
            for (int i = 0; i < E; ++i) {
                double bad;
                RngInlineRK r;
                r.init(a.key, a.first_index + a.base + e[i]);
                out[i] = invalid_h(h[i], bad) ? bad : normal_approx_sample(r, h[i], z[i]);
            }
#pragma unroll
            for (int i = 0; i < E; ++i) {
                if (pos + i * stride < v.count) a.out[a.base + e[i]] = out[i];
            }
#pragma unroll
            for (int i = 0; i < E; ++i) {
                e[i] = nx[i], nx[i] = nn[i];
                h[i] = hn[i], z[i] = zn[i];
            }
        }
        return;
    }
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < v.count; pos += stride) {
        uint32_t e = v.list ? v.list[pos] : pos;
        uint64_t g = a.base + e;
        double h, z, bad;
        load_hz(a, g, h, z);
        if (invalid_h(h, bad)) {
            a.out[g] = bad;
            continue;
        }
        GammaRng rng;
        rng.init(a.key, a.first_index + g);
        a.out[g] = gamma_conv_sample(rng, h, z);
    }
}

// ---------------------------------------------------------------------------------------
// small fills: ONE launch, one thread per element
// ---------------------------------------------------------------------------------------
// The bucketed pipeline costs one bucketing launch, up to fifteen method launches (sized for the whole
// pass, since the bucket sizes stay on the device), a scratch allocation and stream joins: 100-300 us
// per call whatever the size.  A Gibbs sampler calls with 1e3..1e5 elements per sweep, where that
// overhead is the whole cost.  Below kSmallMax elements every thread takes one element through the
// SAME per-bucket code the pipeline would run (same record, same step function, same Philox stream: the
// result is bit-identical), to completion: divergent, but one launch and no scratch.
template <class M>
__device__ __noinline__ double
run_one(uint32_t k0, uint32_t k1, uint64_t index, double h, double z, bool compat)
{
    double rec[M::kFields];
    M::make_record(h, z, compat, rec);
    typename M::Par par;
    typename M::State st;
    if (!M::load(LocalRec{rec}, par, st, compat)) return 0.0;  // devroye with floor(h) == 0
    PhiloxKey key;  // (the samplers' generators use k0, k1 only; the round keys serve the dense normal kernel)
    key.k0 = k0;
    key.k1 = k1;
    typename M::Rng rng;
    rng.init(key, index);
    while (!M::step(par, st, rng, compat)) {
    }
    return st.acc;
}

__global__ void __launch_bounds__(128)
small_kernel(FillArgs a, uint32_t n, int method)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const bool compat = a.compat != 0;
    const uint64_t g = a.base + e;
    double h, z, v;
    load_hz(a, g, h, z);
    if (!invalid_h(h, v)) {
        const int m = (method == kHybrid) ? hybrid_select(h, z) : method;
        const uint32_t k0 = a.key.k0, k1 = a.key.k1;
        const uint64_t index = a.first_index + g;
        switch (bucket_of(m, h, z, compat)) {
            case kBDevroye: v = run_one<DevroyeT<kDevroyePair>>(k0, k1, index, h, z, compat); break;
            case kBDevroyeMSH: v = run_one<DevroyeT<kDevroyeMSH>>(k0, k1, index, h, z, compat); break;
            case kBAlternate: v = run_one<Alternate>(k0, k1, index, h, z, compat); break;
            case kBAlternateLeft: v = run_one<AlternateLeft>(k0, k1, index, h, z, compat); break;
            case kBSaddleFull: v = run_one<SaddleFull>(k0, k1, index, h, z, compat); break;
            case kBSaddleLeft: v = run_one<SaddleLeft>(k0, k1, index, h, z, compat); break;
            case kBSaddleFar:
            case kBSaddleMid: v = run_one<SaddleFar>(k0, k1, index, h, z, compat); break;
            default: {  // kBNormal
                RngInlineRK r;
                r.init(a.key, a.first_index + g);
                v = normal_approx_sample(r, h, z);
            }
        }
    }
    a.out[g] = v;
}

// fills g_guess_uc / g_guess_ur (one thread per table entry)
__global__ void
init_guess_tables_kernel()
{
    using namespace saddle;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kGuessN) return;
    double zz = (double)i / kGuessPerUnit;
    double xl = zz > 0.0 ? tanh(zz) / zz : 1.0;
    double f, fp;
    g_guess_uc[i] = solve(2.75 * xl, table_guess(2.75 * xl), f, fp);
    g_guess_ur[i] = solve(3.0 * xl, table_guess(3.0 * xl), f, fp);
}

__global__ void
fill_value_kernel(double* __restrict__ out, uint64_t n, double value)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        out[i] = value;
    }
}

// any_nonpositive (polyagamma/_polyagamma.pyx:411-430)
__global__ void
check_h_kernel(const double* __restrict__ h, size_t n, int devroye, int* __restrict__ flag)
{
    bool bad = false;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double v = h[i];
        if (v <= 1e-4 || (devroye && v != floor(v))) bad = true;
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// ---------------------------------------------------------------------------------------
// host side: launch plan
// ---------------------------------------------------------------------------------------
struct DeviceInfo {
    int sms = 0;
    // sampling kernels: devroye, alternate, saddle-full, saddle-left, saddle-far, alternate-left,
    // saddle-mid (the far kernel on its second bucket)
    int blocks_per_sm[7] = {0, 0, 0, 0, 0, 0, 0};
};

static DeviceInfo
device_info()
{
    static DeviceInfo cache[64];
    static std::mutex mu;  // first calls from several host threads: one of them fills the entry, the others wait
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    DeviceInfo& d = cache[dev & 63];
    if (d.sms == 0) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[0], sample_kernel<DevroyeT<kDevroyePair>>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[1], sample_kernel<Alternate>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[2], sample_kernel<SaddleFull>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[3], sample_kernel<SaddleLeft>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[4], sample_kernel<SaddleFar>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[5], sample_kernel<AlternateLeft>, kPersistThreads, 0);
        d.blocks_per_sm[6] = d.blocks_per_sm[4];
        for (int i = 0; i < 7; ++i) {
            if (d.blocks_per_sm[i] < 1) d.blocks_per_sm[i] = 1;
        }
        // starting-point tables of the saddle setup (per device, filled by the solver itself)
        init_guess_tables_kernel<<<(saddle::kGuessN + 127) / 128, 128>>>();
        cudaDeviceSynchronize();
        count_launch(1);
        // stream-ordered scratch: keep freed blocks cached in the pool instead of returning
        // them to the OS at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    return d;
}

// helper streams per device: 0 for the overlapped bucketing pre-pass, 1 and 2 for the small
// buckets' kernels (they run beside the big buckets' kernels and fill their tails)
static cudaStream_t
helper_stream(int which)
{
    static cudaStream_t streams[64][3] = {};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    cudaStream_t& st = streams[dev & 63][which];
    if (st == nullptr && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) st = nullptr;
    return st;
}

static inline uint32_t
grid_for(uint64_t full, uint64_t count_hint, uint32_t threads)
{
    uint64_t need = (count_hint + threads - 1) / threads;
    if (need < 1) need = 1;
    return (uint32_t)(need < full ? need : full);
}

// setup + persistent sampling of one bucket with method M (which: 0 devroye, 1 alternate,
// 2 saddle-full, 3 saddle-left, 4 saddle-far, 5 alternate-left, 6 saddle-mid; profiling kinds are
// kKindDevroye / kKindSetup + which)
// Which buckets compute their records inside the sampling kernel (measured on cfg4, ms per 1e9,
// setup + sampling separate -> fused): saddle-far 8.25 -> 6.86, saddle-mid 2.71 -> 2.44, alternate-left
// 2.04 -> 1.55, devroye (cfg3, per 1e8) 5.23 -> 5.04; saddle-left gets slower (3.83 -> 4.22: its setup is
// longer and its lanes refill out of step), and the full saddle / alternate setups are the big ones the
// split exists for.
constexpr bool kFuseFar = true, kFuseMid = true, kFuseLeft = false, kFuseAltLeft = true;
constexpr bool kFuseDevroye = true;  // cfg3 devroye 19.1 -> 19.8e9 with the exponential-free setup (slower with the erfc one)

template <class M, bool FUSED = false>
static void
launch_method(const DeviceInfo& d, int which, const FillArgs& a, const Bucket& b, uint64_t count_hint,
              uint32_t* counter, cudaStream_t s)
{
    const int kind = kKindDevroye + which;
    if (!FUSED) {
        ScopedKernelTimer timer(kKindSetup + which, s);
        uint64_t nrec = b.uniform ? 1 : count_hint;
        setup_kernel<M><<<grid_for((uint64_t)d.sms * 8, nrec, 256), 256, 0, s>>>(a, b);
    }
    {
        ScopedKernelTimer timer(kind, s);
        sample_kernel<M, FUSED><<<grid_for((uint64_t)d.sms * d.blocks_per_sm[which], count_hint, kPersistThreads),
                                  kPersistThreads, 0, s>>>(a, b, counter);
    }
    count_launch(FUSED ? 1 : 2);
}

template <int KIND>
static void
launch_dense(const DeviceInfo& d, const FillArgs& a, const Bucket& b, uint64_t count_hint, cudaStream_t s)
{
    ScopedKernelTimer timer(KIND == kNormal ? kKindNormal : kKindGamma, s);
    dense_kernel<KIND><<<grid_for((uint64_t)d.sms * 8, count_hint, 256), 256, 0, s>>>(a, b);
    count_launch(1);
}

// Elements per pass through the pipeline.  Bounds the scratch (5 B + one record per element)
// and keeps list indices in 32 bits.
// Pass size.  Every kernel of a pass ends with a partially filled last wave, so fewer, larger
// passes are faster (cfg4 per 1e9 elements: 2^24 66.8 ms, 2^25 55.5, 2^26 50.7, 2^27 48.5 ms);
// the per-element path needs 5 B + one record of scratch per element of a pass.  It starts at
// 2^27 elements (or the smallest power of two >= n) and halves the pass while the stream-ordered
// allocation of the scratch fails.  PGM_CUDA_CHUNK_LOG2 in the environment pins the pass size
// (tests use it to cross pass seams).
constexpr int kChunkLog2Min = 20, kChunkLog2Max = 27;
constexpr uint64_t kSmallMaxDefault = 1ull << 16;  // mixed-method fills break even near 1e5, single-method ones far later

static int
env_chunk_log2()
{
    const char* e = getenv("PGM_CUDA_CHUNK_LOG2");
    if (e == nullptr) return 0;
    int v = atoi(e);
    return (v >= 10 && v <= 28) ? v : 0;  // 9 lists x 2^28 entries still index with 32 bits
}

static int
first_chunk_log2(uint64_t n)
{
    if (int forced = env_chunk_log2()) return forced;
    int lg = kChunkLog2Max;
    while (lg > kChunkLog2Min && (1ull << (lg - 1)) >= n) --lg;
    return lg;
}

static inline size_t
align256(size_t v)
{
    return (v + 255) & ~size_t(255);
}

// largest fill that takes the one-launch path (PGM_CUDA_SMALL_MAX overrides; 0 disables it)
static uint64_t
small_max()
{
    const char* e = getenv("PGM_CUDA_SMALL_MAX");  // read per call: the tests switch it
    return e ? (uint64_t)strtoull(e, nullptr, 10) : kSmallMaxDefault;
}

// smallest pass whose buckets are spread over the helper streams (measured: 1e5 mixed elements 251 ->
// 166 us per call, 1e6: 322 -> 210 us; below 2^16 the one-launch kernel runs instead)
static uint64_t
multi_min()
{
    const char* e = getenv("PGM_CUDA_MULTI_MIN_LOG2");
    return 1ull << (e ? atoi(e) : 16);
}

int
launch_fill(FillArgs a, uint64_t n, int method, cudaStream_t s)
{
    if (n == 0) return 0;
    const int m = method & 0xff;
    a.compat = (method & 0x100) ? 1 : 0;
    if (m < kGamma || m > kHybrid) return -1;
    const bool compat = a.compat != 0;
    DeviceInfo d = device_info();
    cudaError_t err;

    // with ONE shape for the whole call the elements fall into two or three buckets at most, a warp of
    // the one-launch kernel serialises correspondingly few code paths, and it wins up to 4x further
    bool one_shape = (a.h == nullptr || a.h_stride == 0);
    if (a.bc.ndim != 0) {
        one_shape = true;
        for (int dd = 0; dd < a.bc.ndim; ++dd) one_shape = one_shape && a.bc.hs[dd] == 0;
    }
    if (m != kGamma && n <= small_max() * (one_shape ? 4 : 1)) {
        ScopedKernelTimer timer(kKindSmall, s);
        small_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, s>>>(a, (uint32_t)n, m);
        count_launch(1);
        return (int)cudaGetLastError();
    }

    const bool scalar_params = (a.h == nullptr && a.z == nullptr && a.bc.ndim == 0);
    if (scalar_params) {
        // pgm_random_polyagamma_fill (src/pgm_random.c:158-170): one (h, z) for all n elements
        double bad;
        if (invalid_h(a.h_scalar, bad)) {
            fill_value_kernel<<<grid_for((uint64_t)d.sms * 8, n, 256), 256, 0, s>>>(a.out + a.base, n, bad);
            count_launch(1);
            return (int)cudaGetLastError();
        }
        const int eff = (m == kHybrid) ? hybrid_select(a.h_scalar, a.z_scalar) : m;
        const int bk = bucket_of(eff, a.h_scalar, a.z_scalar, compat);
        const uint64_t kChunk = 1ull << first_chunk_log2(n);
        const uint64_t nchunks = (n + kChunk - 1) / kChunk;
        char* ws = nullptr;
        const size_t rec_bytes = align256(kMaxRecordFields * sizeof(double));
        if ((err = cudaMallocAsync(&ws, rec_bytes + nchunks * sizeof(uint32_t), s)) != cudaSuccess) return (int)err;
        double* params = (double*)ws;
        uint32_t* counters = (uint32_t*)(ws + rec_bytes);
        cudaMemsetAsync(counters, 0, nchunks * sizeof(uint32_t), s);
        const uint64_t base0 = a.base;
        for (uint64_t c = 0; c < nchunks; ++c) {
            const uint64_t off = c * kChunk;
            const uint32_t cn = (uint32_t)((n - off) < kChunk ? (n - off) : kChunk);
            a.base = base0 + off;
            Bucket b{nullptr, nullptr, bk, cn, 0, params, 1};
            switch (bk) {
                case kBDevroye: launch_method<DevroyeT<kDevroyePair>>(d, 0, a, b, cn, counters + c, s); break;
                case kBDevroyeMSH: launch_method<DevroyeT<kDevroyeMSH>>(d, 0, a, b, cn, counters + c, s); break;
                case kBAlternate: launch_method<Alternate>(d, 1, a, b, cn, counters + c, s); break;
                case kBSaddleFull: launch_method<SaddleFull>(d, 2, a, b, cn, counters + c, s); break;
                case kBSaddleLeft: launch_method<SaddleLeft>(d, 3, a, b, cn, counters + c, s); break;
                case kBSaddleFar: launch_method<SaddleFar>(d, 4, a, b, cn, counters + c, s); break;
                case kBAlternateLeft: launch_method<AlternateLeft>(d, 5, a, b, cn, counters + c, s); break;
                case kBSaddleMid: launch_method<SaddleFar>(d, 6, a, b, cn, counters + c, s); break;
                case kBNormal: launch_dense<kNormal>(d, a, b, cn, s); break;
                default: launch_dense<kGamma>(d, a, b, cn, s); break;
            }
        }
        cudaFreeAsync(ws, s);
        return (int)cudaGetLastError();
    }

    // per-element parameters (fill2).  hybrid, saddle, alternate and devroye go through the bucketing
    // pre-pass (they split into specialised kernels); gamma runs on the identity range.
    const bool bucketed = (m == kHybrid || m == kSaddle || m == kAlternate || m == kDevroye);
    int lg = first_chunk_log2(n);
    // With several passes the bucketing of pass k+1 (HBM- and issue-bound) runs on a helper stream
    // while the sampling kernels of pass k (FP64-bound) run on the caller's stream; the bucketing
    // scratch is double-buffered, the records are not (setup and sampling are serialised on the
    // caller's stream).
    const bool overlap = bucketed && n > (1ull << lg) && getenv("PGM_CUDA_NO_OVERLAP") == nullptr;
    const int nprep = overlap ? 2 : 1;
    uint64_t chunk = 0;
    size_t list_bytes = 0, par_bytes = 0, prep_bytes = 0;
    uint32_t list_stride = 0;
    const size_t ctl_bytes = align256(kCtlWords * sizeof(uint32_t));
    char* ws = nullptr;
    for (;;) {
        chunk = 1ull << lg;
        const uint64_t cmax = n < chunk ? n : chunk;
        // one worst-case list per bucket (the single-pass bucketing cannot know the bucket sizes
        // before it writes): 32 B of scratch per element next to the 72 B of records
        list_stride = (uint32_t)((cmax + 63) & ~uint64_t(63));
        list_bytes = bucketed ? align256((size_t)list_stride * kNumBuckets * sizeof(uint32_t)) : 0;
        par_bytes = (m == kGamma) ? 0 : align256(cmax * kMaxRecordFields * sizeof(double));
        prep_bytes = list_bytes + ctl_bytes;
        err = cudaMallocAsync(&ws, nprep * prep_bytes + par_bytes, s);
        if (err == cudaSuccess) break;
        (void)cudaGetLastError();  // clear the sticky-free allocation error and retry smaller
        if (err != cudaErrorMemoryAllocation || lg <= kChunkLog2Min || env_chunk_log2()) return (int)err;
        --lg;
    }
    struct Prep {
        uint32_t *list, *ctl;
    } prep[2];
    for (int i = 0; i < nprep; ++i) {
        char* base = ws + (size_t)i * prep_bytes;
        prep[i].list = (uint32_t*)base;
        prep[i].ctl = (uint32_t*)(base + list_bytes);
    }
    if (nprep == 1) prep[1] = prep[0];
    double* params = (double*)(ws + (size_t)nprep * prep_bytes);
    const uint64_t cmax_elems = n < chunk ? n : chunk;

    // 16-byte loads in the pre-pass need contiguous, aligned operands (a scalar operand is fine)
    const bool vec = a.bc.ndim == 0 && (a.h == nullptr || (a.h_stride == 1 && ((uintptr_t)(a.h + a.base) & 15) == 0)) &&
                     (a.z == nullptr || (a.z_stride == 1 && ((uintptr_t)(a.z + a.base) & 15) == 0));
    const uint64_t base0 = a.base;

    if (!bucketed) {
        uint32_t* ctl = prep[0].ctl;
        for (uint64_t off = 0; off < n; off += chunk) {
            const uint32_t cn = (uint32_t)((n - off) < chunk ? (n - off) : chunk);
            a.base = base0 + off;
            cudaMemsetAsync(ctl, 0, ctl_bytes, s);
            Bucket b{nullptr, nullptr, 0, cn, 0, params, 0};
            launch_dense<kGamma>(d, a, b, cn, s);
        }
        cudaFreeAsync(ws, s);
        return (int)cudaGetLastError();
    }

    // bucketing pre-pass of one pass, on stream `sb`, into scratch set `pp`
    auto bucketing = [&](const FillArgs& ap, uint32_t cn, const Prep& pp, cudaStream_t sb) {
        const uint32_t nblocks = (cn + kBucketTile * kTilesPerBlock - 1) / (kBucketTile * kTilesPerBlock);
        cudaMemsetAsync(pp.ctl, 0, ctl_bytes, sb);
        {
            ScopedKernelTimer timer(kKindClassify, sb);
            if (vec) bucket_kernel<true><<<nblocks, kBucketThreads, 0, sb>>>(ap, cn, m, pp.ctl, pp.list, list_stride);
            else bucket_kernel<false><<<nblocks, kBucketThreads, 0, sb>>>(ap, cn, m, pp.ctl, pp.list, list_stride);
        }
        count_launch(1);
    };
    // method kernels of one pass, on the caller's stream.  Counts and list bases stay on the
    // device: the kernels read them there, so the host never waits for the histogram (empty
    // buckets cost one near-empty launch).
    // The buckets of a pass are independent (disjoint elements, disjoint record regions), so the
    // small ones run on two helper streams beside the big ones: their kernels, which are mostly
    // launch tail, fill the tails of the others.  s1/s2 == s serialises everything on `s`.
    auto sampling = [&](const FillArgs& ap, uint32_t cn, const Prep& pp, cudaStream_t s1, cudaStream_t s2) {
        uint32_t* ctl = pp.ctl;
        auto bucket = [&](int q) { return Bucket{pp.list, ctl, q, 0, list_stride, params, 0}; };
        if (m == kHybrid || m == kSaddle) {
            launch_method<SaddleFar, kFuseFar>(d, 4, ap, bucket(kBSaddleFar), cn, ctl + kCtlWork + kBSaddleFar, s);
            launch_method<SaddleFar, kFuseMid>(d, 6, ap, bucket(kBSaddleMid), cn, ctl + kCtlWork + kBSaddleMid, s1);
            launch_method<SaddleLeft, kFuseLeft>(d, 3, ap, bucket(kBSaddleLeft), cn, ctl + kCtlWork + kBSaddleLeft, s1);
            launch_method<SaddleFull>(d, 2, ap, bucket(kBSaddleFull), cn, ctl + kCtlWork + kBSaddleFull, s2);
        }
        if (m == kHybrid || m == kAlternate) {
            launch_method<Alternate>(d, 1, ap, bucket(kBAlternate), cn, ctl + kCtlWork + kBAlternate, s2);
            launch_method<AlternateLeft, kFuseAltLeft>(d, 5, ap, bucket(kBAlternateLeft), cn, ctl + kCtlWork + kBAlternateLeft,
                                         m == kAlternate ? s : s1);
        }
        if (m == kHybrid || m == kDevroye) {
            launch_method<DevroyeT<kDevroyePair>, kFuseDevroye>(d, 0, ap, bucket(kBDevroye), cn, ctl + kCtlWork + kBDevroye, s2);
            launch_method<DevroyeT<kDevroyeMSH>, kFuseDevroye>(d, 0, ap, bucket(kBDevroyeMSH), cn, ctl + kCtlWork + kBDevroyeMSH,
                                                 m == kDevroye ? s : s1);
        }
        if (m == kHybrid) launch_dense<kNormal>(d, ap, bucket(kBNormal), cn, s);
    };

    const bool multi = cmax_elems >= multi_min() && getenv("PGM_CUDA_NO_OVERLAP") == nullptr;
    if (!overlap && !multi) {
        for (uint64_t off = 0; off < n; off += chunk) {
            const uint32_t cn = (uint32_t)((n - off) < chunk ? (n - off) : chunk);
            a.base = base0 + off;
            bucketing(a, cn, prep[0], s);
            sampling(a, cn, prep[0], s, s);
        }
        cudaFreeAsync(ws, s);
        return (int)cudaGetLastError();
    }

    // (the caller's stream `s` may be the null handle of the default stream: compare helpers only)
    cudaStream_t hb = overlap ? helper_stream(0) : nullptr;
    cudaStream_t h1 = multi ? helper_stream(1) : nullptr, h2 = multi ? helper_stream(2) : nullptr;
    if ((overlap && hb == nullptr) || (multi && (h1 == nullptr || h2 == nullptr))) {
        cudaFreeAsync(ws, s);
        return (int)cudaErrorUnknown;
    }
    cudaStream_t sb = overlap ? hb : s;
    cudaStream_t s1 = multi ? h1 : s, s2 = multi ? h2 : s;
    // 0 ready, 1-2 prep[buf], 3-4 used[buf], 5 go, 6-7 helper streams done.  One set per host thread
    // and device, created once: a wait captures the record that precedes it, so re-recording an event
    // in the next call cannot disturb the waits of this one, and no other thread touches the set.
    struct EventSet {
        cudaEvent_t ev[8];
        bool ready = false;
    };
    static thread_local EventSet t_events[64];
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    EventSet& es = t_events[cur_dev & 63];
    if (!es.ready) {
        for (auto& e : es.ev) {
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                cudaFreeAsync(ws, s);
                return (int)cudaGetLastError();
            }
        }
        es.ready = true;
    }
    cudaEvent_t* ev = es.ev;
    cudaEvent_t ev_ready = ev[0], *ev_prep = ev + 1, *ev_used = ev + 3, ev_go = ev[5], *ev_done = ev + 6;
    cudaEventRecord(ev_ready, s);  // inputs and the scratch allocation are ordered before this point
    if (overlap) cudaStreamWaitEvent(sb, ev_ready, 0);
    uint64_t k = 0;
    for (uint64_t off = 0; off < n; off += chunk, ++k) {
        const int buf = overlap ? (int)(k & 1) : 0;
        const uint32_t cn = (uint32_t)((n - off) < chunk ? (n - off) : chunk);
        a.base = base0 + off;
        if (overlap) {
            if (k >= 2) cudaStreamWaitEvent(sb, ev_used[buf], 0);  // pass k-2 is done with this scratch set
            bucketing(a, cn, prep[buf], sb);
            cudaEventRecord(ev_prep[buf], sb);
            cudaStreamWaitEvent(s, ev_prep[buf], 0);
        }
        else {
            bucketing(a, cn, prep[buf], s);
        }
        if (multi) {
            // `s` is past the previous pass (it joined the helpers there) and past this pass's bucketing
            cudaEventRecord(ev_go, s);
            cudaStreamWaitEvent(s1, ev_go, 0);
            cudaStreamWaitEvent(s2, ev_go, 0);
        }
        sampling(a, cn, prep[buf], s1, s2);
        if (multi) {
            cudaEventRecord(ev_done[0], s1);
            cudaEventRecord(ev_done[1], s2);
            cudaStreamWaitEvent(s, ev_done[0], 0);
            cudaStreamWaitEvent(s, ev_done[1], 0);
        }
        if (overlap) cudaEventRecord(ev_used[buf], s);
    }
    cudaFreeAsync(ws, s);
    return (int)cudaGetLastError();
}

// DFMA microbenchmark: the FP64 roofline denominator, measured on the box it is quoted on.
__global__ void __launch_bounds__(256)
dfma_peak_kernel(double* sink, int iters, double b, double c)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fma(a0, b, c);
            a1 = fma(a1, b, c);
            a2 = fma(a2, b, c);
            a3 = fma(a3, b, c);
            a4 = fma(a4, b, c);
            a5 = fma(a5, b, c);
            a6 = fma(a6, b, c);
            a7 = fma(a7, b, c);
        }
    }
    double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 12345.678) sink[0] = r;
}

int
measure_fp64_peak(double* tflops, cudaStream_t s)
{
    DeviceInfo d = device_info();
    double* sink = nullptr;
    cudaError_t e;
    if ((e = cudaMallocAsync(&sink, sizeof(double), s)) != cudaSuccess) return (int)e;
    const int iters = 4096, blocks = d.sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, s);
        dfma_peak_kernel<<<blocks, 256, 0, s>>>(sink, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    count_launch(4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFreeAsync(sink, s);
    double flops = (double)blocks * 256.0 * iters * 64.0 * 2.0;
    *tflops = flops / (best * 1e-3) / 1e12;
    return (int)cudaGetLastError();
}

// unit-test hook for the fm:: functions (pgm_cuda_test_math)
__global__ void
test_math_kernel(int which, const double* __restrict__ in, size_t n, double* __restrict__ out)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = in[i];
        double y, t;
        switch (which) {
            case 0: y = fm::rcp(x); break;
            case 1: y = fm::sqrt(x); break;
            case 2: y = fm::log(x); break;
            case 3: y = fm::exp(x); break;
            case 4: y = fm::cos2pi(x); break;
            case 5: y = fm::rsqrt(x); break;
            case 6: fm::sincos2pi(x, y, t); break;
            case 7: fm::sincos2pi(x, t, y); break;
            case 9: y = pgm_lgamma(x); break;
            case 10: y = fm::erfcx_pos(x); break;
            case 11: y = fm::erfcx_pos(x) * fm::exp_neg_sq(x); break;  // erfc(x), x >= 0, as the cdf kernels form it
            default: y = fm::div(1.0, x); break;
        }
        out[i] = y;
    }
}

int
launch_test_math(int which, const double* d_in, size_t n, double* d_out, cudaStream_t s)
{
    if (n == 0) return 0;
    test_math_kernel<<<1184, 256, 0, s>>>(which, d_in, n, d_out);
    count_launch(1);
    return (int)cudaGetLastError();
}

int
launch_check_h(const double* d_h, size_t n, int devroye, int* d_flag, cudaStream_t s)
{
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    check_h_kernel<<<(uint32_t)blocks, 256, 0, s>>>(d_h, n, devroye, d_flag);
    count_launch(1);
    return (int)cudaGetLastError();
}

}  // namespace pgm
