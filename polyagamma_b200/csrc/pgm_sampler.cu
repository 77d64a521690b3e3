// pgm_sampler.cu -- the batched Polya-Gamma fill on sm_100a.
//
// Replaces src/pgm_random.c:144-180 (pgm_random_polyagamma / _fill / _fill2 and the hybrid
// dispatch).  One pass (up to 2^27 elements) of the output index range:
//
//   tile_kernel      persistent blocks, tiles of 2048 elements: load (h, z) once, classify (hybrid rule or the
//                    explicit method), draw the convergent buckets in place -- normal approximation and left-only
//                    saddle, from shared memory -- and append the others as compact (index, h, z) records to
//                    per-bucket lists; the tile's outputs leave as coalesced 16-byte stores
//   per list         setup_kernel<M> (per-element records; alternate and full saddle only) and sample_kernel<M>
//                    on two helper streams, overlapping the next pass's tile_kernel
//   small fills      small_kernel (one launch, one thread per element) up to 2^17 elements, tile_kernel +
//                    drain_kernel (every list in one launch) up to 2^18
//   scalar (h, z)    one record for the whole call, the method's kernels over the identity range
//   gamma            dense_kernel<gamma> over the identity range
//
// sample_kernel: every lane owns one element at a time and advances it by ONE proposal
// attempt per loop trip (pgm_methods.cuh); lanes whose element completed are retired and
// refilled from a per-bucket work counter, found with __ballot_sync/__popc, so a warp never
// idles on another lane's rejection loop.  Nothing here synchronises the host.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pgm_internal.h"
#include "pgm_methods.cuh"

namespace pgm {

// Diagnostic (PGM_CUDA_EVLOG=1): an event after every operation the pipeline enqueues, so that a watchdog can
// ask which of them the device has completed (pgm_cuda_debug_report) without touching the kernels.
struct LoggedOp {
    char what[48];
    cudaEvent_t ev;
};
static std::vector<LoggedOp> g_evlog;
static std::mutex g_evlog_mutex;
static void
evlog(const char* what, int k, cudaStream_t st)
{
    static const bool on = getenv("PGM_CUDA_EVLOG") != nullptr;
    if (!on) return;
    LoggedOp op;
    snprintf(op.what, sizeof(op.what), "%s %d", what, k);
    cudaEventCreateWithFlags(&op.ev, cudaEventDisableTiming);
    cudaEventRecord(op.ev, st);
    std::lock_guard<std::mutex> lock(g_evlog_mutex);
    g_evlog.push_back(op);
}
void
evlog_report()
{
    std::lock_guard<std::mutex> lock(g_evlog_mutex);
    size_t done = 0;
    for (auto& op : g_evlog) done += (cudaEventQuery(op.ev) == cudaSuccess);
    fprintf(stderr, "[pgm] %zu operations logged, %zu complete; incomplete:\n", g_evlog.size(), done);
    int shown = 0;
    for (auto& op : g_evlog) {
        if (cudaEventQuery(op.ev) != cudaSuccess && shown++ < 12) fprintf(stderr, "[pgm]   %s\n", op.what);
    }
}

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
// buckets
// ---------------------------------------------------------------------------------------
// bucket ids (the gamma method never needs the tile pass: it runs on the identity range)
enum : int { kBDevroye = 0, kBAlternate = 1, kBSaddleFull = 2, kBSaddleLeft = 3, kBAlternateLeft = 4, kBNormal = 5,
             kBDevroyeMSH = 6, kNumBuckets = 7, kBGamma = 7 };
constexpr uint32_t kBucketNone = 15;

// ctl layout of one set of element lists (uint32): [0..9] list lengths by bucket, [10..19] work counters of the
// buckets' sampling kernels
constexpr int kCtlCount = 0, kCtlWork = 10, kCtlWords = 32;
static_assert(kNumBuckets <= 10, "ctl layout");

// (written with selects, not branches: the tile kernel classifies eight elements per thread and the lanes
// of a warp meet every bucket)
__host__ __device__ __forceinline__ int
bucket_of(int method, double h, double z, bool compat)
{
    const double az = fabs(z);  // (z = NaN fails every test below: it is sampled as z = 0)
    // |z|/2 < 1/T: exponential-pair proposal, else Michael-Schucany-Haas
    const int bd = (0.5 * az >= 1.5625) ? kBDevroyeMSH : kBDevroye;
    const int ba = alternate_left_only(z) ? kBAlternateLeft : kBAlternate;
    const int bs = saddle_mode(h, z, compat) == kSaddleFull ? kBSaddleFull : kBSaddleLeft;
    return method == kNormal ? kBNormal : method == kSaddle ? bs : method == kDevroye ? bd
           : method == kAlternate ? ba : kBGamma;
}

// bucket_of(hybrid_select(h, z) or the explicit method, h, z) for the tile kernel, which classifies eight elements
// per thread: every predicate is evaluated, nothing branches (the switch / early-return forms above compile
// to a jump table and a dozen divergent branches per element).  `method` is warp-uniform.
__device__ __forceinline__ uint32_t
classify(int method, double h, double z, bool compat)
{
    const double az = fabs(z);  // (z = NaN fails every test: it is sampled as z = 0)
    const bool p_norm = h > 50.0;
    const bool p_sad = (h >= 8.0) | ((h > 4.0) & (z <= 4.0));
    const bool p_dev = (h == 1.0) | ((h == floor(h)) & (z <= 1.0));
    const bool p_full = compat | !(h * fma(0.125, az, 1.0) >= 22.0);
    const uint32_t bs = p_full ? (uint32_t)kBSaddleFull : (uint32_t)kBSaddleLeft;
    const uint32_t ba = (az >= 14.0) ? (uint32_t)kBAlternateLeft : (uint32_t)kBAlternate;
    const uint32_t bd = (az >= 3.125) ? (uint32_t)kBDevroyeMSH : (uint32_t)kBDevroye;
    const uint32_t hyb = p_norm ? (uint32_t)kBNormal : p_sad ? bs : p_dev ? bd : ba;
    const uint32_t expl = (method == kSaddle) ? bs : (method == kDevroye) ? bd : ba;
    return (method == kHybrid) ? hyb : expl;
}

// Where a bucket's elements are drawn.  The near-dense buckets whose proposals cost about the same for every
// element (normal approximation, left-only saddle: 96 % of cfg4) are drawn INSIDE tile_kernel, from shared
// memory, and their outputs leave the SM as part of the tile's coalesced store.  The others go to global
// element lists: slot s of the pass's ElemLists holds (index, h, z) of its elements, compacted, so the
// per-method kernels read contiguously instead of gathering 8-byte operands through 32-byte sectors.
// (devroye was drawn in the tiles for a while: its flattened state machine needs resident lanes, and a tile's
// list is too short for them -- every tile ended in a drain at 16 of 32 lanes, cfg2 ran at 8.7e9 samples/s
// against 1.5e10 with one persistent kernel over the whole pass, and its registers cost the tile kernel a
// fourth resident block.)
__host__ __device__ __forceinline__ int
list_slot(int bucket)
{
    switch (bucket) {
        case kBAlternate: return 0;
        case kBAlternateLeft: return 1;
        case kBDevroye: return 2;
        case kBDevroyeMSH: return 3;
        case kBSaddleFull: return 4;
        default: return -1;  // drawn in the tile
    }
}
constexpr int kNumListSlots = 5;
// An element is in one list at most, but any list can hold the whole pass.  Two slots share a region of
// `stride` entries: the even one grows from its front, the odd one from its back, so five worst-case lists
// take three regions (60 instead of 100 bytes per element of the pass).
constexpr int kNumListRegions = (kNumListSlots + 1) / 2;

struct ElemLists {
    uint32_t* idx;  // entry p of slot s at list_entry(s, stride, p) of each array
    double* h;
    double* z;
    uint32_t stride;
};
__host__ __device__ __forceinline__ size_t
list_entry(int slot, uint32_t stride, uint32_t p)
{
    return (size_t)(slot >> 1) * stride + ((slot & 1) ? stride - 1u - p : p);
}

// ---------------------------------------------------------------------------------------
// work description shared by the per-method kernels
// ---------------------------------------------------------------------------------------
// A bucket is either the identity range [0, count_imm) of the chunk (ctl == nullptr; operands come
// from FillArgs) or one slot of the pass's element lists, whose length and record base live in ctl
// on the device.
struct Bucket {
    const uint32_t* idx;  // list mode: element indices (relative to a.base); nullptr: identity range
    const double* lh;     // list mode: the elements' h and z, compacted in list order
    const double* lz;
    const uint32_t* ctl;
    int bucket;
    uint32_t count_imm;
    double* params;  // record storage (M::kFields doubles per element, structure of arrays)
    int uniform;     // 1: one record (position 0) serves every element (scalar h and z)
    int dir = 1;     // list mode: entry p is at idx[dir * p] (-1: a list that grows from the back of its region)
};

struct BucketView {
    const uint32_t* idx;
    uint32_t count;
    double* recs;     // field f of position p at recs[f * stride + p]
    uint32_t stride;
};

__device__ __forceinline__ BucketView
view_of(const Bucket& b)
{
    BucketView v;
    v.idx = b.idx;
    v.count = b.count_imm;
    v.recs = b.params;
    if (b.ctl) v.count = b.ctl[kCtlCount + b.bucket];
    v.stride = b.uniform ? 1u : v.count;
    return v;
}

// operands of list position pos (list mode) or of element e (identity range)
__device__ __forceinline__ void
bucket_operands(const FillArgs& a, const Bucket& b, uint32_t pos, uint32_t& e, double& h, double& z)
{
    if (b.idx) {
        const ptrdiff_t o = (ptrdiff_t)b.dir * (ptrdiff_t)pos;
        e = b.idx[o];
        h = b.lh[o];
        z = b.lz[o];
    }
    else {
        e = pos;
        load_hz(a, a.base + e, h, z);
    }
}

struct RecView {
    const double* base;
    uint32_t stride, pos;
    __device__ __forceinline__ double operator[](int f) const { return base[(size_t)f * stride + pos]; }
};

// FUSED sampling: the lane computes its element's record itself at refill (registers only) instead of
// reading what setup_kernel stored.
struct LocalRec {
    const double* v;
    __device__ __forceinline__ double operator[](int f) const { return v[f]; }
};

// ---------------------------------------------------------------------------------------
// tile kernel: classify + draw the dense buckets in shared memory + compact the rest
// ---------------------------------------------------------------------------------------
// A persistent block walks over tiles of kTile consecutive elements.  Per tile:
//   1. h and z are read ONCE, with 16-byte coalesced loads, into shared memory; every element gets its
//      bucket (invalid h: the NaN / inf of src/pgm_random.c:137-141 goes straight to the output tile) and
//      its position in the tile's list of that bucket (MATCH.ANY groups the lanes of a bucket, one
//      shared-memory atomic per group);
//   2. the elements of the divergent buckets are appended, as (index, h, z), to the pass's global lists
//      (one global atomic per bucket and tile reserves the range; stores are contiguous per bucket);
//   3. the in-tile buckets are drawn from shared memory, the warps of the block moving through the same
//      phases together (one hot loop at a time in the instruction cache):
//      - left-only saddle (one proposal per trip, acceptance ~0.9): a QUEUE, not resident lanes.  A round
//        gives every thread of the block one queued element; a rejected element goes back to the queue as a
//        record with the word position its Philox stream has reached, and what is left when a tile closes
//        (less than one round) is carried, as records, into the block's next tile -- so every round but the
//        block's very last runs with 32 of 32 lanes in every warp (resident lanes left 18-20 of 32 active: a
//        tile holds ~430 such elements, 1.7 warp-loads per warp);
//      - devroye (a flattened state machine of several trips per draw): resident lanes that retire finished
//        elements and refill from the tile's list (__ballot_sync/__popc), drained per tile;
//      - normal approximation: a work queue of batches of 64 (two elements per lane and trip: their dependency
//        chains interleave);
//   4. the tile's outputs leave as one coalesced, 16-byte vectorised store.  Positions that belong to a
//      divergent bucket, or to a carried saddle element, hold NaN until that bucket's kernel (later in stream
//      order) or this block (after a later barrier) overwrites them.
// HBM traffic: 16 B in + 8 B out per element, + 20 B written for each element of a divergent bucket.
// Which lane draws an element never changes the draw (private Philox stream per element).
#ifndef PGM_TILE_THREADS
#define PGM_TILE_THREADS 256
#endif
// Blocks per SM: 4 (64 registers, 16 B of spills).  While devroye was drawn in the tiles the kernel needed 80
// registers to stay out of local memory and ran 3 blocks per SM (cfg4 tile time 24.3 ms per 1e9; 22.9 now).
#ifndef PGM_TILE_BLOCKS_PER_SM
#define PGM_TILE_BLOCKS_PER_SM (1024 / PGM_TILE_THREADS)
#endif
constexpr int kTileThreads = PGM_TILE_THREADS;
constexpr int kTileItems = 8;
constexpr int kTile = kTileThreads * kTileItems;
constexpr int kCarryCap = 2 * kTileThreads;  // ring of carried saddle records: < one round + one round's rejects

struct __align__(16) TileSmem {
    double hout[kTile];         // h on the way in, the tile's outputs on the way out (an element's h is consumed
    double z[kTile];            // by the thread that later writes its output)
    double2 c_hz[kCarryCap];    // ring of saddle records (rejected proposals, elements of closed tiles): (h, z),
    uint32_t c_idx[kCarryCap];  //   element index inside the pass,
    uint32_t c_pos[kCarryCap];  //   Philox word position
    uint16_t list[kTile];       // local indices, grouped by bucket
    uint32_t cnt2[2][kNumBuckets];   // elements of each bucket in this tile   (two sets: a tile uses the set of its
    uint32_t next2[2][kNumBuckets];  // in-tile buckets: work counter            parity and clears the other one)
    uint32_t off[kNumBuckets];    // start of each bucket's group in list
    uint32_t gbase[kNumBuckets];  // list buckets: start of this tile's range in the global list
    uint32_t c_tail;              // ring: records pushed so far (monotonic; slot = count mod kCarryCap)
};

// Two normal-approximation draws, interleaved
static __device__ __forceinline__ double2
normal_pair(double h0, double z0, double h1, double z1, uint32_t k0, uint32_t k1, uint64_t i0, uint64_t i1)
{
    PhiloxKey key;
    key.k0 = k0;
    key.k1 = k1;
    RngInline r0, r1;
    r0.init(key, i0);
    r1.init(key, i1);
    return make_double2(normal_approx_sample(r0, h0, z0), normal_approx_sample(r1, h1, z1));
}

// ONE proposal of the left-only saddle sampler for an element whose Philox stream stands at word `wpos`.
struct SaddleAttempt {
    double value;
    uint32_t wpos;
    bool accepted;
};

static __device__ __forceinline__ SaddleAttempt
saddle_attempt(double h, double z, bool compat, uint32_t k0, uint32_t k1, uint64_t index, uint32_t wpos)
{
    using M = SaddleLeft;
    double loc[M::kFields];
    M::make_record(h, z, compat, loc);
    M::Par par;
    M::State st;
    M::Rng rng;
    M::load(LocalRec{loc}, par, st, compat);
    PhiloxKey key;  // (the outlined generator uses k0, k1 only)
    key.k0 = k0;
    key.k1 = k1;
    rng.init(key, index);
    rng.seek(wpos);
    SaddleAttempt r;
    r.accepted = M::step(par, st, rng, compat);
    r.value = st.acc;
    r.wpos = rng.word_position();
    return r;
}

// One round of the saddle queue.  The queue is the ring's records c_head .. c_tail-1 followed by the tile's
// first-attempt list entries l_head .. l_count-1; thread t < take takes the t-th queued element and makes ONE
// proposal.  Accepted: the sample goes to the open tile (shared memory) or, for an element of a closed tile,
// straight to global memory.  Rejected: back to the ring with the word position the element's Philox stream
// has reached, so the next attempt continues where this one stopped.  (Block-uniform arguments; the caller
// puts a barrier between rounds.)
__device__ __forceinline__ bool
saddle_round(const FillArgs& a, bool compat, uint32_t tile0, uint32_t tile_n, TileSmem& S, uint32_t take, uint32_t c_head,
             uint32_t c_count, const uint16_t* __restrict__ list, uint32_t l_head, uint32_t tid, uint32_t lt)
{
    const bool mine = tid < take;
    bool rejected = false;
    uint32_t idx = 0, wpos = 0, src = 0;  // idx: element index inside the pass; src: where (h, z) came from
    if (mine) {
        double h, z;
        if (tid < c_count) {
            src = (c_head + tid) % (uint32_t)kCarryCap;
            const double2 hz = S.c_hz[src];
            h = hz.x, z = hz.y;
            idx = S.c_idx[src];
            wpos = S.c_pos[src];
            src |= 0x80000000u;
        }
        else {
            src = list[l_head + (tid - c_count)];
            h = S.hout[src], z = S.z[src];
            idx = tile0 + src;
        }
        const SaddleAttempt r = saddle_attempt(h, z, compat, a.key.k0, a.key.k1, a.first_index + a.base + idx, wpos);
        if (r.accepted) {
            // an element of the open tile goes to shared memory, one of a closed tile straight to global memory
            if (idx - tile0 < tile_n) S.hout[idx - tile0] = r.value;
            else a.out[a.base + idx] = r.value;
        }
        else {
            rejected = true;
            wpos = r.wpos;
        }
    }
    const uint32_t rej = __ballot_sync(0xffffffffu, rejected);
    if (rej != 0) {
        uint32_t base = 0;
        const int leader = __ffs(rej) - 1;
        if ((int)(tid & 31) == leader) base = atomicAdd(&S.c_tail, (uint32_t)__popc(rej));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (rejected) {
            // (h, z) are read back from where they came from: keeping them live through the proposal costs registers
            const double2 hz = (src >> 31) ? S.c_hz[src & 0x7fffffffu] : make_double2(S.hout[src], S.z[src]);
            const uint32_t slot = (base + __popc(rej & lt)) % (uint32_t)kCarryCap;
            S.c_hz[slot] = hz;
            S.c_idx[slot] = idx;
            S.c_pos[slot] = wpos;
        }
    }
    return rejected;
}

// The operands of one full tile of contiguous, 16-byte aligned arrays (or scalars): thread t takes the pairs
// j * T + t, j = 0..3, with 16-byte loads.
static __device__ __forceinline__ void
load_tile_operands(const FillArgs& a, uint64_t g0, uint32_t tid, double (&hv)[8], double (&zv)[8])
{
    if (a.h != nullptr && a.h_stride != 0) {
        const double2* p = reinterpret_cast<const double2*>(a.h + g0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 v = __ldg(p + j * kTileThreads + tid);
            hv[2 * j] = v.x;
            hv[2 * j + 1] = v.y;
        }
    }
    else {  // one value for the whole call: a host scalar, or a one-element device operand
        const double v = a.h ? __ldg(a.h) : a.h_scalar;
#pragma unroll
        for (int k = 0; k < 8; ++k) hv[k] = v;
    }
    if (a.z != nullptr && a.z_stride != 0) {
        const double2* p = reinterpret_cast<const double2*>(a.z + g0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 v = __ldg(p + j * kTileThreads + tid);
            zv[2 * j] = v.x;
            zv[2 * j + 1] = v.y;
        }
    }
    else {
        const double v = a.z ? __ldg(a.z) : a.z_scalar;
#pragma unroll
        for (int k = 0; k < 8; ++k) zv[k] = v;
    }
}

// `method` is the requested sampler (kHybrid: src/pgm_random.c:105-121 decides per element).
// VEC: h and z are contiguous, 16-byte aligned arrays (or scalars) and every tile of the launch is full.
// The launch covers tiles tile_first .. tile_first + tile_count - 1 of the pass.
template <bool VEC>
__global__ void __launch_bounds__(kTileThreads, PGM_TILE_BLOCKS_PER_SM)
tile_kernel(FillArgs a, uint32_t n, int method, uint32_t* __restrict__ ctl, ElemLists el, uint32_t tile_first,
            uint32_t tile_count)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(tile_smem);
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const bool compat = a.compat != 0;
    uint32_t c_head = 0, c_tail = 0;  // ring: records consumed / pushed so far (block-uniform; slot = count mod capacity)
    if (tid == 0) S.c_tail = 0;       // (the shared copy hands out the slots of a round's pushes)
    if (tid < 2 * kNumBuckets) {
        (&S.cnt2[0][0])[tid] = 0;
        (&S.next2[0][0])[tid] = 0;
    }
    __syncthreads();

    // Barriers of a tile: B1 after the classification, B3 after the local lists, one per saddle round, BA after
    // the in-tile sampling, BE after the tile's stores.  The operands of the block's NEXT tile are prefetched into
    // L2 while this one is sampled (holding them in registers across the loop makes ptxas spill them at once);
    // the reservation in the global lists is made by the first warp at the start of the normal-approximation
    // phase, where the other warps do not wait for it.
    uint32_t parity = 0;
    for (uint32_t tile = blockIdx.x; tile < tile_count; tile += gridDim.x, parity ^= 1u) {
        const uint32_t tile0 = (tile_first + tile) * kTile;  // first element of the tile inside the pass
        const uint32_t tn = VEC ? (uint32_t)kTile : min((uint32_t)kTile, n - tile0);
        const uint64_t g0 = a.base + tile0;
        uint32_t* const cnt = S.cnt2[parity];
        uint32_t* const next = S.next2[parity];

        // ---- 1. load + classify.  Thread t owns local elements j * 2T + 2 t + {0, 1}, j = 0..3 ----
        uint32_t where[kTileItems / 2] = {0, 0, 0, 0};  // per owned element: bucket (4 bits) | position in its group (12 bits)
        {
            double hv[kTileItems], zv[kTileItems];
            if (VEC) load_tile_operands(a, g0, tid, hv, zv);
            else {
                // partial tile, or strided / broadcast / unaligned operands: staged through shared memory (one
                // copy of the general index arithmetic)
#pragma unroll 1
                for (int k = 0; k < kTileItems; ++k) {
                    const uint32_t e = (k >> 1) * (2 * kTileThreads) + 2 * tid + (k & 1);
                    double h = 1.0, z = 0.0;
                    if (e < tn) load_hz(a, g0 + e, h, z);
                    S.hout[e] = h;
                    S.z[e] = z;
                }
#pragma unroll
                for (int j = 0; j < kTileItems / 2; ++j) {
                    const double2 vh = reinterpret_cast<const double2*>(S.hout)[j * kTileThreads + tid];
                    const double2 vz = reinterpret_cast<const double2*>(S.z)[j * kTileThreads + tid];
                    hv[2 * j] = vh.x;
                    hv[2 * j + 1] = vh.y;
                    zv[2 * j] = vz.x;
                    zv[2 * j + 1] = vz.y;
                }
            }
            // Bucket of each element, and its position inside the tile's group of that bucket: the lanes of a
            // bucket are found with one MATCH.ANY on the bucket id, the lowest lane of a group takes the group's
            // positions from the bucket's counter in shared memory.
#pragma unroll
            for (int k = 0; k < kTileItems; ++k) {
                const uint32_t e = (k >> 1) * (2 * kTileThreads) + 2 * tid + (k & 1);
                const double hk = hv[k];
                // nan_or_inf (src/pgm_random.c:137-141): h <= 0 or NaN -> NaN, +inf -> +inf (h itself)
                const bool nonpos = !(hk > 0.0);
                const bool inval = nonpos | (hk == INFINITY);
                const uint32_t b = (e < tn && !inval) ? classify(method, hk, zv[k], compat) : kBucketNone;
                hv[k] = nonpos ? NAN : hk;  // invalid h: the output of this element
                const bool valid = b != kBucketNone;
                const uint32_t same = __match_any_sync(0xffffffffu, b);
                const int leader = __ffs(same) - 1;
                uint32_t r = 0;
                if (valid && (int)lane == leader) r = atomicAdd(&cnt[b], (uint32_t)__popc(same));
                r = __shfl_sync(0xffffffffu, r, leader) + __popc(same & lt);
                where[k >> 1] |= (b | (r << 4)) << (16 * (k & 1));
            }
#pragma unroll
            for (int j = 0; j < kTileItems / 2; ++j) {
                reinterpret_cast<double2*>(S.hout)[j * kTileThreads + tid] = make_double2(hv[2 * j], hv[2 * j + 1]);
                reinterpret_cast<double2*>(S.z)[j * kTileThreads + tid] = make_double2(zv[2 * j], zv[2 * j + 1]);
            }
        }
        __syncthreads();

        // ---- 2. group offsets (every warp works them out for itself: identical values, no barrier); clear the
        // other set of counters for the next tile; write the local lists ----
        {
            const uint32_t c = (lane < kNumBuckets) ? cnt[lane] : 0u;
            uint32_t incl = c;  // inclusive scan over the first eight lanes
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            if (lane < kNumBuckets) S.off[lane] = incl - c;
            if (tid < kNumBuckets) {
                S.cnt2[parity ^ 1u][tid] = 0;
                S.next2[parity ^ 1u][tid] = 0;
            }
        }
        __syncwarp();
        if (VEC && tile + gridDim.x < tile_count) {
            // one 128-byte line of the next tile per thread: the first half of the block takes h, the second z
            const double* src = tid < kTileThreads / 2 ? a.h : a.z;
            const bool arr = tid < kTileThreads / 2 ? a.h_stride != 0 : a.z_stride != 0;
            if (src != nullptr && arr) {
                const double* q = src + g0 + (uint64_t)gridDim.x * kTile + 16u * (tid & (kTileThreads / 2 - 1));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
            }
        }
#pragma unroll
        for (int k = 0; k < kTileItems; ++k) {
            const uint32_t e = (k >> 1) * (2 * kTileThreads) + 2 * tid + (k & 1);
            const uint32_t w = (where[k >> 1] >> (16 * (k & 1))) & 0xffffu;
            const uint32_t q = w & 15u;
            if (q != kBucketNone) S.list[S.off[q] + (w >> 4)] = (uint16_t)e;
        }
        __syncthreads();  // B3

        // ---- 3. in-tile buckets ----
        const uint64_t index0 = a.first_index + g0;
        {
            // saddle queue: full rounds while there is a block-load of work, the rest is carried.  c_tail (how
            // many records were ever pushed) is tracked in a register as well: a round's barrier counts its
            // rejections, so the block agrees on the queue length without reading the shared counter.
            const uint32_t m = cnt[kBSaddleLeft];
            const uint16_t* ls = S.list + S.off[kBSaddleLeft];
            uint32_t l_head = 0;
            for (;;) {
                const uint32_t c_count = c_tail - c_head;
                if (c_count + (m - l_head) < (uint32_t)kTileThreads) break;
                const uint32_t from_ring = min(c_count, (uint32_t)kTileThreads);
                const bool rejected = saddle_round(a, compat, tile0, tn, S, kTileThreads, c_head, from_ring, ls, l_head, tid, lt);
                c_head += from_ring;
                l_head += kTileThreads - from_ring;
                c_tail += (uint32_t)__syncthreads_count(rejected);  // (and the round's pushes are visible)
            }
            // close the tile: first-attempt entries that did not fill a round become ring records (their
            // positions in the output tile keep h until this block, from a later tile, writes the sample)
            const uint32_t left = m - l_head;
            if (tid < left) {
                const uint32_t e = ls[l_head + tid];
                const uint32_t slot = (c_tail + tid) % (uint32_t)kCarryCap;
                S.c_hz[slot] = make_double2(S.hout[e], S.z[e]);
                S.c_idx[slot] = tile0 + e;
                S.c_pos[slot] = 0;
            }
            c_tail += left;
            if (tid == 0) S.c_tail = c_tail;  // (the next push is at least two barriers away)
        }
        if (tid < kNumBuckets) {
            // reserve the tile's range in each global list: the first warp waits for the answer while the others
            // are already taking normal-approximation work
            const uint32_t c = cnt[tid];
            uint32_t reserved = 0;
            if (c != 0 && list_slot((int)tid) >= 0) reserved = atomicAdd(&ctl[kCtlCount + tid], c);
            S.gbase[tid] = reserved;
        }
        {
            // normal approximation: a work queue of items of 64 elements (two per lane, interleaved), and of 32
            // elements for the last block-load of them, so that the warps reach the barrier below closer together
            const uint32_t cn = cnt[kBNormal];
            const uint16_t* ln = S.list + S.off[kBNormal];
            const uint32_t nfull = cn > (uint32_t)kTileThreads ? (cn - (uint32_t)kTileThreads) >> 6 : 0u;
            const uint32_t rest0 = nfull << 6;
            const uint32_t nitems = nfull + ((cn - rest0 + 31u) >> 5);
            for (;;) {
                uint32_t it = 0;
                if (lane == 0) it = atomicAdd(&next[kBNormal], 1u);
                it = __shfl_sync(0xffffffffu, it, 0);
                if (it >= nitems) break;
                if (it < nfull) {
                    const uint32_t i = (it << 6) + lane;
                    const uint32_t e0 = ln[i], e1 = ln[i + 32u];
                    const double2 v = normal_pair(S.hout[e0], S.z[e0], S.hout[e1], S.z[e1], a.key.k0, a.key.k1,
                                                  index0 + e0, index0 + e1);
                    S.hout[e1] = v.y;
                    S.hout[e0] = v.x;
                }
                else {
                    const uint32_t i = rest0 + ((it - nfull) << 5) + lane;
                    if (i < cn) {
                        const uint32_t e0 = ln[i];
                        PhiloxKey key;
                        key.k0 = a.key.k0;
                        key.k1 = a.key.k1;
                        RngInline r0;
                        r0.init(key, index0 + e0);
                        S.hout[e0] = normal_approx_sample(r0, S.hout[e0], S.z[e0]);
                    }
                }
            }
        }
        __syncthreads();  // BA

        // ---- 4. divergent buckets: append (index, h, z) to the global lists (their positions in the tile's
        // output are written by the bucket's own kernel, which runs after this one) ----
#pragma unroll 1
        for (int q = 0; q < kNumBuckets; ++q) {
            const int slot = list_slot(q);
            if (slot < 0) continue;
            const uint32_t c = cnt[q];
            if (c == 0) continue;  // block-uniform
            const uint32_t off = S.off[q];
            const uint32_t first = S.gbase[q];
#pragma unroll 1
            for (uint32_t i = tid; i < c; i += kTileThreads) {
                const uint32_t e = S.list[off + i];
                const size_t dst = list_entry(slot, el.stride, first + i);
                el.idx[dst] = tile0 + e;
                el.h[dst] = S.hout[e];
                el.z[dst] = S.z[e];
            }
        }
        // ---- 5. the tile's outputs ----
        double* out = a.out + g0;
        if (tn == (uint32_t)kTile && ((uintptr_t)out & 15) == 0) {
#pragma unroll
            for (int j = 0; j < kTileItems / 2; ++j) {
                reinterpret_cast<double2*>(out)[j * kTileThreads + tid] =
                    reinterpret_cast<const double2*>(S.hout)[j * kTileThreads + tid];
            }
        }
        else {
            for (uint32_t e = tid; e < tn; e += kTileThreads) out[e] = S.hout[e];
        }
        // BE: the appends have read the tile before the next one overwrites it (and these stores are ordered
        // before the block's later stores to carried elements' positions)
        __syncthreads();
    }
    // the block's last carried elements: partial rounds until the ring is empty
    __syncthreads();
    while (c_tail != c_head) {
        const uint32_t take = min(c_tail - c_head, (uint32_t)kTileThreads);
        const bool rejected = saddle_round(a, compat, 0, 0, S, take, c_head, take, nullptr, 0, tid, lt);
        c_head += take;
        c_tail += (uint32_t)__syncthreads_count(rejected);
    }

}

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
        uint32_t e;
        double h, z, bad;
        bucket_operands(a, b, pos, e, h, z);
        double rec[M::kFields];
        if (invalid_h(h, bad)) {
            // only reachable on the identity range (tile_kernel filters invalid h itself)
            if (!b.uniform) a.out[a.base + e] = bad;
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
// every trip, this saves the record's round trip through HBM.

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
                    if (FUSED) {  // (list buckets only: invalid h never reaches a list)
                        uint32_t e;
                        double h, z, loc[M::kFields];
                        bucket_operands(a, b, pos, e, h, z);
                        g = a.base + e;
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
                        g = a.base + (v.idx ? v.idx[(ptrdiff_t)b.dir * (ptrdiff_t)pos] : pos);
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
// Identity range only (scalar parameters, or the gamma method over arrays): the hybrid / saddle / ... fills
// over arrays draw their normal-approximation elements inside tile_kernel.
template <int KIND>
__global__ void __launch_bounds__(256)
dense_kernel(FillArgs a, Bucket b)
{
    const uint32_t count = b.count_imm;
    const uint32_t stride = gridDim.x * blockDim.x;
    if (KIND == kNormal) {
        // two elements per trip: their dependency chains interleave
        for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < count; pos += 2 * stride) {
            const bool two = pos + stride < count;
            const uint64_t ga = a.base + pos, gb = two ? ga + stride : ga;
            double ha, za, hb, zb, bad;
            load_hz(a, ga, ha, za);
            load_hz(a, gb, hb, zb);
            RngInlineRK ra, rb;
            ra.init(a.key, a.first_index + ga);
            rb.init(a.key, a.first_index + gb);
            const double va = invalid_h(ha, bad) ? bad : normal_approx_sample(ra, ha, za);
            const double vb = invalid_h(hb, bad) ? bad : normal_approx_sample(rb, hb, zb);
            a.out[ga] = va;
            if (two) a.out[gb] = vb;
        }
        return;
    }
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < count; pos += stride) {
        uint64_t g = a.base + pos;
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
            default: {  // kBNormal
                RngInlineRK r;
                r.init(a.key, a.first_index + g);
                v = normal_approx_sample(r, h, z);
            }
        }
    }
    a.out[g] = v;
}

// ---------------------------------------------------------------------------------------
// drain kernel: every divergent bucket of a pass in ONE launch -- the path for fills between the one-launch
// kernel and a few 10^5 elements, where seven bucket launches on three streams and their event joins cost more
// than the sampling itself.  The lists of the pass form one index space (list after list); a thread takes the
// entries t, t + threads, ... and runs each to completion with the bucket's make_record / load / step code: the
// draws are the pipeline's, bit for bit.  Good while the lists are short (a hybrid fill with mixed shapes: 4 %
// of the elements); with long lists the lanes of a warp wait for the slowest element and the pipeline's
// refilling kernels win, so the path stops at 2^18 elements (tools/perf_mid.py, profiles/r2_perf_mid.txt; the
// same five lists through the persistent-lane loop inside one kernel were measured too, and lose to both).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
drain_kernel(FillArgs a, ElemLists el, const uint32_t* __restrict__ ctl)
{
    const bool compat = a.compat != 0;
    const uint32_t k0 = a.key.k0, k1 = a.key.k1;
    // end of each list slot's range in the common index space
    uint32_t end[kNumListSlots];
    const int bucket_of_slot[kNumListSlots] = {kBAlternate, kBAlternateLeft, kBDevroye, kBDevroyeMSH, kBSaddleFull};
    uint32_t total = 0;
#pragma unroll
    for (int sl = 0; sl < kNumListSlots; ++sl) {
        total += ctl[kCtlCount + bucket_of_slot[sl]];
        end[sl] = total;
    }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += stride) {
        int sl = 0;
        uint32_t first = 0;
#pragma unroll
        for (int j = 0; j + 1 < kNumListSlots; ++j) {
            if (p >= end[j]) {
                sl = j + 1;
                first = end[j];
            }
        }
        const size_t o = list_entry(sl, el.stride, p - first);
        const uint32_t idx = el.idx[o];
        const double h = el.h[o], z = el.z[o];
        const uint64_t index = a.first_index + a.base + idx;
        double v;
        switch (sl) {  // (list_slot())
            case 0: v = run_one<Alternate>(k0, k1, index, h, z, compat); break;
            case 1: v = run_one<AlternateLeft>(k0, k1, index, h, z, compat); break;
            case 2: v = run_one<DevroyeT<kDevroyePair>>(k0, k1, index, h, z, compat); break;
            case 3: v = run_one<DevroyeT<kDevroyeMSH>>(k0, k1, index, h, z, compat); break;
            default: v = run_one<SaddleFull>(k0, k1, index, h, z, compat); break;
        }
        a.out[a.base + idx] = v;
    }
}
static_assert(kNumListSlots == 5, "drain_kernel's switch follows list_slot()");

// fills the saddle tables (one thread per entry): the starting points of the FULL setup's right solve, and
// the Hermite pieces of Gr and Hr (pgm_methods.cuh)
__global__ void
init_tables_kernel()
{
    using namespace saddle;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < kGuessN) {
        double zz = (double)i / kGuessPerUnit;
        double xl = zz > 0.0 ? tanh(zz) / zz : 1.0;
        double f, fp;
        g_guess_ur[i] = solve(3.0 * xl, table_guess(3.0 * xl), f, fp);
    }
    if (i < kTabCells) {
        const int per = 1 << kTabBits;
        const double w = ldexp(1.0, kTabEmin + (i >> kTabBits) - kTabBits);  // cell width
        const double x0 = ldexp(1.0 + (double)(i & (per - 1)) / per, kTabEmin + (i >> kTabBits));
        double g0, dg0, h0, dh0, g1, dg1, h1, dh1;
        eval_direct(x0, g0, dg0, h0, dh0);
        eval_direct(x0 + w, g1, dg1, h1, dh1);
        // p(t) = f0 + m0 t + (3 df - 2 m0 - m1) t^2 + (m0 + m1 - 2 df) t^3,  m = w f'
        const double dg = g1 - g0, dh = h1 - h0;
        dg0 *= w, dg1 *= w, dh0 *= w, dh1 *= w;
        g_saddle_tab[4 * i] = make_double2(g0, dg0);
        g_saddle_tab[4 * i + 1] = make_double2(3.0 * dg - 2.0 * dg0 - dg1, dg0 + dg1 - 2.0 * dg);
        g_saddle_tab[4 * i + 2] = make_double2(h0, dh0);
        g_saddle_tab[4 * i + 3] = make_double2(3.0 * dh - 2.0 * dh0 - dh1, dh0 + dh1 - 2.0 * dh);
    }
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
    int tile_blocks_per_sm = 1;
    cudaMemPool_t pool = nullptr;  // the library's own stream-ordered scratch pool
    int init_rc = 0;
};

static DeviceInfo g_device_cache[64];
static std::mutex g_device_mutex;  // first calls from several host threads: one fills the entry, the others wait

// One-time, per device: occupancy of the persistent kernels, the saddle starting-point tables (filled by
// the solver itself on a private stream, which this call waits for -- the library's only host
// synchronisation on a device-pointer path, and only on the first call) and a PRIVATE memory pool for the
// stream-ordered scratch (freed blocks stay cached in it; the device's default pool is left alone).
static DeviceInfo
device_info()
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_device_mutex);
    DeviceInfo& d = g_device_cache[dev & 63];
    if (d.sms == 0) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[0], sample_kernel<DevroyeT<kDevroyePair>>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[1], sample_kernel<Alternate>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[2], sample_kernel<SaddleFull>, kPersistThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[3], sample_kernel<SaddleLeft>, kPersistThreads, 0);
        d.blocks_per_sm[4] = 1;  // (round 1: the far saddle kernel)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.blocks_per_sm[5], sample_kernel<AlternateLeft, true>, kPersistThreads, 0);
        d.blocks_per_sm[6] = 1;  // (round 1: the far saddle kernel on 8 <= |z| < 14)
        for (int i = 0; i < 7; ++i) {
            if (d.blocks_per_sm[i] < 1) d.blocks_per_sm[i] = 1;
        }
        cudaFuncSetAttribute(tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        cudaFuncSetAttribute(tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.tile_blocks_per_sm, tile_kernel<true>, kTileThreads, sizeof(TileSmem));
        if (d.tile_blocks_per_sm < 1) d.tile_blocks_per_sm = 1;
        cudaStream_t st = nullptr;
        cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        if (e == cudaSuccess) {
            constexpr int kEntries = saddle::kGuessN > saddle::kTabCells ? saddle::kGuessN : saddle::kTabCells;
            init_tables_kernel<<<(kEntries + 127) / 128, 128, 0, st>>>();
            e = cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
            count_launch(1);
        }
        if (e == cudaSuccess) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&d.pool, &props);
            if (e == cudaSuccess) {
                uint64_t thr = UINT64_MAX;
                cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
        d.init_rc = (int)e;
        if (e != cudaSuccess) (void)cudaGetLastError();
    }
    return d;
}

int
device_init()
{
    return device_info().init_rc;
}

// bytes the private pool holds right now (in use or cached for reuse)
static size_t
scratch_held_bytes()
{
    DeviceInfo d = device_info();
    if (d.init_rc != 0) return 0;
    uint64_t v = 0;
    if (cudaMemPoolGetAttribute(d.pool, cudaMemPoolAttrReservedMemCurrent, &v) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return (size_t)v;
}

int
scratch_alloc(void** p, size_t bytes, cudaStream_t s)
{
    DeviceInfo d = device_info();
    if (d.init_rc != 0) return d.init_rc;
    return (int)cudaMallocFromPoolAsync(p, bytes, d.pool, s);
}

// returns the cached scratch of the current device to the driver (after a device synchronisation)
int
scratch_release()
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_device_mutex);
        pool = g_device_cache[dev & 63].pool;
    }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return (int)e;
    if (pool == nullptr) return 0;
    return (int)cudaMemPoolTrimTo(pool, 0);
}

// helper streams per device: the divergent buckets' kernels of a pass run on them, beside the next
// pass's tile kernel on the caller's stream
static cudaStream_t
helper_stream(int which)
{
    static cudaStream_t streams[64][2] = {};
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
// Which list buckets compute their records inside the sampling kernel (measured on cfg4 in round 1, ms per
// 1e9, setup + sampling separate -> fused): saddle-mid 2.71 -> 2.44, alternate-left 2.04 -> 1.55; saddle-left
// gets slower (3.83 -> 4.22: its setup is longer and its lanes refill out of step), and the full saddle /
// alternate setups are the big ones the split exists for.
constexpr bool kFuseAltLeft = true;

template <class M, bool FUSED = false>
static void
launch_method(const DeviceInfo& d, int which, const FillArgs& a, const Bucket& b, uint64_t count_hint,
              uint32_t* counter, cudaStream_t s)
{
    const int kind = kKindDevroye + which;
    const bool debug = getenv("PGM_CUDA_DEBUG") != nullptr;  // diagnostic: synchronise and report after every launch
    if (!FUSED) {
        ScopedKernelTimer timer(kKindSetup + which, s);
        uint64_t nrec = b.uniform ? 1 : count_hint;
        setup_kernel<M><<<grid_for((uint64_t)d.sms * 8, nrec, 256), 256, 0, s>>>(a, b);
        if (debug) fprintf(stderr, "[pgm] setup_kernel %d -> %d\n", which, (int)cudaStreamSynchronize(s));
        evlog("setup_kernel", which, s);
    }
    {
        ScopedKernelTimer timer(kind, s);
        sample_kernel<M, FUSED><<<grid_for((uint64_t)d.sms * d.blocks_per_sm[which], count_hint, kPersistThreads),
                                  kPersistThreads, 0, s>>>(a, b, counter);
        if (debug) fprintf(stderr, "[pgm] sample_kernel %d -> %d\n", which, (int)cudaStreamSynchronize(s));
        evlog("sample_kernel", which, s);
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

constexpr uint64_t kSmallMaxDefault = 1ull << 17;  // mixed-method fills break even near 1.4e5 (tools/perf_small_max.py), single-method ones far later

static int
env_int(const char* name, int lo, int hi, int fallback)
{
    const char* e = getenv(name);  // read per call: the tests switch these
    if (e == nullptr || *e == '\0') return fallback;
    char* end = nullptr;
    long v = strtol(e, &end, 10);
    if (end == e || v < lo || v > hi) return fallback;
    return (int)v;
}

static int
env_chunk_log2()
{
    return env_int("PGM_CUDA_CHUNK_LOG2", 11, 28, 0);  // a pass holds at least one tile; 2^28 entries index with 32 bits
}

static inline size_t
align256(size_t v)
{
    return (v + 255) & ~size_t(255);
}

// largest fill that takes the one-launch path (PGM_CUDA_SMALL_MAX overrides, clamped to 2^22; 0 disables it)
static uint64_t
small_max()
{
    const char* e = getenv("PGM_CUDA_SMALL_MAX");
    if (e == nullptr || *e == '\0') return kSmallMaxDefault;
    char* end = nullptr;
    unsigned long long v = strtoull(e, &end, 10);
    if (end == e) return kSmallMaxDefault;
    return v > (1ull << 22) ? (1ull << 22) : (uint64_t)v;
}

// smallest pass whose divergent buckets are spread over the helper streams (below 2^16 elements the
// one-launch kernel runs instead)
static uint64_t
multi_min()
{
    return 1ull << env_int("PGM_CUDA_MULTI_MIN_LOG2", 0, 40, 16);
}

// Largest fill that takes the two-launch path (tile kernel + drain kernel).  PGM_CUDA_MID_MAX overrides it
// (0 disables the path); tools/perf_mid.py has the measurements behind the default.
constexpr uint64_t kMidMaxDefault = 1ull << 18;
static uint64_t
mid_max()
{
    const char* e = getenv("PGM_CUDA_MID_MAX");
    if (e == nullptr || *e == '\0') return kMidMaxDefault;
    char* end = nullptr;
    unsigned long long v = strtoull(e, &end, 10);
    if (end == e) return kMidMaxDefault;
    return v > (1ull << 25) ? (1ull << 25) : (uint64_t)v;
}

// fill2 over arrays, n <= 2^31: passes of tile_kernel (caller's stream) + the divergent buckets' kernels (helper
// streams).  There are two sets of element lists: pass k appends to set k & 1 while the bucket kernels of pass
// k - 1 drain the other set beside it; the record regions (one per bucket) are shared by the sets, so a
// bucket's kernels run in launch order on one stream.  Scratch: per set 3 lists x 20 B x pass, + 14 doubles of
// records x pass (the one-pass compaction cannot know a list's length before it writes, so every list has room
// for the whole pass): 232 B per element of a pass, 7.8 GB at 2^25; the pass is halved while the stream-ordered
// allocation fails.  PGM_CUDA_CHUNK_LOG2 in the environment pins the pass size (tests use it to cross pass seams).
// (Tried and dropped: letting the lists accumulate over several passes and draining them at a threshold, so that
// the divergent kernels run over millions of elements instead of a few hundred thousand.  It measured 3 % slower
// on cfg4 -- those kernels are divergence-bound, not tail-bound, and what is left for the forced drain after the
// last pass has nothing to overlap with.)
#ifndef PGM_CHUNK_LOG2_MAX
#define PGM_CHUNK_LOG2_MAX 27
#endif
constexpr int kChunkLog2Min = 20, kChunkLog2Max = PGM_CHUNK_LOG2_MAX;

static int
first_chunk_log2(uint64_t n)
{
    if (int forced = env_chunk_log2()) return forced;
    int lg = kChunkLog2Max;
    while (lg > kChunkLog2Min && (1ull << (lg - 1)) >= n) --lg;
    return lg;
}

static int
fill_arrays(const DeviceInfo& d, FillArgs a, uint64_t n, int m, cudaStream_t s)
{
    cudaError_t err;
    int lg = first_chunk_log2(n);
    const bool no_overlap = getenv("PGM_CUDA_NO_OVERLAP") != nullptr;
    const bool any_buckets = true;
    // one pass, and small enough that launches and joins outweigh the divergence of a run-to-completion kernel:
    // tile kernel + drain kernel on the caller's stream, no records, no helper streams
    const bool mid_ok = !no_overlap && n <= mid_max();
    bool mid = false;
    uint64_t chunk = 0;
    uint32_t cap = 0;
    size_t list_bytes = 0, rec_bytes = 0;
    int nsets = 1;
    const size_t ctl_bytes = align256(kCtlWords * sizeof(uint32_t));
    constexpr int kRecFields = Alternate::kFields + SaddleFull::kFields;
    char* ws = nullptr;
    // Passes above 2^25 elements are worth 4 % on a mixed fill (the divergent buckets' kernels of a pass are
    // latency bound: fewer, longer ones), but their worst-case lists and records are 232 bytes per element of the
    // pass: take them only while the scratch stays below a third of the memory that is free right now.
    if (lg > 25 && !env_chunk_log2()) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            (void)cudaGetLastError();
            free_b = 0;
        }
        const size_t held = scratch_held_bytes();  // (what the pool already holds is reused, not taken again)
        while (lg > 25) {
            const uint64_t c = (n < (1ull << lg)) ? n : (1ull << lg);
            const size_t want = (size_t)c * ((n > c ? 2 : 1) * kNumListRegions * 20 + kRecFields * 8);
            if (want <= held || want - held <= free_b / 3) break;
            --lg;
        }
    }
    for (;;) {
        chunk = 1ull << lg;
        mid = mid_ok && n <= chunk;
        const uint64_t cmax = n < chunk ? n : chunk;
        nsets = (n > chunk && !no_overlap && any_buckets) ? 2 : 1;
        cap = (uint32_t)((cmax + 63) & ~uint64_t(63));
        list_bytes = any_buckets ? align256((size_t)cap * kNumListRegions * (sizeof(uint32_t) + 2 * sizeof(double))) : 0;
        rec_bytes = (any_buckets && !mid) ? align256((size_t)cap * kRecFields * sizeof(double)) : 0;
        err = (cudaError_t)scratch_alloc((void**)&ws, nsets * (ctl_bytes + list_bytes) + rec_bytes, s);
        if (err == cudaSuccess) break;
        (void)cudaGetLastError();  // clear the sticky-free allocation error and retry smaller
        if (err != cudaErrorMemoryAllocation || lg <= kChunkLog2Min || env_chunk_log2()) return (int)err;
        --lg;
    }
    struct ListSet {
        ElemLists el;
        uint32_t* ctl;
    } set[2];
    for (int i = 0; i < nsets; ++i) {
        set[i].ctl = (uint32_t*)(ws + (size_t)i * ctl_bytes);
        char* base = ws + (size_t)nsets * ctl_bytes + (size_t)i * list_bytes;
        // h and z first (8-byte aligned whatever the capacity), then the indices
        set[i].el.h = (double*)base;
        set[i].el.z = set[i].el.h + (size_t)cap * kNumListRegions;
        set[i].el.idx = (uint32_t*)(set[i].el.z + (size_t)cap * kNumListRegions);
        set[i].el.stride = cap;
    }
    if (nsets == 1) set[1] = set[0];
    double* rec_alt = (double*)(ws + (size_t)nsets * (ctl_bytes + list_bytes));
    double* rec_full = rec_alt + (size_t)cap * Alternate::kFields;

    // 16-byte loads in the tile kernel need contiguous, aligned operands (a scalar operand is fine)
    auto vec_ok = [&](const double* p, int64_t stride) {
        return p == nullptr || stride == 0 || (stride == 1 && ((uintptr_t)(p + a.base) & 15) == 0);
    };
    const bool vec = a.bc.ndim == 0 && vec_ok(a.h, a.h_stride) && vec_ok(a.z, a.z_stride);
    const uint64_t base0 = a.base;

    auto tile_pass = [&](const FillArgs& ap, uint32_t cn, const ListSet& ls, cudaStream_t st) {
        const uint32_t ntiles = (cn + kTile - 1) / kTile;
        // full tiles of contiguous, aligned operands take the 16-byte-load kernel; a trailing partial tile (or
        // every tile, for general operands) the general one
        const uint32_t nvec = vec ? cn / kTile : 0, ngen = ntiles - nvec;
        const uint32_t gcap = (uint32_t)d.sms * d.tile_blocks_per_sm;
        const uint32_t gvec = nvec < gcap ? nvec : gcap, ggen = ngen < gcap ? ngen : gcap;
        cudaMemsetAsync(ls.ctl, 0, ctl_bytes, st);
        ScopedKernelTimer timer(kKindTile, st);
        if (gvec) tile_kernel<true><<<gvec, kTileThreads, sizeof(TileSmem), st>>>(ap, cn, m, ls.ctl, ls.el, 0, nvec);
        if (ggen) tile_kernel<false><<<ggen, kTileThreads, sizeof(TileSmem), st>>>(ap, cn, m, ls.ctl, ls.el, nvec, ngen);
        count_launch((gvec ? 1 : 0) + (ggen ? 1 : 0));
        evlog("tile pass, elements", (int)cn, st);
    };
    // The divergent buckets' kernels of one pass.  List lengths stay on the device: the kernels read them
    // there, so the host never waits for a count (an empty bucket costs one near-empty launch).
    auto bucket_pass = [&](const FillArgs& ap, const ListSet& ls, cudaStream_t s1, cudaStream_t s2) {
        auto bucket = [&](int q, double* recs) {
            const int slot = list_slot(q);
            const size_t o = list_entry(slot, cap, 0);  // entry 0: the front of the slot's region, or its back
            return Bucket{ls.el.idx + o, ls.el.h + o, ls.el.z + o, ls.ctl, q, 0, recs, 0, (slot & 1) ? -1 : 1};
        };
        if (m == kHybrid || m == kSaddle)
            launch_method<SaddleFull>(d, 2, ap, bucket(kBSaddleFull, rec_full), cap, ls.ctl + kCtlWork + kBSaddleFull, s2);
        if (m == kHybrid || m == kAlternate) {
            launch_method<Alternate>(d, 1, ap, bucket(kBAlternate, rec_alt), cap, ls.ctl + kCtlWork + kBAlternate, s1);
            launch_method<AlternateLeft, kFuseAltLeft>(d, 5, ap, bucket(kBAlternateLeft, nullptr), cap,
                                                       ls.ctl + kCtlWork + kBAlternateLeft, s2);
        }
        if (m == kHybrid || m == kDevroye) {
            launch_method<DevroyeT<kDevroyePair>, true>(d, 0, ap, bucket(kBDevroye, nullptr), cap, ls.ctl + kCtlWork + kBDevroye, s1);
            launch_method<DevroyeT<kDevroyeMSH>, true>(d, 0, ap, bucket(kBDevroyeMSH, nullptr), cap,
                                                       ls.ctl + kCtlWork + kBDevroyeMSH, s2);
        }
    };

    const uint64_t npass = (n + chunk - 1) / chunk;
    if (mid) {
        tile_pass(a, (uint32_t)n, set[0], s);
        {
            ScopedKernelTimer timer(kKindDrain, s);
            const uint64_t want = (n + 255) / 256, most = (uint64_t)d.sms * 8;
            drain_kernel<<<(unsigned)(want < most ? want : most), 256, 0, s>>>(a, set[0].el, set[0].ctl);
            count_launch(1);
            evlog("drain_kernel", 0, s);
        }
        cudaFreeAsync(ws, s);
        return (int)cudaGetLastError();
    }
    const bool multi = !no_overlap && any_buckets && (nsets == 2 || n >= multi_min());
    if (!multi) {
        for (uint64_t k = 0; k < npass; ++k) {
            const uint64_t off = k * chunk;
            const uint32_t cn = (uint32_t)((n - off) < chunk ? (n - off) : chunk);
            a.base = base0 + off;
            tile_pass(a, cn, set[0], s);
            if (any_buckets) bucket_pass(a, set[0], s, s);
        }
        cudaFreeAsync(ws, s);
        return (int)cudaGetLastError();
    }

    // (the caller's stream `s` may be the null handle of the default stream: compare helpers only)
    cudaStream_t h1 = helper_stream(0), h2 = helper_stream(1);
    if (h1 == nullptr || h2 == nullptr) {
        cudaFreeAsync(ws, s);
        return (int)cudaErrorUnknown;
    }
    // 0-1 tile pass on set, 2-3 helper 1 done with its kernels on set, 4-5 helper 2 likewise.  One set of events
    // per host thread and device, created once: a wait captures the record that precedes it, so re-recording
    // an event in the next call cannot disturb the waits of this one, and no other thread touches the set.
    struct EventSet {
        cudaEvent_t ev[6];
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
    cudaEvent_t *ev_tile = es.ev, *ev_d1 = es.ev + 2, *ev_d2 = es.ev + 4;
    for (uint64_t k = 0; k < npass; ++k) {
        const int buf = (nsets == 2) ? (int)(k & 1) : 0;
        const uint64_t off = k * chunk;
        const uint32_t cn = (uint32_t)((n - off) < chunk ? (n - off) : chunk);
        a.base = base0 + off;
        if (k >= (uint64_t)nsets) {  // the bucket kernels of this set's previous pass are done with it
            cudaStreamWaitEvent(s, ev_d1[buf], 0);
            cudaStreamWaitEvent(s, ev_d2[buf], 0);
        }
        tile_pass(a, cn, set[buf], s);
        cudaEventRecord(ev_tile[buf], s);
        cudaStreamWaitEvent(h1, ev_tile[buf], 0);
        cudaStreamWaitEvent(h2, ev_tile[buf], 0);
        bucket_pass(a, set[buf], h1, h2);
        cudaEventRecord(ev_d1[buf], h1);
        cudaEventRecord(ev_d2[buf], h2);
    }
    // join: each set's last bucket kernels (earlier ones precede them in stream order)
    for (int b2 = 0; b2 < (npass >= 2 ? nsets : 1); ++b2) {
        const int buf = (nsets == 2) ? (int)((npass - 1 - b2) & 1) : 0;
        cudaStreamWaitEvent(s, ev_d1[buf], 0);
        cudaStreamWaitEvent(s, ev_d2[buf], 0);
    }
    cudaFreeAsync(ws, s);
    return (int)cudaGetLastError();
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
    if (d.init_rc != 0) return d.init_rc;
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
        const uint64_t kChunk = 1ull << (env_chunk_log2() ? env_chunk_log2() : 27);
        const uint64_t nchunks = (n + kChunk - 1) / kChunk;
        char* ws = nullptr;
        const size_t rec_bytes = align256(kMaxRecordFields * sizeof(double));
        if ((err = (cudaError_t)scratch_alloc((void**)&ws, rec_bytes + nchunks * sizeof(uint32_t), s)) != cudaSuccess) return (int)err;
        double* params = (double*)ws;
        uint32_t* counters = (uint32_t*)(ws + rec_bytes);
        cudaMemsetAsync(counters, 0, nchunks * sizeof(uint32_t), s);
        const uint64_t base0 = a.base;
        for (uint64_t c = 0; c < nchunks; ++c) {
            const uint64_t off = c * kChunk;
            const uint32_t cn = (uint32_t)((n - off) < kChunk ? (n - off) : kChunk);
            a.base = base0 + off;
            Bucket b{nullptr, nullptr, nullptr, nullptr, bk, cn, params, 1};
            switch (bk) {
                case kBDevroye: launch_method<DevroyeT<kDevroyePair>>(d, 0, a, b, cn, counters + c, s); break;
                case kBDevroyeMSH: launch_method<DevroyeT<kDevroyeMSH>>(d, 0, a, b, cn, counters + c, s); break;
                case kBAlternate: launch_method<Alternate>(d, 1, a, b, cn, counters + c, s); break;
                case kBSaddleFull: launch_method<SaddleFull>(d, 2, a, b, cn, counters + c, s); break;
                case kBSaddleLeft: launch_method<SaddleLeft>(d, 3, a, b, cn, counters + c, s); break;
                case kBAlternateLeft: launch_method<AlternateLeft>(d, 5, a, b, cn, counters + c, s); break;
                case kBNormal: launch_dense<kNormal>(d, a, b, cn, s); break;
                default: launch_dense<kGamma>(d, a, b, cn, s); break;
            }
        }
        cudaFreeAsync(ws, s);
        return (int)cudaGetLastError();
    }

    if (m == kGamma) {
        // 200 gamma draws per element, no per-element setup worth splitting off: identity range, no scratch
        const uint64_t chunk = 1ull << 27;
        const uint64_t base0 = a.base;
        for (uint64_t off = 0; off < n; off += chunk) {
            const uint32_t cn = (uint32_t)((n - off) < chunk ? (n - off) : chunk);
            a.base = base0 + off;
            Bucket b{nullptr, nullptr, nullptr, nullptr, kBGamma, cn, nullptr, 0};
            launch_dense<kGamma>(d, a, b, cn, s);
        }
        return (int)cudaGetLastError();
    }

    // per-element parameters (fill2), every other method
    return fill_arrays(d, a, n, m, s);
}

// DFMA microbenchmark: the FP64 roofline denominator, measured on the box it is quoted on.  Sixteen
// independent chains per thread (eight left the pipe 8 % short of what tools/dmma_peak.cu reaches: 34.2
// against 37.0 TFLOP/s).
__global__ void __launch_bounds__(256)
dfma_peak_kernel(double* sink, int iters, double b, double c)
{
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fma(acc[k], b, c);
        }
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) r += acc[k];
    if (r == 12345.678) sink[0] = r;
}

int
measure_fp64_peak(double* tflops, cudaStream_t s)
{
    DeviceInfo d = device_info();
    double* sink = nullptr;
    cudaError_t e;
    if ((e = (cudaError_t)scratch_alloc((void**)&sink, sizeof(double), s)) != cudaSuccess) return (int)e;
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
            case 12: y = fm::exp_poly(x); break;
            case 13: y = fm::cos2pi_poly(x); break;
            case 10: y = fm::erfcx_pos(x); break;
            case 11: y = fm::erfcx_pos(x) * fm::exp_neg_sq(x); break;  // erfc(x), x >= 0, as the cdf kernels form it
            case 20: saddle::lookup<true>(x, y, t); break;               // Gr(x), tabulated
            case 21: saddle::lookup<true>(x, t, y); break;               // Hr(x), tabulated
            case 22: { double d0, d1; saddle::eval_direct(x, y, d0, t, d1); break; }  // Gr(x), direct solve
            case 23: { double d0, d1; saddle::eval_direct(x, t, d0, y, d1); break; }  // Hr(x), direct solve
            default: y = fm::div(1.0, x); break;
        }
        out[i] = y;
    }
}

int
launch_test_math(int which, const double* d_in, size_t n, double* d_out, cudaStream_t s)
{
    if (n == 0) return 0;
    if (int rc = device_init()) return rc;  // (the saddle tables)
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
