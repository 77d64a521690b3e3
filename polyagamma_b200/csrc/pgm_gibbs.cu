// pgm_gibbs.cu -- the logistic-regression Gibbs omega-step around the Polya-Gamma fill
// (the caller pattern of BASELINE cfg2; Polson, Scott & Windle 2013, sec. 3):
//
//     psi = X beta                         xbeta_kernel          (HBM-bound: X read once)
//     omega_i ~ PG(n_i, psi_i)             launch_fill, hybrid   (the hot path of this library)
//     G = X^T diag(omega) X                weighted_gram_kernel  (+ gram_reduce_kernel)
//
// X is n x p, row-major with leading dimension ldx, p <= 64.  omega is drawn with exactly the
// stream discipline of pgm_cuda_random_polyagamma_fill2 on (n_trials, psi): element i uses
// (seed, offset, first_index + i), so a row shard reproduces its part of the single-GPU draw.
// G is reduced in a fixed order (per-block partial sums, then one pass over the blocks): the
// result does not depend on scheduling.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "pgm_internal.h"

namespace pgm {

constexpr int kGibbsMaxP = 64;

// psi[r] = sum_j X[r][j] beta[j].  W = lanes per row (power of two >= p, capped at 32): a warp
// reads 32 / W consecutive rows at once (coalesced), each lane multiplies its columns and the
// partial sums are folded inside the W-lane segment.  Eight row groups per trip keep eight
// independent loads per lane in flight.
template <int W>
__global__ void __launch_bounds__(256)
xbeta_kernel(const double* __restrict__ X, size_t n, int p, size_t ldx, const double* __restrict__ beta,
             double* __restrict__ psi)
{
    constexpr int kRowsPerWarp = 32 / W;
    constexpr int kUnroll = 8;
    const int lane = threadIdx.x & 31;
    const int sub = lane / W, col0 = lane % W;
    const double b0 = col0 < p ? beta[col0] : 0.0;
    const double b1 = (col0 + 32 < p) ? beta[col0 + 32] : 0.0;  // only reachable when W == 32
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r0 = warp * (kRowsPerWarp * kUnroll); r0 < n; r0 += nwarps * (kRowsPerWarp * kUnroll)) {
        double x0[kUnroll], x1[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const size_t r = r0 + u * kRowsPerWarp + sub;
            x0[u] = (r < n && col0 < p) ? __ldcs(X + r * ldx + col0) : 0.0;
            x1[u] = (W == 32 && r < n && col0 + 32 < p) ? __ldcs(X + r * ldx + col0 + 32) : 0.0;
        }
        double v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = (W == 32) ? fma(x1[u], b1, x0[u] * b0) : x0[u] * b0;
        // fold over the W lanes of a row.  While a lane still holds more than one row group, a level
        // is a transposing exchange -- the lane keeps one half of its values and receives the partner's
        // sums for them -- so the eight folds cost 4 + 2 + 1 shuffles instead of 8 per level.
        int cnt = kUnroll, ubase = 0, same = 0;
#pragma unroll
        for (int off = W / 2; off > 0; off >>= 1) {
            if (cnt > 1) {
                const int half = cnt / 2;
                const bool up = (col0 & off) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const double send = up ? v[i] : v[i + half];
                    const double keep = up ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
                if (up) ubase += half;
                cnt = half;
            } else {
                v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
                same |= off;
            }
        }
        if ((col0 & same) == 0) {
#pragma unroll
            for (int i = 0; i < cnt; ++i) {
                const size_t r = r0 + (size_t)(ubase + i) * kRowsPerWarp + sub;
                if (r < n) psi[r] = v[i];
            }
        }
    }
}

// ---- G = X^T diag(omega) X ------------------------------------------------------------------
// A rank-n update in FP64: this is the one GEMM-shaped piece of the path, so it runs on the FP64
// tensor-core instruction (DMMA m8n8k4; measured 37.0 TFLOP/s on a B200, the same peak as DFMA,
// at 1/8 of the issue slots and 1/4 of the shared-memory traffic of a register-tiled DFMA kernel,
// which was shared-memory bound at 10 TFLOP/s -- profiles/r1h_gram64_dfma.txt).
//
// The columns are padded to PP = 8 NB and G is cut into 8 x 8 tiles; a WARP owns all NB (NB + 1) / 2
// tiles on or above the diagonal (2 accumulators per tile and lane).  Each warp runs its own
// kGramStages-deep cp.async pipeline over chunks of RW rows in a private slice of shared memory (no
// block-wide barrier in the main loop); rows past n and columns past p are zero-filled by the copy.
// Per 4 rows a lane loads ONE fragment element per column block, F_b = X[r0 + q][8 b + g] with
// g = lane / 4, q = lane % 4 -- it is the B fragment of column block b and, times omega[r0 + q], the
// A fragment of row block b -- and issues one DMMA per tile: NB loads for NB (NB + 1) / 2 DMMAs.
// The row stride in shared memory is PP + 4 doubles, which spreads the 4 rows x 4 doubles a
// half-warp reads over all 32 banks.
constexpr int kGramStages = 3;
constexpr int kGramWarps = 8;
constexpr int kGramStageBytes = 8 * 1024;  // per warp and stage (half of it, two blocks per SM, for NB <= 4)

struct GramShape {
    int NB, ntiles, S, RW, blocks_per_sm;
    size_t stage_doubles;  // per warp and stage: RW * S + RW
    size_t smem_bytes;
};

static GramShape
gram_shape(int p)
{
    GramShape g;
    g.NB = (p + 7) / 8;
    g.ntiles = g.NB * (g.NB + 1) / 2;
    g.S = 8 * g.NB + 4;
    g.blocks_per_sm = g.NB <= 4 ? 2 : 1;
    g.RW = std::max(4, (kGramStageBytes / g.blocks_per_sm / (g.S * 8)) & ~3);
    g.stage_doubles = (size_t)g.RW * g.S + g.RW;
    g.smem_bytes = std::max((size_t)kGramWarps * kGramStages * g.stage_doubles, (size_t)g.ntiles * 64) * sizeof(double);
    return g;
}

__device__ __forceinline__ void
cp_async16(void* dst, const void* src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void
cp_async8(void* dst, const void* src, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void
dmma_8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NB, bool VEC16>
__global__ void __launch_bounds__(kGramWarps * 32, NB <= 4 ? 2 : 1)
weighted_gram_kernel(const double* __restrict__ X, size_t n, int p, size_t ldx, const double* __restrict__ omega,
                     double* __restrict__ partial, GramShape gs)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int PP = 8 * NB, S = PP + 4, NTILES = NB * (NB + 1) / 2;
    constexpr int U = VEC16 ? PP / 2 : PP;  // copy units per row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int RW = gs.RW;
    double* mine = smem + (size_t)warp * kGramStages * gs.stage_doubles;

    const size_t nchunks = (n + RW - 1) / RW;
    auto issue = [&](size_t chunk, int stage) {
        double* xs = mine + (size_t)stage * gs.stage_doubles;
        double* ws = xs + RW * S;
        if (chunk < nchunks) {
            const size_t r0 = chunk * RW;
            for (int u = lane; u < RW * U; u += 32) {
                const int rr = u / U, cc = (u - rr * U) * (VEC16 ? 2 : 1);
                const size_t r = r0 + rr;
                int bytes = 0;
                if (r < n && cc < p) bytes = VEC16 ? (cc + 1 < p ? 16 : 8) : 8;
                const double* src = bytes ? X + r * ldx + cc : X;
                if (VEC16) cp_async16(xs + rr * S + cc, src, bytes);
                else cp_async8(xs + rr * S + cc, src, bytes);
            }
            for (int rr = lane; rr < RW; rr += 32) {
                const size_t r = r0 + rr;
                cp_async8(ws + rr, r < n ? omega + r : omega, r < n ? 8 : 0);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    double acc[NTILES][2];
#pragma unroll
    for (int t = 0; t < NTILES; ++t) acc[t][0] = acc[t][1] = 0.0;

    const size_t nwarps = (size_t)gridDim.x * kGramWarps;
    size_t chunk = (size_t)blockIdx.x * kGramWarps + warp;
    for (int s = 0; s < kGramStages - 1; ++s) issue(chunk + s * nwarps, s);
    for (int it = 0; chunk < nchunks; chunk += nwarps, ++it) {
        issue(chunk + (kGramStages - 1) * nwarps, (it + kGramStages - 1) % kGramStages);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(kGramStages - 1) : "memory");
        __syncwarp();
        const double* xs = mine + (size_t)(it % kGramStages) * gs.stage_doubles;
        const double* ws = xs + RW * S;
        const double* frag = xs + q * S + g;
#pragma unroll 1
        for (int r0 = 0; r0 < RW; r0 += 4) {
            const double w = ws[r0 + q];
            double f[NB], a[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) f[b] = frag[r0 * S + 8 * b];
#pragma unroll
            for (int b = 0; b < NB; ++b) a[b] = w * f[b];
            int t = 0;
#pragma unroll
            for (int bj = 0; bj < NB; ++bj)
#pragma unroll
                for (int bk = bj; bk < NB; ++bk, ++t) dmma_8x8x4(acc[t][0], acc[t][1], a[bj], f[bk]);
        }
        __syncwarp();
    }
    // fold the warps in a fixed order through shared memory; the block's partial sums are stored
    // tile-major in fragment order: partial[block][tile][lane][2]
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    for (int w = 1; w < kGramWarps; ++w) {
        __syncthreads();
        if (warp == w) {
#pragma unroll
            for (int t = 0; t < NTILES; ++t) {
                smem[t * 64 + lane * 2] = acc[t][0];
                smem[t * 64 + lane * 2 + 1] = acc[t][1];
            }
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int t = 0; t < NTILES; ++t) {
                acc[t][0] += smem[t * 64 + lane * 2];
                acc[t][1] += smem[t * 64 + lane * 2 + 1];
            }
        }
    }
    if (warp == 0) {
        double* dst = partial + (size_t)blockIdx.x * NTILES * 64;
#pragma unroll
        for (int t = 0; t < NTILES; ++t) {
            dst[t * 64 + lane * 2] = acc[t][0];
            dst[t * 64 + lane * 2 + 1] = acc[t][1];
        }
    }
}

// G = sum over blocks, in block order.  Accumulator c of lane (g, q) of tile (bj, bk) is
// G[8 bj + g][8 bk + 2 q + c]; only entries on or above the diagonal are taken and mirrored, so G is
// exactly symmetric.
__global__ void
gram_reduce_kernel(const double* __restrict__ partial, int nblocks, int p, int NB, double* __restrict__ gram)
{
    const int ntiles = NB * (NB + 1) / 2;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ntiles * 64) return;
    int bj = 0, bk = e >> 6;
    while (bk >= NB - bj) {
        bk -= NB - bj;
        ++bj;
    }
    bk += bj;
    const int lane = (e >> 1) & 31;
    const int j = 8 * bj + (lane >> 2), k = 8 * bk + 2 * (lane & 3) + (e & 1);
    if (j > k || k >= p) return;
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partial[(size_t)b * ntiles * 64 + e];
    gram[(size_t)j * p + k] = v;
    gram[(size_t)k * p + j] = v;
}

template <int NB>
static void
launch_gram(bool vec16, int blocks, const GramShape& gs, const double* X, size_t n, int p, size_t ldx,
            const double* omega, double* partial, cudaStream_t s)
{
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs.smem_bytes);
        kern<<<blocks, kGramWarps * 32, gs.smem_bytes, s>>>(X, n, p, ldx, omega, partial, gs);
    };
    if (vec16) go(weighted_gram_kernel<NB, true>);
    else go(weighted_gram_kernel<NB, false>);
}

int
launch_logistic_omega_step(uint64_t seed, uint64_t offset, const double* d_X, size_t n, int p, size_t ldx,
                           const double* d_beta, const double* d_ntrials, double ntrials_scalar, int method,
                           double* d_gram, double* d_omega, double* d_psi, uint64_t first_index, cudaStream_t s)
{
    if (p < 1 || p > kGibbsMaxP || ldx < (size_t)p) return -1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t err;
    if (n == 0) {
        return (int)cudaMemsetAsync(d_gram, 0, sizeof(double) * p * p, s);
    }
    const bool vec16 = (ldx % 2 == 0) && (((uintptr_t)d_X) % 16 == 0);
    const GramShape gs = gram_shape(p);
    const int gram_blocks = (int)std::min<size_t>((size_t)sms * gs.blocks_per_sm,
                                                 ((n + gs.RW - 1) / gs.RW + kGramWarps - 1) / kGramWarps);
    char* ws = nullptr;
    const size_t omega_bytes = d_omega ? 0 : ((n * sizeof(double) + 255) & ~size_t(255));
    const size_t psi_bytes = d_psi ? 0 : ((n * sizeof(double) + 255) & ~size_t(255));
    const size_t part_bytes = (size_t)gram_blocks * gs.ntiles * 64 * sizeof(double);
    if ((err = (cudaError_t)scratch_alloc((void**)&ws, omega_bytes + psi_bytes + part_bytes, s)) != cudaSuccess) return (int)err;
    double* omega = d_omega ? d_omega : (double*)ws;
    double* psi = d_psi ? d_psi : (double*)(ws + omega_bytes);
    double* partial = (double*)(ws + omega_bytes + psi_bytes);

    {
        ScopedKernelTimer timer(kKindGibbs, s);
        const uint32_t grid = (uint32_t)std::min<size_t>((size_t)sms * 8, (n + 63) / 64);
        if (p <= 4) xbeta_kernel<4><<<grid, 256, 0, s>>>(d_X, n, p, ldx, d_beta, psi);
        else if (p <= 8) xbeta_kernel<8><<<grid, 256, 0, s>>>(d_X, n, p, ldx, d_beta, psi);
        else if (p <= 16) xbeta_kernel<16><<<grid, 256, 0, s>>>(d_X, n, p, ldx, d_beta, psi);
        else xbeta_kernel<32><<<grid, 256, 0, s>>>(d_X, n, p, ldx, d_beta, psi);
        count_launch(1);
    }
    FillArgs a;
    memset(&a, 0, sizeof(a));
    a.key = derive_call_key(seed, offset);
    a.out = omega;
    a.first_index = first_index;
    a.h = d_ntrials;
    a.h_scalar = ntrials_scalar;
    a.h_stride = 1;
    a.z = psi;
    a.z_stride = 1;
    int rc = launch_fill(a, n, method, s);
    if (rc != 0) {
        cudaFreeAsync(ws, s);
        return rc;
    }
    {
        ScopedKernelTimer timer(kKindGibbs + 1, s);
        switch (gs.NB) {
        case 1: launch_gram<1>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        case 2: launch_gram<2>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        case 3: launch_gram<3>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        case 4: launch_gram<4>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        case 5: launch_gram<5>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        case 6: launch_gram<6>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        case 7: launch_gram<7>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        default: launch_gram<8>(vec16, gram_blocks, gs, d_X, n, p, ldx, omega, partial, s); break;
        }
        const int nred = gs.ntiles * 64;
        gram_reduce_kernel<<<(nred + 127) / 128, 128, 0, s>>>(partial, gram_blocks, p, gs.NB, d_gram);
        count_launch(2);
    }
    cudaFreeAsync(ws, s);
    return (int)cudaGetLastError();
}

}  // namespace pgm
