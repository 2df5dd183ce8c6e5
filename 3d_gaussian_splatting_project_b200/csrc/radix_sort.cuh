// Stable least-significant-digit radix sort of (uint32 key, uint32 payload) pairs, hand written
// for sm_100a (no library).  One pass = histogram kernel + row-scan kernel + scatter kernel:
//
//   tile     = THREADS x ITEMS consecutive pairs; warp w of the tile owns a contiguous run of
//              32 x ITEMS pairs and walks it in rounds of 32 consecutive pairs (lane = pair), so
//              loads are fully coalesced and the order inside a tile is (warp, round, lane).
//   ranking  = in a round, __match_any_sync groups the lanes holding the same digit; a lane's
//              rank is the per-warp running count of its digit plus the number of lower lanes
//              of its group -- no atomics, and the rank order is the input order (stability).
//   scatter  = through shared memory: the tile is put in sorted order there and written out by
//              consecutive threads, so runs of one digit leave as contiguous stores.
//   bases    = hist[digit][tile]: one CTA per digit scans its row and adds the number of pairs with
//              a smaller digit (digit totals are accumulated by the histogram kernel).
//
// Used by the viewer-side depth sort (viewer.cu: 17-bit bucket keys, two 9-bit passes) and by
// the k-nearest-neighbour grid of region_growing.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace gsl {
namespace radix {

constexpr int kThreads = 512;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;      // 8192 pairs per tile
constexpr int kWarps = kThreads / 32;

inline int n_tiles(int64_t N) { return (int)((N + kTile - 1) / kTile); }
// uint32 words of histogram scratch one pass of `bits` bits needs.
inline size_t hist_words(int64_t N, int bits) { return ((size_t)n_tiles(N) << bits) + ((size_t)1 << bits); }

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Per-warp digit counts of the tile's rounds; on return cnt[warp][d] holds the warp's total of
// digit d and rank[r] the rank of the lane's pair of round r among the warp's pairs of that digit.
template <int BITS>
__device__ __forceinline__ void rank_warp_run(const uint32_t *__restrict__ keys, int64_t N, int64_t warp_begin,
                                              int shift, uint32_t *cnt_warp, uint32_t (&key)[kItems],
                                              uint32_t (&rank)[kItems])
{
    constexpr uint32_t DIG = 1u << BITS;
    const unsigned lane = threadIdx.x & 31u, lt = lanemask_lt();
    // all loads first: the ranking rounds below are separated by warp barriers, which would otherwise
    // put one memory round trip in front of every round
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = warp_begin + r * 32 + lane;
        key[r] = i < N ? keys[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = warp_begin + r * 32 + lane;
        const bool valid = i < N;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        rank[r] = 0;
        if (valid) {
            const uint32_t d = (key[r] >> shift) & (DIG - 1u);
            // lanes holding the same digit: one ballot per digit bit (measured faster than match.any here)
            unsigned peers = act;
#pragma unroll
            for (int b = 0; b < BITS; ++b) {
                const unsigned bal = __ballot_sync(act, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? bal : ~bal;
            }
            const uint32_t pre = cnt_warp[d];
            __syncwarp(act);
            if ((peers & lt) == 0) cnt_warp[d] = pre + __popc(peers);
            __syncwarp(act);
            rank[r] = pre + __popc(peers & lt);
        }
    }
}

template <int BITS>
__global__ void __launch_bounds__(kThreads) histogram_kernel(const uint32_t *__restrict__ keys, int64_t N, int shift,
                                                             uint32_t *__restrict__ hist, uint32_t *__restrict__ total, int tiles)
{
    constexpr int DIG = 1 << BITS;
    // two private copies (even / odd warps) halve the collisions of the shared-memory atomics
    __shared__ uint32_t cnt[2][DIG];
    for (int d = threadIdx.x; d < 2 * DIG; d += kThreads) (&cnt[0][0])[d] = 0;
    __syncthreads();
    const int tile = blockIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int64_t warp_begin = (int64_t)tile * kTile + (int64_t)warp * 32 * kItems;
    uint32_t *mine = cnt[warp & 1];
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = warp_begin + r * 32 + lane;
        if (i < N) atomicAdd(&mine[(keys[i] >> shift) & (uint32_t)(DIG - 1)], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < DIG; d += kThreads) {
        const uint32_t c = cnt[0][d] + cnt[1][d];
        hist[(size_t)d * tiles + tile] = c;
        if (c) atomicAdd(&total[d], c);
    }
}

// hist[d][0..tiles) <- exclusive scan of the row, offset by the number of pairs with a smaller digit.
// One CTA per digit.
template <int BITS>
__global__ void __launch_bounds__(256) scan_rows_kernel(uint32_t *__restrict__ hist, const uint32_t *__restrict__ total, int tiles)
{
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t carry;
    const int d = blockIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t below = 0;
    for (int j = threadIdx.x; j < d; j += 256) below += total[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) warp_sum[warp] = below;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t b = 0;
        for (int w = 0; w < 8; ++w) b += warp_sum[w];
        carry = b;
    }
    __syncthreads();
    uint32_t *row = hist + (size_t)d * tiles;
    for (int base = 0; base < tiles; base += 256) {
        const int i = base + (int)threadIdx.x;
        const uint32_t x = i < tiles ? row[i] : 0u;
        uint32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (unsigned)o) inc += t;
        }
        __syncthreads();                      // previous round's warp_sum / carry reads are done
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        uint32_t before = carry;
        for (unsigned w = 0; w < warp; ++w) before += warp_sum[w];
        if (i < tiles) row[i] = before + inc - x;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + inc;
    }
}

// keys_out / pay_out may be NULL (not needed after the last pass / keys-only).  pay_in == NULL
// means the payload is the pair's own position (first pass of an index sort).  A pair whose FULL
// key equals drop_key gets `drop_payload` instead of its payload (viewer.cu: the typed-array
// out-of-range quirk of the reference's counting sort); pass 0xffffffff to disable.
//
// The tile's pairs are first put in sorted order in shared memory (position = first slot of the
// digit in the tile + pairs of that digit in lower warps + rank inside the warp) and then written
// out by consecutive threads: a digit's run inside a tile is contiguous in the output, so a warp's
// store covers a few sectors instead of one per lane.
template <int BITS>
struct ScatterSmem {
    static constexpr int DIG = 1 << BITS;
    uint32_t cnt[kWarps][DIG];      // per-warp digit counts, then pairs of the digit in lower warps
    uint32_t lstart[DIG];           // first slot of the digit in the tile's sorted order
    uint32_t gdelta[DIG];           // output position of that slot - lstart
    uint32_t wsum[kWarps];
    uint32_t key[kTile];
    uint32_t pay[kTile];
};

template <int BITS>
__global__ void __launch_bounds__(kThreads) scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ pay_in,
                                                           uint32_t *__restrict__ keys_out, uint32_t *__restrict__ pay_out,
                                                           int64_t N, int shift, const uint32_t *__restrict__ base, int tiles,
                                                           uint32_t drop_key, uint32_t drop_payload)
{
    constexpr int DIG = 1 << BITS;
    static_assert(DIG <= kThreads, "one thread per digit");
    extern __shared__ __align__(16) unsigned char radix_smem[];
    ScatterSmem<BITS> &S = *reinterpret_cast<ScatterSmem<BITS> *>(radix_smem);
    for (int d = threadIdx.x; d < kWarps * DIG; d += kThreads) (&S.cnt[0][0])[d] = 0;
    __syncthreads();
    const int tile = blockIdx.x;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int64_t tile_begin = (int64_t)tile * kTile;
    const int64_t warp_begin = tile_begin + (int64_t)warp * 32 * kItems;
    const int tile_n = (int)(N - tile_begin < kTile ? N - tile_begin : kTile);
    uint32_t key[kItems], rank[kItems];
    rank_warp_run<BITS>(keys, N, warp_begin, shift, S.cnt[warp], key, rank);
    __syncthreads();
    // thread d: pairs of digit d in the tile, and per warp the pairs in lower warps
    uint32_t mine = 0;
    if (threadIdx.x < DIG) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = S.cnt[w][threadIdx.x];
            S.cnt[w][threadIdx.x] = mine;
            mine += c;
        }
    }
    // exclusive scan of `mine` over the digits (one per thread)
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) S.wsum[warp] = inc;
    __syncthreads();
    if (threadIdx.x < DIG) {
        uint32_t before = 0;
        for (unsigned w = 0; w < warp; ++w) before += S.wsum[w];
        const uint32_t ls = before + inc - mine;
        S.lstart[threadIdx.x] = ls;
        S.gdelta[threadIdx.x] = base[(size_t)threadIdx.x * tiles + tile] - ls;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const int64_t i = warp_begin + r * 32 + lane;
        if (i < N) {
            const uint32_t d = (key[r] >> shift) & (uint32_t)(DIG - 1);
            const uint32_t at = S.lstart[d] + S.cnt[warp][d] + rank[r];
            S.key[at] = key[r];
            if (pay_out) {
                uint32_t p = pay_in ? pay_in[i] : (uint32_t)i;
                if (key[r] == drop_key) p = drop_payload;
                S.pay[at] = p;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tile_n; i += kThreads) {
        const uint32_t k = S.key[i];
        const uint32_t at = S.gdelta[(k >> shift) & (uint32_t)(DIG - 1)] + (uint32_t)i;
        if (keys_out) keys_out[at] = k;
        if (pay_out) pay_out[at] = S.pay[i];
    }
}

// One stable pass on `st`.  hist: hist_words(N, BITS) words of scratch.
template <int BITS>
inline int pass(const uint32_t *keys, const uint32_t *pay_in, uint32_t *keys_out, uint32_t *pay_out, int64_t N,
                int shift, uint32_t *hist, uint32_t drop_key, uint32_t drop_payload, cudaStream_t st)
{
    if (N <= 0) return GSL_OK;
    const int tiles = n_tiles(N);
    uint32_t *total = hist + ((size_t)tiles << BITS);
    GSL_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(uint32_t) << BITS, st));
    histogram_kernel<BITS><<<tiles, kThreads, 0, st>>>(keys, N, shift, hist, total, tiles);
    GSL_LAUNCH_CHECK("radix::histogram_kernel");
    scan_rows_kernel<BITS><<<1 << BITS, 256, 0, st>>>(hist, total, tiles);
    GSL_LAUNCH_CHECK("radix::scan_rows_kernel");
    GSL_CUDA_TRY(cudaFuncSetAttribute(scatter_kernel<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem<BITS>)));
    scatter_kernel<BITS><<<tiles, kThreads, sizeof(ScatterSmem<BITS>), st>>>(keys, pay_in, keys_out, pay_out, N, shift, hist, tiles,
                                                                             drop_key, drop_payload);
    GSL_LAUNCH_CHECK("radix::scatter_kernel");
    return GSL_OK;
}

}  // namespace radix
}  // namespace gsl
