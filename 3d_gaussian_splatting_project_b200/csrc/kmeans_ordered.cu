// Reference-order centroid update for sm_100a.
//
// NumPy evaluates `data[labels == c].mean(axis=0)` (3D_clustering/k_means.py:125-128) as a
// float32 SEQUENTIAL sum over the members in index order, then one float64 division rounded
// to float32 (see oracle/gsl_oracle.c: orc_kmeans_update).  Float32 addition is not
// associative, so reproducing the reference bit for bit means keeping that order: one serial
// chain per (cluster, dimension).  The K*D chains are independent, so the work is:
//
//   ordered_count_kernel     per 1024-row tile, member count per cluster
//   ordered_scan_kernel      per cluster, exclusive scan of the tile counts (+ totals)
//   ordered_start_kernel     exclusive scan of the totals -> segment starts
//   ordered_scatter_kernel   stable scatter of row indices into per-cluster member lists
//                            (warp match_any ranks keep ascending row order)
//   ordered_chain_kernel     CTA per cluster: consumer warps (lane = dimension) add the member rows in
//                            order out of a shared-memory ring that the producer warps fill, batches
//                            handed over with mbarriers
//   shift_kernel             ||new - old||_F
//
// This is the parity mode (single device).  The throughput mode is gsl_kmeans_step's float64
// segmented reduction, which is sharded and all-reduced.
#include "common.cuh"

namespace gsl {

constexpr int kOrdTile = 1024;
constexpr int kOrdThreads = 256;

__global__ void __launch_bounds__(kOrdThreads)
ordered_count_kernel(const int32_t *__restrict__ labels, int64_t N, int K, int32_t *__restrict__ tile_counts,
                     int *__restrict__ bad_label)
{
    extern __shared__ int hist[];
    for (int i = threadIdx.x; i < K; i += kOrdThreads) hist[i] = 0;
    __syncthreads();
    const int64_t row0 = (int64_t)blockIdx.x * kOrdTile;
    for (int r = threadIdx.x; r < kOrdTile; r += kOrdThreads) {
        const int64_t row = row0 + r;
        if (row < N) {
            const int lab = labels[row];
            if ((unsigned)lab < (unsigned)K) atomicAdd(&hist[lab], 1);
            else *bad_label = 1;                         // skipped here and in the scatter; the shift comes back NaN
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += kOrdThreads) tile_counts[(size_t)blockIdx.x * K + i] = hist[i];
}

// Block k: tile_counts[:, k] -> exclusive offsets in place, total[k].
__global__ void __launch_bounds__(kOrdThreads)
ordered_scan_kernel(int32_t *__restrict__ tile_counts, int n_tiles, int K, int64_t *__restrict__ total)
{
    __shared__ int64_t part[kOrdThreads];
    const int k = blockIdx.x, t = threadIdx.x;
    const int seg = (n_tiles + kOrdThreads - 1) / kOrdThreads;
    const int lo = min(t * seg, n_tiles), hi = min(lo + seg, n_tiles);
    int64_t s = 0;
    for (int i = lo; i < hi; ++i) s += tile_counts[(size_t)i * K + k];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < kOrdThreads; ++i) { const int64_t v = part[i]; part[i] = run; run += v; }
        total[k] = run;
    }
    __syncthreads();
    int64_t run = part[t];
    for (int i = lo; i < hi; ++i) {
        const int c = tile_counts[(size_t)i * K + k];
        tile_counts[(size_t)i * K + k] = (int32_t)run;   // within-cluster offset (< N < 2^31)
        run += c;
    }
}

__global__ void ordered_start_kernel(const int64_t *__restrict__ total, int K, int64_t *__restrict__ start)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int64_t run = 0;
        for (int k = 0; k < K; ++k) { start[k] = run; run += total[k]; }
    }
}

__global__ void __launch_bounds__(kOrdThreads)
ordered_scatter_kernel(const int32_t *__restrict__ labels, int64_t N, int K,
                       const int32_t *__restrict__ tile_offsets, const int64_t *__restrict__ start,
                       int32_t *__restrict__ members)
{
    extern __shared__ int wc[];                 // [8][K] counts, then absolute bases (as int64 would
    int64_t *base = reinterpret_cast<int64_t *>(wc + 8 * K + ((8 * K) & 1));  // overflow int): [8][K] int64
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int i = t; i < 8 * K; i += kOrdThreads) wc[i] = 0;
    __syncthreads();
    const int64_t row0 = (int64_t)blockIdx.x * kOrdTile + w * 128;
    int lab[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int64_t row = row0 + c * 32 + lane;
        lab[c] = row < N ? labels[row] : -1;
        if ((unsigned)lab[c] >= (unsigned)K) lab[c] = -1;        // out-of-range labels belong to no cluster
        if (lab[c] >= 0) atomicAdd(&wc[w * K + lab[c]], 1);
    }
    __syncthreads();
    for (int k = t; k < K; k += kOrdThreads) {
        int64_t run = start[k] + tile_offsets[(size_t)blockIdx.x * K + k];
        for (int ww = 0; ww < 8; ++ww) { base[ww * K + k] = run; run += wc[ww * K + k]; }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int key = lab[c] >= 0 ? lab[c] : -1 - lane;     // invalid rows never match
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (lab[c] >= 0) {
            const int rank = __popc(m & ((1u << lane) - 1u));
            const int64_t b = base[w * K + lab[c]];
            members[b + rank] = (int32_t)(row0 + c * 32 + lane);
        }
        __syncwarp();
        if (lab[c] >= 0 && lane == __ffs(m) - 1) base[w * K + lab[c]] += __popc(m);
        __syncwarp();
    }
}

// One CTA of 512 threads per cluster.  The float32 sum of a cluster's members is ONE dependent
// chain per dimension (4 cycles per add): nothing can shorten it, so everything else is arranged
// never to make it wait.
//   consumers  ceil(D / 32) warps, lane = dimension: they own the accumulators and add member rows
//              strictly in index order, straight out of a ring of row batches in shared memory;
//   producers  the other warps: warp p fetches batches p, p + P, ... -- 32 member indices in one
//              coalesced load, then the 32 rows (every row a couple of coalesced 128-byte loads, all
//              of them in flight before the first store), written into the batch's ring slot.
// A batch is handed over with a pair of mbarriers (full / empty) per ring slot, so the consumers
// never wait on a block barrier and a producer that is merely issuing loads never holds them up
// (round 1 used __syncthreads per batch and 4-byte cp.async: 5.1 ms at 6 M x 59, K = 64, against a
// chain floor of 1.5 ms for the largest cluster).
constexpr int kChainThreads = 512;
constexpr int kChainBatch = 32;        // rows per batch == rows a producer warp fetches at a time
constexpr int kChainSlots = 16;        // batches in the ring (fewer when D > 96: the ring must fit shared memory)

__device__ __forceinline__ uint32_t chain_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void chain_bar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(chain_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void chain_bar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(chain_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void chain_bar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "CW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CD_%=;\n\t"
        "bra CW_%=;\n\t"
        "CD_%=:\n\t"
        "}" ::"r"(chain_smem_u32(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(kChainThreads, 1)
ordered_chain_kernel(const float *__restrict__ data, int D, const int32_t *__restrict__ members,
                     const int64_t *__restrict__ start, const int64_t *__restrict__ total,
                     const float *__restrict__ old_c, float *__restrict__ new_c, int n_slots)
{
    extern __shared__ __align__(16) float ring[];          // [n_slots][kChainBatch][dp]
    __shared__ unsigned long long full_bar[kChainSlots], empty_bar[kChainSlots];
    const int k = blockIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_cons = (D + 31) / 32, n_prod = kChainThreads / 32 - n_cons;
    const int dp = n_cons * 32;                            // floats per row slot
    const int64_t n = total[k], s0 = start[k];
    if (n == 0) {                                          // km:126 else-branch: an empty cluster keeps its centroid
        for (int d = threadIdx.x; d < D; d += kChainThreads) new_c[(size_t)k * D + d] = old_c[(size_t)k * D + d];
        return;
    }
    if (threadIdx.x < n_slots) {
        chain_bar_init(&full_bar[threadIdx.x], 32);                        // the 32 lanes of the producer warp that filled it
        chain_bar_init(&empty_bar[threadIdx.x], 32u * (unsigned)n_cons);   // every consumer lane
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int64_t n_batches = (n + kChainBatch - 1) / kChainBatch;
    if (w < n_cons) {
        // ===== consumer: dimension d, batches in order =====
        const int d = w * 32 + lane;
        float acc = -0.0f;              // (-0) + x == x for every x: same as starting from the first row
        for (int64_t b = 0; b < n_batches; ++b) {
            const int slot = (int)(b % n_slots);
            chain_bar_wait(&full_bar[slot], (unsigned)((b / n_slots) & 1));
            const int cnt = (int)min((int64_t)kChainBatch, n - b * kChainBatch);
            const float *src = ring + ((size_t)slot * kChainBatch) * dp + d;
            if (cnt == kChainBatch) {
                float x[kChainBatch];
#pragma unroll
                for (int j = 0; j < kChainBatch; ++j) x[j] = src[j * dp];  // all loads first: the chain never waits for shared memory
#pragma unroll
                for (int j = 0; j < kChainBatch; ++j) acc = __fadd_rn(acc, x[j]);
            } else {
                for (int j = 0; j < cnt; ++j) acc = __fadd_rn(acc, src[j * dp]);
            }
            chain_bar_arrive(&empty_bar[slot]);
        }
        if (d < D) new_c[(size_t)k * D + d] = (float)((double)acc / (double)n);
    } else {
        // ===== producer =====
        const int p = w - n_cons;
        for (int64_t b = p; b < n_batches; b += n_prod) {
            const int slot = (int)(b % n_slots);
            const int64_t use = b / n_slots;
            const int64_t mi = b * kChainBatch + lane;
            const int my_idx = mi < n ? __ldg(members + s0 + mi) : -1;     // before the wait: in flight while the slot drains
            if (use > 0) chain_bar_wait(&empty_bar[slot], (unsigned)((use - 1) & 1));
            float *dst = ring + ((size_t)slot * kChainBatch) * dp + lane;
            // 8 rows at a time: all their loads are issued before the first store
#pragma unroll 1
            for (int j0 = 0; j0 < kChainBatch; j0 += 8) {
                float v[8][8];
                int idx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    idx[j] = __shfl_sync(0xffffffffu, my_idx, j0 + j);
                    const float *row = data + (size_t)(idx[j] < 0 ? 0 : idx[j]) * D;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < n_cons) v[j][c] = (idx[j] >= 0 && c * 32 + lane < D) ? __ldg(row + c * 32 + lane) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < n_cons) dst[(size_t)(j0 + j) * dp + c * 32] = v[j][c];
            }
            chain_bar_arrive(&full_bar[slot]);
        }
    }
}

__global__ void __launch_bounds__(256)
shift_kernel(const float *__restrict__ a, const float *__restrict__ b, int n, float *__restrict__ shift,
             const int *__restrict__ bad_label)
{
    __shared__ double red[8];
    double sq = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float df = a[i] - b[i];
        sq += (double)df * (double)df;
    }
    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        *shift = (bad_label && *bad_label) ? __int_as_float(0x7fc00000) : (float)sqrt(s);
    }
}

struct OrdWs { size_t tile_counts, total, start, members, bytes; };

static OrdWs ordered_layout(int64_t N, int K)
{
    OrdWs o;
    const size_t n_tiles = (size_t)((N + kOrdTile - 1) / kOrdTile);
    size_t off = 256;
    o.tile_counts = off; off += align_up(n_tiles * K * sizeof(int32_t), 256);
    o.total = off;       off += align_up((size_t)K * sizeof(int64_t), 256);
    o.start = off;       off += align_up((size_t)K * sizeof(int64_t), 256);
    o.members = off;     off += align_up((size_t)N * sizeof(int32_t), 256);
    o.bytes = off;
    return o;
}

size_t ordered_workspace_bytes(int64_t N, int K) { return ordered_layout(N, K).bytes; }

}  // namespace gsl

using namespace gsl;

extern "C" int gsl_kmeans_update_ordered(const float *data, const int32_t *labels, int64_t N, int D, int K,
                                         const float *old_centroids, float *new_centroids, float *shift,
                                         void *ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (N < 0 || N >= ((int64_t)1 << 31)) return fail(GSL_EINVAL, "gsl_kmeans_update_ordered: N out of range");
    if (D < 1 || D > GSL_KMEANS_MAX_D || K < 1 || K > GSL_KMEANS_MAX_K) return fail(GSL_EINVAL, "gsl_kmeans_update_ordered: bad K/D");
    if (!old_centroids || !new_centroids || !shift) return fail(GSL_EINVAL, "gsl_kmeans_update_ordered: null pointer");
    if (N > 0 && (!data || !labels || !ws)) return fail(GSL_EINVAL, "gsl_kmeans_update_ordered: null pointer");
    const OrdWs L = ordered_layout(N, K);
    if (ws_bytes < L.bytes) return fail(GSL_EWORKSPACE, "gsl_kmeans_update_ordered: workspace %zu < %zu", ws_bytes, L.bytes);
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    int32_t *tile_counts = reinterpret_cast<int32_t *>(base + L.tile_counts - 256);
    int64_t *total = reinterpret_cast<int64_t *>(base + L.total - 256);
    int64_t *start = reinterpret_cast<int64_t *>(base + L.start - 256);
    int32_t *members = reinterpret_cast<int32_t *>(base + L.members - 256);
    const int n_tiles = (int)((N + kOrdTile - 1) / kOrdTile);
    int *bad_label = reinterpret_cast<int *>(base);          // first word of the workspace header
    GSL_CUDA_TRY(cudaMemsetAsync(bad_label, 0, sizeof(int), st));

    if (N == 0) {
        GSL_CUDA_TRY(cudaMemsetAsync(total, 0, sizeof(int64_t) * (size_t)K, st));
        GSL_CUDA_TRY(cudaMemsetAsync(start, 0, sizeof(int64_t) * (size_t)K, st));
    } else {
        ordered_count_kernel<<<n_tiles, kOrdThreads, K * sizeof(int), st>>>(labels, N, K, tile_counts, bad_label);
        GSL_LAUNCH_CHECK("ordered_count_kernel");
        ordered_scan_kernel<<<K, kOrdThreads, 0, st>>>(tile_counts, n_tiles, K, total);
        GSL_LAUNCH_CHECK("ordered_scan_kernel");
        ordered_start_kernel<<<1, 32, 0, st>>>(total, K, start);
        GSL_LAUNCH_CHECK("ordered_start_kernel");
        const size_t sc_smem = (size_t)(8 * K + ((8 * K) & 1)) * sizeof(int) + (size_t)8 * K * sizeof(int64_t);
        GSL_CUDA_TRY(cudaFuncSetAttribute(ordered_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
        ordered_scatter_kernel<<<n_tiles, kOrdThreads, sc_smem, st>>>(labels, N, K, tile_counts, start, members);
        GSL_LAUNCH_CHECK("ordered_scatter_kernel");
    }
    const size_t batch_bytes = (size_t)kChainBatch * (size_t)((D + 31) / 32 * 32) * sizeof(float);
    int n_slots = (int)((size_t)200 * 1024 / batch_bytes);
    if (n_slots > kChainSlots) n_slots = kChainSlots;
    const size_t ch_smem = (size_t)n_slots * batch_bytes;
    GSL_CUDA_TRY(cudaFuncSetAttribute(ordered_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ch_smem));
    ordered_chain_kernel<<<(unsigned)K, kChainThreads, ch_smem, st>>>(data, D, members, start, total, old_centroids, new_centroids, n_slots);
    GSL_LAUNCH_CHECK("ordered_chain_kernel");
    shift_kernel<<<1, 256, 0, st>>>(new_centroids, old_centroids, K * D, shift, bad_label);
    GSL_LAUNCH_CHECK("shift_kernel");
    return GSL_OK;
}
