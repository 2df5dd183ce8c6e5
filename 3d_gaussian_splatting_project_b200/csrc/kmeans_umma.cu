// K-means assignment + per-cluster sums with the screening contraction on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in tensor memory), sm_100a.  K <= 64, 8 <= D <= 64.
//
// Same three-stage screening as kmeans_tc.cu (every stage with a rigorous bound, so the labels
// are exactly those of scipy's float64 scan, 3D_clustering/k_means.py:116-122):
//   A  x'.c'_k for the 128 rows of a tile and all 64 centroids is ONE 128 x 64 x 64 TF32
//      contraction: eight tcgen05.mma (M = 128, N = 64, K = 8 each) issued by one thread, both
//      operands in shared memory in the canonical K-major layout, the 128 x 64 float32 result in
//      tensor memory, read back with tcgen05.ld (32 lanes x 64 columns per warp: a thread gets
//      the whole row of its own data row, so minimum, threshold and candidate mask need no
//      shuffles).  Operands are rounded to TF32 with cvt.rna when they are laid out, so the bound
//      of kmeans_tc.cu (2^-11 per operand) carries over unchanged.
//   B  float32 distance on the original values for the candidates (1.03 - 1.17 per row),
//   C  scipy-order float64 for near ties.
//
// One persistent CTA per SM, warp specialised:
//   warp 8          producer: the [128][D] row block of a tile is contiguous in memory; one elected
//                   thread fetches it with a TMA bulk copy (cp.async.bulk + mbarrier) into a ring of
//                   three raw tiles
//   warps 0-3, 4-7  two consumer groups of 128 threads (thread = row) that take alternate tiles:
//                   re-lay the raw tile (centred on the centroid mean, TF32-rounded) as operand A,
//                   issue the MMAs, read the accumulator, build the candidate set, refine, write
//                   the labels, sort the tile's rows by label and add them into the CTA's float64
//                   accumulators.  While one group waits for its MMA or walks its clusters, the
//                   other one lays out or screens the next tile.  The accumulators are shared; the
//                   groups add their tiles in tile order (a named-barrier hand-off), so the sums
//                   stay bit-reproducible.
// Rows past the last full tile (and matrices that are not 16-byte aligned) take kmeans_tc.cu.
#include <stdlib.h>

#include "common.cuh"
#include "kmeans_common.cuh"
#include "kmeans_screen.cuh"

namespace gsl {

constexpr int kURows = 128;                     // rows per tile == UMMA M
constexpr int kUGroup = 128;                    // threads per consumer group (thread = row)
constexpr int kUThreads = 2 * kUGroup + 32;     // two groups + the producer warp
constexpr int kURing = 3;                       // raw tiles in flight
constexpr int kUPad = 64;                       // K and D are padded to 64 (UMMA N and K extent)
constexpr int kUABytes = kURows * kUPad * 4;    // operand A of one group
constexpr int kUBBytes = kUPad * kUPad * 4;     // operand B
constexpr int kUTmemCols = 128;                 // two 128 x 64 float32 accumulators

struct USmem {
    size_t raw, a, b, acc, cn2, eab, mean, scratch, bits, cstart, order, bars, tmem, total;
    size_t raw_tile;       // bytes of one raw tile
    size_t bits_g, cstart_g;   // per-group strides of `bits` and `cstart`
};

static inline USmem umma_layout(int D, int K, bool accumulate)
{
    USmem s;
    size_t o = 0;
    s.raw_tile = (size_t)kURows * D * sizeof(float);                  // a multiple of 512 bytes
    s.raw = o;     o += kURing * s.raw_tile;
    s.a = o;       o += 2 * (size_t)kUABytes;
    s.b = o;       o += kUBBytes;
    s.acc = o;     o += accumulate ? align_up((size_t)K * (D + 1) * sizeof(double), 16) : 0;
    s.cn2 = o;     o += kUPad * sizeof(float);
    s.eab = o;     o += 2 * kUPad * sizeof(float);
    s.mean = o;    o += kUPad * sizeof(float);
    s.scratch = o; o += 8 * 256;                                       // per warp: the refinement's pair / distance lists
    s.bits_g = align_up((size_t)K * 4 * sizeof(unsigned), 16);
    s.cstart_g = align_up((size_t)(K + 2) * 2, 16);
    s.bits = o;    o += accumulate ? 2 * s.bits_g : 0;
    s.cstart = o;  o += accumulate ? 2 * s.cstart_g : 0;
    s.order = o;   o += accumulate ? 2 * (size_t)kURows * 2 : 0;
    s.bars = o;    o += 8 * 8;                                         // raw_full[3], raw_empty[3], mma_done[2]
    s.tmem = o;    o += 16;
    s.total = o;
    return s;
}

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_wait_u(void *bar, unsigned parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// same, for a thread that may wait long (the producer): back off between polls
__device__ __forceinline__ void mbar_wait_backoff(void *bar, unsigned parity)
{
    unsigned done = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor of a K-major, non-swizzled operand (canonical layout: 8 rows x
// 16 bytes core matrices; `lbo` = byte distance between the two 16-byte K chunks of one MMA,
// `sbo` = byte distance between consecutive groups of 8 rows).  Bits: [0,14) address >> 4,
// [16,30) lbo >> 4, [32,46) sbo >> 4, [46,48) version = 1, [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ unsigned long long umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo)
{
    return (unsigned long long)((smem_addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor: D = F32 (bits 4-5 = 1), A and B = TF32 (bits 7-9 and 10-12 = 2), both
// K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28.
constexpr uint32_t kUIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kUPad >> 3) << 17) | ((uint32_t)(kURows >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, unsigned long long adesc, unsigned long long bdesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kUIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(void *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One row of the accumulator: 64 consecutive columns of this thread's tensor-memory lane.
__device__ __forceinline__ void tmem_load_row64(uint32_t taddr, float (&v)[64])
{
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// Byte offset of element (row r, 16-byte chunk c) of a K-major operand with `groups` groups of 8 rows.
__device__ __forceinline__ uint32_t operand_offset(int r, int c, int groups) { return (uint32_t)(c * groups * 128 + (r >> 3) * 128 + (r & 7) * 16); }

template <bool kAccumulate, bool kCheck>
__global__ void __launch_bounds__(kUThreads, 1)
kmeans_step_umma_kernel(const float *__restrict__ data, int64_t n_tiles, int D, const float *__restrict__ centroids,
                        int K, int32_t *__restrict__ labels, double *__restrict__ partials, USmem L, float eps,
                        unsigned long long *__restrict__ check_out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float *raw = reinterpret_cast<float *>(smem + L.raw);
    double *acc = reinterpret_cast<double *>(smem + L.acc);
    float *cn2 = reinterpret_cast<float *>(smem + L.cn2);
    float *eab = reinterpret_cast<float *>(smem + L.eab);          // interleaved {ea_k, eb_k}, see screen_bound
    float *mean = reinterpret_cast<float *>(smem + L.mean);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + L.bars);
    void *raw_full = bars, *raw_empty = bars + kURing, *mma_done = bars + 2 * kURing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.tmem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- per-CTA prologue: centre the centroids, norms, operand B (TF32, K-major layout), barriers, tensor memory
    if (tid < kUPad) {
        float s = 0.f;
        if (tid < D)
            for (int k = 0; k < K; ++k) s += centroids[k * D + tid];
        mean[tid] = tid < D ? s / (float)K : 0.f;
    }
    __syncthreads();
    if (tid < kUPad) {
        float s = 0.f;
        if (tid < K)
            for (int d = 0; d < D; ++d) {
                const float v = centroids[tid * D + d] - mean[d];
                s = fmaf(v, v, s);
            }
        cn2[tid] = tid < K ? s : INFINITY;              // padded centroids can never be candidates
        const float ncu = tid < K ? sqrtf(s) * 1.0001f : 0.f;
        eab[2 * tid] = 2.9296875e-3f * ncu;             // 1.5 * 2^-9 |c'_k|
        eab[2 * tid + 1] = 4.76837158203125e-7f * ncu * ncu * 1.0001f;   // 2^-21 |c'_k|^2
    }
    for (int i = tid; i < kUPad * kUPad / 4; i += kUThreads) {         // one 16-byte chunk of operand B per step
        const int k = i & (kUPad - 1), c = i / kUPad;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = 4 * c + j;
            v[j] = (k < K && d < D) ? __uint_as_float(to_tf32(centroids[k * D + d] - mean[d])) : 0.f;
        }
        *reinterpret_cast<float4 *>(smem + L.b + operand_offset(k, c, kUPad / 8)) = make_float4(v[0], v[1], v[2], v[3]);
    }
    for (int i = tid; i < 2 * kUABytes / 16; i += kUThreads)           // operand A: the chunks past D stay zero for good
        reinterpret_cast<float4 *>(smem + L.a)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kAccumulate)
        for (int i = tid; i < K * (D + 1); i += kUThreads) acc[i] = 0.0;
    if (tid == 0) {
        for (int i = 0; i < kURing; ++i) {
            mbar_init(reinterpret_cast<unsigned long long *>(raw_full) + i, 1);
            mbar_init(reinterpret_cast<unsigned long long *>(raw_empty) + i, 1);
        }
        mbar_init(reinterpret_cast<unsigned long long *>(mma_done), 1);
        mbar_init(reinterpret_cast<unsigned long long *>(mma_done) + 1, 1);
    }
    if (warp == 8) {                                    // the producer warp owns the tensor-memory allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kUTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // operand B was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    float ea_max = 0.f, eb_max = 0.f;                   // row-wide bound: the largest ea_k, eb_k over the real centroids
    for (int k = 0; k < K; ++k) { ea_max = fmaxf(ea_max, eab[2 * k]); eb_max = fmaxf(eb_max, eab[2 * k + 1]); }

    // tiles of this CTA: blockIdx.x, + gridDim.x, ...; n counts them
    const int64_t n_cta = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const unsigned tile_bytes = (unsigned)L.raw_tile;
    const int tile_floats = kURows * D;
    unsigned long long n_viol = 0, n_cand = 0;

    if (warp == 8) {
        // ===== producer: one TMA bulk copy per tile into the ring =====
        if (lane == 0) {
            for (int64_t n = 0; n < n_cta; ++n) {
                const int slot = (int)(n % kURing);
                if (n >= kURing) mbar_wait_backoff(reinterpret_cast<unsigned long long *>(raw_empty) + slot, (unsigned)((n / kURing - 1) & 1));
                const int64_t tl = blockIdx.x + n * gridDim.x;
                bulk_load_tile(raw + (size_t)slot * tile_floats, data + tl * tile_floats, tile_bytes,
                               reinterpret_cast<unsigned long long *>(raw_full) + slot);
            }
        }
        __syncwarp();
    } else {
        // ===== consumer groups =====
        const int g = warp >> 2, wg = warp & 3, r = tid - g * kUGroup;           // group, warp in group, row in tile
        const int grp_bar = 1 + g;
        unsigned char *a_op = smem + L.a + (size_t)g * kUABytes;
        const uint32_t a_addr = smem_u32(a_op), b_addr = smem_u32(smem + L.b);
        unsigned *bits = reinterpret_cast<unsigned *>(smem + L.bits + (size_t)g * L.bits_g);
        unsigned short *cstart = reinterpret_cast<unsigned short *>(smem + L.cstart + (size_t)g * L.cstart_g);
        unsigned short *order = reinterpret_cast<unsigned short *>(smem + L.order + (size_t)g * kURows * 2);
        float *distbuf = reinterpret_cast<float *>(smem + L.scratch + (size_t)warp * 256);
        unsigned short *pairbuf = reinterpret_cast<unsigned short *>(distbuf + kTcPairs);
        const unsigned long long kmask = K >= 64 ? ~0ull : ((1ull << K) - 1ull);
        const uint32_t tmem_acc = tmem_base + (uint32_t)g * kUPad;               // this group's accumulator: 64 columns
        unsigned uses = 0;                                                        // MMAs this group has committed
        for (int64_t n = g; n < n_cta; n += 2) {
            const int slot = (int)(n % kURing);
            const int64_t row0 = (blockIdx.x + n * gridDim.x) * (int64_t)kURows;
            const float *xt = raw + (size_t)slot * tile_floats;
            // one thread waits for the tile (and later for the MMA); the others sleep at the group barrier
            if (r == 0) mbar_wait_u(reinterpret_cast<unsigned long long *>(raw_full) + slot, (unsigned)((n / kURing) & 1));
            if (kAccumulate) zero_member_bits(bits, K, 4, r, kUGroup);
            bar_sync(grp_bar, kUGroup);

            // ---- operand A: this thread's row, centred on the centroid mean, TF32, 16 bytes per step
            const float *x = xt + r * D;
            float nx2 = 0.f;
            const int c_full = D >> 2;
#pragma unroll 2
            for (int c = 0; c < c_full; ++c) {
                const float4 m4 = *reinterpret_cast<const float4 *>(mean + 4 * c);
                const float x0 = x[4 * c] - m4.x, x1 = x[4 * c + 1] - m4.y, x2 = x[4 * c + 2] - m4.z, x3 = x[4 * c + 3] - m4.w;
                nx2 = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, nx2))));
                *reinterpret_cast<uint4 *>(a_op + operand_offset(r, c, kURows / 8)) = make_uint4(to_tf32(x0), to_tf32(x1), to_tf32(x2), to_tf32(x3));
            }
            if (D & 3) {                                                           // the ragged last chunk
                const float4 m4 = *reinterpret_cast<const float4 *>(mean + 4 * c_full);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                uint32_t v[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (4 * c_full + j < D) {
                        const float xv = x[4 * c_full + j] - mm[j];
                        nx2 = fmaf(xv, xv, nx2);
                        v[j] = to_tf32(xv);
                    }
                *reinterpret_cast<uint4 *>(a_op + operand_offset(r, c_full, kURows / 8)) = make_uint4(v[0], v[1], v[2], v[3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to the tensor core
            tc_fence_before();
            bar_sync(grp_bar, kUGroup);
            if (r == 0) {
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < kUPad / 8; ++ks)                              // K = 8 per MMA: two 16-byte chunks
                    umma_tf32(tmem_acc, umma_desc(a_addr + ks * 2 * (kURows / 8) * 128, (kURows / 8) * 128, 128),
                              umma_desc(b_addr + ks * 2 * (kUPad / 8) * 128, (kUPad / 8) * 128, 128), ks > 0 ? 1u : 0u);
                umma_commit(reinterpret_cast<unsigned long long *>(mma_done) + g);
                mbar_wait_u(reinterpret_cast<unsigned long long *>(mma_done) + g, uses & 1u);
            }
            ++uses;
            bar_sync(grp_bar, kUGroup);
            tc_fence_after();

            // ---- stage A epilogue: the row's 64 ranking values, their minimum, the candidates
            float gk[64];
            tmem_load_row64(tmem_acc + ((uint32_t)(wg * 32) << 16), gk);
            const float nx = sqrtf(nx2) * 1.0001f;
            const float nx_term = 4.76837158203125e-7f * nx * nx * 1.0001f;        // 2^-21 |x'|^2
            float gmin = INFINITY;
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                gk[k] = cn2[k] - 2.f * gk[k];                                       // padded centroids: +inf
                gmin = fminf(gmin, gk[k]);
            }
            const float thr = gmin + 2.f * (screen_bound(nx, nx_term, ea_max, eb_max) * 1.000001f);
            unsigned m_lo = 0u, m_hi = 0u;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (gk[k] <= thr) m_lo |= 1u << k;
                if (gk[k + 32] <= thr) m_hi |= 1u << k;
            }
            unsigned long long mask = (((unsigned long long)m_hi << 32) | m_lo) & kmask;
            if (mask == 0 || !(nx < INFINITY)) mask = kmask;       // NaN/Inf/overflowing rows: everything is a candidate
            if (kCheck) {
                // bound check: (g_k - g_0) vs float64 (d2_k - d2_0), tolerance E_k + E_0
                n_cand += __popcll(mask);
                const double d0 = sqdist_scipy(centroids, x, D);
                const float e0 = screen_bound(nx, nx_term, eab[0], eab[1]);
                for (int k = 0; k < K; ++k) {
                    const double dk = sqdist_scipy(centroids + (size_t)k * D, x, D);
                    float gkk = 0.f, g00 = 0.f;
#pragma unroll
                    for (int q = 0; q < 64; ++q) { if (q == k) gkk = gk[q]; if (q == 0) g00 = gk[q]; }
                    const double err = fabs(((double)gkk - (double)g00) - (dk - d0));
                    const double tol = (double)screen_bound(nx, nx_term, eab[2 * k], eab[2 * k + 1]) + (double)e0;
                    if (!(err <= tol)) ++n_viol;
                }
            }
            // ---- stages B, C on the original values (lane = row of this warp)
            const int mine = refine_warp(mask, xt + wg * 32 * D, D, centroids, D, eps, pairbuf, distbuf, lane);
            labels[row0 + r] = mine;

            if (kAccumulate) {
                // counting sort of the tile's rows by label, then each warp walks its clusters
                const unsigned same = tile_member_bits(bits, mine, 4, lane, wg);
                bar_sync(grp_bar, kUGroup);
                if (wg == 0) tile_cluster_starts(bits, cstart, K, 4, lane);
                bar_sync(grp_bar, kUGroup);
                tile_row_order(bits, cstart, order, mine, same, 4, r, lane, wg);
                bar_sync(grp_bar, kUGroup);
                // the accumulators are shared by the two groups: tiles are added in tile order
                if (n >= 1) bar_sync(3 + (1 - g), 2 * kUGroup);
                accumulate_rows(acc, xt, D, cstart, order, K, D, kURows, lane, wg, 4);
                __threadfence_block();
                bar_sync(grp_bar, kUGroup);
                if (n + 1 < n_cta) bar_arrive(3 + g, 2 * kUGroup);
            } else {
                bar_sync(grp_bar, kUGroup);                        // every reader of the raw tile is done
            }
            tc_fence_before();
            if (r == 0) mbar_arrive(reinterpret_cast<unsigned long long *>(raw_empty) + slot);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kAccumulate) {
        double *out = partials + (size_t)blockIdx.x * K * (D + 1);
        for (int i = tid; i < K * (D + 1); i += kUThreads) out[i] = acc[i];
    }
    if (kCheck && warp < 8) {
        atomicAdd(check_out, n_viol);
        atomicAdd(check_out + 1, n_cand);
    }
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kUTmemCols) : "memory");
    }
}

bool umma_supported(int D, int K)
{
    return K >= 16 && K <= 64 && D >= 8 && D <= 64;
}

template <bool kAcc, bool kCheck>
static int launch_umma_t(const float *data, int64_t n_tiles, int D, const float *centroids, int K, int32_t *labels,
                         double *partials, int grid, unsigned long long *check_out, cudaStream_t st)
{
    const USmem L = umma_layout(D, K, kAcc);
    if (L.total > 227 * 1024) return fail(GSL_EINVAL, "kmeans (tcgen05 screening): %zu B of shared memory", L.total);
    GSL_CUDA_TRY(cudaFuncSetAttribute(kmeans_step_umma_kernel<kAcc, kCheck>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const float eps = (float)((D + 12) * 5.9604644775390625e-08);
    kmeans_step_umma_kernel<kAcc, kCheck><<<grid, kUThreads, L.total, st>>>(data, n_tiles, D, centroids, K, labels, partials, L, eps, check_out);
    GSL_LAUNCH_CHECK("kmeans_step_umma_kernel");
    return GSL_OK;
}

// Assignment (+ per-CTA sums) of the FULL 128-row tiles of a 16-byte aligned matrix; returns the
// number of rows it covered in *rows_done (0: not applicable, the caller takes another path) and
// the number of partial-sum blocks it wrote in *n_parts.
int launch_step_umma(bool accumulate, const float *data, int64_t N, int D, const float *centroids, int K,
                     int32_t *labels, double *partials, cudaStream_t st, int64_t *rows_done, int *n_parts)
{
    *rows_done = 0;
    *n_parts = 0;
    const int64_t n_tiles = N / kURows;
    if (!umma_supported(D, K) || n_tiles == 0 || ((uintptr_t)data & 15)) return GSL_OK;
    const USmem L = umma_layout(D, K, accumulate);
    if (L.total > 227 * 1024) return GSL_OK;
    const int64_t cap = sm_count();
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    const int rc = accumulate ? launch_umma_t<true, false>(data, n_tiles, D, centroids, K, labels, partials, grid, nullptr, st)
                              : launch_umma_t<false, false>(data, n_tiles, D, centroids, K, labels, partials, grid, nullptr, st);
    if (rc != GSL_OK) return rc;
    *rows_done = n_tiles * kURows;
    *n_parts = accumulate ? grid : 0;
    return GSL_OK;
}

int launch_umma_selftest(const float *data, int64_t N, int D, const float *centroids, int K, int32_t *labels,
                         unsigned long long *out2, cudaStream_t st, int64_t *rows_done)
{
    *rows_done = 0;
    const int64_t n_tiles = N / kURows;
    if (!umma_supported(D, K) || n_tiles == 0 || ((uintptr_t)data & 15)) return GSL_OK;
    const int64_t cap = sm_count();
    const int rc = launch_umma_t<false, true>(data, n_tiles, D, centroids, K, labels, nullptr, (int)(n_tiles < cap ? n_tiles : cap), out2, st);
    if (rc == GSL_OK) *rows_done = n_tiles * kURows;
    return rc;
}

}  // namespace gsl
