// Device helpers shared by the K-means kernels (kmeans.cu, kmeans_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gsl {

// scipy ckdtree sqeuclidean_distance_double on float64 copies of float32 values.
__device__ __forceinline__ double sqdist_scipy(const float *__restrict__ c, const float *__restrict__ x, int D)
{
    double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
    int i = 0;
    for (; i + 4 <= D; i += 4) {
        const double d0 = (double)c[i] - (double)x[i];
        const double d1 = (double)c[i + 1] - (double)x[i + 1];
        const double d2 = (double)c[i + 2] - (double)x[i + 2];
        const double d3 = (double)c[i + 3] - (double)x[i + 3];
        a0 += d0 * d0; a1 += d1 * d1; a2 += d2 * d2; a3 += d3 * d3;
    }
    double s = a0 + a1 + a2 + a3;
    for (; i < D; ++i) {
        const double d = (double)c[i] - (double)x[i];
        s += d * d;
    }
    return s;
}

// Stage the contiguous [rows x D] float32 block at `src` into shared memory with `pitch` floats
// per row.  `vec` = the block is 16-byte aligned and a whole number of float4.
__device__ __forceinline__ void stage_tile(float *__restrict__ tile, int pitch, const float *__restrict__ src,
                                           int rows, int D, bool vec, int t, int n_threads)
{
    const int n_el = rows * D;
    if (vec) {
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        for (int i = t; i < (n_el >> 2); i += n_threads) {
            const float4 v = __ldcs(src4 + i);
            const int e = i << 2;
            int r = e / D, d = e - r * D;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                tile[r * pitch + d] = vv[j];
                if (++d == D) { d = 0; ++r; }
            }
        }
    } else {
        for (int i = t; i < n_el; i += n_threads) {
            const int r = i / D, d = i - r * D;
            tile[r * pitch + d] = __ldcs(src + i);
        }
    }
}

// Segmented reduction of one staged tile into the CTA's [K][D+1] float64 accumulator.
// Phase 1 (tile_member_bits): every warp groups its 32 rows by label with one match_any and the
// lowest lane of each group publishes the group's lane mask, bits[k][w] = rows of warp w with
// label k.  Phase 2 (accumulate_tile): warp w owns the clusters k = w (mod n_warps), walks their
// member bits in ascending row order and adds the rows lane-per-dimension (two dimension chunks
// per walk).  No atomics, fixed order: bit-reproducible.
// `bits` is [K][n_warps] uint32 in shared memory; lab = label of this thread's row or -1.
// Call order: zero_member_bits -> __syncthreads -> tile_member_bits -> __syncthreads ->
// accumulate_tile (the caller's next __syncthreads protects the buffers).
__device__ __forceinline__ void zero_member_bits(unsigned *__restrict__ bits, int K, int n_warps, int t, int n_threads)
{
    for (int i = t; i < K * n_warps; i += n_threads) bits[i] = 0u;
}

__device__ __forceinline__ void tile_member_bits(unsigned *__restrict__ bits, int lab, int n_warps, int lane, int warp)
{
    const unsigned same = __match_any_sync(0xffffffffu, lab);
    if (lab >= 0 && lane == __ffs(same) - 1) bits[lab * n_warps + warp] = same;
}

__device__ __forceinline__ void accumulate_tile(double *__restrict__ acc, const float *__restrict__ tile, int pitch,
                                                const unsigned *__restrict__ bits, int K, int D,
                                                int lane, int warp, int n_warps)
{
    for (int k = warp; k < K; k += n_warps) {
        for (int d0 = 0; d0 < D; d0 += 64) {
            const int da = d0 + lane, db = da + 32;
            const bool ina = da < D, inb = db < D;
            double sa = 0.0, sb = 0.0;
            int cnt = 0;
            for (int w = 0; w < n_warps; ++w) {
                unsigned m = bits[k * n_warps + w];
                cnt += __popc(m);
                const float *base = tile + (w * 32) * pitch;
                while (m) {
                    const float *row = base + (__ffs(m) - 1) * pitch;
                    m &= m - 1;
                    if (ina) sa += (double)row[da];
                    if (inb) sb += (double)row[db];
                }
            }
            if (cnt) {
                if (ina) acc[k * (D + 1) + da] += sa;
                if (inb) acc[k * (D + 1) + db] += sb;
                if (d0 == 0 && lane == 0) acc[k * (D + 1) + D] += (double)cnt;
            }
        }
    }
}

}  // namespace gsl
