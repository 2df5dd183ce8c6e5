// Viewer-side consumers of the lifted labels: the two per-Gaussian
// loops the reference's WebGL viewer runs in its worker (Web_Viewer_Gaussians_Selection/
// gaussians_selection.js, `gs` below) on every camera move / click:
//
//   gsl_viewer_depth_sort   gs:417-462  runSort: 16-bit counting sort of the view-space depths
//   gsl_viewer_hit_test     gs:361-395  performHitTesting: nearest projected Gaussian within 10 px
//
// JavaScript numbers are IEEE float64 and `| 0` is ECMAScript ToInt32, so both loops are restated
// in float64 with the reference's association order (no FMA: the library is built with
// -fmad=false and the products below are spelled with __dmul_rn/__dadd_rn).  Math.hypot is
// implementation-approximated by the standard; this file follows V8's builtin (Chrome / Node:
// scale by the larger magnitude, sum of squares, sqrt, rescale), which uses correctly rounded
// operations only and can therefore be reproduced bit for bit.
#include <float.h>
#include <limits.h>
#include <math.h>
#include <math_constants.h>

#include <algorithm>

#include "common.cuh"
#include "radix_sort.cuh"

namespace gsl {

// ECMAScript ToInt32 (the `| 0` of gs:437, :446).
__device__ __forceinline__ int32_t js_to_int32(double d)
{
    if (!isfinite(d)) return 0;
    const double t = trunc(d);
    if (fabs(t) < 2147483648.0) return (int32_t)t;
    double m = fmod(t, 4294967296.0);
    if (m < 0) m += 4294967296.0;
    return (int32_t)(uint32_t)(unsigned long long)m;
}

// gs:436-441: depth = ((vp[2]*x + vp[6]*y + vp[10]*z) * 4096) | 0, with running min / max.
__global__ void __launch_bounds__(256) viewer_depth_kernel(const float *__restrict__ pos, int64_t N, int stride, double a,
                                                           double b, double c, int32_t *__restrict__ depth,
                                                           int *__restrict__ minmax)
{
    int lo = INT_MAX, hi = INT_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = pos + i * stride;
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(a, (double)p[0]), __dmul_rn(b, (double)p[1])),
                                   __dmul_rn(c, (double)p[2]));
        const int32_t d = js_to_int32(__dmul_rn(s, 4096.0));
        depth[i] = d;
        lo = min(lo, d);
        hi = max(hi, d);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ int slo[8], shi[8];
    if ((threadIdx.x & 31) == 0) {
        slo[threadIdx.x >> 5] = lo;
        shi[threadIdx.x >> 5] = hi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            lo = min(lo, slo[w]);
            hi = max(hi, shi[w]);
        }
        if (lo <= hi) {
            atomicMin(&minmax[0], lo);
            atomicMax(&minmax[1], hi);
        }
    }
}

__global__ void viewer_init_kernel(int *__restrict__ minmax)
{
    minmax[0] = INT_MAX;
    minmax[1] = INT_MIN;
}

// gs:443-447: bucket = ((depth - minDepth) * depthInv) | 0 with depthInv = 65536 / (max - min).
// The result lies in [0, 65536]; 65536 (the deepest Gaussians, when the product does not round
// below it) indexes past the reference's 65536-entry typed arrays -- see viewer_depth_sort.
__global__ void __launch_bounds__(256) viewer_bucket_kernel(int32_t *__restrict__ depth_to_key, int64_t N,
                                                            const int *__restrict__ minmax)
{
    const double lo = (double)minmax[0], hi = (double)minmax[1];
    const double inv = __ddiv_rn(65536.0, __dsub_rn(hi, lo));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = __dsub_rn((double)depth_to_key[i], lo);
        depth_to_key[i] = js_to_int32(__dmul_rn(d, inv));
    }
}

struct HitPick {
    double dist, depth;      // the (dist, depth) of the candidate; dist = +inf when empty
    int64_t idx;
};
// Partial result of the hit test over a set of Gaussians, mergeable in any grouping:
//   first  = the lowest-index candidate among those at the smallest distance
//   best   = the (dist, depth, index)-smallest candidate among those whose depth is not NaN
// The reference's sequential scan (gs:385-390) ends on `first` when its depth is NaN (a NaN depth
// is never displaced on a distance tie and never displaces), else on `best`.
struct HitState {
    HitPick first, best;
};

__device__ __forceinline__ bool first_less(const HitPick &a, const HitPick &b)
{
    return a.dist < b.dist || (a.dist == b.dist && a.idx < b.idx);
}
__device__ __forceinline__ bool best_less(const HitPick &a, const HitPick &b)
{
    if (a.dist != b.dist) return a.dist < b.dist;
    if (a.depth != b.depth) return a.depth < b.depth;
    return a.idx < b.idx;
}
__device__ __forceinline__ void hit_merge(HitState &s, const HitState &o)
{
    if (first_less(o.first, s.first)) s.first = o.first;
    if (best_less(o.best, s.best)) s.best = o.best;
}
__device__ __forceinline__ HitPick pick_shfl(const HitPick &p, int o)
{
    HitPick q;
    q.dist = __shfl_xor_sync(0xffffffffu, p.dist, o);
    q.depth = __shfl_xor_sync(0xffffffffu, p.depth, o);
    q.idx = __shfl_xor_sync(0xffffffffu, p.idx, o);
    return q;
}

// V8's Math.hypot for two arguments (builtins/math.tq, MathHypot): with the Kahan compensation
// written out, two terms reduce to sqrt(n0*n0 + n1*n1) * max, n_i = |v_i| / max.
__device__ __forceinline__ double js_hypot2(double dx, double dy)
{
    const double ax = fabs(dx), ay = fabs(dy);
    if (isinf(ax) || isinf(ay)) return CUDART_INF;
    if (isnan(ax) || isnan(ay)) return CUDART_NAN;
    const double m = ax > ay ? ax : ay;
    if (m == 0.0) return 0.0;
    const double n0 = __ddiv_rn(ax, m), n1 = __ddiv_rn(ay, m);
    const double sum = __dadd_rn(__dmul_rn(n0, n0), __dmul_rn(n1, n1));
    return __dmul_rn(__dsqrt_rn(sum), m);
}

struct HitParams {
    double m[16];            // combined matrix, gs:364
    double x, y, vw, vh;
};

__device__ __forceinline__ HitState hit_empty()
{
    HitState s;
    s.first.dist = s.best.dist = CUDART_INF;
    s.first.depth = s.best.depth = CUDART_INF;
    s.first.idx = s.best.idx = LLONG_MAX;
    return s;
}

__device__ __forceinline__ HitState hit_block_reduce(HitState s, HitState *sh)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        HitState t;
        t.first = pick_shfl(s.first, o);
        t.best = pick_shfl(s.best, o);
        hit_merge(s, t);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) hit_merge(s, sh[w]);
    return s;
}

__global__ void __launch_bounds__(256) viewer_hit_kernel(const float *__restrict__ pos, int64_t N, int stride, HitParams P,
                                                         HitState *__restrict__ partial)
{
    HitState s = hit_empty();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = pos + i * stride;
        const double x = p[0], y = p[1], z = p[2];
        double r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)      // gs:400-402, pos[3] = 1.0
            r[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, P.m[k]), __dmul_rn(y, P.m[k + 4])), __dmul_rn(z, P.m[k + 8])),
                             __dmul_rn(1.0, P.m[k + 12]));
        if (r[3] <= 0.0) continue;       // gs:403
        const double sx = __dmul_rn(__dmul_rn(__dadd_rn(__ddiv_rn(r[0], r[3]), 1.0), 0.5), P.vw);   // gs:379
        const double sy = __dmul_rn(__dmul_rn(__dadd_rn(__ddiv_rn(r[1], r[3]), 1.0), 0.5), P.vh);   // gs:380
        const double depth = __ddiv_rn(r[2], r[3]);                                                 // gs:381
        const double dist = js_hypot2(__dsub_rn(sx, P.x), __dsub_rn(sy, P.y));                      // gs:384
        if (!(dist < 10.0)) continue;    // gs:387
        HitPick c{dist, depth, i};
        if (first_less(c, s.first)) s.first = c;
        if (!isnan(depth) && best_less(c, s.best)) s.best = c;
    }
    __shared__ HitState sh[8];
    s = hit_block_reduce(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) viewer_hit_final_kernel(const HitState *__restrict__ partial, int n,
                                                               const int32_t *__restrict__ labels, int32_t no_selection,
                                                               int32_t *__restrict__ label_out, int64_t *__restrict__ index_out)
{
    HitState s = hit_empty();
    for (int i = threadIdx.x; i < n; i += blockDim.x) hit_merge(s, partial[i]);
    __shared__ HitState sh[8];
    s = hit_block_reduce(s, sh);
    if (threadIdx.x == 0) {
        int64_t win = -1;
        if (s.first.idx != LLONG_MAX) win = isnan(s.first.depth) ? s.first.idx : s.best.idx;
        if (label_out) *label_out = win >= 0 ? labels[win] : no_selection;
        if (index_out) *index_out = win;
    }
}

struct SortPlan {
    size_t keys_a, keys_b, idx_a, hist, minmax, total;
};
static SortPlan sort_plan(int64_t N)
{
    SortPlan p;
    size_t o = 0;
    const size_t n = (size_t)(N > 0 ? N : 0);
    p.keys_a = o; o = align_up(o + n * 4, 256);
    p.keys_b = o; o = align_up(o + n * 4, 256);
    p.idx_a = o;  o = align_up(o + n * 4, 256);
    p.hist = o;   o = align_up(o + radix::hist_words(N, 9) * 4, 256);
    p.minmax = o; o = align_up(o + 8, 256);
    p.total = o;
    return p;
}

}  // namespace gsl

using namespace gsl;

extern "C" size_t gsl_viewer_sort_workspace_bytes(int64_t N) { return sort_plan(N).total; }

extern "C" int gsl_viewer_depth_sort(const float *pos, int64_t N, int stride, const double *view_proj,
                                     uint32_t *depth_index, void *ws, size_t ws_bytes, void *stream)
{
    if (N < 0 || stride < 3 || !view_proj || (N > 0 && (!pos || !depth_index || !ws)))
        return fail(GSL_EINVAL, "gsl_viewer_depth_sort: bad argument");
    if (N > 0xffffffffLL) return fail(GSL_EINVAL, "gsl_viewer_depth_sort: vertex count exceeds a Uint32Array index");
    if (N == 0) return GSL_OK;
    const SortPlan p = sort_plan(N);
    if (ws_bytes < p.total) return fail(GSL_EWORKSPACE, "gsl_viewer_depth_sort: workspace %zu < %zu", ws_bytes, p.total);
    cudaStream_t st = (cudaStream_t)stream;
    char *w = (char *)ws;
    int32_t *keys_a = (int32_t *)(w + p.keys_a);
    uint32_t *keys_b = (uint32_t *)(w + p.keys_b), *idx_a = (uint32_t *)(w + p.idx_a), *hist = (uint32_t *)(w + p.hist);
    int *minmax = (int *)(w + p.minmax);
    viewer_init_kernel<<<1, 1, 0, st>>>(minmax);          // (a copy from host memory would synchronise the stream)
    GSL_LAUNCH_CHECK("viewer_init_kernel");
    const int grid = (int)std::min<int64_t>((N + 255) / 256, (int64_t)sm_count() * 8);
    viewer_depth_kernel<<<grid, 256, 0, st>>>(pos, N, stride, view_proj[2], view_proj[6], view_proj[10], keys_a, minmax);
    GSL_LAUNCH_CHECK("viewer_depth_kernel");
    viewer_bucket_kernel<<<grid, 256, 0, st>>>(keys_a, N, minmax);
    GSL_LAUNCH_CHECK("viewer_bucket_kernel");
    // Stable sort by bucket = the reference's counting sort (gs:449-457: starts0 from the counts,
    // then a scatter in index order).  Buckets use 17 bits: two 9-bit passes.  A Gaussian in
    // bucket 65536 is past the end of counts0/starts0 (Uint32Array(65536)): `counts0[65536]++`
    // is a no-op, `starts0[65536]++` yields NaN and `depthIndex[NaN] = i` is a no-op as well, so
    // the Gaussian is missing from depthIndex and the tail keeps its initial zeros.
    int rc = radix::pass<9>((const uint32_t *)keys_a, nullptr, keys_b, idx_a, N, 0, hist, 0xffffffffu, 0, st);
    if (rc) return rc;
    return radix::pass<9>(keys_b, idx_a, nullptr, depth_index, N, 9, hist, 65536u, 0u, st);
}

extern "C" size_t gsl_viewer_hit_workspace_bytes(void) { return (size_t)4096 * sizeof(HitState); }

extern "C" int gsl_viewer_hit_test(const float *pos, const int32_t *labels, int64_t N, int stride, const double *matrix,
                                   double x, double y, double viewport_w, double viewport_h, int32_t no_selection,
                                   int32_t *label_out, int64_t *index_out, void *ws, size_t ws_bytes, void *stream)
{
    if (N < 0 || stride < 3 || !matrix || !ws || (N > 0 && (!pos || !labels)) || (!label_out && !index_out))
        return fail(GSL_EINVAL, "gsl_viewer_hit_test: bad argument");
    if (ws_bytes < gsl_viewer_hit_workspace_bytes())
        return fail(GSL_EWORKSPACE, "gsl_viewer_hit_test: workspace %zu < %zu", ws_bytes, gsl_viewer_hit_workspace_bytes());
    cudaStream_t st = (cudaStream_t)stream;
    HitParams P;
    for (int i = 0; i < 16; ++i) P.m[i] = matrix[i];
    P.x = x; P.y = y; P.vw = viewport_w; P.vh = viewport_h;
    int grid = (int)std::min<int64_t>(std::max<int64_t>((N + 255) / 256, 1), std::min<int64_t>((int64_t)sm_count() * 8, 4096));
    viewer_hit_kernel<<<grid, 256, 0, st>>>(pos, N, stride, P, (HitState *)ws);
    GSL_LAUNCH_CHECK("viewer_hit_kernel");
    viewer_hit_final_kernel<<<1, 256, 0, st>>>((const HitState *)ws, grid, labels, no_selection, label_out, index_out);
    GSL_LAUNCH_CHECK("viewer_hit_final_kernel");
    return GSL_OK;
}
