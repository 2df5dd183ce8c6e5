// Spatial ordering and the per-(tile, view) verdicts of the lifting sweep, sm_100a.
//
// The vote of a Gaussian depends only on its own position, so the order in which Gaussians are
// processed is free.  Processing them in Morton order makes (i) the 32 Gaussians of a warp
// project into a small 2-D patch of every view, so that a gather request touches a few lines
// of the packed label map instead of 32, (ii) every 128-Gaussian tile of the sweep a small box in
// space, so that a whole (tile, view) can be proven invisible -- behind the camera or outside the
// image for every point of the box -- and skipped, or proven to lie in front of the camera with a
// float32 error bound that holds for the whole tile, which is what makes the sweep's fast path two
// compares per pair, and (iii) the tiles resident on the GPU at a time neighbours in space, so the
// parts of the label maps they touch stay in L2.  Labels are unaffected: the sweep still evaluates
// the reference's exact test for every pair it does not skip, and every verdict is conservative
// (margins ten times the float32 rounding error of its own evaluation, never the other way).
//
//   order_stats_kernel    min/max (ordered-uint atomics) and first two moments (float64
//                         atomics, one per CTA) of the finite coordinates
//   order_key_kernel      256^3 grid over [mean - 4 sigma, mean + 4 sigma] clipped to the
//                         bounding box (outliers clamp to the border cells), key = 24-bit Morton
//                         code; non-finite positions take the last key
//   sort_cells            (key, row) pairs by key (lift_sort.cu)
//   order_permute_kernel  pos_sorted[i] = pos[perm[i]], and the bounding box (+ non-finite flag)
//                         of each run of 256 sorted Gaussians (a CTA is one run)
//   fill_view_planes      (host) per view, the five half-spaces of the visibility test as linear
//                         forms; uploaded with the other view tables
//   order_verdict_kernel  16 bits per (tile, view): cull / fast with its bound / general / exact
// Eight stream operations besides the sort's five: at 750 K Gaussians per rank (8 GPUs) the pass
// is launch-bound, so the small kernels of round 1 (grid, tile boxes, planes) were folded away.
// The order inside a cell follows the input order (the sort is stable), so the whole ordering is
// deterministic.  (A 1024^3 grid with 30-bit keys was measured: one more sort pass, same gather time.)
#include <algorithm>

#include "common.cuh"
#include "lift_internal.cuh"

namespace gsl {

__device__ __forceinline__ unsigned enc_f32(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned e)
{
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// stats layout (all 8-byte slots): u[0..2] min xyz, u[3..5] max xyz (encoded float32 in the low
// word), d[6..8] sum, d[9..11] sum of squares, d[12] count of finite rows, then the derived grid:
// f32 lo[3], inv[3] at byte 128.
constexpr int kStatsBytes = 256;

__global__ void __launch_bounds__(256)
order_stats_kernel(const float *__restrict__ pos, int64_t N, unsigned long long *__restrict__ stats)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    double s1[3] = {0.0, 0.0, 0.0}, s2[3] = {0.0, 0.0, 0.0}, cnt = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        const float v[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        if (fabsf(v[0]) < INFINITY && fabsf(v[1]) < INFINITY && fabsf(v[2]) < INFINITY) {
            cnt += 1.0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                lo[a] = fminf(lo[a], v[a]); hi[a] = fmaxf(hi[a], v[a]);
                s1[a] += (double)v[a]; s2[a] += (double)v[a] * (double)v[a];
            }
        }
    }
    __shared__ float redf[6][8];
    __shared__ double redd[7][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            s1[a] += __shfl_xor_sync(0xffffffffu, s1[a], o);
            s2[a] += __shfl_xor_sync(0xffffffffu, s2[a], o);
        }
        if (lane == 0) { redf[a][warp] = lo[a]; redf[3 + a][warp] = hi[a]; redd[a][warp] = s1[a]; redd[3 + a][warp] = s2[a]; }
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) redd[6][warp] = cnt;
    __syncthreads();
    if (threadIdx.x < 6) {                       // one atomic per CTA and bound
        float v = redf[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, redf[threadIdx.x][w]) : fmaxf(v, redf[threadIdx.x][w]);
        // minima are kept as the complement of their code: every slot starts at 0 and only grows
        unsigned *slot = reinterpret_cast<unsigned *>(stats + threadIdx.x);
        atomicMax(slot, threadIdx.x < 3 ? ~enc_f32(v) : enc_f32(v));
    } else if (threadIdx.x >= 32 && threadIdx.x < 39) {
        const int j = threadIdx.x - 32;
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += redd[j][w];
        atomicAdd(reinterpret_cast<double *>(stats + 6 + j), v);
    }
}

// Grid of the ordering: per axis [max(min, mean - 4 sigma), min(max, mean + 4 sigma)] in 256 cells.
// Only the quality of the ordering depends on it, never a result.  Every CTA of the key kernel
// derives it from the statistics itself (three threads, a few float64 operations).
__device__ __forceinline__ void order_grid_axis(const unsigned long long *__restrict__ stats, int a, float &lo_out, float &inv_out)
{
    const double *d = reinterpret_cast<const double *>(stats);
    const double n = d[12];
    float lo = dec_f32(~(unsigned)stats[a]), hi = dec_f32((unsigned)stats[3 + a]);
    if (n > 0.0) {
        const double mean = d[6 + a] / n;
        const double var = fmax(d[9 + a] / n - mean * mean, 0.0);
        const double sd = sqrt(var);
        if (mean - 4.0 * sd > (double)lo) lo = (float)(mean - 4.0 * sd);
        if (mean + 4.0 * sd < (double)hi) hi = (float)(mean + 4.0 * sd);
    }
    const float ext = hi - lo;
    lo_out = lo;
    inv_out = (ext > 0.f && ext < INFINITY) ? 256.f / ext : 0.f;
}

__device__ __forceinline__ unsigned spread8(unsigned v)      // 8 bits -> every third bit
{
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256)
order_key_kernel(const float *__restrict__ pos, int64_t N, const unsigned long long *__restrict__ stats,
                 uint32_t *__restrict__ keys, int32_t *__restrict__ idx)
{
    __shared__ float grid[6];
    if (threadIdx.x < 3) order_grid_axis(stats, threadIdx.x, grid[threadIdx.x], grid[3 + threadIdx.x]);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    unsigned q[3];
    bool finite = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float v = pos[3 * i + a];
        finite = finite && (fabsf(v) < INFINITY);
        const float c = fminf(fmaxf((v - grid[a]) * grid[3 + a], 0.f), 255.f);     // NaN -> 0
        q[a] = (unsigned)(int)c;
    }
    keys[i] = finite ? (spread8(q[0]) | (spread8(q[1]) << 1) | (spread8(q[2]) << 2)) : 0x00ffffffu;
    idx[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
order_iota_kernel(int32_t *__restrict__ perm, int64_t N)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) perm[i] = (int32_t)i;
}

// pos_sorted[i] = pos[perm[i]] as float4 (one 16-byte load per Gaussian in the sweep), and, since a
// CTA is exactly one tile of kTile sorted Gaussians, the tile's box = {lo xyz, hi xyz, nonfinite, 0}.
static_assert(kTile == 256, "order_permute_kernel: one CTA of 256 threads per tile");
__global__ void __launch_bounds__(256)
order_permute_kernel(const float *__restrict__ pos, int64_t N, const int32_t *__restrict__ perm, float4 *__restrict__ pos_sorted,
                     float *__restrict__ box)
{
    __shared__ float red[8][8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int bad = 0;
    if (i < N) {
        const int64_t s = perm[i];
        const float pv[3] = {pos[3 * s + 0], pos[3 * s + 1], pos[3 * s + 2]};
        pos_sorted[i] = make_float4(pv[0], pv[1], pv[2], 0.f);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (fabsf(pv[a]) < INFINITY) { lo[a] = pv[a]; hi[a] = pv[a]; } else bad = 1;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    bad = __any_sync(0xffffffffu, bad);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        red[warp][0] = lo[0]; red[warp][1] = lo[1]; red[warp][2] = lo[2];
        red[warp][3] = hi[0]; red[warp][4] = hi[1]; red[warp][5] = hi[2];
        red[warp][6] = bad ? 1.f : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int c = threadIdx.x;
        float v = c == 7 ? 0.f : red[0][c];
        if (c < 7)
            for (int w = 1; w < 8; ++w) v = c < 3 ? fminf(v, red[w][c]) : fmaxf(v, red[w][c]);     // the flag is a max too
        box[(int64_t)blockIdx.x * 8 + c] = v;
    }
}

// The five half-spaces that bound "can pass the reference's visibility test" (dls:72, :80), as
// linear forms a.X + c of the world position, per view, in float32:
//   p0  cz                                   visible needs  > 0
//   p1  fx*cx + half_w*cz                    x >= 0      <=>  p1 >= 0      (cz > 0)
//   p2  fx*cx + (half_w - width)*cz          x <  width  <=>  p2 <  0
//   p3  fy*cy + half_h*cz                    y >= 0      <=>  p3 >= 0
//   p4  fy*cy + (half_h - height)*cz         y <  height <=>  p4 <  0
// planes[v][p] = {a0, a1, a2, c}.  A view with a non-finite parameter gets NaN planes, which
// never cull.
// (host side: the planes travel with the other view tables in the one upload of gsl_lift_prepare)
void fill_view_planes(const GslView &w, float4 (&planes)[5])
{
    const double kx[5] = {0.0, w.fx, w.fx, 0.0, 0.0};
    const double ky[5] = {0.0, 0.0, 0.0, w.fy, w.fy};
    const double kz[5] = {1.0, w.half_w, w.half_w - w.width, w.half_h, w.half_h - w.height};
    for (int p = 0; p < 5; ++p) {
        float4 o;
        o.x = (float)(kx[p] * w.R[0] + ky[p] * w.R[3] + kz[p] * w.R[6]);
        o.y = (float)(kx[p] * w.R[1] + ky[p] * w.R[4] + kz[p] * w.R[7]);
        o.z = (float)(kx[p] * w.R[2] + ky[p] * w.R[5] + kz[p] * w.R[8]);
        o.w = (float)(kx[p] * w.t[0] + ky[p] * w.t[1] + kz[p] * w.t[2]);
        planes[p] = o;
    }
}

// range of a.X + c over the box, and the magnitude that scales its rounding error
__device__ __forceinline__ void lin_range(const float4 pl, const float (&lo)[3], const float (&hi)[3],
                                          float &mn, float &mx, float &mag)
{
    const float a[3] = {pl.x, pl.y, pl.z};
    mn = pl.w; mx = pl.w; mag = fabsf(pl.w);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float p = a[j] * lo[j], q = a[j] * hi[j];
        mn += fminf(p, q); mx += fmaxf(p, q);
        mag += fmaxf(fabsf(p), fabsf(q));
    }
}

// Can any point of the box pass the visibility test in this view?  Conservative: float32
// evaluation (coefficients rounded once, four roundings per form: error < 1e-6 * mag) against
// margins of 1e-5 * mag plus 1e-6 px, so "no" is only ever said with room to spare; NaN says yes.
// cz_lo = a lower bound of cz over the box (same margin).
__device__ __forceinline__ bool box_may_be_visible(const float4 *__restrict__ pl, const float (&lo)[3], const float (&hi)[3],
                                                   float &cz_lo, bool &interior)
{
    const float rel = 1e-5f, px = 1e-6f;
    float mn, mx, mag;
    interior = false;
    lin_range(pl[0], lo, hi, mn, mx, mag);
    cz_lo = mn - (rel * mag + 1e-30f);
    if (mx < -(rel * mag + 1e-30f)) return false;                          // every point has z <= 0
    const float zpos = fmaxf(mx, 0.f) * px;
    const float zin = fmaxf(cz_lo, 0.f);
    // `interior`: every point of the box is in front of the camera and projects to
    // -15 <= x <= width + 1, -7 <= y <= height + 1, i.e. at least half a pixel inside the ring of the
    // packed maps, so the sweep's clamp of the ring coordinates cannot act.  With cz >= cz_lo > 0:
    //   x >= -15  <=>  p1 + 15 cz >= 0  <=  min p1 + 15 cz_lo >= 0,   x <= width + 1  <=  max p2 - cz_lo <= 0
    lin_range(pl[1], lo, hi, mn, mx, mag);
    if (mx < -(zpos + rel * mag)) return false;                            // x < 0 everywhere
    bool in = cz_lo > 0.f && mn + 15.f * zin >= zpos + rel * mag;
    lin_range(pl[2], lo, hi, mn, mx, mag);
    if (mn > zpos + rel * mag) return false;                               // x >= width everywhere
    in = in && mx - zin <= -(zpos + rel * mag);
    lin_range(pl[3], lo, hi, mn, mx, mag);
    if (mx < -(zpos + rel * mag)) return false;                            // y < 0 everywhere
    in = in && mn + 7.f * zin >= zpos + rel * mag;
    lin_range(pl[4], lo, hi, mn, mx, mag);
    if (mn > zpos + rel * mag) return false;                               // y >= height everywhere
    in = in && mx - zin <= -(zpos + rel * mag);
    interior = in;
    return true;
}

// Thread per (tile, view): the verdict the sweep acts on (lift_internal.cuh).  For the fast path
// the error bound of lift.cu (screen_pair) is evaluated once for the whole tile: with
// a >= |X|+|Y|+|Z| over the box and cz >= cz_lo > 0,
//     k_max = (g_rm a + g_tm) / cz_lo,   E = k_max (FXH + span) + 3.03 u span + c0,
// every factor rounded up; 1/2 - E, rounded down to a multiple of 2^-16, is what a pair's offset
// from the pixel centre is compared with; bit 0 of the verdict says the tile is `interior`
// (box_may_be_visible), which lets the sweep skip the clamp of the ring coordinates.
// A thread owns a VIEW (its five planes and facts stay in registers) and walks tiles; the eight floats of
// a tile's box are two broadcast loads.  (One thread per (tile, view), each re-reading the view's ~150
// bytes of tables, took 75 us at 6 M x 300 -- bound by L1 traffic.)
__global__ void __launch_bounds__(1024)
order_verdict_kernel(const float *__restrict__ box, int64_t n_tiles, const float4 *__restrict__ planes,
                     const ViewFacts *__restrict__ facts, int V, int v_pad, int cull, int exact_only,
                     uint16_t *__restrict__ verdict)
{
    const int v = blockIdx.y * blockDim.x + threadIdx.x;
    if (v >= v_pad) return;
    const bool real = v < V;
    ViewFacts f = {};
    float4 pl[5] = {};
    if (real) {
        f = facts[v];
#pragma unroll
        for (int j = 0; j < 5; ++j) pl[j] = planes[(size_t)v * 5 + j];
    }
    const unsigned slow = ((f.flags & kViewScreen) && !exact_only) ? kVerdictGeneral : kVerdictF64;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    unsigned out = kVerdictCull;
    if (real) {
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(box + tile * 8)), b1 = __ldg(reinterpret_cast<const float4 *>(box + tile * 8) + 1);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        out = slow;
        if (b[6] == 0.f) {                                                  // a non-finite member: never cull, never fast
            const float lo[3] = {b[0], b[1], b[2]}, hi[3] = {b[3], b[4], b[5]};
            float cz_lo;
            bool interior;
            const bool vis = box_may_be_visible(pl, lo, hi, cz_lo, interior);
            if (!vis && cull) {
                out = kVerdictCull;
            } else if (slow == kVerdictGeneral && (f.flags & kViewBorder) && cz_lo > 0.f) {
                const float a = (fmaxf(fabsf(lo[0]), fabsf(hi[0])) + fmaxf(fabsf(lo[1]), fabsf(hi[1])) + fmaxf(fabsf(lo[2]), fabsf(hi[2]))) * 1.000002f;
                const float ec = (f.g_rm * a + f.g_tm) * 1.000001f;
                const float k = (ec / cz_lo) * 1.000002f;                   // covers r <= (1 / cz)(1 + 2.1 u)
                const float E = (k * (f.fxh + f.span) * 1.000002f + 1.8119812e-07f * f.span + f.c0) * 1.000001f;
                const float room = 0.5f - E;                                // exact or rounded to nearest: inside c0's 1e-6
                if (room >= 0.25f && a < 1e15f) {                           // false for NaN
                    unsigned q = (unsigned)(room * 131072.f) - 1u;          // floor, minus one step for the rounding of `room`
                    q = q > 65533u ? 65533u : q;
                    out = (q & ~1u) | (interior ? 1u : 0u);                 // bit 0: the ring clamp cannot act (room >= 1/4: q >= 2)
                }
            }
        }
    }
    verdict[tile * v_pad + v] = (uint16_t)out;
    }
}

OrderWs order_layout(int64_t N, int V)
{
    OrderWs o;
    const int64_t n_pad = (N + kTile - 1) / kTile * kTile;
    const int64_t n_tiles = n_pad / kTile;
    const int64_t v_pad = (V + kWin - 1) / kWin * kWin;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
    o.sheet = take((size_t)((V + 3) / 4) * (size_t)n_pad * sizeof(uint32_t));
    o.pos_sorted = take((size_t)n_pad * sizeof(float4));
    o.perm = take((size_t)n_pad * sizeof(int32_t));
    o.keys = take((size_t)n_pad * sizeof(uint32_t));
    o.keys_sorted = take((size_t)n_pad * sizeof(uint32_t));
    o.idx = take((size_t)n_pad * sizeof(int32_t));
    o.sort_temp_bytes = sort_temp_capacity(N);
    o.sort_temp = take(o.sort_temp_bytes);
    o.stats = take(kStatsBytes);
    o.tilebox = take((size_t)n_tiles * 8 * sizeof(float));
    o.verdict = take((size_t)n_tiles * (size_t)(v_pad > 0 ? v_pad : kWin) * sizeof(uint16_t));
    o.views = take((size_t)(V > 0 ? V : 1) * sizeof(GslView));
    o.facts = take((size_t)(V > 0 ? V : 1) * sizeof(ViewFacts));
    o.hot = take((size_t)(v_pad > 0 ? v_pad : kWin) * sizeof(HotView));
    o.planes = take((size_t)(V > 0 ? V : 1) * 5 * sizeof(float4));
    o.bytes = off + 256;
    return o;
}

// Sort positions into Morton order (sort = false: keep the caller's order), box the tiles, and
// decide per (tile, view) how the sweep treats it.  All launches on `st`; the view tables are
// already in the workspace (gsl_lift_prepare).
int order_gaussians(const float *pos, int64_t N, int V, bool sort, bool exact_only, unsigned char *base, const OrderWs &L, cudaStream_t st)
{
    float4 *pos_sorted = reinterpret_cast<float4 *>(base + L.pos_sorted);
    int32_t *perm = reinterpret_cast<int32_t *>(base + L.perm);
    uint32_t *keys = reinterpret_cast<uint32_t *>(base + L.keys);
    uint32_t *keys_sorted = reinterpret_cast<uint32_t *>(base + L.keys_sorted);
    int32_t *idx = reinterpret_cast<int32_t *>(base + L.idx);
    unsigned long long *stats = reinterpret_cast<unsigned long long *>(base + L.stats);
    float *tilebox = reinterpret_cast<float *>(base + L.tilebox);
    uint16_t *verdict = reinterpret_cast<uint16_t *>(base + L.verdict);
    const ViewFacts *d_facts = reinterpret_cast<const ViewFacts *>(base + L.facts);
    float4 *planes = reinterpret_cast<float4 *>(base + L.planes);
    const int64_t n_tiles = (N + kTile - 1) / kTile;
    const int v_pad = (V + kWin - 1) / kWin * kWin;

    const unsigned rows_grid = (unsigned)((N + 255) / 256);
    if (sort) {
        GSL_CUDA_TRY(cudaMemsetAsync(stats, 0x00, kStatsBytes, st));
        int64_t blocks = (N + 256 * 8 - 1) / (256 * 8);
        const int64_t cap = (int64_t)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        order_stats_kernel<<<(unsigned)blocks, 256, 0, st>>>(pos, N, stats);
        GSL_LAUNCH_CHECK("order_stats_kernel");
        order_key_kernel<<<rows_grid, 256, 0, st>>>(pos, N, stats, keys, idx);
        GSL_LAUNCH_CHECK("order_key_kernel");
        if (int rc = sort_cells(keys, keys_sorted, idx, perm, N, base + L.sort_temp, L.sort_temp_bytes, st)) return rc;
    } else {
        order_iota_kernel<<<rows_grid, 256, 0, st>>>(perm, N);
        GSL_LAUNCH_CHECK("order_iota_kernel");
    }
    order_permute_kernel<<<rows_grid, 256, 0, st>>>(pos, N, perm, pos_sorted, tilebox);       // rows_grid == n_tiles
    GSL_LAUNCH_CHECK("order_permute_kernel");
    const int vthreads = v_pad < 1024 ? (v_pad + 31) / 32 * 32 : 1024;                          // a thread per view
    const dim3 vgrid((unsigned)std::min<int64_t>(n_tiles, (int64_t)sm_count() * (2048 / vthreads) * 2), (unsigned)((v_pad + vthreads - 1) / vthreads));
    order_verdict_kernel<<<vgrid, vthreads, 0, st>>>(tilebox, n_tiles, planes, d_facts, V, v_pad, sort ? 1 : 0, exact_only ? 1 : 0, verdict);
    GSL_LAUNCH_CHECK("order_verdict_kernel");
    return GSL_OK;
}

}  // namespace gsl
