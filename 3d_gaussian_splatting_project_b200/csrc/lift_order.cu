// Spatial ordering and per-tile frustum culling for the lifting kernels, sm_100a.
//
// The vote of a Gaussian depends only on its own position, so the order in which Gaussians are
// processed is free.  Processing them in a spatially coherent order makes every 256-Gaussian
// tile of the gather kernel a small box in space, and then a whole (tile, view) can be proven
// invisible -- behind the camera or outside the image for every point of the box -- with a few
// interval evaluations, and skipped.  On the reference's own cameras (bundled cameras.json)
// only ~13 % of (Gaussian, view) pairs are visible, so this removes most of the work; on the
// synthetic look-at scenes it removes the ~25 % that is invisible.  Labels are unaffected: the
// gather kernel still evaluates the reference's exact test for every pair it does not skip, and
// the cull is conservative (margins ten times its own float32 rounding error, never the other way).
//
//   order_bbox_kernel     min/max of the finite coordinates (ordered-uint atomics)
//   order_cell_kernel     16^3 grid over the box, cell id = 12-bit Morton code (non-finite
//                         positions go to an extra last cell); per-CTA shared-memory histogram
//                         of a 16 K-row chunk, flushed with one global atomic per occupied cell
//   order_scan_kernel     exclusive scan of the 4097 counts (one CTA)
//   order_scatter_kernel  counting-sort scatter, same chunks: a CTA reserves a range per cell
//                         with one global atomic and ranks its rows inside it in shared memory.
//                         The order inside a cell depends on atomic arrival; results do not.
//   order_tilebox_kernel  bounding box (+ non-finite flag) of each run of 256 sorted Gaussians
//   order_planes_kernel   per view, the five half-spaces of the visibility test as linear forms
//   order_cull_kernel     one bit per (tile, view), 16 views to a mask word, all views in one launch
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

__global__ void __launch_bounds__(256)
order_bbox_kernel(const float *__restrict__ pos, int64_t N, unsigned *__restrict__ bbox)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pos[3 * i + a];
            if (fabsf(v) < INFINITY) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
        }
    }
    __shared__ float red[6][8];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { red[a][threadIdx.x >> 5] = lo[a]; red[3 + a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {                       // one atomic per CTA and bound
        float v = red[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        if (threadIdx.x < 3) atomicMin(bbox + threadIdx.x, enc_f32(v));
        else atomicMax(bbox + threadIdx.x, enc_f32(v));
    }
}

__device__ __forceinline__ unsigned spread4(unsigned v)      // 4 bits -> every third bit
{
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6);
}

__device__ __forceinline__ unsigned cell_of(const float *__restrict__ p, const float (&lo)[3], const float (&inv)[3])
{
    unsigned q[3];
    bool finite = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float v = p[a];
        finite = finite && (fabsf(v) < INFINITY);
        q[a] = (unsigned)min(max((int)((v - lo[a]) * inv[a]), 0), 15);
    }
    return finite ? (spread4(q[0]) | (spread4(q[1]) << 1) | (spread4(q[2]) << 2)) : (unsigned)kOrderCells;
}

__device__ __forceinline__ void load_grid(const unsigned *__restrict__ bbox, float (&lo)[3], float (&inv)[3])
{
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = dec_f32(bbox[a]);
        const float ext = dec_f32(bbox[3 + a]) - lo[a];
        inv[a] = ext > 0.f && ext < INFINITY ? 16.f / ext : 0.f;
    }
}

constexpr int kOrderChunk = 16384;     // rows per CTA in the cell / scatter kernels

__global__ void __launch_bounds__(512)
order_cell_kernel(const float *__restrict__ pos, int64_t N, const unsigned *__restrict__ bbox,
                  uint16_t *__restrict__ cell, unsigned *__restrict__ hist)
{
    __shared__ unsigned local[kOrderCells + 1];
    for (int i = threadIdx.x; i <= kOrderCells; i += blockDim.x) local[i] = 0u;
    float lo[3], inv[3];
    load_grid(bbox, lo, inv);
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * kOrderChunk;
    const int64_t r1 = min(r0 + kOrderChunk, N);
    for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
        const unsigned c = cell_of(pos + 3 * i, lo, inv);
        cell[i] = (uint16_t)c;
        atomicAdd(local + c, 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= kOrderCells; i += blockDim.x)
        if (local[i]) atomicAdd(hist + i, local[i]);
}

// One CTA of 1024 threads; hist[0 .. kOrderCells] -> exclusive offsets in place.
__global__ void __launch_bounds__(1024)
order_scan_kernel(unsigned *__restrict__ hist)
{
    __shared__ unsigned part[1024];
    constexpr int n = kOrderCells + 1;
    constexpr int seg = (n + 1023) / 1024;
    const int t = threadIdx.x;
    const int lo = min(t * seg, n), hi = min(lo + seg, n);
    unsigned s = 0;
    for (int i = lo; i < hi; ++i) s += hist[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // Hillis-Steele inclusive scan
        const unsigned v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned run = part[t] - s;
    for (int i = lo; i < hi; ++i) {
        const unsigned c = hist[i];
        hist[i] = run;
        run += c;
    }
}

__global__ void __launch_bounds__(512)
order_scatter_kernel(const float *__restrict__ pos, int64_t N, const uint16_t *__restrict__ cell,
                     unsigned *__restrict__ cursor, float *__restrict__ pos_sorted, int32_t *__restrict__ perm)
{
    __shared__ unsigned local[kOrderCells + 1];     // count of this chunk per cell, then its reserved base
    for (int i = threadIdx.x; i <= kOrderCells; i += blockDim.x) local[i] = 0u;
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * kOrderChunk;
    const int64_t r1 = min(r0 + kOrderChunk, N);
    for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) atomicAdd(local + cell[i], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i <= kOrderCells; i += blockDim.x)
        if (local[i]) local[i] = atomicAdd(cursor + i, local[i]);      // reserve [base, base + count)
    __syncthreads();
    for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
        const unsigned dst = atomicAdd(local + cell[i], 1u);
        pos_sorted[3 * (int64_t)dst + 0] = pos[3 * i + 0];
        pos_sorted[3 * (int64_t)dst + 1] = pos[3 * i + 1];
        pos_sorted[3 * (int64_t)dst + 2] = pos[3 * i + 2];
        perm[dst] = (int32_t)i;
    }
}

// One warp per tile of kSheetTile sorted Gaussians: box[tile] = {lo xyz, hi xyz, nonfinite, 0}.
__global__ void __launch_bounds__(256)
order_tilebox_kernel(const float *__restrict__ pos_sorted, int64_t N, int64_t n_tiles, float *__restrict__ box)
{
    const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (tile >= n_tiles) return;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int bad = 0;
    for (int r = lane; r < kSheetTile; r += 32) {
        const int64_t g = tile * kSheetTile + r;
        if (g >= N) break;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pos_sorted[3 * g + a];
            if (fabsf(v) < INFINITY) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); } else bad = 1;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        float *b = box + tile * 8;
        b[0] = lo[0]; b[1] = lo[1]; b[2] = lo[2]; b[3] = hi[0]; b[4] = hi[1]; b[5] = hi[2];
        b[6] = bad ? 1.f : 0.f; b[7] = 0.f;
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
__global__ void __launch_bounds__(128)
order_planes_kernel(const GslView *__restrict__ views, int V, float4 *__restrict__ planes)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const GslView w = views[v];
    const double kx[5] = {0.0, w.fx, w.fx, 0.0, 0.0};
    const double ky[5] = {0.0, 0.0, 0.0, w.fy, w.fy};
    const double kz[5] = {1.0, w.half_w, w.half_w - w.width, w.half_h, w.half_h - w.height};
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        float4 o;
        o.x = (float)(kx[p] * w.R[0] + ky[p] * w.R[3] + kz[p] * w.R[6]);
        o.y = (float)(kx[p] * w.R[1] + ky[p] * w.R[4] + kz[p] * w.R[7]);
        o.z = (float)(kx[p] * w.R[2] + ky[p] * w.R[5] + kz[p] * w.R[8]);
        o.w = (float)(kx[p] * w.t[0] + ky[p] * w.t[1] + kz[p] * w.t[2]);
        planes[v * 5 + p] = o;
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
__device__ __forceinline__ bool box_may_be_visible(const float4 *__restrict__ pl, const float (&lo)[3], const float (&hi)[3])
{
    const float rel = 1e-5f, px = 1e-6f;
    float mn, mx, mag;
    lin_range(pl[0], lo, hi, mn, mx, mag);
    if (mx < -(rel * mag + 1e-30f)) return false;                          // every point has z <= 0
    const float zpos = fmaxf(mx, 0.f) * px;
    lin_range(pl[1], lo, hi, mn, mx, mag);
    if (mx < -(zpos + rel * mag)) return false;                            // x < 0 everywhere
    lin_range(pl[2], lo, hi, mn, mx, mag);
    if (mn > zpos + rel * mag) return false;                               // x >= width everywhere
    lin_range(pl[3], lo, hi, mn, mx, mag);
    if (mx < -(zpos + rel * mag)) return false;                            // y < 0 everywhere
    lin_range(pl[4], lo, hi, mn, mx, mag);
    if (mn > zpos + rel * mag) return false;                               // y >= height everywhere
    return true;
}

// Thread per (tile, view); 16 consecutive lanes share a tile and fill one 16-bit mask word:
// masks[tile * n_words16 + v / 16], bit v % 16.
__global__ void __launch_bounds__(256)
order_cull_kernel(const float *__restrict__ box, int64_t n_tiles, const float4 *__restrict__ planes, int V,
                  uint16_t *__restrict__ masks, int n_words16)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int vpad = n_words16 * 16;
    const int64_t tile = idx / vpad;
    const int v = (int)(idx - tile * vpad);
    bool vis = false;
    if (tile < n_tiles && v < V) {
        const float *b = box + tile * 8;
        if (b[6] != 0.f) {
            vis = true;                                                     // non-finite member: never cull
        } else {
            const float lo[3] = {b[0], b[1], b[2]}, hi[3] = {b[3], b[4], b[5]};
            vis = box_may_be_visible(planes + (size_t)v * 5, lo, hi);
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, vis);
    const int lane = threadIdx.x & 31;
    if (tile < n_tiles && (lane & 15) == 0)
        masks[tile * n_words16 + (v >> 4)] = (uint16_t)((bal >> (lane & 16)) & 0xffffu);
}

OrderWs order_layout(int64_t N, int V)
{
    OrderWs o;
    const int64_t n_pad = (N + kSheetTile - 1) / kSheetTile * kSheetTile;
    const int64_t n_tiles = n_pad / kSheetTile;
    const int64_t n_words16 = (V + 15) / 16;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
    o.sheet = take((size_t)((V + 3) / 4) * (size_t)n_pad * sizeof(uint32_t));
    o.pos_sorted = take((size_t)n_pad * 3 * sizeof(float));
    o.perm = take((size_t)n_pad * sizeof(int32_t));
    o.cell = take((size_t)n_pad * sizeof(uint16_t));
    o.hist = take((size_t)(kOrderCells + 1) * sizeof(unsigned));
    o.bbox = take(8 * sizeof(unsigned));
    o.tilebox = take((size_t)n_tiles * 8 * sizeof(float));
    o.masks = take((size_t)n_tiles * (size_t)n_words16 * sizeof(uint16_t));
    o.views = take((size_t)(V > 0 ? V : 1) * sizeof(GslView));
    o.planes = take((size_t)(V > 0 ? V : 1) * 5 * sizeof(float4));
    o.bytes = off + 256;
    return o;
}

// Sort positions into cells, box the tiles, and decide per (tile, view) whether the view must be
// swept.  All launches on `st`; `views` is the caller's host table.
int order_gaussians(const float *pos, int64_t N, const GslView *views, int V, unsigned char *base,
                    const OrderWs &L, cudaStream_t st)
{
    float *pos_sorted = reinterpret_cast<float *>(base + L.pos_sorted);
    int32_t *perm = reinterpret_cast<int32_t *>(base + L.perm);
    uint16_t *cell = reinterpret_cast<uint16_t *>(base + L.cell);
    unsigned *hist = reinterpret_cast<unsigned *>(base + L.hist);
    unsigned *bbox = reinterpret_cast<unsigned *>(base + L.bbox);
    float *tilebox = reinterpret_cast<float *>(base + L.tilebox);
    uint16_t *masks = reinterpret_cast<uint16_t *>(base + L.masks);
    GslView *d_views = reinterpret_cast<GslView *>(base + L.views);
    float4 *planes = reinterpret_cast<float4 *>(base + L.planes);
    const int64_t n_tiles = (N + kSheetTile - 1) / kSheetTile;
    const int n_words16 = (V + 15) / 16;

    // pageable source: the runtime stages the table before returning
    GSL_CUDA_TRY(cudaMemcpyAsync(d_views, views, sizeof(GslView) * (size_t)V, cudaMemcpyHostToDevice, st));
    GSL_CUDA_TRY(cudaMemsetAsync(bbox, 0xff, 3 * sizeof(unsigned), st));
    GSL_CUDA_TRY(cudaMemsetAsync(bbox + 3, 0x00, 3 * sizeof(unsigned), st));
    GSL_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)(kOrderCells + 1) * sizeof(unsigned), st));
    int64_t blocks = (N + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    order_bbox_kernel<<<(unsigned)blocks, 256, 0, st>>>(pos, N, bbox);
    GSL_LAUNCH_CHECK("order_bbox_kernel");
    const unsigned chunks = (unsigned)((N + kOrderChunk - 1) / kOrderChunk);
    order_cell_kernel<<<chunks, 512, 0, st>>>(pos, N, bbox, cell, hist);
    GSL_LAUNCH_CHECK("order_cell_kernel");
    order_scan_kernel<<<1, 1024, 0, st>>>(hist);
    GSL_LAUNCH_CHECK("order_scan_kernel");
    order_scatter_kernel<<<chunks, 512, 0, st>>>(pos, N, cell, hist, pos_sorted, perm);
    GSL_LAUNCH_CHECK("order_scatter_kernel");
    order_tilebox_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, st>>>(pos_sorted, N, n_tiles, tilebox);
    GSL_LAUNCH_CHECK("order_tilebox_kernel");
    order_planes_kernel<<<(V + 127) / 128, 128, 0, st>>>(d_views, V, planes);
    GSL_LAUNCH_CHECK("order_planes_kernel");
    const int64_t threads = n_tiles * (int64_t)n_words16 * 16;
    order_cull_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(tilebox, n_tiles, planes, V, masks, n_words16);
    GSL_LAUNCH_CHECK("order_cull_kernel");
    return GSL_OK;
}

}  // namespace gsl
