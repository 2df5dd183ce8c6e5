// Declarations shared by lift.cu, lift_order.cu, lift_sort.cu and host_stage.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gslift.h"

namespace gsl {

constexpr int kLiftThreads = 64;       // threads of a sweep CTA
constexpr int kLiftPer = 4;            // Gaussians per thread (two packed float32x2 pairs)
constexpr int kTile = kLiftThreads * kLiftPer;    // Gaussians per tile == per sweep CTA == vote-sheet tile
constexpr int kWin = 16;               // views per window (staging unit of the view table)

// ---------------------------------------------------------------------------------------
// Packed label maps are stored in STRIPS: a map of seg_w x seg_h codes is cut into vertical
// strips 16 pixels wide; inside a strip the rows follow each other, 16 bytes each.  There are
// strips_x = ceil(seg_w / 16) + 2 strips of rows_pad = 8 * (ceil(seg_h / 8) + 2) rows: one ring
// strip on the left, at least one on the right, 8 ring rows above and at least 8 below, all
// holding code 0 ("no vote"), so a pixel up to 16 columns / 8 rows outside the map is fetched
// like any other and votes for nothing.  Byte offset of pixel (x, y), with X = x + 16, Y = y + 8:
//     off = (X >> 4) * (16 * rows_pad) + 16 * Y + (X & 15) = X + 16 Y + (X >> 4) * (16 rows_pad - 16)
// A 128-byte line is a 16 x 8-pixel block (8 aligned rows of one strip): the lanes of a warp
// (spatially sorted Gaussians) project into a small 2-D patch, which this layout turns into a
// handful of lines instead of one line per image row, and the offset is three integer
// operations on values the float32 screening has in registers anyway (lift.cu).
//
// Behind the strips of a map lies its COARSE table: one byte per 8 x 8-pixel cell (half a 128-byte
// line), the cell's code if all 64 pixels carry the same code, kMixed otherwise; coarse_w =
// 2 * strips_x cells per row, rows_pad / 8 rows, row-major:  offc = CY * coarse_w + CX.
// Label maps are piecewise constant, so most lookups are answered by this table, which is 64
// times smaller than the map: the 32 lookups of a warp fall into a few sectors, and the tables of
// all views of a scene (10 MB at 300 x 1920 x 1080) stay in L2.  A kMixed cell sends the lookup
// to the full-resolution strips.  Results are identical by construction.
// (The table was also tried in strips of 16 cells -- a 128-byte line = 16 x 8 cells: 2.0 tags per
// lookup, no change in kernel time, one more multiply-add per lookup; measured in round 2.)
// ---------------------------------------------------------------------------------------
constexpr uint32_t kMixed = 255u;      // coarse-table marker; label codes are 1..254 (GSL_MAX_CODES)
__host__ __device__ inline uint32_t map_strips_x(int seg_w) { return (uint32_t)((seg_w + 15) >> 4) + 2u; }
__host__ __device__ inline uint32_t map_rows_pad(int seg_h) { return ((uint32_t)((seg_h + 7) >> 3) + 2u) * 8u; }
__host__ __device__ inline int64_t map_fine_bytes(int seg_w, int seg_h)
{
    return (int64_t)map_strips_x(seg_w) * (int64_t)map_rows_pad(seg_h) * 16;
}
__host__ __device__ inline int64_t map_coarse_bytes(int seg_w, int seg_h)
{
    return ((int64_t)(2u * map_strips_x(seg_w)) * (int64_t)(map_rows_pad(seg_h) / 8u) + 15) / 16 * 16;
}
inline int64_t packed_map_bytes(int seg_w, int seg_h)
{
    return map_fine_bytes(seg_w, seg_h) + map_coarse_bytes(seg_w, seg_h);
}
// byte offset of pixel (xs, ys), xs in [-16, seg_w + 15], ys in [-8, seg_h + 7]; strip = 16 * rows_pad
__device__ __forceinline__ uint32_t strip_offset(uint32_t strip, int xs, int ys)
{
    const uint32_t xb = (uint32_t)(xs + 16), yb = (uint32_t)(ys + 8);
    return (xb >> 4) * strip + (yb << 4) + (xb & 15u);
}

// The pack kernels work one warp per 4 strips x 8 rows (lane = strip * 8 + row, `w` = the lane's 16
// codes): the 8 lanes of a strip hold one 128-byte line, i.e. two coarse cells.  Returns, for the
// lane's strip, the cell values of its left and right half (valid in every lane of the strip).
// All 32 lanes must call it.
__device__ __forceinline__ void coarse_cells_of_line(uint4 w, unsigned lane, uint32_t &left, uint32_t &right)
{
    const unsigned gmask = 0xffu << (lane & 24u);
    const uint32_t bl = w.x & 0xffu, br = w.z & 0xffu;
    const bool ul = w.x == bl * 0x01010101u && w.y == w.x, ur = w.z == br * 0x01010101u && w.w == w.z;
    const unsigned okl = __ballot_sync(0xffffffffu, ul), okr = __ballot_sync(0xffffffffu, ur);
    const unsigned ml = __match_any_sync(0xffffffffu, bl), mr = __match_any_sync(0xffffffffu, br);
    left = ((okl & gmask) == gmask && (ml & gmask) == gmask) ? bl : kMixed;
    right = ((okr & gmask) == gmask && (mr & gmask) == gmask) ? br : kMixed;
}

// Decodes thread index i into (map, strip, row) of the strip layout.  A warp's 32 consecutive
// indices are one group of 4 strips x 8 rows; strip may lie past strips_x in the last group of a row.
__device__ __forceinline__ void pack_coords(int64_t i, uint32_t strips_x, uint32_t rows_pad,
                                            int64_t &m, uint32_t &strip, uint32_t &row)
{
    const uint32_t groups_x = (strips_x + 3u) >> 2, groups_y = rows_pad >> 3;
    const int64_t per_map = (int64_t)groups_x * groups_y * 32;
    m = i / per_map;
    const uint32_t rem = (uint32_t)(i - m * per_map);
    const uint32_t grp = rem >> 5, lane = rem & 31u;
    const uint32_t gy = grp / groups_x, gx = grp - gy * groups_x;
    strip = gx * 4u + (lane >> 3);
    row = gy * 8u + (lane & 7u);
}

// Stores the lane's 16 codes and, from the first lane of every strip, the two coarse cells of the
// strip's 128-byte line.  Called by all 32 lanes of a warp (lanes past the map store nothing).
__device__ __forceinline__ void store_packed_row(uint8_t *__restrict__ packed, int64_t map_bytes, int64_t fine_bytes, int64_t m,
                                                 uint32_t strips_x, uint32_t rows_pad, uint32_t strip, uint32_t row, uint4 w, bool live)
{
    uint32_t left, right;
    coarse_cells_of_line(w, threadIdx.x & 31u, left, right);
    if (!live) return;
    uint8_t *base = packed + m * map_bytes;
    reinterpret_cast<uint4 *>(base)[(int64_t)strip * rows_pad + row] = w;
    if ((row & 7u) == 0u)
        *reinterpret_cast<unsigned short *>(base + fine_bytes + (int64_t)(row >> 3) * (2u * strips_x) + 2u * strip) = (unsigned short)(left | right << 8);
}

// Everything the float32 sweep reads per (Gaussian, view) pair: 96 bytes, 16-byte aligned, the
// views of a window back to back in shared memory (six broadcast LDS.128 per view).
struct alignas(16) HotView {
    // the camera rounded to float32, rows 0 and 1 pre-multiplied by fx, fy
    float R[9], t[3];
    float hwp, hhp;         // width/2 - 1/2 + 16, height/2 - 1/2 + 8: image coordinate minus 1/2, in ring coordinates
    uint32_t xmax_bits;     // float bits of width + 18 / height + 10: unsigned-min clamp of the ring coordinate
    uint32_t ymax_bits;
    uint32_t strip_m16;     // 16 * rows_pad - 16
    uint32_t addr_k;        // folded constant of the float-derived offset (lift.cu: fast_pair), modulo 2^32
    uint64_t map;           // byte offset of the view's packed map; the sweep's staged copy holds its address
    uint32_t coarse_w;      // cells per row of the coarse table
    uint32_t caddr_k;       // spare
    uint64_t cmap;          // byte offset (staged copy: address) of the view's coarse table, minus the folded
                            // constant of the float-derived cell offset (lift.cu: fast_pair2)
};
static_assert(sizeof(HotView) == 96, "HotView is six 16-byte words");
constexpr int kHotWords = sizeof(HotView) / 16;

// What the other code paths need of a view (per-pair bound, rescaled maps, float64 evaluation):
// the public GslView plus facts the host derives once.
struct ViewFacts {
    float g_rm, g_tm;       // ec = g_rm * (|X|+|Y|+|Z|) + g_tm bounds the error of a camera coordinate
    float fxh;              // >= max(5, |fx| + width/2, |fy| + height/2) (1 + 2e-6)
    float c0;               // 2.01 u max(half_w, half_h) + 1e-6 (+ ring shift)
    float span;             // max(width, height) + 18: largest |ring coordinate| after the clamp
    int wi, hi;             // width, height as integers (0 when the frame is not integral)
    int flags;              // kViewScreen | kViewBorder
    uint32_t strip;         // 16 * rows_pad of the view's packed map
    uint32_t pad_;
};
constexpr int kViewScreen = 1;    // float32 screening applies: integer frame below 2^21 px, finite parameters
constexpr int kViewBorder = 2;    // unit scale and the map IS the camera frame: out-of-frame pixels read the zero ring

// Per (tile, view) verdict of the culling pass, 16 bits:
//   0        no Gaussian of the tile can be visible in the view: skipped
//   1..65533 fast path: every Gaussian of the tile is in front of the camera and the float32 error
//            of an image coordinate is below E for all of them; value = floor((1/2 - E) * 2^17)
//   65534    exact path: float64 expressions for every pair (views the screening does not cover)
//   65535    general path: float32 screening with a per-pair bound
constexpr unsigned kVerdictCull = 0u, kVerdictF64 = 65534u, kVerdictGeneral = 65535u;

// Byte offsets of the pieces of the lifting workspace (each 256-byte aligned).
struct OrderWs {
    size_t sheet, pos_sorted, perm, keys, keys_sorted, idx, sort_temp, sort_temp_bytes, stats, tilebox, verdict, views, facts, hot, planes, bytes;
};

void fill_view_planes(const GslView &w, float4 (&planes)[5]);
OrderWs order_layout(int64_t N, int V);
int order_gaussians(const float *pos, int64_t N, int V, bool sort, bool exact_only, unsigned char *base, const OrderWs &L, cudaStream_t st);

// lift_sort.cu: stable radix sort of (24-bit cell key, row index) pairs.
size_t sort_temp_capacity(int64_t N);
int sort_cells(const uint32_t *keys_in, uint32_t *keys_out, const int32_t *idx_in, int32_t *idx_out,
               int64_t N, void *temp, size_t temp_bytes, cudaStream_t st);

}  // namespace gsl
