// Declarations shared by lift.cu, lift_order.cu and lift_sort.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gslift.h"

namespace gsl {

constexpr int kSheetTile = 256;        // Gaussians per vote-sheet tile == gather block size

// ---------------------------------------------------------------------------------------
// Packed label maps are TILED: a map of seg_w x seg_h codes is stored as 16 x 8-pixel tiles of
// 128 bytes (one L1 line each; row r of a tile = 16 consecutive bytes), tile rows of
// tiles_x = ceil(seg_w / 16) + 2 tiles, tiles_y = ceil(seg_h / 8) + 2 tile rows.  The extra
// ring of tiles and the padding up to the tile multiple hold code 0 ("no vote"), so a pixel
// up to 16 columns / 8 rows outside the map can be fetched like any other and votes for
// nothing.  The lanes of a warp (spatially sorted Gaussians) project into a small 2-D patch;
// tiling turns that patch into a handful of lines instead of one line per image row.
// ---------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t map_tiles_x(int seg_w) { return (uint32_t)((seg_w + 15) >> 4) + 2u; }
__host__ __device__ inline uint32_t map_tiles_y(int seg_h) { return (uint32_t)((seg_h + 7) >> 3) + 2u; }
inline int64_t packed_map_bytes(int seg_w, int seg_h)
{
    return (int64_t)map_tiles_x(seg_w) * (int64_t)map_tiles_y(seg_h) * 128;
}
// byte offset of pixel (xs, ys), xs in [-16, seg_w + 15], ys in [-8, seg_h + 7]; pitch = tiles_x * 128
__device__ __forceinline__ uint32_t tiled_offset(uint32_t pitch, int xs, int ys)
{
    const uint32_t xb = (uint32_t)(xs + 16), yb = (uint32_t)(ys + 8);
    return (yb >> 3) * pitch + ((xb >> 4) << 7) + ((yb & 7u) << 4) + (xb & 15u);
}

// Device-side view, in two parts.  HotView is everything the float32 sweep reads per pair: 80
// bytes, 16-byte aligned, the views of a window back to back, so a view arrives in five wide
// uniform constant loads and the whole window is 1.3 KB of constant cache.
struct alignas(16) HotView {
    // the camera rounded to float32, rows 0 and 1 pre-multiplied by fx, fy; half_w / half_h hold
    // width/2 - 1/2, height/2 - 1/2 (lift.cu: screen_pair)
    float R[9], t[3], half_w, half_h;
    float x_hi, y_hi;       // width + 2, height + 2: clamp range of the screened coordinate
    int64_t map_offset;     // byte offset of the view's packed map
    uint32_t pitch_m128;    // pitch - 128 and the folded constant of the float-derived offset
    uint32_t addr_k;        //   (lift.cu: screen_pair)
};

// ColdView: the public GslView (float64 path) plus facts the host derives once per view.
struct ColdView {
    GslView g;
    uint32_t pitch;    // bytes per tile row of this view's packed map
    int wi, hi;        // width, height as integers
    int unit_scale;    // scale_x == 1 && scale_y == 1: the rescale of dls:281-282 is the identity
    int no_clamp;      // unit scale and the map covers the camera frame: dls:285-286 cannot fire
    int screen_ok;     // float32 screening applies: integer frame below 2^21 px, finite parameters
    int border_ok;     // unit scale and the map IS the camera frame: pixels just outside the frame
                       // land in the zero ring of the tiled map, no bounds test needed
};

// A window of views as the float64 kernel (lift_gather_kernel) takes it: by value, as a kernel
// parameter (constant bank 0, compile-time offsets).
template <int VW>
struct ViewWindow {
    ColdView c[VW];
};

// One window of the float32 sweep as it lies in device memory (workspace): a single launch covers
// many windows, block (tile, window) reads its window's entry.  gsl_lift_prepare uploads the table
// of all 16-view windows once (entries [0, ceil(V / 16))), before any label map is in flight, so
// the sweeps need no host-to-device copy of their own (it would queue behind the map uploads of a
// pipelined caller); 8-view windows (view_window <= 8) are built per call behind that table.
struct alignas(16) WinDev {
    HotView h[16];
    // constants of the screening bounds, valid for all views of the window (lift.cu: screen_pair):
    // ec = g_rm * (|X|+|Y|+|Z|) + g_tm  (camera coordinate) and
    // 1/2 - E = k * fxh_neg + room0 - (k + 3.04 u) |x|,  k = ec / cz  (image coordinate)
    float g_rm, g_tm, fxh_neg, room0;
    int n_live;        // views of the window (<= 16)
    int border;        // every view has border_ok
    int word0;         // first sheet word of the window = first_view / 4
    int first_view;
};

// Byte offsets of the pieces of the lifting workspace (each 256-byte aligned).
struct OrderWs {
    size_t sheet, pos_sorted, perm, keys, keys_sorted, idx, sort_temp, sort_temp_bytes, stats, tilebox, masks, views, planes, wins, bytes;
};

OrderWs order_layout(int64_t N, int V);
int order_gaussians(const float *pos, int64_t N, int V, unsigned char *base, const OrderWs &L, cudaStream_t st);

// lift_sort.cu: stable radix sort of (24-bit cell key, row index) pairs.
size_t sort_temp_capacity(int64_t N);
int sort_cells(const uint32_t *keys_in, uint32_t *keys_out, const int32_t *idx_in, int32_t *idx_out,
               int64_t N, void *temp, size_t temp_bytes, cudaStream_t st);

}  // namespace gsl
