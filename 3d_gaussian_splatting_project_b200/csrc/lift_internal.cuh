// Declarations shared by lift.cu and lift_order.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gslift.h"

namespace gsl {

constexpr int kSheetTile = 256;        // Gaussians per vote-sheet tile == gather block size
constexpr int kOrderCells = 4096;      // 16^3 ordering grid (+1 extra cell for non-finite positions)

// Device-side view: the public GslView plus facts the host derives once per view.
struct DevView {
    GslView g;
    int unit_scale;    // scale_x == 1 && scale_y == 1: the rescale of dls:281-282 is the identity
    int no_clamp;      // unit scale and the map covers the camera frame: dls:285-286 cannot fire
    // float32 screening (lift.cu: screen_pair): the camera rounded to float32 and the two
    // constants of the error bound  Ec = g_rm * (|X|+|Y|+|Z|) + g_tm  on a camera coordinate
    int screen_ok;     // bounds are integers below 2^21 and every parameter is finite
    float R[9], t[3], fx, fy, half_w, half_h;
    float fx_abs, fy_abs, g_rm, g_tm;
    int wi, hi;        // width, height as integers
};

// A window of views travels as a kernel parameter (constant bank 0, compile-time offsets).
template <int VW>
struct ViewWindow {
    DevView v[VW];
};

// Byte offsets of the pieces of the lifting workspace (each 256-byte aligned).
struct OrderWs {
    size_t sheet, pos_sorted, perm, cell, hist, bbox, tilebox, masks, views, planes, bytes;
};

OrderWs order_layout(int64_t N, int V);
int order_gaussians(const float *pos, int64_t N, const GslView *views, int V, unsigned char *base,
                    const OrderWs &L, cudaStream_t st);

}  // namespace gsl
