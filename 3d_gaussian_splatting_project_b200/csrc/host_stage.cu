// Hybrid label-map staging.  lift_labels from HOST int32 maps is bound by
// the PCIe transfer of 4 bytes per pixel.  The host cores can narrow maps to the 1-byte codes
// while the DMA engine moves other maps as int32, so that both resources work in parallel:
//
//   gsl_host_pack_labels   HOST helper (no device work): int32 maps -> uint8 codes, row-major, on
//                          n_threads host threads; also reports the value range it saw
//   gsl_tile_codes         device: row-major uint8 codes -> the packed strip layout of
//                          lift_internal.cuh (what gsl_pack_labels produces from int32 maps)
//
// The votes themselves are computed on the device either way; this only changes which side turns
// an int32 pixel into a code before it crosses the bus.
#include <limits.h>
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"
#include "lift_internal.cuh"

namespace {

void narrow_scalar(const int32_t *in, uint8_t *out, int64_t n, int label_min, int n_classes, int &b, int &lo, int &hi)
{
    for (int64_t i = 0; i < n; ++i) {
        const int v = in[i];
        const unsigned c = (unsigned)v - (unsigned)label_min;      // wraps, like the vector path
        b |= c >= (unsigned)n_classes;
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
        out[i] = (uint8_t)(c < (unsigned)n_classes ? c + 1u : 0u);
    }
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
// 32 pixels per iteration: four 8 x int32 vectors -> one 32 x uint8 vector.
__attribute__((target("avx2")))
int64_t narrow_avx2(const int32_t *in, uint8_t *out, int64_t n, int label_min, int n_classes, int &b, int &lo, int &hi)
{
    const __m256i vmin = _mm256_set1_epi32(label_min), one = _mm256_set1_epi32(1);
    const __m256i sign = _mm256_set1_epi32((int)0x80000000u);
    const __m256i lim = _mm256_set1_epi32((int)((unsigned)n_classes ^ 0x80000000u));     // unsigned c < n  <=>  (c ^ sign) < (n ^ sign) signed
    const __m256i fix = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    __m256i vlo = _mm256_set1_epi32(lo), vhi = _mm256_set1_epi32(hi), vbad = _mm256_setzero_si256();
    int64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i code[4];
        for (int j = 0; j < 4; ++j) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(in + i + 8 * j));
            vlo = _mm256_min_epi32(vlo, v);
            vhi = _mm256_max_epi32(vhi, v);
            const __m256i c = _mm256_sub_epi32(v, vmin);
            const __m256i ok = _mm256_cmpgt_epi32(lim, _mm256_xor_si256(c, sign));
            vbad = _mm256_or_si256(vbad, _mm256_andnot_si256(ok, one));
            code[j] = _mm256_and_si256(_mm256_add_epi32(c, one), ok);
        }
        const __m256i ab = _mm256_packus_epi32(code[0], code[1]), cd = _mm256_packus_epi32(code[2], code[3]);
        const __m256i bytes = _mm256_permutevar8x32_epi32(_mm256_packus_epi16(ab, cd), fix);
        _mm256_storeu_si256(reinterpret_cast<__m256i *>(out + i), bytes);
    }
    alignas(32) int t[8];
    _mm256_store_si256(reinterpret_cast<__m256i *>(t), vlo);
    for (int j = 0; j < 8; ++j) lo = t[j] < lo ? t[j] : lo;
    _mm256_store_si256(reinterpret_cast<__m256i *>(t), vhi);
    for (int j = 0; j < 8; ++j) hi = t[j] > hi ? t[j] : hi;
    _mm256_store_si256(reinterpret_cast<__m256i *>(t), vbad);
    for (int j = 0; j < 8; ++j) b |= t[j];
    return i;
}
#define GSL_HAVE_AVX2 1
#endif

void narrow_range(const int32_t *in, uint8_t *out, int64_t n, int label_min, int n_classes, int *bad, int *lo_out, int *hi_out)
{
    int b = 0, lo = INT_MAX, hi = INT_MIN;
    int64_t done = 0;
#ifdef GSL_HAVE_AVX2
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) done = narrow_avx2(in, out, n, label_min, n_classes, b, lo, hi);
#endif
    narrow_scalar(in + done, out + done, n - done, label_min, n_classes, b, lo, hi);
    *bad |= b;
    if (lo < *lo_out) *lo_out = lo;
    if (hi > *hi_out) *hi_out = hi;
}

}  // namespace

// maps[m] = host pointer to n_px[m] int32 values; codes of map m are written at out + sum(n_px[0..m)).
// minmax[2] (host, in/out): running min and max of the values seen.  *bad (host, out): 1 if a value
// fell outside [label_min, label_min + n_classes) (its code is 0).
extern "C" int gsl_host_pack_labels(const int32_t *const *maps, const int64_t *n_px, int n_maps, int label_min,
                                    int n_classes, uint8_t *out, int n_threads, int *minmax, int *bad)
{
    if (n_maps < 0 || (n_maps > 0 && (!maps || !n_px || !out)) || !minmax || !bad)
        return gsl::fail(GSL_EINVAL, "gsl_host_pack_labels: null pointer or negative count");
    if (n_classes < 1 || n_classes > GSL_MAX_CODES) return gsl::fail(GSL_EINVAL, "gsl_host_pack_labels: n_classes %d not in [1, %d]", n_classes, GSL_MAX_CODES);
    std::vector<int64_t> start((size_t)n_maps + 1, 0);
    for (int m = 0; m < n_maps; ++m) {
        if (n_px[m] < 0 || (n_px[m] > 0 && !maps[m])) return gsl::fail(GSL_EINVAL, "gsl_host_pack_labels: map %d is null or has a negative size", m);
        start[(size_t)m + 1] = start[(size_t)m] + n_px[m];
    }
    const int64_t total = start[(size_t)n_maps];
    *bad = 0;
    if (total == 0) return GSL_OK;
    if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > total / 65536 + 1) n_threads = (int)(total / 65536 + 1);
    std::vector<int> t_bad((size_t)n_threads, 0), t_lo((size_t)n_threads, INT_MAX), t_hi((size_t)n_threads, INT_MIN);
    auto work = [&](int t) {
        const int64_t a = total * t / n_threads, b = total * (t + 1) / n_threads;
        int m = 0;
        while (start[(size_t)m + 1] <= a) ++m;
        for (int64_t at = a; at < b;) {
            const int64_t end = start[(size_t)m + 1] < b ? start[(size_t)m + 1] : b;
            narrow_range(maps[m] + (at - start[(size_t)m]), out + at, end - at, label_min, n_classes,
                         &t_bad[(size_t)t], &t_lo[(size_t)t], &t_hi[(size_t)t]);
            at = end;
            ++m;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (int t = 0; t < n_threads; ++t) {
        *bad |= t_bad[(size_t)t];
        if (t_lo[(size_t)t] < minmax[0]) minmax[0] = t_lo[(size_t)t];
        if (t_hi[(size_t)t] > minmax[1]) minmax[1] = t_hi[(size_t)t];
    }
    return GSL_OK;
}

namespace gsl {

// One thread per 16-byte row of the output, a warp = 4 adjacent strips x 8 rows (same mapping as
// pack_labels_kernel in lift.cu): 8 runs of 64 consecutive codes in, four 128-byte lines and their
// eight coarse cells out.
__global__ void __launch_bounds__(256)
tile_codes_kernel(const uint8_t *__restrict__ codes, uint8_t *__restrict__ packed, int n_maps, int seg_w, int seg_h,
                  uint32_t strips_x, uint32_t rows_pad, int64_t total, int vec_ok)
{
    const int64_t fine_bytes = map_fine_bytes(seg_w, seg_h), map_bytes = fine_bytes + map_coarse_bytes(seg_w, seg_h);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {      // total % 32 == 0: warps stay whole
        int64_t m;
        uint32_t strip, row;
        pack_coords(i, strips_x, rows_pad, m, strip, row);
        const int y = (int)row - 8, x0 = (int)(strip * 16) - 16;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if (y >= 0 && y < seg_h && x0 >= 0 && x0 < seg_w) {
            const uint8_t *src = codes + (m * seg_h + y) * (int64_t)seg_w + x0;
            if (vec_ok && x0 + 16 <= seg_w) {
                w = __ldcs(reinterpret_cast<const uint4 *>(src));
            } else {
                uint32_t v[4] = {0u, 0u, 0u, 0u};
                for (int j = 0; j < 16; ++j)
                    if (x0 + j < seg_w) v[j >> 2] |= (uint32_t)src[j] << (8 * (j & 3));
                w = make_uint4(v[0], v[1], v[2], v[3]);
            }
        }
        store_packed_row(packed, map_bytes, fine_bytes, m, strips_x, rows_pad, strip, row, w, strip < strips_x);
    }
}

}  // namespace gsl

extern "C" int gsl_tile_codes(const uint8_t *codes, int n_maps, int seg_w, int seg_h, uint8_t *packed, void *stream)
{
    using namespace gsl;
    if (n_maps < 0 || seg_w < 1 || seg_h < 1) return fail(GSL_EINVAL, "gsl_tile_codes: negative count or empty map shape");
    if (n_maps == 0) return GSL_OK;
    if (!codes || !packed) return fail(GSL_EINVAL, "gsl_tile_codes: null pointer");
    if ((uintptr_t)packed & 15) return fail(GSL_EINVAL, "gsl_tile_codes: packed must be 16-byte aligned");
    if (packed_map_bytes(seg_w, seg_h) > 0x7fffffffLL) return fail(GSL_EINVAL, "gsl_tile_codes: map of %d x %d exceeds 2^31 packed bytes", seg_w, seg_h);
    const uint32_t sx = map_strips_x(seg_w), rp = map_rows_pad(seg_h);
    const int64_t total = (int64_t)((sx + 3) / 4) * (rp / 8) * 32 * n_maps;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    const int vec_ok = ((uintptr_t)codes & 15) == 0 && (seg_w & 15) == 0;      // every 16-pixel run starts 16-byte aligned
    tile_codes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(codes, packed, n_maps, seg_w, seg_h, sx, rp, total, vec_ok);
    GSL_LAUNCH_CHECK("tile_codes_kernel");
    return GSL_OK;
}
