/*
 * viewer_oracle.c -- CPU restatement of the viewer-side label consumers.  TEST INFRASTRUCTURE ONLY
 * (same rule as gsl_oracle.c: only tests/, smoke() and bench.py's CPU legs may load it).
 *
 * Restates, statement by statement, two loops of the reference's viewer worker
 * (/root/reference/Web_Viewer_Gaussians_Selection/gaussians_selection.js, `gs`):
 *     runSort            gs:417-462   (without the early-out of gs:421-425)
 *     performHitTesting  gs:361-395   with project() gs:398-405
 *
 * PARITY UNPINNED BY EXECUTION: the build container has no JavaScript engine (node, browsers:
 * absent) and the reference ships no test vectors for these loops, so this file cannot be checked
 * against a run of the reference.  It is pinned instead to the language semantics it relies on
 * (ECMA-262: Number = IEEE-754 binary64, `|` applies ToInt32, integer-indexed exotic objects ignore
 * out-of-range and non-integer keys, `x++` on undefined yields NaN) through hand-computed known
 * answers in tests/test_viewer_oracle.py and an independent pure-Python restatement
 * (oracle/oracle.py: viewer_depth_sort_py / viewer_hit_test_py).
 * Math.hypot is implementation-approximated in ECMA-262; V8's builtin (src/builtins/math.tq,
 * MathHypot: scale by the largest magnitude, Kahan-summed squares, sqrt, rescale) is restated here.
 *
 * -ffp-contract=off (oracle/Makefile) keeps a*b + c*d unfused, as JavaScript evaluates it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ECMA-262 7.1.6 ToInt32. */
static int32_t js_to_int32(double d)
{
    if (!isfinite(d)) return 0;
    double t = trunc(d);
    double m = fmod(t, 4294967296.0);
    if (m < 0) m += 4294967296.0;
    return (int32_t)(uint32_t)(uint64_t)m;
}

/*
 * runSort, gs:427-457.  pos: Gaussian i at pos + i*stride floats (f_buffer[8*i + 0..2], gs:437).
 * depth_index: Uint32Array(vertexCount), zero-initialised by its constructor (gs:453).
 */
void orc_viewer_depth_sort(const float *pos, int64_t n, int stride, const double *view_proj, uint32_t *depth_index)
{
    double max_depth = -INFINITY, min_depth = INFINITY;                      /* gs:432-433 */
    int32_t *size_list = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));   /* gs:434 */
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pos + i * stride;
        double s = view_proj[2] * (double)p[0] + view_proj[6] * (double)p[1] + view_proj[10] * (double)p[2];
        int32_t depth = js_to_int32(s * 4096);                               /* gs:437 */
        size_list[i] = depth;
        if (depth > max_depth) max_depth = depth;                            /* gs:439-440 */
        if (depth < min_depth) min_depth = depth;
    }
    double depth_inv = (256 * 256) / (max_depth - min_depth);                /* gs:443 */
    uint32_t *counts0 = (uint32_t *)calloc(256 * 256, sizeof(uint32_t));     /* gs:444 */
    for (int64_t i = 0; i < n; ++i) {
        size_list[i] = js_to_int32(((double)size_list[i] - min_depth) * depth_inv);     /* gs:446 */
        /* gs:447 counts0[sizeList[i]]++ : a typed array ignores an index outside [0, length) */
        if (size_list[i] >= 0 && size_list[i] < 256 * 256) counts0[size_list[i]]++;
    }
    uint32_t *starts0 = (uint32_t *)calloc(256 * 256, sizeof(uint32_t));     /* gs:450 */
    for (int i = 1; i < 256 * 256; ++i) starts0[i] = starts0[i - 1] + counts0[i - 1];   /* gs:451 */
    memset(depth_index, 0, (size_t)(n > 0 ? n : 0) * sizeof(uint32_t));      /* gs:453 */
    for (int64_t i = 0; i < n; ++i) {
        /* gs:455-456: starts0[k]++ with k outside the array reads undefined -> NaN, stores nothing,
         * and depthIndex[NaN] = i stores nothing either */
        if (size_list[i] >= 0 && size_list[i] < 256 * 256) {
            uint32_t sorted_index = starts0[size_list[i]]++;
            depth_index[sorted_index] = (uint32_t)i;
        }
    }
    free(size_list);
    free(counts0);
    free(starts0);
}

/* The bucket of every Gaussian (gs:446), for tests that want to see the intermediate. */
void orc_viewer_buckets(const float *pos, int64_t n, int stride, const double *view_proj, int32_t *bucket)
{
    double max_depth = -INFINITY, min_depth = INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pos + i * stride;
        double s = view_proj[2] * (double)p[0] + view_proj[6] * (double)p[1] + view_proj[10] * (double)p[2];
        bucket[i] = js_to_int32(s * 4096);
        if (bucket[i] > max_depth) max_depth = bucket[i];
        if (bucket[i] < min_depth) min_depth = bucket[i];
    }
    double depth_inv = (256 * 256) / (max_depth - min_depth);
    for (int64_t i = 0; i < n; ++i) bucket[i] = js_to_int32(((double)bucket[i] - min_depth) * depth_inv);
}

/* V8 MathHypot for the two arguments of gs:384. */
static double js_hypot2(double a, double b)
{
    double abs_values[2] = { fabs(a), fabs(b) };
    int one_arg_is_nan = 0;
    double max = 0;
    for (int i = 0; i < 2; ++i) {
        if (isnan(abs_values[i])) one_arg_is_nan = 1;
        else if (abs_values[i] > max) max = abs_values[i];
    }
    if (max == INFINITY) return INFINITY;
    if (one_arg_is_nan) return NAN;
    if (max == 0) return 0;
    double sum = 0, compensation = 0;
    for (int i = 0; i < 2; ++i) {
        double v = abs_values[i] / max;
        double summand = (v * v) - compensation;
        double preliminary = sum + summand;
        compensation = (preliminary - sum) - summand;
        sum = preliminary;
    }
    return sqrt(sum) * max;
}

/* multiply4, gs:110-123 (column-major 4x4, result[4*r + c] = sum_k b[4*r + k] * a[c + 4*k]). */
void orc_multiply4(const double *a, const double *b, double *out)
{
    for (int row = 0; row < 16; row += 4)
        for (int col = 0; col < 4; ++col)
            out[row + col] = b[row] * a[col] + b[row + 1] * a[col + 4] + b[row + 2] * a[col + 8] + b[row + 3] * a[col + 12];
}

/*
 * performHitTesting, gs:361-395, with combinedMatrix = multiply4(projectionMatrix, viewMatrix)
 * already formed (gs:364).  Returns the selected label, *index = the selected Gaussian or -1.
 */
int32_t orc_viewer_hit_test(const float *pos, const int32_t *labels, int64_t n, int stride, const double *matrix,
                            double x, double y, double viewport_w, double viewport_h, int32_t no_selection,
                            int64_t *index)
{
    double closest_dist = INFINITY, closest_depth = INFINITY;                /* gs:365-366 */
    int32_t selected = no_selection;
    int64_t sel_idx = -1;
    for (int64_t i = 0; i < n; ++i) {
        const float *p = pos + i * stride;
        double v[4] = { (double)p[0], (double)p[1], (double)p[2], 1.0 };    /* gs:372 */
        double r[4];
        for (int k = 0; k < 4; ++k)                                          /* gs:400-402 */
            r[k] = v[0] * matrix[k] + v[1] * matrix[k + 4] + v[2] * matrix[k + 8] + v[3] * matrix[k + 12];
        if (r[3] <= 0) continue;                                             /* gs:403, :376 */
        double screen_x = (r[0] / r[3] + 1) * 0.5 * viewport_w;              /* gs:379 */
        double screen_y = (r[1] / r[3] + 1) * 0.5 * viewport_h;              /* gs:380 */
        double depth = r[2] / r[3];                                          /* gs:381 */
        double dist = js_hypot2(screen_x - x, screen_y - y);                 /* gs:384 */
        if (dist < 10 && (dist < closest_dist || (dist == closest_dist && depth < closest_depth))) {   /* gs:387 */
            closest_dist = dist;
            closest_depth = depth;
            selected = labels[i];
            sel_idx = i;
        }
    }
    if (index) *index = sel_idx;
    return selected;
}
