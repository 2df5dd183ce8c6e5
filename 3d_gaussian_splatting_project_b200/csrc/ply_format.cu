// Host-side ASCII formatter for the labelled PLY that 3D_clustering/k_means.py:169-194 writes
// through plyfile with text=True: every field of a vertex goes through '%.18g' of its float64
// value, fields separated by one space, one vertex per line.  At 6 M vertices x 63 columns the
// Python formatter needs minutes; this one formats row ranges on all host threads.
// (SURVEY.md 8f, row N1.  Pure host code: no device work here.)
#include <stdio.h>
#include <string.h>

#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

enum FieldType { F_I8 = 0, F_U8, F_I16, F_U16, F_I32, F_U32, F_F32, F_F64 };

inline double load_field(const unsigned char *p, int type)
{
    switch (type) {
        case F_I8:  { int8_t v;   memcpy(&v, p, 1); return v; }
        case F_U8:  { uint8_t v;  memcpy(&v, p, 1); return v; }
        case F_I16: { int16_t v;  memcpy(&v, p, 2); return v; }
        case F_U16: { uint16_t v; memcpy(&v, p, 2); return v; }
        case F_I32: { int32_t v;  memcpy(&v, p, 4); return v; }
        case F_U32: { uint32_t v; memcpy(&v, p, 4); return v; }
        case F_F32: { float v;    memcpy(&v, p, 4); return v; }
        default:    { double v;   memcpy(&v, p, 8); return v; }
    }
}

void format_rows(const unsigned char *rec, int64_t r0, int64_t r1, int record_size, int n_fields,
                 const int *types, const int *offsets, std::string &out)
{
    char tmp[64];
    out.reserve((size_t)(r1 - r0) * n_fields * 12);
    for (int64_t r = r0; r < r1; ++r) {
        const unsigned char *row = rec + r * record_size;
        for (int f = 0; f < n_fields; ++f) {
            const double v = load_field(row + offsets[f], types[f]);
            int n;
            if (std::isnan(v)) { memcpy(tmp, "nan", 3); n = 3; }                  // Python prints nan, never -nan
            else n = snprintf(tmp, sizeof(tmp), "%.18g", v);
            out.append(tmp, (size_t)n);
            out.push_back(f + 1 == n_fields ? '\n' : ' ');
        }
    }
}

}  // namespace

// Formats rows [0, n_rows) of packed little-endian records into `out` (capacity out_cap bytes).
// types[f] in {0:i1 1:u1 2:i2 3:u2 4:i4 5:u4 6:f4 7:f8}, offsets[f] = byte offset in the record.
// Returns the number of bytes written, or a negative GSL_E* code (GSL_EWORKSPACE: out too small;
// 40 bytes per field always suffice).
extern "C" int64_t gsl_ply_format_ascii(const void *records, int64_t n_rows, int record_size, int n_fields,
                                        const int *types, const int *offsets, char *out, int64_t out_cap,
                                        int n_threads)
{
    if (!records || !types || !offsets || !out || n_rows < 0 || n_fields < 1 || record_size < 1)
        return gsl::fail(GSL_EINVAL, "gsl_ply_format_ascii: bad argument");
    for (int f = 0; f < n_fields; ++f)
        if (types[f] < 0 || types[f] > F_F64 || offsets[f] < 0 || offsets[f] >= record_size)
            return gsl::fail(GSL_EINVAL, "gsl_ply_format_ascii: bad field %d", f);
    if (n_threads < 1) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_rows / 1024 + 1) n_threads = (int)(n_rows / 1024 + 1);
    std::vector<std::string> parts((size_t)n_threads);
    std::vector<std::thread> pool;
    const unsigned char *rec = static_cast<const unsigned char *>(records);
    for (int t = 0; t < n_threads; ++t) {
        const int64_t r0 = n_rows * t / n_threads, r1 = n_rows * (t + 1) / n_threads;
        pool.emplace_back(format_rows, rec, r0, r1, record_size, n_fields, types, offsets, std::ref(parts[(size_t)t]));
    }
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (auto &p : parts) total += (int64_t)p.size();
    if (total > out_cap) return gsl::fail(GSL_EWORKSPACE, "gsl_ply_format_ascii: output %lld > capacity %lld", (long long)total, (long long)out_cap);
    int64_t at = 0;
    for (auto &p : parts) { memcpy(out + at, p.data(), p.size()); at += (int64_t)p.size(); }
    return total;
}
