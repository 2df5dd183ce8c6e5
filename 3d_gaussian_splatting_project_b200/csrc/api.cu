// Library-wide pieces of the C ABI: version, error string, device facts.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace gsl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count()
{
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace gsl

extern "C" int gsl_version(void) { return GSL_ABI_VERSION; }

extern "C" const char *gsl_last_error(void) { return gsl::g_err; }

extern "C" unsigned long long gsl_launch_count(void) { return gsl::g_launches.load(std::memory_order_relaxed); }

extern "C" int gsl_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return gsl::fail(GSL_ECUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }
    return n;
}
