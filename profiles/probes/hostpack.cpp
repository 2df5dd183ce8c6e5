#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
static void pack(const int32_t *in, uint8_t *out, int64_t n, int lmin, int ncls, int *bad, int *mn, int *mx)
{
    int b = 0, lo = INT32_MAX, hi = INT32_MIN;
    for (int64_t i = 0; i < n; ++i) {
        const int v = in[i];
        const unsigned c = (unsigned)(v - lmin);
        b |= c >= (unsigned)ncls;
        lo = v < lo ? v : lo; hi = v > hi ? v : hi;
        out[i] = (uint8_t)(c < (unsigned)ncls ? c + 1 : 0);
    }
    *bad = b; *mn = lo; *mx = hi;
}
int main(int argc, char **argv)
{
    const int64_t n = 300LL * 1080 * 1920;
    int32_t *in = (int32_t *)malloc(n * 4);
    uint8_t *out = (uint8_t *)malloc(n);
    for (int64_t i = 0; i < n; ++i) in[i] = (int)(i % 151) - 1;
    for (int64_t i = 0; i < n; i += 4096) out[i] = 0;
    printf("hw threads %u\n", std::thread::hardware_concurrency());
    for (int nt : {4, 8, 16, 32, 64}) {
        for (int rep = 0; rep < 2; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            std::vector<int> bad(nt), mn(nt), mx(nt);
            for (int t = 0; t < nt; ++t) {
                const int64_t a = n * t / nt, b = n * (t + 1) / nt;
                pool.emplace_back(pack, in + a, out + a, b - a, -1, 151, &bad[t], &mn[t], &mx[t]);
            }
            for (auto &th : pool) th.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rep) printf("threads %2d: %.1f ms, %.1f GB/s read\n", nt, dt * 1e3, n * 4 / dt / 1e9);
        }
    }
    return 0;
}
