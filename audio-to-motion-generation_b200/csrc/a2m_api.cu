// Library-level state: version, thread-local error string, launch counter, device properties.
#include <atomic>
#include <cstring>
#include "a2m_common.cuh"

namespace {
thread_local char g_error[1024] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void a2m_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void a2m_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int a2m_num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" int a2m_version(void) { return A2M_VERSION; }
extern "C" const char* a2m_last_error(void) { return g_error; }
extern "C" int64_t a2m_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void a2m_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }
