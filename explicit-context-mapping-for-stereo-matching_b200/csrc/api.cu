// Library-level entry points of libcmfb200.so: ABI version, thread-local error string, launch counter.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cmfb200 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

}  // namespace cmfb200

extern "C" int cmfb200_abi_version(void) { return CMFB200_ABI_VERSION; }
extern "C" const char* cmfb200_last_error(void) { return cmfb200::g_err; }
extern "C" unsigned long long cmfb200_launch_count(void) {
    return cmfb200::g_launches.load(std::memory_order_relaxed);
}
