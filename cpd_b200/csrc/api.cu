// Library-wide state of libcpd_b200.so: version, last error text, launch counter.
#include <atomic>
#include <stdarg.h>
#include "common.cuh"

namespace cpd {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace cpd

extern "C" int32_t cpd_version(void) { return 100; }
extern "C" const char *cpd_last_error_string(void) { return cpd::g_err; }
extern "C" int64_t cpd_launch_count(void) { return cpd::g_launches.load(std::memory_order_relaxed); }
