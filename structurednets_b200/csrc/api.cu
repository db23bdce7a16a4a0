// Library-wide entry points: version, thread-local error string, launch counter.
#include <atomic>
#include <stdarg.h>
#include "common.cuh"

namespace snb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace snb

extern "C" {
int sn_version(void) { return SN_VERSION; }
const char* sn_last_error_string(void) { return snb::g_err; }
long long sn_launch_count(void) { return snb::g_launches.load(); }
void sn_reset_launch_count(void) { snb::g_launches.store(0); }
}
