// Library-level entry points: version, build info, thread-local error text.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void crn_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t crn_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

void crn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int crn_version(void) { return 100; }
extern "C" const char* crn_build_arch(void) { return "sm_100a"; }
extern "C" const char* crn_last_error(void) { return g_err; }

static std::atomic<int> g_flags{0};
int crn_get_flags() { return g_flags.load(std::memory_order_relaxed); }
extern "C" void crn_set_flags(int flags) { g_flags.store(flags, std::memory_order_relaxed); }
