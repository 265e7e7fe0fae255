// core.cu -- error string, version and launch counter of libseb200.
#include <stdarg.h>
#include <atomic>
#include "common.cuh"

namespace seb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace seb

extern "C" int seb200_version(void) { return SEB200_ABI_VERSION; }
extern "C" const char* seb200_last_error_string(void) { return seb::g_err; }
extern "C" long long seb200_launch_count(void) { return seb::g_launches.load(std::memory_order_relaxed); }
