// Library-level state of libvps_b200.so: last-error string, launch counter, version.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace vps {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace vps

extern "C" const char* vps_last_error(void) { return vps::g_err; }
extern "C" int vps_version(void) { return 100; }
extern "C" int64_t vps_launch_count(void) { return vps::g_launches.load(std::memory_order_relaxed); }
// kernels replayed through a captured CUDA graph do not pass through the launch wrappers: the host adds them here
extern "C" void vps_add_launch_count(int64_t n) { vps::count_launch((int)n); }
