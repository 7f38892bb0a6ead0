// Error channel, ABI version and launch counter of the C ABI (include/airpose_b200.h).
#include <atomic>

#include "common.cuh"

namespace airpose {

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace airpose

extern "C" const char* airpose_last_error(void) { return airpose::g_err; }
extern "C" int airpose_abi_version(void) { return AIRPOSE_B200_ABI_VERSION; }
extern "C" int64_t airpose_launch_count(void) { return airpose::g_launches.load(); }
