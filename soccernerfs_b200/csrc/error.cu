// Error reporting for the C-ABI: every entry point returns non-zero on failure and leaves a message here.
#include <stdarg.h>

#include "common.cuh"

namespace kp {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace kp

extern "C" const char* kp_last_error(void) { return kp::g_err; }
extern "C" int kp_abi_version(void) { return KP_ABI_VERSION; }
extern "C" long long kp_launch_count(void) { return kp::g_launches.load(); }
