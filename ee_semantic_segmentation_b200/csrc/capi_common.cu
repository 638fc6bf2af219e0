// Error string, ABI version and launch counter of libeeseg_b200.
#include <stdarg.h>

#include "common.cuh"

namespace eeseg {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace eeseg

extern "C" int eeseg_abi_version(void) { return EESEG_ABI_VERSION; }
extern "C" const char* eeseg_last_error(void) { return eeseg::g_err; }
extern "C" int64_t eeseg_launch_count(void) { return eeseg::g_launches.load(); }
