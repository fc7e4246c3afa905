// Library identity and the thread-local error string of the C-ABI.
#include <stdarg.h>

#include "common.cuh"

namespace desire {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace desire

extern "C" int desire_version(void) { return DESIRE_ABI_VERSION; }
extern "C" const char* desire_last_error(void) { return desire::g_err; }
