// C-ABI plumbing: thread-local error message, ABI version.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace phifem {
namespace {
thread_local char g_error[512] = "";
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace phifem

extern "C" const char* phifem_last_error(void) { return phifem::g_error; }
extern "C" int phifem_abi_version(void) { return PHIFEM_B200_ABI_VERSION; }
