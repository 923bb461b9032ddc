// Thread-local last-error string of the library (C ABI: egm_last_error()).
#include <stdarg.h>
#include <stdio.h>

#include "egm_gemm.h"

namespace egm {
namespace {
thread_local char g_err[512] = "";
}
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace egm
