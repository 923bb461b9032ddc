// Thread-local last-error string (C ABI: egm_last_error()) and the process-wide count of
// kernels this library launched (C ABI: egm_launch_count(), the benchmark's gpu_launches).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "egm_gemm.h"

namespace egm {
namespace {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
}
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace egm
