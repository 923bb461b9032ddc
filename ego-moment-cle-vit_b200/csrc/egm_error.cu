// Thread-local last-error string (C ABI: egm_last_error()) and the process-wide count of
// kernels this library launched (C ABI: egm_launch_count(), the benchmark's gpu_launches).
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

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

// ---- optional per-launch timing of the GEMM engine (C ABI: egm_prof_*) -------------------------
// When enabled, every tcgen05 GEMM launch is bracketed by two CUDA events on its own stream; the
// benchmark reads the durations back after synchronising. Off by default: no events, no cost.
namespace {
struct ProfRec { cudaEvent_t e0, e1; double flops; int dims[6]; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
std::atomic<int> g_prof_on{0};
}
// level 1: every launch is bracketed; level 2: only the groups the chains open (a Newton-Schulz chain =
// one record), so the launches inside keep their programmatic-dependent-launch overlap
int prof_level() { return g_prof_on.load(std::memory_order_relaxed); }
bool prof_enabled() { return prof_level() != 0; }
void prof_enable(int level) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on.store(level < 0 ? 0 : (level > 2 ? 2 : level));
}
namespace {
struct ProfGroup { int id = -1; double flops = 0; int launches = 0; int dims[6] = {0, 0, 0, 0, 0, 0}; };
thread_local ProfGroup t_group;
}
bool prof_group_open() { return t_group.id >= 0; }
void prof_group_begin(cudaStream_t st) {
  if (prof_level() != 2 || t_group.id >= 0) return;
  const int zero[6] = {0, 0, 0, 0, 0, 0};
  t_group = ProfGroup();
  t_group.id = prof_begin(st, 0.0, zero);
}
void prof_group_note(double flops, const int dims[6]) {
  if (t_group.id < 0) return;
  t_group.flops += flops;
  if (t_group.launches++ == 0)
    for (int i = 0; i < 6; ++i) t_group.dims[i] = dims[i];
}
void prof_group_end(cudaStream_t st) {
  if (t_group.id < 0) return;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (t_group.id < (int)g_prof.size()) {
      ProfRec& r = g_prof[t_group.id];
      r.flops = t_group.flops;
      for (int i = 0; i < 6; ++i) r.dims[i] = t_group.dims[i];
      r.dims[3] = -t_group.launches;   // a negative "K of term 1" marks a group of that many launches
      cudaEventRecord(r.e1, st);
    }
  }
  t_group = ProfGroup();
}
void prof_reset() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
}
int prof_begin(cudaStream_t st, double flops, const int dims[6]) {
  ProfRec r;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
  r.flops = flops;
  for (int i = 0; i < 6; ++i) r.dims[i] = dims[i];
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void prof_end(int id, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (id >= 0 && id < (int)g_prof.size()) cudaEventRecord(g_prof[id].e1, st);
}
int prof_count() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  return (int)g_prof.size();
}
int prof_read(int i, float* ms, double* flops, int* dims) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (i < 0 || i >= (int)g_prof.size()) return -1;
  if (cudaEventSynchronize(g_prof[i].e1) != cudaSuccess) return -2;
  if (cudaEventElapsedTime(ms, g_prof[i].e0, g_prof[i].e1) != cudaSuccess) return -2;
  *flops = g_prof[i].flops;
  for (int k = 0; k < 6; ++k) dims[k] = g_prof[i].dims[k];
  return 0;
}
}  // namespace egm
