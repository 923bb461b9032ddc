// Helpers shared by the operator chains (egm_api.cu, egm_lowrank.cu): bump allocator over the
// caller's workspace, GEMM dispatch by precision mode, argument / CUDA error macros.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/egm_b200.h"
#include "egm_gemm.h"
#include "egm_kernels.cuh"

namespace egm {
namespace chain {

struct Arena {
  uint8_t* p;
  size_t cap, used;
  Arena(void* base, size_t bytes) : p(static_cast<uint8_t*>(base)), cap(bytes), used(0) {}
  void* take(size_t bytes) {
    const size_t a = (used + 255) & ~size_t(255);
    used = a + bytes;
    return (used <= cap && p) ? p + a : nullptr;
  }
};
inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

inline bool prec_ok(int prec) { return prec == PREC_FP32_SIMT || prec == PREC_BF16X3 || prec == PREC_BF16; }

inline cudaError_t run_gemm(const GemmProblem& g, int prec, cudaStream_t st) {
  if (prec == PREC_FP32_SIMT) return gemm_simt(g, st);
  return gemm_tc(g, prec == PREC_BF16X3 ? 3 : 1, st);
}
// route a working-matrix output / addend to the right slot of the problem
inline void out_w(GemmProblem& g, const W& w, int prec) {
  if (prec == PREC_FP32_SIMT) g.Cf = w_mat(w, prec); else g.Cp = w_mat(w, prec);
}
inline void addend_w(GemmProblem& g, const W& w, float gamma, int prec) {
  g.E = w_mat(w, prec);
  g.e_planes = (prec != PREC_FP32_SIMT);
  g.gamma = gamma;
}
inline GemmTerm term(const W& A, int tA, const W& B, int tB, int K, int prec) {
  GemmTerm t;
  t.A = w_mat(A, prec); t.transA = tA; t.B = w_mat(B, prec); t.transB = tB; t.K = K;
  return t;
}

#define EGM_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      set_error(__VA_ARGS__);          \
      return code;                     \
    }                                  \
  } while (0)
#define EGM_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      char prev_[256];                /* set_error formats INTO the buffer last_error() returns */ \
      snprintf(prev_, sizeof(prev_), "%s", last_error());                           \
      set_error("%s: %s [%s]", #expr, cudaGetErrorString(e_), prev_);               \
      return EGM_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)
#define EGM_LAUNCHED() EGM_CUDA(cudaGetLastError())


// pooling prologue / backward tail shared by the dense and the low-rank moment paths
struct PoolVecs {
  float *s, *deg, *w, *wdiag, *t, *sw;
  PoolVecs(float* vecs, int B, int N)
      : s(vecs), deg(vecs + (size_t)B * N), w(vecs + (size_t)2 * B * N), wdiag(vecs + (size_t)3 * B * N),
        t(vecs + (size_t)4 * B * N), sw(vecs + (size_t)4 * B * N + B) {}
};

}  // namespace chain
}  // namespace egm
