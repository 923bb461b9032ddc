// C ABI of the library (include/egm_b200.h): the chains of GEMM-engine launches and
// bandwidth kernels that make up each operator of the moment-pooling path.
//
// Arithmetic follows SURVEY.md Appendix A. Newton-Schulz uses the commuting form
//   P_k = Z_k Y_k,  T_k = 1.5 I - 0.5 P_k,  Y_{k+1} = Y_k T_k,  Z_{k+1} = T_k Z_k
// (Y_k, Z_k are polynomials in A, so Y_k Z_k = Z_k Y_k for ANY input A): 3K-3 GEMMs forward
// and 6K-6 backward for K >= 2 instead of the reference's 4K / 8K (moment_head.py:53-64).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "egm_chain.h"

using namespace egm;

using namespace egm::chain;

namespace {

// -------------------------------------------------------------- NS state layout
// K iterations store: A, T_0..T_{K-1}, Z_1..Z_{K-1}, Y_2..Y_{K-1}  (Y_1 == T_0)
struct NsState {
  int B, D, K;
  uint8_t* base;
  size_t each;
  NsState(void* s, int B_, int D_, int K_) : B(B_), D(D_), K(K_), base(static_cast<uint8_t*>(s)) {
    each = pad256(w_bytes(B, D, D));
  }
  static int count(int K) { return K <= 0 ? 1 : 1 + K + (K > 1 ? K - 1 : 0) + (K > 2 ? K - 2 : 0); }
  W slot(int i) const { return make_w(base + (size_t)i * each, B, D, D); }
  W A() const { return slot(0); }
  W T(int k) const { return slot(1 + k); }
  W Z(int k) const { return slot(1 + K + (k - 1)); }                  // k in [1, K-1]
  W Y(int k) const { return k == 1 ? T(0) : slot(1 + K + (K - 1) + (k - 2)); }  // k in [1, K-1]
};


// Products of the forward chain given A and T_0 in the state (K >= 2):
//   Z_1 = T_0 A; k = 1..K-1: T_k = 1.5 I - 0.5 Z_k Y_k, Y_{k+1} = Y_k T_k, Z_{k+1} = T_k Z_k;
// the last product leaves as  post * Y_{K-1} T_{K-1}  either in fp32 (O) or as the packed upper
// triangle in bf16 planes (X: the operand of second_net's Linear; lower tiles are not computed).
// `sym`: A is symmetric, so every matrix of the chain is; all of them are kept in the engine's
// symmetric block storage (GemmTerm::symA) and only the upper tiles of each product are evaluated.
int ns_products_fwd_impl(const NsState& S, int prec, const float* post, float* O, const W* X,
                         cudaStream_t st, bool sym) {
  const int B = S.B, D = S.D, K = S.K;
  const long long dd = (long long)D * D;
  auto prod = [&](const W& a, const W& b) {
    GemmTerm t = term(a, 0, b, 0, D, prec);
    t.symA = t.symB = sym ? 1 : 0;
    return t;
  };
  {
    GemmProblem g;  // Z_1 = T_0 A
    g.M = D; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = prod(S.T(0), S.A());
    g.sym_out = sym;
    out_w(g, S.Z(1), prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  for (int k = 1; k <= K - 1; ++k) {
    {
      GemmProblem g;  // T_k = 1.5 I - 0.5 Z_k Y_k
      g.M = D; g.N = D; g.batch = B; g.nterms = 1;
      g.t[0] = prod(S.Z(k), S.Y(k));
      g.alpha = -0.5f; g.beta_eye = 1.5f;
      g.sym_out = sym;
      out_w(g, S.T(k), prec);
      EGM_CUDA(run_gemm(g, prec, st));
    }
    if (k < K - 1) {
      {
        GemmProblem g;  // Y_{k+1} = Y_k T_k
        g.M = D; g.N = D; g.batch = B; g.nterms = 1;
        g.t[0] = prod(S.Y(k), S.T(k));
        g.sym_out = sym;
        out_w(g, S.Y(k + 1), prec);
        EGM_CUDA(run_gemm(g, prec, st));
      }
      {
        GemmProblem g;  // Z_{k+1} = T_k Z_k
        g.M = D; g.N = D; g.batch = B; g.nterms = 1;
        g.t[0] = prod(S.T(k), S.Z(k));
        g.sym_out = sym;
        out_w(g, S.Z(k + 1), prec);
        EGM_CUDA(run_gemm(g, prec, st));
      }
    } else {
      GemmProblem g;  // O = post * Y_{K-1} T_{K-1}
      g.M = D; g.N = D; g.batch = B; g.nterms = 1;
      g.t[0] = prod(S.Y(k), S.T(k));
      g.alpha_b = post;
      if (X) {
        g.X = w_mat(*X, prec);
      } else {
        g.Cf = f32_mat(O, D, D, D, dd);
      }
      EGM_CUDA(run_gemm(g, prec, st));
    }
  }
  return EGM_OK;
}

// Where the last backward product (dA) goes: fp32, or planes together with <dA, A>.
struct NsFinal {
  float* dA_f32 = nullptr;
  const W* dA_w = nullptr;
  float* dot_out = nullptr;   // <dA, A> per image (tensor-core modes, with dA_w)
  float* dot_ws = nullptr;
};

// Backward products (K >= 2). buf[0] holds dY_K = post * dO on entry; buf[1..5] are scratch.
int ns_products_bwd_impl(const NsState& S, int prec, W* buf, const NsFinal& fin, cudaStream_t st) {
  const int B = S.B, D = S.D, K = S.K;
  const long long dd = (long long)D * D;
  W dY = buf[0], dYn = buf[1], dZ = buf[2], dZn = buf[3], dP = buf[4], X = buf[5];
  bool have_dZ = false;
  for (int k = K - 1; k >= 1; --k) {
    {
      GemmProblem g;  // dP = -0.5 (Y_k^T dY [+ dZ Z_k^T])
      g.M = D; g.N = D; g.batch = B;
      g.t[0] = term(S.Y(k), 1, dY, 0, D, prec);
      g.nterms = 1;
      if (have_dZ) { g.t[1] = term(dZ, 0, S.Z(k), 1, D, prec); g.nterms = 2; }
      g.alpha = -0.5f;
      out_w(g, dP, prec);
      EGM_CUDA(run_gemm(g, prec, st));
    }
    {
      GemmProblem g;  // dY_k = dY T_k^T + Z_k^T dP
      g.M = D; g.N = D; g.batch = B; g.nterms = 2;
      g.t[0] = term(dY, 0, S.T(k), 1, D, prec);
      g.t[1] = term(S.Z(k), 1, dP, 0, D, prec);
      out_w(g, dYn, prec);
      EGM_CUDA(run_gemm(g, prec, st));
    }
    {
      GemmProblem g;  // dZ_k = [T_k^T dZ +] dP Y_k^T
      g.M = D; g.N = D; g.batch = B;
      g.t[0] = term(dP, 0, S.Y(k), 1, D, prec);
      g.nterms = 1;
      if (have_dZ) { g.t[1] = term(S.T(k), 1, dZ, 0, D, prec); g.nterms = 2; }
      out_w(g, dZn, prec);
      EGM_CUDA(run_gemm(g, prec, st));
    }
    W tmp = dY; dY = dYn; dYn = tmp;
    tmp = dZ; dZ = dZn; dZn = tmp;
    have_dZ = true;
  }
  // k = 0: Y_1 = T_0, Z_1 = T_0 A, T_0 = 1.5 I - 0.5 A
  {
    GemmProblem g;  // X = dT_0 = dY_1 + dZ_1 A^T
    g.M = D; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(dZ, 0, S.A(), 1, D, prec);
    addend_w(g, dY, 1.f, prec);
    out_w(g, X, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  {
    GemmProblem g;  // dA = T_0^T dZ_1 - 0.5 dT_0
    g.M = D; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(S.T(0), 1, dZ, 0, D, prec);
    addend_w(g, X, -0.5f, prec);
    if (fin.dA_w) {
      out_w(g, *fin.dA_w, prec);
      if (fin.dot_out) {
        g.F = w_mat(S.A(), prec);
        g.f_planes = 1;
        g.dot_out = fin.dot_out;
        g.dot_ws = fin.dot_ws;
      }
    } else {
      g.Cf = f32_mat(fin.dA_f32, D, D, D, dd);
    }
    EGM_CUDA(run_gemm(g, prec, st));
  }
  return EGM_OK;
}


// Backward of the chain for a SYMMETRIC A (K >= 2), evaluated as a forward tangent.
// Y_K is a fixed polynomial in A, so its Frechet derivative L(A, .) is self-adjoint for symmetric A:
// the gradient L^*(A, dY_K) equals the tangent L(A, dY_K). It also commutes with transposition, so
// the symmetric part of dA (all the rest of the path needs when the graph is symmetric) is the
// tangent in the direction sym(dY_K) - and along a symmetric direction every tangent Y'_k, Z'_k,
// T'_k is itself symmetric: each of the 3K-3 two-term products below evaluates its upper tiles only.
//   Y'_1 = T'_0 = -E/2,  Z'_1 = T'_0 A + T_0 E = -3 Y'_1 + (Y'_1 A + A Y'_1)
//   T'_k = -(Z'_k Y_k + Z_k Y'_k)/2,  Y'_{k+1} = Y'_k T_k + Y_k T'_k,  Z'_{k+1} = T'_k Z_k + T_k Z'_k
// buf[0] holds Y'_1 = -post*sym(dO)/2 on entry (symmetric block storage); buf[1..4] are scratch.
// The result dA = Y'_K goes to fin.dA_w with <dA, A> in fin.dot_out.
int ns_tangent_bwd_impl(const NsState& S, int prec, W* buf, const NsFinal& fin, cudaStream_t st) {
  const int B = S.B, D = S.D, K = S.K;
  W Yd = buf[0], Zd = buf[1], Td = buf[2], Ydn = buf[3], Zdn = buf[4];
  auto two = [&](GemmProblem& g, const W& a0, const W& b0, const W& a1, const W& b1) {
    g.M = D; g.N = D; g.batch = B; g.nterms = 2;
    g.t[0] = term(a0, 0, b0, 0, D, prec);
    g.t[1] = term(a1, 0, b1, 0, D, prec);
    g.t[0].symA = g.t[0].symB = g.t[1].symA = g.t[1].symB = 1;
    g.sym_out = 1;
  };
  {
    GemmProblem g;  // Z'_1 = -3 Y'_1 + Y'_1 A + A Y'_1
    two(g, Yd, S.A(), S.A(), Yd);
    addend_w(g, Yd, -3.f, prec);
    out_w(g, Zd, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  for (int k = 1; k <= K - 1; ++k) {
    {
      GemmProblem g;  // T'_k = -(Z'_k Y_k + Z_k Y'_k)/2
      two(g, Zd, S.Y(k), S.Z(k), Yd);
      g.alpha = -0.5f;
      out_w(g, Td, prec);
      EGM_CUDA(run_gemm(g, prec, st));
    }
    if (k < K - 1) {
      {
        GemmProblem g;  // Y'_{k+1} = Y'_k T_k + Y_k T'_k
        two(g, Yd, S.T(k), S.Y(k), Td);
        out_w(g, Ydn, prec);
        EGM_CUDA(run_gemm(g, prec, st));
      }
      {
        GemmProblem g;  // Z'_{k+1} = T'_k Z_k + T_k Z'_k
        two(g, Td, S.Z(k), S.T(k), Zd);
        out_w(g, Zdn, prec);
        EGM_CUDA(run_gemm(g, prec, st));
      }
      W tmp = Yd; Yd = Ydn; Ydn = tmp;
      tmp = Zd; Zd = Zdn; Zdn = tmp;
    } else {
      GemmProblem g;  // dA = Y'_K = Y'_{K-1} T_{K-1} + Y_{K-1} T'_{K-1}
      two(g, Yd, S.T(k), S.Y(k), Td);
      out_w(g, *fin.dA_w, prec);
      g.F = w_mat(S.A(), prec);
      g.f_planes = 1;
      g.dot_out = fin.dot_out;
      g.dot_ws = fin.dot_ws;
      EGM_CUDA(run_gemm(g, prec, st));
    }
  }
  return EGM_OK;
}


// The chains as the benchmark's roofline sees them: under egm_prof_enable(2) all launches of one chain
// share one pair of CUDA events (the launches in between keep their dependent-launch overlap).
struct ProfGroupScope {
  cudaStream_t st;
  explicit ProfGroupScope(cudaStream_t s) : st(s) { prof_group_begin(st); }
  ~ProfGroupScope() { prof_group_end(st); }
};
int ns_products_fwd(const NsState& S, int prec, const float* post, float* O, const W* X, cudaStream_t st,
                    bool sym = false) {
  ProfGroupScope scope(st);
  return ns_products_fwd_impl(S, prec, post, O, X, st, sym);
}
int ns_products_bwd(const NsState& S, int prec, W* buf, const NsFinal& fin, cudaStream_t st) {
  ProfGroupScope scope(st);
  return ns_products_bwd_impl(S, prec, buf, fin, st);
}
int ns_tangent_bwd(const NsState& S, int prec, W* buf, const NsFinal& fin, cudaStream_t st) {
  ProfGroupScope scope(st);
  return ns_tangent_bwd_impl(S, prec, buf, fin, st);
}

// Tail of the pooling backward given V1 = Zc dM^T and V2 = Zc dM:
//   dZc = Wn V1 + Wn^T V2, dW = V2 Zc^T, then centring / weighted mean / degree normalisation.
int pool_bwd_tail(const W& Wn, const W& Zc, const W& V1, const W& V2, const float* du, const float* Z,
                  const float* G, const float* u, const float* vecs, const float* mu, int B, int N, int D,
                  float eps, float* dZc, float* dW, long long ldW, float* dmu, float* dw, float* ds,
                  float* dt, float* dZ, float* dG, int prec, cudaStream_t st, bool sym = false) {
  const float* s = vecs;
  const float* deg = vecs + (size_t)B * N;
  const float* w = vecs + (size_t)2 * B * N;
  const float* t = vecs + (size_t)4 * B * N;
  const float* sw = t + B;
  if (sym) {
    GemmProblem g;  // symmetric graph and dM: V1 == V2 and Wn == Wn^T, dZc = 2 Wn V
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Wn, 0, V1, 0, N, prec);
    g.alpha = 2.f;
    g.Cf = f32_mat(dZc, N, D, D, (long long)N * D);
    EGM_CUDA(run_gemm(g, prec, st));
  } else {
    GemmProblem g;  // dZc = Wn V1 + Wn^T V2
    g.M = N; g.N = D; g.batch = B; g.nterms = 2;
    g.t[0] = term(Wn, 0, V1, 0, N, prec);
    g.t[1] = term(Wn, 1, V2, 0, N, prec);
    g.Cf = f32_mat(dZc, N, D, D, (long long)N * D);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  {
    GemmProblem g;  // dW = V2 Zc^T
    g.M = N; g.N = N; g.batch = B; g.nterms = 1;
    g.t[0] = term(V2, 0, Zc, 1, D, prec);
    g.Cf = f32_mat(dW, N, N, ldW, (long long)N * ldW);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  k::pool_bwd_dmu(dZc, du, sw, t, B, N, D, eps, dmu, st);
  k::pool_bwd_rows(dZc, Z, Zc, w, t, mu, u, du, dmu, B, N, D, eps, dZ, dw, dt, prec, st);
  k::pool_bwd_ds(dW, ldW, dw, dt, G, s, B, N, sym ? 1 : 0, ds, st);
  k::pool_bwd_dG(dW, ldW, dw, dt, s, deg, ds, B, N, eps, sym ? 1 : 0, dG, st);
  EGM_LAUNCHED();
  return EGM_OK;
}

}  // namespace

extern "C" {

int egm_version(void) { return 100; }
const char* egm_last_error(void) { return last_error(); }
unsigned long long egm_launch_count(void) { return launch_count(); }
void egm_prof_enable(int on) { prof_enable(on); }
void egm_prof_reset(void) { prof_reset(); }
int egm_prof_count(void) { return prof_count(); }
int egm_prof_read(int i, float* ms, double* flops, int* dims) { return prof_read(i, ms, flops, dims); }

// =========================================================================== GPF
long long egm_gpf_ldr(int N) { return ((long long)N + 3) / 4 * 4; }

size_t egm_gpf_fwd_workspace(int B, int N, int D, int prec) {
  (void)prec;
  return 2 * pad256(w_bytes(B, N, D)) + 512;
}
size_t egm_gpf_state_bytes(int B, int N, int D, int prec) {
  (void)prec;
  return 2 * pad256(w_bytes(B, N, D)) + 256;
}

int egm_gpf_fused_ok(int N, int D, int P, int Q, int prec) {
  static const float dummy[4] __attribute__((aligned(16))) = {0.f, 0.f, 0.f, 0.f};
  return prec != PREC_FP32_SIMT && k::gpf_fused_supported(N, D, P, Q, dummy, dummy) ? 1 : 0;
}

int egm_gpf_raw_planes_ok(int N, int D, int P, int Q, int prec) {
  // the fused forward can hand the backward the RAW token planes iff the backward can fold F.normalize
  // into E (degrees <= 3: the specialised polynomial kernel)
  const W probe = make_w(nullptr, 1, N, N);
  return egm_gpf_fused_ok(N, D, P, Q, prec) && k::gpf_poly_bwd_can_fold(P, Q, N, egm_gpf_ldr(N), probe) ? 1 : 0;
}

int egm_gpf_fwd(const float* a, const float* p, const float* coef, int B, int N, int D, int P, int Q,
                int cosine, float eps, int symmetric, float* G, float* Ra, float* Rp, float* nrm_a,
                float* nrm_p, void* xn_state, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_gpf_fwd: unknown precision mode %d", prec);
  EGM_REQUIRE(a && p && coef && G && nrm_a && nrm_p, EGM_ERR_ARG, "egm_gpf_fwd: null pointer");
  EGM_REQUIRE((Ra == nullptr) == (Rp == nullptr), EGM_ERR_ARG, "egm_gpf_fwd: Ra and Rp go together");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0 && P >= 0 && Q >= 0 && P <= 15 && Q <= 15, EGM_ERR_ARG,
              "egm_gpf_fwd: bad sizes B=%d N=%d D=%d P=%d Q=%d (degrees 0..15)", B, N, D, P, Q);
  // One fused pass (egm_gpf_fused.cu) whenever the tensor-core modes can address the tokens: nothing
  // but G, the norms and (on request) R_a / R_p is written; the normalised operand planes are not
  // produced (xn_state stays unused and the backward re-derives them).
  const bool raw_ok = cosine && egm_gpf_raw_planes_ok(N, D, P, Q, prec);
  if (prec != PREC_FP32_SIMT && (!xn_state || raw_ok) && k::gpf_fused_supported(N, D, P, Q, a, p)) {
    (void)symmetric;   // the fused result is symmetric by construction: F_ij is evaluated once per pair
    // xn_state (cosine, degrees <= 3): the RAW tokens as operand planes, written by the converter threads;
    // the caller passes `symmetric | 4` to egm_gpf_bwd, which then folds F.normalize's backward into E
    Arena xs(xn_state, xn_state ? egm_gpf_state_bytes(B, N, D, prec) : 0);
    const W Ar = make_w(xs.take(w_bytes(B, N, D)), B, N, D), Pr = make_w(xs.take(w_bytes(B, N, D)), B, N, D);
    EGM_CUDA(k::gpf_fused_fwd(a, p, coef, B, N, D, P, Q, cosine, eps, G, Ra, Rp, egm_gpf_ldr(N), nrm_a,
                              nrm_p, prec == PREC_BF16X3 ? 3 : 1, st, xn_state ? &Ar : nullptr,
                              xn_state ? &Pr : nullptr));
    return EGM_OK;
  }
  EGM_REQUIRE(Ra && Rp, EGM_ERR_ARG, "egm_gpf_fwd: the staged path needs Ra and Rp");
  // the normalised tokens (GEMM operand planes) go to `xn_state` when the caller keeps them for the
  // backward, else to scratch
  if (xn_state) { ws = xn_state; ws_bytes = egm_gpf_state_bytes(B, N, D, prec); }
  Arena ar(ws, ws_bytes);
  void* wa = ar.take(w_bytes(B, N, D));
  void* wp = ar.take(w_bytes(B, N, D));
  EGM_REQUIRE(wa && wp, EGM_ERR_WORKSPACE, "egm_gpf_fwd: workspace %zu < %zu", ws_bytes,
              egm_gpf_fwd_workspace(B, N, D, prec));
  const long long ldR = egm_gpf_ldr(N);
  const W An = make_w(wa, B, N, D), Pn = make_w(wp, B, N, D);
  k::rownorm(a, B, N, D, eps, cosine, nrm_a, An, prec, st);
  k::rownorm(p, B, N, D, eps, cosine, nrm_p, Pn, prec, st);
  EGM_LAUNCHED();
  for (int v = 0; v < 2; ++v) {
    GemmProblem g;
    g.M = N; g.N = N; g.batch = B; g.nterms = 1;
    g.t[0] = term(v ? Pn : An, 0, v ? Pn : An, 1, D, prec);
    g.Cf = f32_mat(v ? Rp : Ra, N, N, ldR, (long long)N * ldR);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  k::gpf_poly_fwd(Ra, Rp, ldR, coef, P, Q, symmetric, B, N, G, st);
  EGM_LAUNCHED();
  return EGM_OK;
}

size_t egm_gpf_bwd_workspace(int B, int N, int D, int P, int Q, int prec) {
  (void)prec;
  return pad256(w_bytes(B, N, D)) + 2 * pad256(w_bytes(B, N, N)) + pad256((size_t)B * N * D * 4) +
         pad256((size_t)k::gpf_poly_bwd_blocks(B, N) * (P + 1) * (Q + 1) * 4) +
         pad256(k::gpf_poly_bwd_rowpart_floats(B, N) * 4) + pad256((size_t)B * N * 4) + 1024;
}

int egm_gpf_bwd(const float* dG, const float* a, const float* p, const float* coef, const float* Ra,
                const float* Rp, const float* nrm_a, const float* nrm_p, const void* xn_state, int B,
                int N, int D, int P, int Q, int cosine, float eps, int symmetric, float* da, float* dp,
                float* dcoef, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_gpf_bwd: unknown precision mode %d", prec);
  EGM_REQUIRE(dG && a && p && coef && Ra && Rp && nrm_a && nrm_p && da && dp && dcoef, EGM_ERR_ARG,
              "egm_gpf_bwd: null pointer");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0 && P >= 0 && Q >= 0 && P <= 15 && Q <= 15, EGM_ERR_ARG,
              "egm_gpf_bwd: bad sizes");
  Arena ar(ws, ws_bytes);
  void* wx = ar.take(w_bytes(B, N, D));
  void* wea = ar.take(w_bytes(B, N, N));
  void* wep = ar.take(w_bytes(B, N, N));
  float* dxn = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  const int nblocks = k::gpf_poly_bwd_blocks(B, N);
  float* partial = static_cast<float*>(ar.take((size_t)nblocks * (P + 1) * (Q + 1) * 4));
  float* rowpart = static_cast<float*>(ar.take(k::gpf_poly_bwd_rowpart_floats(B, N) * 4));
  float* nrm_scratch = static_cast<float*>(ar.take((size_t)B * N * 4));
  EGM_REQUIRE(wx && wea && wep && dxn && partial && rowpart && nrm_scratch, EGM_ERR_WORKSPACE,
              "egm_gpf_bwd: workspace %zu < %zu", ws_bytes, egm_gpf_bwd_workspace(B, N, D, P, Q, prec));
  const long long ldR = egm_gpf_ldr(N);
  const W Ea = make_w(wea, B, N, N), Ep = make_w(wep, B, N, N), Xn = make_w(wx, B, N, D);
  // Without kept operand planes (always after the fused forward) and with cosine similarity, the
  // backward of F.normalize is folded into E (k::gpf_poly_bwd): the product with the RAW token planes
  // is the token gradient itself - no [B,N,D] rownorm_bwd pass, no fp32 round trip of d x-hat.
  static const bool fold_off = []() { const char* e = getenv("EGM_GPF_FOLD"); return e && e[0] == '0'; }();
  // bit 2 of `symmetric`: xn_state holds the RAW token planes (fused forward in training mode)
  const bool raw_planes = xn_state && (symmetric & 4);
  symmetric &= 3;
  EGM_REQUIRE(!raw_planes || (cosine && k::gpf_poly_bwd_can_fold(P, Q, N, ldR, Ea)), EGM_ERR_ARG,
              "egm_gpf_bwd: raw token planes need cosine similarity and degrees <= 3");
  const bool fold = raw_planes ||
                    (cosine && !xn_state && !fold_off && k::gpf_poly_bwd_can_fold(P, Q, N, ldR, Ea));
  k::gpf_poly_bwd(dG, Ra, Rp, ldR, coef, P, Q, symmetric, B, N, Ea, Ep, partial, nblocks, dcoef, prec, st, nrm_a,
                  nrm_p, eps, fold ? rowpart : nullptr);
  EGM_LAUNCHED();
  Arena xs(const_cast<void*>(xn_state), xn_state ? egm_gpf_state_bytes(B, N, D, prec) : 0);
  const W An = make_w(xs.take(w_bytes(B, N, D)), B, N, D), Pn = make_w(xs.take(w_bytes(B, N, D)), B, N, D);
  for (int v = 0; v < 2; ++v) {
    const float* x = v ? p : a;
    const float* nrm = v ? nrm_p : nrm_a;
    float* dx = v ? dp : da;
    if (!xn_state) {
      // re-materialise the token operand planes: raw tokens when the normalisation is folded into E,
      // else the normalised ones (memory-saving mode of the staged path)
      k::rownorm(x, B, N, D, eps, fold ? 0 : cosine, nrm_scratch, Xn, prec, st);
      EGM_LAUNCHED();
    }
    const W& Xv = xn_state ? (v ? Pn : An) : Xn;
    GemmProblem g;  // d An = (dR + dR^T) An, or dx = E'' X
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(v ? Ep : Ea, 0, Xv, 0, N, prec);
    g.Cf = f32_mat((cosine && !fold) ? dxn : dx, N, D, D, (long long)N * D);
    EGM_CUDA(run_gemm(g, prec, st));
    if (cosine && !fold) {
      k::rownorm_bwd(x, nrm, dxn, B, N, D, eps, dx, st);
      EGM_LAUNCHED();
    }
  }
  return EGM_OK;
}

// ========================================================================== pool
size_t egm_pool_state_bytes(int B, int N, int D, int prec) {
  (void)prec;
  return pad256(w_bytes(B, N, N)) + pad256(w_bytes(B, N, D)) + 256;
}
size_t egm_pool_fwd_workspace(int B, int N, int D, int prec) {
  (void)prec;
  return pad256(w_bytes(B, N, D)) + 512;
}

int egm_pool_fwd(const float* Z, const float* G, int B, int N, int D, float eps, float* M2, float* u,
                 float* vecs, float* mu, void* state, int prec, void* ws, size_t ws_bytes,
                 egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_pool_fwd: unknown precision mode %d", prec);
  EGM_REQUIRE(Z && G && M2 && vecs && mu && state, EGM_ERR_ARG, "egm_pool_fwd: null pointer");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0, EGM_ERR_ARG, "egm_pool_fwd: bad sizes");
  Arena sa(state, egm_pool_state_bytes(B, N, D, prec));
  const W Wn = make_w(sa.take(w_bytes(B, N, N)), B, N, N);
  const W Zc = make_w(sa.take(w_bytes(B, N, D)), B, N, D);
  Arena ar(ws, ws_bytes);
  void* wu = ar.take(w_bytes(B, N, D));
  EGM_REQUIRE(wu, EGM_ERR_WORKSPACE, "egm_pool_fwd: workspace %zu < %zu", ws_bytes,
              egm_pool_fwd_workspace(B, N, D, prec));
  const W U = make_w(wu, B, N, D);
  float* s = vecs;
  float* deg = vecs + (size_t)B * N;
  float* w = vecs + (size_t)2 * B * N;
  float* wdiag = vecs + (size_t)3 * B * N;
  float* t = vecs + (size_t)4 * B * N;
  float* sw = t + B;
  k::degree(G, B, N, eps, deg, s, st);
  k::weight(G, s, B, N, Wn, w, wdiag, prec, st);
  k::mean_center(Z, w, wdiag, B, N, D, eps, t, sw, mu, u, Zc, prec, st);
  EGM_LAUNCHED();
  {
    GemmProblem g;  // U = Wn Zc
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Wn, 0, Zc, 0, N, prec);
    out_w(g, U, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  {
    GemmProblem g;  // M2 = Zc^T U
    g.M = D; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 1, U, 0, N, prec);
    g.Cf = f32_mat(M2, D, D, D, (long long)D * D);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  return EGM_OK;
}

size_t egm_pool_bwd_workspace(int B, int N, int D, int prec) {
  (void)prec;
  return pad256(w_bytes(B, D, D)) + 2 * pad256(w_bytes(B, N, D)) + pad256((size_t)B * N * D * 4) +
         pad256((size_t)B * N * egm_gpf_ldr(N) * 4) + pad256((size_t)B * D * 4) +
         3 * pad256((size_t)B * N * 4) + 2048;
}

int egm_pool_bwd(const float* dM2, const float* du, const float* Z, const float* G, const float* u,
                 const float* vecs, const float* mu, const void* state, int B, int N, int D, float eps,
                 float* dZ, float* dG, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_pool_bwd: unknown precision mode %d", prec);
  EGM_REQUIRE(dM2 && Z && G && vecs && mu && state && dZ && dG, EGM_ERR_ARG, "egm_pool_bwd: null pointer");
  EGM_REQUIRE(!du || u, EGM_ERR_ARG, "egm_pool_bwd: du given without u");
  Arena sa(const_cast<void*>(state), egm_pool_state_bytes(B, N, D, prec));
  const W Wn = make_w(sa.take(w_bytes(B, N, N)), B, N, N);
  const W Zc = make_w(sa.take(w_bytes(B, N, D)), B, N, D);
  Arena ar(ws, ws_bytes);
  void* wdm = ar.take(w_bytes(B, D, D));
  void* wv1 = ar.take(w_bytes(B, N, D));
  void* wv2 = ar.take(w_bytes(B, N, D));
  float* dZc = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  const long long ldW = egm_gpf_ldr(N);
  float* dW = static_cast<float*>(ar.take((size_t)B * N * ldW * 4));
  float* dmu = static_cast<float*>(ar.take((size_t)B * D * 4));
  float* dw = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* ds = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* dt = static_cast<float*>(ar.take((size_t)B * N * 4));
  EGM_REQUIRE(wdm && wv1 && wv2 && dZc && dW && dmu && dw && ds && dt, EGM_ERR_WORKSPACE,
              "egm_pool_bwd: workspace %zu < %zu", ws_bytes, egm_pool_bwd_workspace(B, N, D, prec));
  const float* s = vecs;
  const float* deg = vecs + (size_t)B * N;
  const float* w = vecs + (size_t)2 * B * N;
  const float* t = vecs + (size_t)4 * B * N;
  const float* sw = t + B;
  const W dMw = make_w(wdm, B, D, D), V1 = make_w(wv1, B, N, D), V2 = make_w(wv2, B, N, D);
  k::affine(dM2, D, (long long)D * D, B, D, D, nullptr, 1.f, 0.f, dMw, 0.f, 0.f, nullptr, prec, st);
  EGM_LAUNCHED();
  {
    GemmProblem g;  // V1 = Zc dM^T
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 0, dMw, 1, D, prec);
    out_w(g, V1, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  {
    GemmProblem g;  // V2 = Zc dM
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 0, dMw, 0, D, prec);
    out_w(g, V2, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  return pool_bwd_tail(Wn, Zc, V1, V2, du, Z, G, u, vecs, mu, B, N, D, eps, dZc, dW, ldW, dmu, dw, ds, dt, dZ,
                       dG, prec, st);
}

// ============================================================================ NS
size_t egm_ns_state_bytes(int B, int D, int iters, int prec) {
  (void)prec;
  return (size_t)NsState::count(iters) * pad256(w_bytes(B, D, D)) + 256;
}
size_t egm_ns_fwd_workspace(int B, int D, int iters, int prec) {
  (void)iters; (void)prec;
  return 2 * pad256(w_bytes(B, D, D)) + 512;
}

int egm_ns_fwd(const float* M, int B, int D, int iters, float eps, int post_mode, float* O,
               float* scal, void* state, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_ns_fwd: unknown precision mode %d", prec);
  EGM_REQUIRE(M && O && scal && state, EGM_ERR_ARG, "egm_ns_fwd: null pointer");
  EGM_REQUIRE(B > 0 && D > 0 && iters >= 0 && iters <= 64, EGM_ERR_ARG, "egm_ns_fwd: bad sizes");
  EGM_REQUIRE(post_mode == 0 || post_mode == 1, EGM_ERR_ARG, "egm_ns_fwd: post_mode must be 0 or 1");
  const int K = iters;
  NsState S(state, B, D, K);
  Arena ar(ws, ws_bytes);
  void* wy = ar.take(w_bytes(B, D, D));
  EGM_REQUIRE(wy, EGM_ERR_WORKSPACE, "egm_ns_fwd: workspace %zu < %zu", ws_bytes,
              egm_ns_fwd_workspace(B, D, iters, prec));
  float* tr = scal;
  float* inv = scal + B;
  float* post = scal + 2 * B;
  const long long dd = (long long)D * D;
  k::trace_scales(M, B, D, eps, post_mode, tr, inv, post, st);
  // A = M/(tr+eps) and T_0 = 1.5 I - 0.5 A in one pass over M
  const W A = S.A();
  if (K >= 1) {
    const W T0 = S.T(0);
    k::affine(M, D, dd, B, D, D, inv, 1.f, 0.f, A, -0.5f, 1.5f, &T0, prec, st);
  } else {
    k::affine(M, D, dd, B, D, D, inv, 1.f, 0.f, A, 0.f, 0.f, nullptr, prec, st);
  }
  EGM_LAUNCHED();
  if (K <= 1) {
    // no product needed: Y_0 = I, Y_1 = 1.5 I - 0.5 A.  O = post * Y_K in fp32.
    W Yf;
    Yf.base = wy; Yf.rows = D; Yf.cols = D; Yf.ld = D; Yf.batch = B;
    W Ow = Yf;
    Ow.base = O;
    k::affine(M, D, dd, B, D, D, inv, K == 1 ? -0.5f : 0.f, K == 1 ? 1.5f : 1.f, Yf, 0.f, 0.f, nullptr,
              PREC_FP32_SIMT, st);
    k::affine(static_cast<const float*>(wy), D, dd, B, D, D, post, 1.f, 0.f, Ow, 0.f, 0.f, nullptr,
              PREC_FP32_SIMT, st);
    EGM_LAUNCHED();
    return EGM_OK;
  }
  return ns_products_fwd(S, prec, post, O, nullptr, st);
}

size_t egm_ns_bwd_workspace(int B, int D, int iters, int prec) {
  (void)iters; (void)prec;
  return 6 * pad256(w_bytes(B, D, D)) + pad256((size_t)B * D * D * 4) + 2 * pad256((size_t)B * 4) + 2048;
}

int egm_ns_bwd(const float* dO, const float* O, const float* M, const float* scal, const void* state,
               int B, int D, int iters, float eps, int post_mode, float* dM, int prec, void* ws, size_t ws_bytes,
               egm_stream_t stream) {
  (void)eps;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_ns_bwd: unknown precision mode %d", prec);
  EGM_REQUIRE(dO && O && M && scal && state && dM, EGM_ERR_ARG, "egm_ns_bwd: null pointer");
  EGM_REQUIRE(B > 0 && D > 0 && iters >= 0 && iters <= 64, EGM_ERR_ARG, "egm_ns_bwd: bad sizes");
  const int K = iters;
  NsState S(const_cast<void*>(state), B, D, K);
  Arena ar(ws, ws_bytes);
  W buf[6];
  for (int i = 0; i < 6; ++i) buf[i] = make_w(ar.take(w_bytes(B, D, D)), B, D, D);
  float* dA = static_cast<float*>(ar.take((size_t)B * D * D * 4));
  float* dotO = static_cast<float*>(ar.take((size_t)B * 4));
  float* dotA = static_cast<float*>(ar.take((size_t)B * 4));
  EGM_REQUIRE(buf[5].base && dA && dotO && dotA, EGM_ERR_WORKSPACE, "egm_ns_bwd: workspace %zu < %zu",
              ws_bytes, egm_ns_bwd_workspace(B, D, iters, prec));
  const float* inv = scal + B;
  const float* post = scal + 2 * B;
  const long long dd = (long long)D * D;
  // O = post * Y_K :  dY_K = post * dO ;  d tau (through post) = c * <dO,O> / (tau+eps),
  //   c = -1/2 for post = (tau+eps)^-1/2, +1/2 for post = (tau+eps)^+1/2
  const float coef_tau = post_mode == 0 ? -0.5f : 0.5f;
  k::batch_dot(dO, O, B, dd, dotO, st);
  EGM_LAUNCHED();
  if (K == 0) {
    EGM_CUDA(cudaMemsetAsync(dA, 0, (size_t)B * dd * 4, st));
  } else if (K == 1) {
    // Y_1 = 1.5 I - 0.5 A : dA = -0.5 * post * dO
    W dAw;
    dAw.base = dA; dAw.rows = D; dAw.cols = D; dAw.ld = D; dAw.batch = B;
    k::affine(dO, D, dd, B, D, D, post, -0.5f, 0.f, dAw, 0.f, 0.f, nullptr, PREC_FP32_SIMT, st);
    EGM_LAUNCHED();
  } else {
    k::affine(dO, D, dd, B, D, D, post, 1.f, 0.f, buf[0], 0.f, 0.f, nullptr, prec, st);
    EGM_LAUNCHED();
    NsFinal fin;
    fin.dA_f32 = dA;
    const int rc = ns_products_bwd(S, prec, buf, fin, st);
    if (rc != EGM_OK) return rc;
  }
  // A = inv * M with inv = 1/(tau+eps):  dM = inv dA + (d tau) I,
  //   d tau = coef_tau <dO,O> inv - <dA,M> inv^2
  k::batch_dot(dA, M, B, dd, dotA, st);
  k::ns_bwd_finish(dA, inv, dotO, dotA, coef_tau, B, D, dM, st);
  EGM_LAUNCHED();
  return EGM_OK;
}

// ============================================================ fused dense moment head
// pool -> trace normalisation -> Newton-Schulz (D x D) -> packed half-vector, with every
// intermediate kept in the GEMM engine's operand format: no fp32 M2 / O / dO / dM round trips.
//   tau = tr(Zc^T Wn Zc) = <Zc, U>      (a by-product of the U = Wn Zc product)
//   A = Zc^T U / (tau+eps) and T_0 = 1.5 I - 0.5 A leave the same accumulator
//   the last product writes post * Y_K as the packed upper triangle = operand of the Linear
// Backward: dY_K = post * unpack(dv); <dO,O> is supplied by the caller (for y = x W^T + b it is
// <dy, y - b>); <dA,A> is a by-product of the last product; dM = inv dA + dtau I is never formed:
// V1 = Zc dM^T = inv Zc dA^T + dtau Zc, V2 = Zc dM = inv Zc dA + dtau Zc.
size_t egm_mhd_state_bytes(int B, int N, int D, int iters, int prec) {
  return egm_pool_state_bytes(B, N, D, prec) + egm_ns_state_bytes(B, D, iters, prec) + 256;
}
size_t egm_mhd_fwd_workspace(int B, int N, int D, int iters, int prec) {
  (void)iters; (void)prec;
  return pad256(w_bytes(B, N, D)) + pad256((size_t)B * 4) +
         pad256((size_t)B * ((N + 127) / 128 + 1) * 8 * ((D + 255) / 256) * 4) + 2048;
}

int egm_mhd_fwd(const float* Z, const float* G, int B, int N, int D, int iters, float eps, int flags,
                void* x_planes, float* u, float* vecs, float* mu, float* scal, void* state, int prec,
                void* ws, size_t ws_bytes, egm_stream_t stream) {
  const bool sym = (flags & EGM_MHD_SYMMETRIC_GRAPH) != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec == PREC_BF16X3 || prec == PREC_BF16, EGM_ERR_ARG,
              "egm_mhd_fwd: tensor-core precision modes only (got %d)", prec);
  EGM_REQUIRE(Z && G && x_planes && vecs && mu && scal && state, EGM_ERR_ARG, "egm_mhd_fwd: null pointer");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0 && iters >= 2 && iters <= 64, EGM_ERR_ARG,
              "egm_mhd_fwd: bad sizes (needs iters >= 2)");
  const int K = iters;
  const size_t pool_bytes = pad256(egm_pool_state_bytes(B, N, D, prec));
  Arena sa(state, pool_bytes);
  const W Wn = make_w(sa.take(w_bytes(B, N, N)), B, N, N);
  const W Zc = make_w(sa.take(w_bytes(B, N, D)), B, N, D);
  NsState S(static_cast<uint8_t*>(state) + pool_bytes, B, D, K);
  Arena ar(ws, ws_bytes);
  const W U = make_w(ar.take(w_bytes(B, N, D)), B, N, D);
  float* tau = static_cast<float*>(ar.take((size_t)B * 4));
  GemmProblem gu;  // U = Wn Zc, tau = <U, Zc>
  gu.M = N; gu.N = D; gu.batch = B; gu.nterms = 1;
  gu.t[0] = term(Wn, 0, Zc, 0, N, prec);
  out_w(gu, U, prec);
  gu.F = w_mat(Zc, prec);
  gu.f_planes = 1;
  gu.dot_out = tau;
  float* dot_ws = static_cast<float*>(ar.take(gemm_tc_dot_ws_floats(gu) * 4));
  gu.dot_ws = dot_ws;
  EGM_REQUIRE(U.base && tau && dot_ws, EGM_ERR_WORKSPACE, "egm_mhd_fwd: workspace %zu < %zu", ws_bytes,
              egm_mhd_fwd_workspace(B, N, D, iters, prec));
  PoolVecs pv(vecs, B, N);
  k::degree(G, B, N, eps, pv.deg, pv.s, st);
  k::weight(G, pv.s, B, N, Wn, pv.w, pv.wdiag, prec, st);
  k::mean_center(Z, pv.w, pv.wdiag, B, N, D, eps, pv.t, pv.sw, mu, u, Zc, prec, st);
  EGM_LAUNCHED();
  EGM_CUDA(run_gemm(gu, prec, st));
  k::mh_scalars_fwd(tau, B, eps, scal, st);
  EGM_LAUNCHED();
  const float* inv = scal + B;
  const float* post = scal + 2 * B;
  {
    GemmProblem g;  // A = inv Zc^T U ; T_0 = 1.5 I - 0.5 A
    g.M = D; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 1, U, 0, N, prec);
    g.alpha_b = inv;
    g.sym_out = sym;   // Zc^T Wn Zc is symmetric when the graph is
    out_w(g, S.A(), prec);
    g.Cp2 = w_mat(S.T(0), prec);
    g.c2_scale = -0.5f; g.c2_eye = 1.5f;
    EGM_CUDA(run_gemm(g, prec, st));
  }
  const long long L = (long long)D * (D + 1) / 2;
  W X;
  X.base = x_planes; X.rows = B; X.cols = (int)L; X.ld = w_ld((int)L); X.batch = 1;
  return ns_products_fwd(S, prec, post, nullptr, &X, st, sym);
}

size_t egm_mhd_bwd_workspace(int B, int N, int D, int iters, int prec) {
  (void)iters; (void)prec;
  return 7 * pad256(w_bytes(B, D, D)) + 2 * pad256(w_bytes(B, N, D)) + pad256((size_t)B * N * D * 4) +
         pad256((size_t)B * N * egm_gpf_ldr(N) * 4) + pad256((size_t)B * D * 4) +
         3 * pad256((size_t)B * N * 4) + 2 * pad256((size_t)B * 4) +
         pad256((size_t)B * 16 * 4 * ((D + 255) / 256) * ((D + 255) / 256)) + 4096;
}

int egm_mhd_bwd(const float* dv, const float* dotO, const float* du, const float* Z, const float* G,
                const float* u, const float* vecs, const float* mu, const float* scal, const void* state,
                int B, int N, int D, int iters, float eps, int flags, float* dZ, float* dG, int prec,
                void* ws, size_t ws_bytes, egm_stream_t stream) {
  const bool sym = (flags & EGM_MHD_SYMMETRIC_GRAPH) != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec == PREC_BF16X3 || prec == PREC_BF16, EGM_ERR_ARG,
              "egm_mhd_bwd: tensor-core precision modes only (got %d)", prec);
  EGM_REQUIRE(dv && dotO && Z && G && vecs && mu && scal && state && dZ && dG, EGM_ERR_ARG,
              "egm_mhd_bwd: null pointer");
  EGM_REQUIRE(!du || u, EGM_ERR_ARG, "egm_mhd_bwd: du given without u");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0 && iters >= 2 && iters <= 64, EGM_ERR_ARG, "egm_mhd_bwd: bad sizes");
  const int K = iters;
  const size_t pool_bytes = pad256(egm_pool_state_bytes(B, N, D, prec));
  Arena sa(const_cast<void*>(state), pool_bytes);
  const W Wn = make_w(sa.take(w_bytes(B, N, N)), B, N, N);
  const W Zc = make_w(sa.take(w_bytes(B, N, D)), B, N, D);
  NsState S(static_cast<uint8_t*>(const_cast<void*>(state)) + pool_bytes, B, D, K);
  Arena ar(ws, ws_bytes);
  W buf[6];
  for (int i = 0; i < 6; ++i) buf[i] = make_w(ar.take(w_bytes(B, D, D)), B, D, D);
  const W dAw = make_w(ar.take(w_bytes(B, D, D)), B, D, D);
  const W V1 = make_w(ar.take(w_bytes(B, N, D)), B, N, D), V2 = make_w(ar.take(w_bytes(B, N, D)), B, N, D);
  float* dZc = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  const long long ldW = egm_gpf_ldr(N);
  float* dW = static_cast<float*>(ar.take((size_t)B * N * ldW * 4));
  float* dmu = static_cast<float*>(ar.take((size_t)B * D * 4));
  float* dw = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* ds = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* dt = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* dotA = static_cast<float*>(ar.take((size_t)B * 4));
  float* dtau = static_cast<float*>(ar.take((size_t)B * 4));
  const int tl = (D + 255) / 256;
  float* dot_ws = static_cast<float*>(ar.take((size_t)B * 16 * 4 * tl * tl));
  EGM_REQUIRE(buf[5].base && dAw.base && V1.base && V2.base && dZc && dW && dmu && dw && ds && dt && dotA &&
                  dtau && dot_ws,
              EGM_ERR_WORKSPACE, "egm_mhd_bwd: workspace %zu < %zu", ws_bytes,
              egm_mhd_bwd_workspace(B, N, D, iters, prec));
  const float* inv = scal + B;
  const float* post = scal + 2 * B;
  const long long L = (long long)D * (D + 1) / 2;
  NsFinal fin;
  fin.dA_w = &dAw;
  fin.dot_out = dotA;
  fin.dot_ws = dot_ws;
  if (sym) {
    // symmetric graph: sym(dA) by the tangent chain (upper tiles only), then dM = inv dA + dtau I is
    // symmetric, V1 == V2, and the dG returned is the symmetric part of the reference's - the only
    // part a symmetric-graph producer (GraphPolynomialFusion's symmetrisation) lets through
    k::triu_unpack_sym_planes(dv, L, B, D, post, -0.5f, buf[0], prec, st);   // Y'_1 = -post sym(dO)/2
    EGM_LAUNCHED();
    const int rc = ns_tangent_bwd(S, prec, buf, fin, st);
    if (rc != EGM_OK) return rc;
    k::mh_dtau(scal, B, dotO, dotA, dtau, st);
    EGM_LAUNCHED();
    GemmProblem g;  // V = Zc dM = inv Zc dA + dtau Zc
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 0, dAw, 0, D, prec);
    g.t[0].symB = 1;
    g.alpha_b = inv;
    addend_w(g, Zc, 1.f, prec);
    g.gamma_b = dtau;
    out_w(g, V1, prec);
    EGM_CUDA(run_gemm(g, prec, st));
    return pool_bwd_tail(Wn, Zc, V1, V1, du, Z, G, u, vecs, mu, B, N, D, eps, dZc, dW, ldW, dmu, dw, ds, dt,
                         dZ, dG, prec, st, true);
  }
  k::triu_unpack_planes(dv, L, B, D, post, buf[0], nullptr, prec, st);   // dY_K = post * dO
  EGM_LAUNCHED();
  const int rc = ns_products_bwd(S, prec, buf, fin, st);
  if (rc != EGM_OK) return rc;
  k::mh_dtau(scal, B, dotO, dotA, dtau, st);
  EGM_LAUNCHED();
  {
    GemmProblem g;  // V1 = Zc dM^T = inv Zc dA^T + dtau Zc
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 0, dAw, 1, D, prec);
    g.alpha_b = inv;
    addend_w(g, Zc, 1.f, prec);
    g.gamma_b = dtau;
    out_w(g, V1, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  {
    GemmProblem g;  // V2 = Zc dM = inv Zc dA + dtau Zc
    g.M = N; g.N = D; g.batch = B; g.nterms = 1;
    g.t[0] = term(Zc, 0, dAw, 0, D, prec);
    g.alpha_b = inv;
    addend_w(g, Zc, 1.f, prec);
    g.gamma_b = dtau;
    out_w(g, V2, prec);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  return pool_bwd_tail(Wn, Zc, V1, V2, du, Z, G, u, vecs, mu, B, N, D, eps, dZc, dW, ldW, dmu, dw, ds, dt,
                       dZ, dG, prec, st);
}

int egm_rowdot_bias(const float* dy, const float* y, const float* bias, int B, int n, float* out,
                    egm_stream_t stream) {
  EGM_REQUIRE(dy && y && out && B > 0 && n > 0, EGM_ERR_ARG, "egm_rowdot_bias: bad argument");
  k::rowdot_bias(dy, y, bias, B, n, out, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}

// ========================================================================== misc
int egm_triu_pack(const float* O, int B, int D, float* v, egm_stream_t stream) {
  EGM_REQUIRE(O && v && B > 0 && D > 0, EGM_ERR_ARG, "egm_triu_pack: bad argument");
  k::triu_pack(O, B, D, v, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}
int egm_triu_unpack(const float* dv, int B, int D, float* dO, egm_stream_t stream) {
  EGM_REQUIRE(dv && dO && B > 0 && D > 0, EGM_ERR_ARG, "egm_triu_unpack: bad argument");
  k::triu_unpack(dv, B, D, dO, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}
int egm_sketch_fwd(const float* x, int B, int D, int S, const int* off, const int* idx,
                   const float* sgn, float* cs, float* out, egm_stream_t stream) {
  EGM_REQUIRE(x && off && idx && sgn && cs && out && B > 0 && D > 0 && S > 0, EGM_ERR_ARG,
              "egm_sketch_fwd: bad argument");
  k::sketch_fwd(x, B, D, S, off, idx, sgn, cs, out, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}
int egm_sketch_bwd(const float* dout, const float* cs, int B, int D, int S, const long long* hash,
                   const long long* sign, float* dx, egm_stream_t stream) {
  EGM_REQUIRE(dout && cs && hash && sign && dx && B > 0 && D > 0 && S > 0, EGM_ERR_ARG,
              "egm_sketch_bwd: bad argument");
  k::sketch_bwd(dout, cs, B, D, S, hash, sign, dx, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}

size_t egm_gram_workspace(int B, int N, int D, int prec) {
  (void)prec;
  return pad256(w_bytes(B, N, D)) + pad256(w_bytes(B, N, N)) + pad256((size_t)B * N * D * 4) + 1024;
}
int egm_gram_fwd(const float* x, int B, int N, int D, int cosine, float eps, float* R, float* nrm,
                 int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec) && x && R && nrm && B > 0 && N > 0 && D > 0, EGM_ERR_ARG, "egm_gram_fwd: bad argument");
  Arena ar(ws, ws_bytes);
  void* wx = ar.take(w_bytes(B, N, D));
  EGM_REQUIRE(wx, EGM_ERR_WORKSPACE, "egm_gram_fwd: workspace too small");
  const W Xn = make_w(wx, B, N, D);
  k::rownorm(x, B, N, D, eps, cosine, nrm, Xn, prec, st);
  EGM_LAUNCHED();
  GemmProblem g;
  g.M = N; g.N = N; g.batch = B; g.nterms = 1;
  g.t[0] = term(Xn, 0, Xn, 1, D, prec);
  g.Cf = f32_mat(R, N, N, N, (long long)N * N);
  EGM_CUDA(run_gemm(g, prec, st));
  return EGM_OK;
}
int egm_gram_bwd(const float* dR, const float* x, const float* nrm, int B, int N, int D, int cosine,
                 float eps, float* dx, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec) && dR && x && nrm && dx && B > 0 && N > 0 && D > 0, EGM_ERR_ARG, "egm_gram_bwd: bad argument");
  Arena ar(ws, ws_bytes);
  void* wx = ar.take(w_bytes(B, N, D));
  void* wr = ar.take(w_bytes(B, N, N));
  float* dxn = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  EGM_REQUIRE(wx && wr && dxn, EGM_ERR_WORKSPACE, "egm_gram_bwd: workspace too small");
  const W Xn = make_w(wx, B, N, D), dRw = make_w(wr, B, N, N);
  k::rownorm(x, B, N, D, eps, cosine, dxn, Xn, prec, st);
  k::affine(dR, N, (long long)N * N, B, N, N, nullptr, 1.f, 0.f, dRw, 0.f, 0.f, nullptr, prec, st);
  EGM_LAUNCHED();
  GemmProblem g;  // d Xn = dR Xn + dR^T Xn
  g.M = N; g.N = D; g.batch = B; g.nterms = 2;
  g.t[0] = term(dRw, 0, Xn, 0, N, prec);
  g.t[1] = term(dRw, 1, Xn, 0, N, prec);
  g.Cf = f32_mat(cosine ? dxn : dx, N, D, D, (long long)N * D);
  EGM_CUDA(run_gemm(g, prec, st));
  if (cosine) {
    k::rownorm_bwd(x, nrm, dxn, B, N, D, eps, dx, st);
    EGM_LAUNCHED();
  }
  return EGM_OK;
}
int egm_normalize_graph(const float* G, int B, int N, int method, float eps, float* out, float* deg,
                        egm_stream_t stream) {
  EGM_REQUIRE(G && out && deg && B > 0 && N > 0 && (method == 0 || method == 1), EGM_ERR_ARG,
              "egm_normalize_graph: bad argument");
  k::normalize_graph(G, B, N, method, eps, out, deg, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}
int egm_batch_trace(const float* M, int B, int D, float* tr, egm_stream_t stream) {
  EGM_REQUIRE(M && tr && B > 0 && D > 0, EGM_ERR_ARG, "egm_batch_trace: bad argument");
  k::batch_trace(M, B, D, tr, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}

// ============================================================= graph alignment loss
int egm_align_fwd(const float* G, const long long* labels, int B, int N, float* g, float* dg,
                  float* rowloss, float* loss, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(G && labels && g && dg && rowloss && loss && B > 0 && N > 0, EGM_ERR_ARG,
              "egm_align_fwd: bad argument");
  k::graph_mean(G, B, (long long)N * N, g, st);
  k::align_rows(g, labels, B, rowloss, dg, st);
  k::sum_partials(rowloss, B, 1, loss, st);
  EGM_LAUNCHED();
  return EGM_OK;
}
int egm_align_bwd(const float* dg, const float* dloss, int B, int N, float* dG, egm_stream_t stream) {
  EGM_REQUIRE(dg && dloss && dG && B > 0 && N > 0, EGM_ERR_ARG, "egm_align_bwd: bad argument");
  k::align_bwd(dg, dloss, B, (long long)N * N, dG, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}

// ======================================================================== linear
// y = x W^T + bias on the tcgen05 engine (the Linear(D(D+1)/2 -> d) of second_net has K = 295 296:
// one 256 x 256 output tile, so K is sliced across all CTA pairs and the partial sums reduced).
static int linear_splits(int M, int N, int K) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ((M + 255) / 256) * ((N + 255) / 256);
  const int nkb = (K + 63) / 64;
  int s = (sms / 2) / (tiles > 0 ? tiles : 1);
  if (s > nkb) s = nkb;
  if (s > 160) s = 160;
  return s < 1 ? 1 : s;
}
size_t egm_linear_state_bytes(int M, int N, int K, int prec) {
  (void)prec;
  return pad256(w_bytes(1, M, K)) + pad256(w_bytes(1, N, K)) + 256;
}
size_t egm_linear_fwd_workspace(int M, int N, int K, int prec) {
  (void)K; (void)prec;
  return pad256((size_t)160 * M * N * 4) + 512;
}
int egm_linear_fwd(const float* x, const float* Wt, const float* bias, int M, int N, int K, float* y,
                   void* state, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec == PREC_BF16X3 || prec == PREC_BF16, EGM_ERR_ARG,
              "egm_linear_fwd: tensor-core precision modes only (got %d)", prec);
  EGM_REQUIRE(Wt && y && state && M > 0 && N > 0 && K > 0, EGM_ERR_ARG, "egm_linear_fwd: bad argument");
  Arena sa(state, egm_linear_state_bytes(M, N, K, prec));
  const W Xw = make_w(sa.take(w_bytes(1, M, K)), 1, M, K);
  const W Ww = make_w(sa.take(w_bytes(1, N, K)), 1, N, K);
  const int S = linear_splits(M, N, K);
  Arena ar(ws, ws_bytes);
  float* partial = static_cast<float*>(ar.take((size_t)S * M * N * 4));
  EGM_REQUIRE(partial, EGM_ERR_WORKSPACE, "egm_linear_fwd: workspace too small");
  // x == NULL: the producer (egm_mhd_fwd / egm_mlr_fwd) already wrote the operand planes into `state`
  if (x) k::affine(x, K, (long long)M * K, 1, M, K, nullptr, 1.f, 0.f, Xw, 0.f, 0.f, nullptr, prec, st);
  k::affine(Wt, K, (long long)N * K, 1, N, K, nullptr, 1.f, 0.f, Ww, 0.f, 0.f, nullptr, prec, st);
  EGM_LAUNCHED();
  GemmProblem g;
  g.M = M; g.N = N; g.batch = 1; g.nterms = 1; g.split_k = S;
  g.t[0] = term(Xw, 0, Ww, 1, K, prec);
  g.Cf = f32_mat(partial, M, N, N, (long long)M * N);
  EGM_CUDA(run_gemm(g, prec, st));
  k::reduce_splits(partial, S, M, N, bias, y, st);
  EGM_LAUNCHED();
  return EGM_OK;
}
size_t egm_linear_bwd_workspace(int M, int N, int K, int prec) {
  (void)K; (void)prec;
  return pad256(w_bytes(1, M, N)) + 512;
}
int egm_linear_bwd(const float* dy, const void* state, int M, int N, int K, float* dx, float* dW,
                   float* dbias, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec == PREC_BF16X3 || prec == PREC_BF16, EGM_ERR_ARG,
              "egm_linear_bwd: tensor-core precision modes only (got %d)", prec);
  EGM_REQUIRE(dy && state && M > 0 && N > 0 && K > 0, EGM_ERR_ARG, "egm_linear_bwd: bad argument");
  Arena sa(const_cast<void*>(state), egm_linear_state_bytes(M, N, K, prec));
  const W Xw = make_w(sa.take(w_bytes(1, M, K)), 1, M, K);
  const W Ww = make_w(sa.take(w_bytes(1, N, K)), 1, N, K);
  Arena ar(ws, ws_bytes);
  void* wdy = ar.take(w_bytes(1, M, N));
  EGM_REQUIRE(wdy, EGM_ERR_WORKSPACE, "egm_linear_bwd: workspace too small");
  const W dYw = make_w(wdy, 1, M, N);
  k::affine(dy, N, (long long)M * N, 1, M, N, nullptr, 1.f, 0.f, dYw, 0.f, 0.f, nullptr, prec, st);
  EGM_LAUNCHED();
  if (dx) {  // dx = dy W
    GemmProblem g;
    g.M = M; g.N = K; g.batch = 1; g.nterms = 1;
    g.t[0] = term(dYw, 0, Ww, 0, N, prec);
    g.Cf = f32_mat(dx, M, K, K, (long long)M * K);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  if (dW) {  // dW = dy^T x
    GemmProblem g;
    g.M = N; g.N = K; g.batch = 1; g.nterms = 1;
    g.t[0] = term(dYw, 1, Xw, 0, M, prec);
    g.Cf = f32_mat(dW, N, K, K, (long long)N * K);
    EGM_CUDA(run_gemm(g, prec, st));
  }
  if (dbias) {
    k::colsum(dy, M, N, dbias, st);
    EGM_LAUNCHED();
  }
  return EGM_OK;
}

// ============================================================ feature-net tail (BN + GELU + Dropout)
int egm_feature_tail_fwd(const float* y, int M, int N, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, int training, float momentum, float bn_eps,
                         float drop_p, unsigned long long seed, float* out, float* save_mean, float* save_rstd,
                         egm_stream_t stream) {
  EGM_REQUIRE(y && out && save_mean && save_rstd && M > 0 && N > 0, EGM_ERR_ARG, "egm_feature_tail_fwd: bad argument");
  EGM_REQUIRE(training || (running_mean && running_var), EGM_ERR_ARG,
              "egm_feature_tail_fwd: eval mode needs the running statistics");
  EGM_REQUIRE((running_mean == nullptr) == (running_var == nullptr), EGM_ERR_ARG,
              "egm_feature_tail_fwd: running_mean and running_var go together");
  EGM_REQUIRE(drop_p >= 0.f && drop_p <= 1.f, EGM_ERR_ARG, "egm_feature_tail_fwd: dropout p=%f", (double)drop_p);
  k::feature_tail_fwd(y, M, N, gamma, beta, running_mean, running_var, training, momentum, bn_eps, drop_p, seed,
                      out, save_mean, save_rstd, static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}
int egm_feature_tail_bwd(const float* dout, const float* y, const float* gamma, const float* beta,
                         const float* save_mean, const float* save_rstd, int M, int N, int training,
                         float drop_p, unsigned long long seed, float* dy, float* dgamma, float* dbeta,
                         egm_stream_t stream) {
  EGM_REQUIRE(dout && y && save_mean && save_rstd && dy && M > 0 && N > 0, EGM_ERR_ARG,
              "egm_feature_tail_bwd: bad argument");
  k::feature_tail_bwd(dout, y, M, N, gamma, beta, save_mean, save_rstd, training, drop_p, seed, dy, dgamma, dbeta,
                      static_cast<cudaStream_t>(stream));
  EGM_LAUNCHED();
  return EGM_OK;
}

size_t egm_bmm_workspace(int B, int M, int N, int K, int prec) {
  (void)prec;
  // either orientation of each operand (the row padding depends on which extent is the row length)
  auto both = [&](int r, int c) { const size_t x = w_bytes(B, r, c), y = w_bytes(B, c, r); return x > y ? x : y; };
  return pad256(both(M, K)) + pad256(both(K, N)) + 1024;
}
int egm_bmm(const float* A, int transA, const float* Bm, int transB, int B, int M, int N, int K,
            float alpha, float* C, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec) && A && Bm && C && B > 0 && M > 0 && N > 0 && K > 0, EGM_ERR_ARG, "egm_bmm: bad argument");
  const int ar_ = transA ? K : M, ac = transA ? M : K, br = transB ? N : K, bc = transB ? K : N;
  Arena ar(ws, ws_bytes);
  void* wa = ar.take(w_bytes(B, ar_, ac));
  void* wb = ar.take(w_bytes(B, br, bc));
  EGM_REQUIRE(wa && wb, EGM_ERR_WORKSPACE, "egm_bmm: workspace too small");
  const W Aw = make_w(wa, B, ar_, ac), Bw = make_w(wb, B, br, bc);
  k::affine(A, ac, (long long)ar_ * ac, B, ar_, ac, nullptr, 1.f, 0.f, Aw, 0.f, 0.f, nullptr, prec, st);
  k::affine(Bm, bc, (long long)br * bc, B, br, bc, nullptr, 1.f, 0.f, Bw, 0.f, 0.f, nullptr, prec, st);
  EGM_LAUNCHED();
  GemmProblem g;
  g.M = M; g.N = N; g.batch = B; g.nterms = 1;
  g.t[0] = term(Aw, transA, Bw, transB, K, prec);
  g.alpha = alpha;
  g.Cf = f32_mat(C, M, N, N, (long long)M * N);
  EGM_CUDA(run_gemm(g, prec, st));
  return EGM_OK;
}

}  // extern "C"
