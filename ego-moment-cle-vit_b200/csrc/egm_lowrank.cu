// Low-rank evaluation of pooling + iSQRT-COV for MomentHead when N < D.
//
// M2 = C^T W C has rank <= N (C = centred tokens, N x D; W = normalised graph, N x N), and every
// Newton-Schulz iterate is a polynomial in A = M2/tau':  f(A) = f(0) I + C^T W g(H) C / tau' with
// H = C C^T W / tau' (N x N) and g(x) = (f(x) - f(0))/x. Writing Y_k = a_k I + L(Q_k),
// Z_k = L(R_k), L(Q) = C^T W Q C / tau', and yh_k = a_k I + H Q_k, the coupled iteration of
// moment_head.py:53-64 becomes, entirely on N x N matrices,
//     S_k = -1/2 R_k yh_k,   Th_k = 3/2 I + H S_k,
//     Q_{k+1} = a_k S_k + Q_k Th_k,   R_{k+1} = R_k Th_k,   yh_{k+1} = yh_k Th_k,   a_{k+1} = 3/2 a_k
// (Q_1 = -1/2 I, R_1 = yh_1 = Th_0 = 3/2 I - 1/2 H), and the output is
//     O = tau'^-1/2 ( a_K I + C^T (W Q_K C) / tau' ).
// Same function of (Z, G) as the dense chain (no assumption on W: it may be non-symmetric), about
// 20x fewer flops at (N, D) = (197, 768). The backward is the reverse mode of exactly these
// recurrences (verified against autograd of the reference's dense loop to 2e-14 in fp64).
#include "egm_chain.h"

using namespace egm;
using namespace egm::chain;

namespace {

struct Gb {  // small builder around GemmProblem
  GemmProblem g;
  int prec;
  Gb(int M, int N, int batch, int prec_) : prec(prec_) { g.M = M; g.N = N; g.batch = batch; g.nterms = 0; }
  Gb& t(const W& A, int tA, const W& B, int tB, int K) { g.t[g.nterms++] = term(A, tA, B, tB, K, prec); return *this; }
  Gb& alpha(float a, const float* ab = nullptr) { g.alpha = a; g.alpha_b = ab; return *this; }
  Gb& eye(float b, const float* bb = nullptr) { g.beta_eye = b; g.beta_b = bb; return *this; }
  Gb& addw(const W& E, float gamma, const float* gb = nullptr) { addend_w(g, E, gamma, prec); g.gamma_b = gb; return *this; }
  Gb& addf(const float* E, int rows, int cols, long long ld, float gamma) {
    g.E = f32_mat(E, rows, cols, ld, (long long)rows * ld); g.e_planes = 0; g.gamma = gamma; return *this;
  }
  Gb& outw(const W& w) { out_w(g, w, prec); return *this; }
  Gb& outf(float* p, int rows, int cols, long long ld) { g.Cf = f32_mat(p, rows, cols, ld, (long long)rows * ld); return *this; }
  cudaError_t run(cudaStream_t st) { return run_gemm(g, prec, st); }
};

struct MlrState {
  int B, N, D, K;
  uint8_t* base;
  size_t nn, nd;
  MlrState(void* s, int B_, int N_, int D_, int K_) : B(B_), N(N_), D(D_), K(K_), base(static_cast<uint8_t*>(s)) {
    nn = pad256(w_bytes(B, N, N));
    nd = pad256(w_bytes(B, N, D));
  }
  static size_t bytes(int B, int N, int D, int K) {
    const size_t nn = pad256(w_bytes(B, N, N)), nd = pad256(w_bytes(B, N, D));
    return 2 * nd + (size_t)(4 + count_k(K)) * nn + 256;
  }
  // per-iteration N x N matrices: T[0..K-1], S[1..K-1], Q[1..K], R[2..K-1], Y[2..K-1]
  static int count_k(int K) { return K + (K - 1) + K + 2 * (K > 2 ? K - 2 : 0); }
  W nnw(size_t i) const { return make_w(base + 2 * nd + i * nn, B, N, N); }
  W Wn() const { return nnw(0); }
  W Gm() const { return nnw(1); }
  W H() const { return nnw(2); }
  W X() const { return nnw(3); }
  W Zc() const { return make_w(base, B, N, D); }
  W V() const { return make_w(base + nd, B, N, D); }
  W T(int k) const { return nnw(4 + k); }                                   // k in [0, K-1]
  W S(int k) const { return nnw(4 + K + (k - 1)); }                         // k in [1, K-1]
  W Q(int k) const { return nnw(4 + K + (K - 1) + (k - 1)); }               // k in [1, K]
  W R(int k) const { return k == 1 ? T(0) : nnw(4 + 3 * K - 1 + (k - 2)); }  // k in [1, K-1]
  W Y(int k) const { return k == 1 ? T(0) : nnw(4 + 3 * K - 1 + (K > 2 ? K - 2 : 0) + (k - 2)); }
};

}  // namespace

extern "C" {

size_t egm_mlr_state_bytes(int B, int N, int D, int iters, int prec) {
  (void)prec;
  return MlrState::bytes(B, N, D, iters < 1 ? 1 : iters);
}
size_t egm_mlr_fwd_workspace(int B, int N, int D, int iters, int prec) {
  (void)N; (void)D; (void)iters; (void)prec;
  return pad256((size_t)B * 4) + 512;
}

int egm_mlr_fwd(const float* Z, const float* G, int B, int N, int D, int iters, float eps, float* O,
                void* x_planes, float* u, float* vecs, float* mu, float* scal, void* state, int prec,
                void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_mlr_fwd: unknown precision mode %d", prec);
  EGM_REQUIRE(Z && G && (O || x_planes) && vecs && mu && scal && state, EGM_ERR_ARG, "egm_mlr_fwd: null pointer");
  EGM_REQUIRE(!x_planes || prec != PREC_FP32_SIMT, EGM_ERR_ARG,
              "egm_mlr_fwd: the packed-planes output needs a tensor-core precision mode");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0 && iters >= 1 && iters <= 64, EGM_ERR_ARG,
              "egm_mlr_fwd: bad sizes (needs iters >= 1)");
  const int K = iters;
  MlrState S(state, B, N, D, K);
  Arena ar(ws, ws_bytes);
  float* tau = static_cast<float*>(ar.take((size_t)B * 4));
  EGM_REQUIRE(tau, EGM_ERR_WORKSPACE, "egm_mlr_fwd: workspace too small");
  PoolVecs pv(vecs, B, N);
  const W Wn = S.Wn(), Zc = S.Zc();
  k::degree(G, B, N, eps, pv.deg, pv.s, st);
  k::weight(G, pv.s, B, N, Wn, pv.w, pv.wdiag, prec, st);
  k::mean_center(Z, pv.w, pv.wdiag, B, N, D, eps, pv.t, pv.sw, mu, u, Zc, prec, st);
  EGM_LAUNCHED();
  // Gm = Zc Zc^T ; tau = tr(Zc^T Wn Zc) = <Wn, Gm>
  EGM_CUDA(Gb(N, N, B, prec).t(Zc, 0, Zc, 1, D).outw(S.Gm()).run(st));
  k::w_dot(Wn, S.Gm(), tau, prec, st);
  float aK = 1.f;
  for (int k = 0; k < K; ++k) aK *= 1.5f;
  k::mlr_scalars_fwd(tau, B, eps, aK, scal, st);
  EGM_LAUNCHED();
  const float* inv = scal + B;
  const float* c1 = scal + 3 * B;
  const float* betaK = scal + 4 * B;
  // H = Gm Wn / tau' ;  Th_0 = 1.5 I - 0.5 H ;  Q_1 = -0.5 I
  EGM_CUDA(Gb(N, N, B, prec).t(S.Gm(), 0, Wn, 0, N).alpha(1.f, inv).outw(S.H()).run(st));
  EGM_CUDA(Gb(N, N, B, prec).t(S.Gm(), 0, Wn, 0, N).alpha(-0.5f, inv).eye(1.5f).outw(S.T(0)).run(st));
  k::w_fill_eye(S.Q(1), -0.5f, prec, st);
  EGM_LAUNCHED();
  float ak = 1.5f;  // a_1
  for (int k = 1; k <= K - 1; ++k) {
    EGM_CUDA(Gb(N, N, B, prec).t(S.R(k), 0, S.Y(k), 0, N).alpha(-0.5f).outw(S.S(k)).run(st));
    EGM_CUDA(Gb(N, N, B, prec).t(S.H(), 0, S.S(k), 0, N).eye(1.5f).outw(S.T(k)).run(st));
    EGM_CUDA(Gb(N, N, B, prec).t(S.Q(k), 0, S.T(k), 0, N).addw(S.S(k), ak).outw(S.Q(k + 1)).run(st));
    if (k < K - 1) {
      EGM_CUDA(Gb(N, N, B, prec).t(S.R(k), 0, S.T(k), 0, N).outw(S.R(k + 1)).run(st));
      EGM_CUDA(Gb(N, N, B, prec).t(S.Y(k), 0, S.T(k), 0, N).outw(S.Y(k + 1)).run(st));
    }
    ak *= 1.5f;
  }
  // X = Wn Q_K ; V = X Zc ; O = c1 Zc^T V + betaK I
  EGM_CUDA(Gb(N, N, B, prec).t(Wn, 0, S.Q(K), 0, N).outw(S.X()).run(st));
  EGM_CUDA(Gb(N, D, B, prec).t(S.X(), 0, Zc, 0, N).outw(S.V()).run(st));
  if (x_planes) {
    // the result leaves as its packed upper triangle in operand planes (the Linear's x)
    const long long L = (long long)D * (D + 1) / 2;
    W X;
    X.base = x_planes; X.rows = B; X.cols = (int)L; X.ld = w_ld((int)L); X.batch = 1;
    Gb g(D, D, B, prec);
    g.t(Zc, 1, S.V(), 0, N).alpha(1.f, c1).eye(1.f, betaK);
    g.g.X = w_mat(X, prec);
    EGM_CUDA(g.run(st));
  } else {
    EGM_CUDA(Gb(D, D, B, prec).t(Zc, 1, S.V(), 0, N).alpha(1.f, c1).eye(1.f, betaK).outf(O, D, D, D).run(st));
  }
  return EGM_OK;
}

size_t egm_mlr_bwd_workspace(int B, int N, int D, int iters, int prec) {
  (void)iters; (void)prec;
  const size_t nn = pad256(w_bytes(B, N, N)), nd = pad256(w_bytes(B, N, D));
  return pad256(w_bytes(B, D, D)) + nd + 14 * nn + 3 * pad256((size_t)B * N * D * 4) +
         pad256((size_t)B * N * egm_gpf_ldr(N) * 4) + pad256((size_t)B * D * 4) +
         3 * pad256((size_t)B * N * 4) + 4 * pad256((size_t)B * 4) +
         pad256((size_t)B * k::triu_unpack_blocks(D) * 4) + 4096;
}

int egm_mlr_bwd(const float* dO, const float* dv, const float* dotOO_in, const float* du, const float* Z,
                const float* G, const float* O, const float* u, const float* vecs, const float* mu,
                const float* scal, const void* state, int B, int N, int D, int iters, float eps,
                float* dZ, float* dG, int prec, void* ws, size_t ws_bytes, egm_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EGM_REQUIRE(prec_ok(prec), EGM_ERR_ARG, "egm_mlr_bwd: unknown precision mode %d", prec);
  EGM_REQUIRE(Z && G && vecs && mu && scal && state && dZ && dG, EGM_ERR_ARG, "egm_mlr_bwd: null pointer");
  EGM_REQUIRE((dO && O) || (dv && dotOO_in), EGM_ERR_ARG,
              "egm_mlr_bwd: needs either (dO, O) or the packed gradient (dv, <dO,O>)");
  EGM_REQUIRE(!du || u, EGM_ERR_ARG, "egm_mlr_bwd: du given without u");
  EGM_REQUIRE(B > 0 && N > 0 && D > 0 && iters >= 1 && iters <= 64, EGM_ERR_ARG, "egm_mlr_bwd: bad sizes");
  const int K = iters;
  MlrState S(const_cast<void*>(state), B, N, D, K);
  Arena ar(ws, ws_bytes);
  const W dOw = make_w(ar.take(w_bytes(B, D, D)), B, D, D);
  const W dV = make_w(ar.take(w_bytes(B, N, D)), B, N, D);
  W nn[14];
  for (int i = 0; i < 14; ++i) nn[i] = make_w(ar.take(w_bytes(B, N, N)), B, N, N);
  float* P1 = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  float* P2 = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  float* dZc = static_cast<float*>(ar.take((size_t)B * N * D * 4));
  const long long ldW = egm_gpf_ldr(N);
  float* dW = static_cast<float*>(ar.take((size_t)B * N * ldW * 4));
  float* dmu = static_cast<float*>(ar.take((size_t)B * D * 4));
  float* dw = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* ds = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* dt = static_cast<float*>(ar.take((size_t)B * N * 4));
  float* trdO = static_cast<float*>(ar.take((size_t)B * 4));
  float* dotOO = static_cast<float*>(ar.take((size_t)B * 4));
  float* dotHs = static_cast<float*>(ar.take((size_t)B * 4));
  float* dtaup = static_cast<float*>(ar.take((size_t)B * 4));
  float* trp = static_cast<float*>(ar.take((size_t)B * k::triu_unpack_blocks(D) * 4));
  EGM_REQUIRE(trp && dOw.base && dV.base && nn[13].base && P1 && P2 && dZc && dW && dmu && dw && ds && dt && trdO &&
                  dotOO && dotHs && dtaup,
              EGM_ERR_WORKSPACE, "egm_mlr_bwd: workspace %zu < %zu", ws_bytes,
              egm_mlr_bwd_workspace(B, N, D, iters, prec));
  const W Wn = S.Wn(), Zc = S.Zc();
  const float* inv = scal + B;
  const float* c1 = scal + 3 * B;
  float aK = 1.f;
  for (int k = 0; k < K; ++k) aK *= 1.5f;
  const long long dd = (long long)D * D;

  if (dv) {
    // gradient of the packed half-vector: unpack straight into operand planes
    const long long L = (long long)D * (D + 1) / 2;
    k::triu_unpack_planes(dv, L, B, D, nullptr, dOw, trp, prec, st);
    k::sum_partials(trp, k::triu_unpack_blocks(D), B, trdO, st);
    dotOO = const_cast<float*>(dotOO_in);
  } else {
    k::batch_trace(dO, B, D, trdO, st);
    k::batch_dot(dO, O, B, dd, dotOO, st);
    k::affine(dO, D, dd, B, D, D, nullptr, 1.f, 0.f, dOw, 0.f, 0.f, nullptr, prec, st);
  }
  EGM_LAUNCHED();
  // O = betaK I + c1 Zc^T V
  EGM_CUDA(Gb(N, D, B, prec).t(Zc, 0, dOw, 0, D).alpha(1.f, c1).outw(dV).run(st));             // dV = c1 Zc dO
  EGM_CUDA(Gb(N, D, B, prec).t(S.V(), 0, dOw, 1, D).alpha(1.f, c1).outf(P1, N, D, D).run(st));  // c1 V dO^T
  W dX = nn[0], dQ = nn[1], dQn = nn[2], dR = nn[3], dRn = nn[4], dY = nn[5], dYn = nn[6], dH = nn[7],
    dHn = nn[8], dT = nn[9], tmp = nn[10], dSh = nn[11], dHs = nn[12], dGm = nn[13];
  EGM_CUDA(Gb(N, N, B, prec).t(dV, 0, Zc, 1, D).outw(dX).run(st));                             // dX = dV Zc^T
  EGM_CUDA(Gb(N, N, B, prec).t(Wn, 1, dX, 0, N).outw(dQ).run(st));                             // dQ_K = Wn^T dX
  bool have_rt = false, have_h = false;
  float ak = aK / 1.5f;  // a_{K-1}
  for (int k = K - 1; k >= 1; --k) {
    // dTh = Q_k^T dQ [+ R_k^T dR + yh_k^T dY]
    if (have_rt) {
      EGM_CUDA(Gb(N, N, B, prec).t(S.Q(k), 1, dQ, 0, N).t(S.R(k), 1, dR, 0, N).outw(tmp).run(st));
      EGM_CUDA(Gb(N, N, B, prec).t(S.Y(k), 1, dY, 0, N).addw(tmp, 1.f).outw(dT).run(st));
    } else {
      EGM_CUDA(Gb(N, N, B, prec).t(S.Q(k), 1, dQ, 0, N).outw(dT).run(st));
    }
    // dH += dTh S_k^T
    {
      Gb g(N, N, B, prec);
      g.t(dT, 0, S.S(k), 1, N);
      if (have_h) g.addw(dH, 1.f);
      EGM_CUDA(g.outw(dHn).run(st));
      W x = dH; dH = dHn; dHn = x;
      have_h = true;
    }
    // dSh = -1/2 (a_k dQ + H^T dTh)
    EGM_CUDA(Gb(N, N, B, prec).t(S.H(), 1, dT, 0, N).alpha(-0.5f).addw(dQ, -0.5f * ak).outw(dSh).run(st));
    // dQ_k = dQ Th_k^T ; dR_k = [dR Th_k^T +] dSh yh_k^T ; dyh_k = [dY Th_k^T +] R_k^T dSh
    EGM_CUDA(Gb(N, N, B, prec).t(dQ, 0, S.T(k), 1, N).outw(dQn).run(st));
    {
      Gb g(N, N, B, prec);
      g.t(dSh, 0, S.Y(k), 1, N);
      if (have_rt) g.t(dR, 0, S.T(k), 1, N);
      EGM_CUDA(g.outw(dRn).run(st));
    }
    {
      Gb g(N, N, B, prec);
      g.t(S.R(k), 1, dSh, 0, N);
      if (have_rt) g.t(dY, 0, S.T(k), 1, N);
      EGM_CUDA(g.outw(dYn).run(st));
    }
    W x = dQ; dQ = dQn; dQn = x;
    x = dR; dR = dRn; dRn = x;
    x = dY; dY = dYn; dYn = x;
    have_rt = true;
    ak /= 1.5f;
  }
  // k = 0: R_1 = yh_1 = Th_0 = 1.5 I - 0.5 H  =>  dH_total = dH - 1/2 (dR_1 + dyh_1);  dHs = inv * dH_total
  if (have_rt) {
    k::w_lincomb(dHs, 1.f, dH, -0.5f, &dR, -0.5f, &dY, inv, prec, st);
    k::w_dot(dHs, S.H(), dotHs, prec, st);
  } else {
    EGM_CUDA(cudaMemsetAsync(dHs.base, 0, w_bytes(B, N, N), st));
    EGM_CUDA(cudaMemsetAsync(dotHs, 0, (size_t)B * 4, st));
  }
  k::mlr_scalars_bwd(scal, B, aK, dotOO, trdO, dotHs, dtaup, st);
  EGM_LAUNCHED();
  // dGm = dHs Wn^T + dtaup Wn ;  dW = dX Q_K^T + Gm dHs + dtaup Gm
  EGM_CUDA(Gb(N, N, B, prec).t(dHs, 0, Wn, 1, N).addw(Wn, 1.f, dtaup).outw(dGm).run(st));
  EGM_CUDA(Gb(N, N, B, prec).t(dX, 0, S.Q(K), 1, N).t(S.Gm(), 0, dHs, 0, N).addw(S.Gm(), 1.f, dtaup)
               .outf(dW, N, N, ldW).run(st));
  // dZc = c1 V dO^T + X^T dV + (dGm + dGm^T) Zc
  EGM_CUDA(Gb(N, D, B, prec).t(S.X(), 1, dV, 0, N).t(dGm, 0, Zc, 0, N).addf(P1, N, D, D, 1.f).outf(P2, N, D, D).run(st));
  EGM_CUDA(Gb(N, D, B, prec).t(dGm, 1, Zc, 0, N).addf(P2, N, D, D, 1.f).outf(dZc, N, D, D).run(st));
  // shared pooling tail: centring, weighted mean, degree normalisation
  PoolVecs pv(const_cast<float*>(vecs), B, N);
  k::pool_bwd_dmu(dZc, du, pv.sw, pv.t, B, N, D, eps, dmu, st);
  k::pool_bwd_rows(dZc, Z, Zc, pv.w, pv.t, mu, u, du, dmu, B, N, D, eps, dZ, dw, dt, prec, st);
  k::pool_bwd_ds(dW, ldW, dw, dt, G, pv.s, B, N, 0, ds, st);
  k::pool_bwd_dG(dW, ldW, dw, dt, pv.s, pv.deg, ds, B, N, eps, 0, dG, st);
  EGM_LAUNCHED();
  return EGM_OK;
}

}  // extern "C"
