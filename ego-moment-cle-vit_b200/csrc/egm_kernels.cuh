// Bandwidth-bound kernels of the moment-pooling path (everything that is not a GEMM):
// token normalisation, Hadamard-power polynomial + symmetrise + clamp, degree
// normalisation, graph-weighted mean / centring, trace, triu pack, count-sketch, and the
// matching backward passes. Launchers only; definitions in egm_kernels.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "egm_gemm.h"

namespace egm {

// A "working matrix": batched row-major [batch][rows][ld] in the native operand format of
// the active precision mode - fp32 (PREC_FP32_SIMT) or bf16 hi/lo planes (tensor-core
// modes; the lo plane sits `batch*rows*ld` bf16 elements after the hi plane).
struct W {
  void* base = nullptr;
  int rows = 0, cols = 0;
  long long ld = 0;
  int batch = 1;
};
inline long long w_ld(int cols) { return ((long long)cols + 7) / 8 * 8; }
inline size_t w_bytes(int batch, int rows, int cols) {
  return (size_t)batch * rows * w_ld(cols) * 4;
}
inline W make_w(void* base, int batch, int rows, int cols) {
  W w;
  w.base = base; w.rows = rows; w.cols = cols; w.ld = w_ld(cols); w.batch = batch;
  return w;
}
// View of a working matrix as a GEMM operand.
inline Mat w_mat(const W& w, int prec) {
  Mat m;
  m.p0 = w.base;
  m.rows = w.rows; m.cols = w.cols; m.ld = w.ld; m.bstride = (long long)w.rows * w.ld;
  if (prec == PREC_BF16X3)
    m.p1 = static_cast<__nv_bfloat16*>(w.base) + (long long)w.batch * w.rows * w.ld;
  return m;
}
inline Mat f32_mat(const void* p, int rows, int cols, long long ld, long long bstride) {
  Mat m;
  m.p0 = const_cast<void*>(p); m.rows = rows; m.cols = cols; m.ld = ld; m.bstride = bstride;
  return m;
}

namespace k {

// out_i = a_i * s[b] * x + b_i * I  for i in {1,2}  (out2 optional; s optional)
void affine(const float* x, long long ld, long long bs, int batch, int rows, int cols,
            const float* s, float a1, float b1, const W& out1, float a2, float b2, const W* out2,
            int prec, cudaStream_t st);
// working matrix -> fp32
void export_f32(const W& in, float* out, long long ld, long long bs, int prec, cudaStream_t st);

// nrm[b,i] = ||x_i||;  xn = x / max(nrm, eps) (cosine) or x (dot)  -> working matrix
void rownorm(const float* x, int batch, int n, int d, float eps, int cosine, float* nrm,
             const W& xn, int prec, cudaStream_t st);
// da = (dAn - An*(An.dAn)) / max(nrm,eps)  [nrm >= eps]  else dAn/eps
void rownorm_bwd(const float* x, const float* nrm, const float* dxn, int batch, int n, int d,
                 float eps, float* dx, cudaStream_t st);

// G = max(sym(sum_pq c_pq f_p(Ra) f_q(Rp)), 0)
void gpf_poly_fwd(const float* Ra, const float* Rp, long long ldR, const float* coef, int P, int Q,
                  int symmetric, int batch, int n, float* G, cudaStream_t st);
// Ea = dRa + dRa^T, Ep = dRp + dRp^T (working matrices), dcoef[(P+1)(Q+1)]
// `symmetric`: bit 0 = symmetric_enforce, bit 1 = R_a / R_p symmetric bit for bit (fused forward).
// With `rowpart` (gpf_poly_bwd_rowpart_floats) and gpf_poly_bwd_can_fold(): the backward of F.normalize is
// folded in - Ea, Ep become E'' with dx = E'' X on the RAW token planes (see gpf_poly3_bwd_kernel).
void gpf_poly_bwd(const float* dG, const float* Ra, const float* Rp, long long ldR,
                  const float* coef, int P, int Q, int symmetric, int batch, int n, const W& Ea,
                  const W& Ep, float* partial, int nblocks, float* dcoef, int prec,
                  cudaStream_t st, const float* nrm_a = nullptr, const float* nrm_p = nullptr,
                  float eps = 0.f, float* rowpart = nullptr);
int gpf_poly_bwd_blocks(int batch, int n);
bool gpf_poly_bwd_can_fold(int P, int Q, int n, long long ldR, const W& Ea);
size_t gpf_poly_bwd_rowpart_floats(int batch, int n);

// GraphPolynomialFusion.forward in one pass over the tokens (egm_gpf_fused.cu): Gram matrices of both
// views on tcgen05 from fp32 tokens converted on the fly, cosine scaling + polynomial + clamp in the
// epilogue. Ra / Rp (optional, [B][n][ldR]) are written when the backward will need them.
bool gpf_fused_supported(int n, int d, int P, int Q, const float* a, const float* p);
// xa_raw / xp_raw (optional working matrices [B][n][d]): the RAW tokens as operand planes, for a backward
// that folds the normalisation into E (gpf_poly_bwd with rowpart).
cudaError_t gpf_fused_fwd(const float* a, const float* p, const float* coef, int batch, int n, int d, int P,
                          int Q, int cosine, float eps, float* G, float* Ra, float* Rp, long long ldR,
                          float* nrm_a, float* nrm_p, int npass, cudaStream_t st, const W* xa_raw = nullptr,
                          const W* xp_raw = nullptr);

// deg = G 1 ; s = rsqrt(max(deg, eps))
void degree(const float* G, int batch, int n, float eps, float* deg, float* s, cudaStream_t st);
// Wn_ij = s_i G_ij s_j (working matrix) ; w = Wn 1 ; wdiag_i = Wn_ii
void weight(const float* G, const float* s, int batch, int n, const W& Wn, float* w, float* wdiag,
            int prec, cudaStream_t st);
// t = sum wdiag ; sw = sum w ; mu = Z^T w/(t+eps) ; Zc = Z - mu ; u = Zc^T w /(t+eps)
void mean_center(const float* Z, const float* w, const float* wdiag, int batch, int n, int d,
                 float eps, float* t, float* sw, float* mu, float* u, const W& Zc, int prec,
                 cudaStream_t st);

// tr = trace(M); inv = 1/(tr+eps); post = (tr+eps)^(-1/2) [mode 0] or (tr+eps)^(1/2) [mode 1]
void trace_scales(const float* M, int batch, int d, float eps, int post_mode, float* tr,
                  float* inv, float* post, cudaStream_t st);
// out[b] = <X[b], Y[b]>  (fp32 matrices, same layout)
void batch_dot(const float* X, const float* Y, int batch, long long n_per, float* out,
               cudaStream_t st);
// dM = dA*inv + (coef_tau*dotO - dotA*inv)*inv I   with dotO = <dO,O>, dotA = <dA,M>
void ns_bwd_finish(const float* dA, const float* inv, const float* dotO, const float* dotA,
                   float coef_tau, int batch, int d, float* dM, cudaStream_t st);

// utils.ops.normalize_graph: deg = max(G 1, eps); method 0: G_ij/sqrt(deg_i deg_j); 1: G_ij/deg_i
void normalize_graph(const float* G, int batch, int n, int method, float eps, float* out,
                     float* deg, cudaStream_t st);
void batch_trace(const float* M, int batch, int d, float* tr, cudaStream_t st);

void triu_pack(const float* O, int batch, int d, float* v, cudaStream_t st);
void triu_unpack(const float* dv, int batch, int d, float* dO, cudaStream_t st);

// count sketch in gather form (CSR inverse of the hash): cs_k[b,s] = sum_{j in bucket s} sign*x[b,j]
void sketch_fwd(const float* x, int batch, int d, int S, const int* off, const int* idx,
                const float* sgn, float* cs /*[3,B,S]*/, float* out, cudaStream_t st);
void sketch_bwd(const float* dout, const float* cs, int batch, int d, int S, const long long* hash,
                const long long* sign /*[3,d] each, int64 as in the state_dict*/, float* dx,
                cudaStream_t st);

// ---- working-matrix helpers for the low-rank Newton-Schulz chain (N x N per image)
void w_fill_eye(const W& out, float c, int prec, cudaStream_t st);                 // out = c I
// out = s_b[b] * (a X + b Y + c Z)   (Y, Z, s_b optional)
void w_lincomb(const W& out, float a, const W& X, float b, const W* Y, float c, const W* Z,
               const float* s_b, int prec, cudaStream_t st);
void w_dot(const W& X, const W& Y, float* out, int prec, cudaStream_t st);        // out[b] = <X_b, Y_b>
// scal rows: 0 taup = tau+eps, 1 inv = 1/taup, 2 post = taup^-1/2, 3 c1 = post*inv, 4 betaK = post*aK
void mlr_scalars_fwd(const float* tau, int batch, float eps, float aK, float* scal, cudaStream_t st);
// dc1 = (dotOO - betaK trdO)/c1 ; dtaup = -inv^2 (post dc1 + dotHs taup^2) - 0.5 taup^-1.5 (aK trdO + inv dc1)
void mlr_scalars_bwd(const float* scal, int batch, float aK, const float* dotOO, const float* trdO,
                     const float* dotHs, float* dtaup, cudaStream_t st);

// ---- split-K epilogue / bias helpers for the Linear layers
// y[m,n] = sum_s partial[s,m,n] + bias[n]
void reduce_splits(const float* partial, int splits, int m, int n, const float* bias, float* y,
                   cudaStream_t st);
void colsum(const float* x, int m, int n, float* out, cudaStream_t st);   // out[n] = sum_m x[m,n]
// BatchNorm1d -> GELU(erf) -> Dropout after a feature net's Linear (moment_head.py:186-191,195-200):
// y [m,n] is the Linear's output; training: batch statistics (and the running-stat update when
// run_mean / run_var are given), eval: running statistics. save_mean / save_rstd [n] feed the backward.
void feature_tail_fwd(const float* y, int m, int n, const float* gamma, const float* beta, float* run_mean,
                      float* run_var, int training, float momentum, float bn_eps, float drop_p,
                      unsigned long long seed, float* out, float* save_mean, float* save_rstd, cudaStream_t st);
void feature_tail_bwd(const float* dout, const float* y, int m, int n, const float* gamma, const float* beta,
                      const float* save_mean, const float* save_rstd, int training, float drop_p,
                      unsigned long long seed, float* dy, float* dgamma, float* dbeta, cudaStream_t st);

// ---- fused moment head (pool -> iSQRT-COV -> packed half-vector as the Linear's operand)
// out = s_b[b] * unpack(dv[b, :L]) as a D x D working matrix, zero lower triangle;
// tr_partial (optional) [batch][triu_unpack_blocks(d)] partial traces of unpack(dv)
int triu_unpack_blocks(int d);
void triu_unpack_planes(const float* dv, long long ld_dv, int batch, int d, const float* s_b,
                        const W& out, float* tr_partial, int prec, cudaStream_t st);
// out = scale * s_b[b] * (unpack(dv[b]) + unpack(dv[b])^T)/2 in the GEMM engine's symmetric block
// storage (upper 256 x 256 blocks only; see GemmTerm::symA)
void triu_unpack_sym_planes(const float* dv, long long ld_dv, int batch, int d, const float* s_b,
                            float scale, const W& out, int prec, cudaStream_t st);
// out (a [batch, d(d+1)/2] working matrix with batch == 1 image of `batch` rows) = triu(O)
void triu_pack_planes(const float* O, int batch, int d, const W& out, int prec, cudaStream_t st);
// out[b] = sum_n dy[b,n] (y[b,n] - bias[n])
void rowdot_bias(const float* dy, const float* y, const float* bias, int batch, int n, float* out,
                 cudaStream_t st);
// scal rows: 0 tau, 1 inv = 1/(tau+eps), 2 post = (tau+eps)^-1/2
void mh_scalars_fwd(const float* tau, int batch, float eps, float* scal, cudaStream_t st);
// dtau = (-1/2 dotO - dotA) * inv
void mh_dtau(const float* scal, int batch, const float* dotO, const float* dotA, float* dtau,
             cudaStream_t st);
void sum_partials(const float* partial, int nper, int batch, float* out, cudaStream_t st);

// ---- graph alignment loss (ego_moment_clevit.py:278-316)
void graph_mean(const float* G, int batch, long long per, float* g, cudaStream_t st);   // g[b] = mean(G[b])
// rowloss_i = sum_j (sigmoid(g_i g_j) - [l_i == l_j])^2 / B^2 and dg = d(sum rowloss)/dg
void align_rows(const float* g, const long long* labels, int batch, float* rowloss, float* dg,
                cudaStream_t st);
void align_bwd(const float* dg, const float* dloss, int batch, long long per, float* dG, cudaStream_t st);

// ---- pooling backward pieces
// dmu = -(colsum(dZc) + sw*du/(t+eps))
void pool_bwd_dmu(const float* dZc, const float* du, const float* sw, const float* t, int batch,
                  int n, int d, float eps, float* dmu, cudaStream_t st);
// dZ = dZc + w (du + dmu)^T/(t+eps); dw = (Zc du + Z dmu)/(t+eps); dt = -(u.du + mu.dmu)/(t+eps)
void pool_bwd_rows(const float* dZc, const float* Z, const W& Zc, const float* w, const float* t,
                   const float* mu, const float* u, const float* du, const float* dmu, int batch,
                   int n, int d, float eps, float* dZ, float* dw, float* dt, int prec,
                   cudaStream_t st);
// ds_i = sum_j (dWf_ij G_ij s_j + dWf_ji G_ji s_j),  dWf = dW + dw 1^T + dt I
// sym: G and dW are symmetric - no mirrored reads
void pool_bwd_ds(const float* dW, long long ldW, const float* dw, const float* dt, const float* G,
                 const float* s, int batch, int n, int sym, float* ds, cudaStream_t st);
// dG_ij = s_i dWf_ij s_j + ddeg_i ;  ddeg = -0.5 * s^3 * ds * [deg >= eps]
// sym: the symmetric part (dG + dG^T)/2 of that gradient (dW symmetric)
void pool_bwd_dG(const float* dW, long long ldW, const float* dw, const float* dt, const float* s,
                 const float* deg, const float* ds, int batch, int n, float eps, int sym, float* dG,
                 cudaStream_t st);

}  // namespace k
}  // namespace egm
