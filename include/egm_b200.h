/* egm_b200.h - C ABI of the B200-native moment-pooling path of EGO-Moment-CLE-ViT.
 *
 * The reference (hibana2077/EGO-Moment-CLE-ViT) has no FFI: the path lives behind the Python
 * nn.Module API of src/models. This header is the boundary a binding
 * for that API sits on: every entry point takes plain device pointers, sizes and a CUDA
 * stream, never allocates, never synchronises, and returns 0 or a negative error code
 * (EGM_ERR_*); egm_last_error() gives the message. The caller owns every buffer, including
 * the scratch workspace whose size the matching *_workspace() call returns.
 *
 * All tensors are contiguous row-major fp32 unless stated. `prec` selects how the dense
 * contractions are evaluated (EGM_PREC_*). Buffers marked "state" are opaque (their layout
 * depends on `prec`); they only need to be passed back unchanged to the matching *_bwd.
 *
 * Reference interface each entry point replaces (file:line in the reference tree):
 *   egm_gpf_fwd/bwd      GraphPolynomialFusion.forward          src/models/gpf_kernel.py:117-159
 *                        (_compute_similarity :75-94, _hadamard_power :96-115) and its autograd
 *   egm_pool_fwd/bwd     MomentHead._normalize_weight_matrix    src/models/moment_head.py:246-266
 *                        MomentHead._graph_weighted_mean        src/models/moment_head.py:222-244
 *                        centring + M2 = Zc^T W Zc              src/models/moment_head.py:288-293
 *                        third-order weighted mean u            src/models/moment_head.py:305-311
 *   egm_ns_fwd/bwd       NewtonSchulzSqrtm.forward              src/models/moment_head.py:28-70
 *                        matrix_sqrt_newton_schulz (post_mode 1) src/utils/ops.py:122-165
 *   egm_mlr_fwd/bwd      the same pooling + iSQRT-COV lines, low-rank evaluation (opt-in)
 *   egm_mhd_fwd/bwd      MomentHead.forward 2nd-order branch    src/models/moment_head.py:279-300
 *                        (pool -> NewtonSchulzSqrtm -> _half_vectorize fused, dense D x D chain)
 *   egm_linear_fwd/bwd   nn.Linear of second_net / third_net    src/models/moment_head.py:186-200
 *   egm_triu_pack/unpack MomentHead._half_vectorize             src/models/moment_head.py:202-220
 *                        half_vectorize_symmetric               src/utils/ops.py:100-119
 *   egm_sketch_fwd/bwd   TensorSketch.forward/_count_sketch     src/models/moment_head.py:100-133
 *   egm_align_fwd/bwd    EGOMomentCLEViT._graph_alignment_loss  src/models/ego_moment_clevit.py:278-316
 *   egm_gram_fwd/bwd     cosine_similarity_matrix               src/utils/ops.py:355-381
 *   egm_normalize_graph  normalize_graph                        src/utils/ops.py:238-271
 *   egm_batch_trace      batch_trace                            src/utils/ops.py:316-326
 */
#ifndef EGM_B200_H_
#define EGM_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* egm_stream_t; /* cudaStream_t */

enum {
  EGM_PREC_FP32_SIMT = 0, /* fp32 FFMA on CUDA cores                                   */
  EGM_PREC_BF16X3 = 1,    /* fp32 emulated on tcgen05: bf16 hi/lo split, 3 MMAs/product */
  EGM_PREC_BF16 = 2       /* single bf16 tcgen05 MMA, fp32 accumulate                   */
};
enum {
  EGM_OK = 0,
  EGM_ERR_ARG = -1,       /* bad argument (null pointer, size, enum)     */
  EGM_ERR_WORKSPACE = -2, /* workspace / state buffer too small          */
  EGM_ERR_CUDA = -3       /* a CUDA call or launch failed                */
};

int egm_version(void);
const char* egm_last_error(void);
/* kernels launched by this library in this process so far (monotonic; bench.py's gpu_launches) */
unsigned long long egm_launch_count(void);
/* Optional timing of the tcgen05 GEMM engine (bench.py's roofline). level 1: every engine launch is
 * bracketed by two CUDA events on its own stream; level 2: one pair of events around each
 * Newton-Schulz chain instead (the launches inside keep their dependent-launch overlap); 0: off.
 * egm_prof_read(i) synchronises on record i and returns its duration, its algorithmic flops and
 * dims = {M, N, K of term 0, K of term 1 (0 if absent), batch, MMA passes per product}; for a chain
 * record dims[3] = -(number of launches) and the other dims are those of its first launch. */
void egm_prof_enable(int level);
void egm_prof_reset(void);
int egm_prof_count(void);
int egm_prof_read(int i, float* ms, double* flops, int* dims);

/* ---- Graph Polynomial Fusion ------------------------------------------------------------
 * a, p [B,N,D]; coef [(P+1)*(Q+1)] = softplus(alpha) on the device; G [B,N,N].
 * Saved for backward: Ra, Rp [B,N,ldR] with ldR = egm_gpf_ldr(N); nrm_a, nrm_p [B,N]; optionally
 * xn_state (egm_gpf_state_bytes; opaque: the normalised tokens as GEMM operands). xn_state == NULL
 * in both calls selects the memory-saving mode: the backward re-normalises the tokens instead.
 * egm_gpf_fused_ok(N, D, P, Q, prec) != 0: egm_gpf_fwd called with xn_state == NULL runs as ONE fused
 * pass over the tokens (both Gram matrices on tcgen05 from fp32 tokens converted on the fly, cosine
 * scaling + polynomial + clamp in the epilogue; gpf_kernel.py:117-159). Then Ra == Rp == NULL is
 * allowed (forward-only call: nothing but G and the norms is written) and `ws` is unused. R_a / R_p
 * written by the fused pass are symmetric bit for bit; pass `symmetric | 2` to egm_gpf_bwd to let the
 * backward evaluate each (i,j)/(j,i) pair once. */
int egm_gpf_fused_ok(int N, int D, int P, int Q, int prec);
/* != 0: egm_gpf_fwd (fused, cosine similarity) may also be given xn_state; it then receives the RAW tokens
 * as GEMM operand planes (written by the threads that convert them for the Gram products), and the caller
 * passes `symmetric | 2 | 4` to egm_gpf_bwd, which folds the backward of F.normalize into E and needs
 * neither a re-normalisation pass nor a rownorm_bwd pass over [B,N,D]. */
int egm_gpf_raw_planes_ok(int N, int D, int P, int Q, int prec);
long long egm_gpf_ldr(int N);
size_t egm_gpf_state_bytes(int B, int N, int D, int prec);
size_t egm_gpf_fwd_workspace(int B, int N, int D, int prec);
int egm_gpf_fwd(const float* a, const float* p, const float* coef, int B, int N, int D, int P, int Q,
                int cosine, float eps, int symmetric, float* G, float* Ra, float* Rp, float* nrm_a,
                float* nrm_p, void* xn_state, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);
size_t egm_gpf_bwd_workspace(int B, int N, int D, int P, int Q, int prec);
int egm_gpf_bwd(const float* dG, const float* a, const float* p, const float* coef, const float* Ra,
                const float* Rp, const float* nrm_a, const float* nrm_p, const void* xn_state, int B,
                int N, int D, int P, int Q, int cosine, float eps, int symmetric, float* da, float* dp,
                float* dcoef, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);

/* ---- graph-weighted second-order pooling ----------------------------------------------------
 * Z [B,N,D] tokens, G [B,N,N] graph (any real matrix) -> M2 [B,D,D], optional u [B,D].
 * Saved: vecs [B, 4N+2] (s, deg, w, wdiag | t, sw), mu [B,D], state (egm_pool_state_bytes). */
size_t egm_pool_state_bytes(int B, int N, int D, int prec);
size_t egm_pool_fwd_workspace(int B, int N, int D, int prec);
int egm_pool_fwd(const float* Z, const float* G, int B, int N, int D, float eps, float* M2, float* u,
                 float* vecs, float* mu, void* state, int prec, void* ws, size_t ws_bytes,
                 egm_stream_t stream);
size_t egm_pool_bwd_workspace(int B, int N, int D, int prec);
int egm_pool_bwd(const float* dM2, const float* du, const float* Z, const float* G, const float* u,
                 const float* vecs, const float* mu, const void* state, int B, int N, int D, float eps,
                 float* dZ, float* dG, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);

/* ---- iSQRT-COV: trace pre-normalisation, Newton-Schulz, trace post-compensation -------------
 * M [B,D,D] -> O [B,D,D]. post_mode 0: O = Y_K / sqrt(tr+eps) (NewtonSchulzSqrtm);
 * post_mode 1: O = Y_K * sqrt(tr+eps) (utils.ops.matrix_sqrt_newton_schulz).
 * Saved: scal [3,B] (tr, 1/(tr+eps), post scale), state (egm_ns_state_bytes). */
size_t egm_ns_state_bytes(int B, int D, int iters, int prec);
size_t egm_ns_fwd_workspace(int B, int D, int iters, int prec);
int egm_ns_fwd(const float* M, int B, int D, int iters, float eps, int post_mode, float* O,
               float* scal, void* state, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);
size_t egm_ns_bwd_workspace(int B, int D, int iters, int prec);
int egm_ns_bwd(const float* dO, const float* O, const float* M, const float* scal, const void* state,
               int B, int D, int iters, float eps, int post_mode, float* dM, int prec, void* ws,
               size_t ws_bytes, egm_stream_t stream);

/* ---- fused pooling + iSQRT-COV in low-rank form (MomentHead when N < D) ----
 * Same function of (Z, G) as egm_pool_fwd -> egm_ns_fwd(post_mode 0), evaluated with every
 * Newton-Schulz product on N x N matrices (M2 = Zc^T W Zc has rank <= N): replaces
 * moment_head.py:279-296. iters >= 1. Saved: vecs, mu (as egm_pool_fwd), scal [5,B], state. */
size_t egm_mlr_state_bytes(int B, int N, int D, int iters, int prec);
size_t egm_mlr_fwd_workspace(int B, int N, int D, int iters, int prec);
/* Output: O [B,D,D] fp32, or (x_planes != NULL, tensor-core modes) the packed upper triangle of O
 * written as operand planes straight into the `state` of egm_linear_fwd (see egm_mhd_fwd).
 * Backward takes either (dO, O) or the packed gradient dv [B, D(D+1)/2] with dotOO = <dO,O> [B]. */
int egm_mlr_fwd(const float* Z, const float* G, int B, int N, int D, int iters, float eps, float* O,
                void* x_planes, float* u, float* vecs, float* mu, float* scal, void* state, int prec,
                void* ws, size_t ws_bytes, egm_stream_t stream);
size_t egm_mlr_bwd_workspace(int B, int N, int D, int iters, int prec);
int egm_mlr_bwd(const float* dO, const float* dv, const float* dotOO, const float* du, const float* Z,
                const float* G, const float* O, const float* u, const float* vecs, const float* mu,
                const float* scal, const void* state, int B, int N, int D, int iters, float eps,
                float* dZ, float* dG, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);

/* ---- fused dense moment head: pooling -> iSQRT-COV -> packed half-vector -----------------------
 * MomentHead.forward up to the input of second_net's Linear (moment_head.py:279-300), D x D
 * Newton-Schulz chain, every intermediate kept in the GEMM engine's operand format (no fp32
 * M2 / O / dO / dM round trips; trace(M2) and <dA,A> are by-products of GEMM epilogues).
 * x_planes: the operand region of egm_linear_fwd's `state` for (M=B, K=D(D+1)/2): the packed upper
 * triangle of iSQRT-COV(M2) is written there, then egm_linear_fwd is called with x == NULL.
 * Backward: dv [B, D(D+1)/2] = gradient of the half-vector (egm_linear_bwd's dx), dotO [B] =
 * <dv_b, v_b> (for y = v W^T + b this equals <dy_b, y_b - b>: egm_rowdot_bias).
 * iters >= 2, tensor-core precision modes only. Saved: vecs, mu (as egm_pool_fwd), scal [3,B], state.
 * flags: EGM_MHD_SYMMETRIC_GRAPH promises G == G^T exactly (the output of GraphPolynomialFusion with
 * symmetric_enforce, gpf_kernel.py:150-152). Every matrix of the Newton-Schulz chain is then symmetric:
 * only the upper 256 x 256 blocks of each product are evaluated and stored (6 of 9 tiles at D = 768),
 * and the backward runs as a symmetric forward tangent (the Frechet derivative of a matrix polynomial
 * at a symmetric point is self-adjoint). dZ is the same; dG is the SYMMETRIC PART (dG + dG^T)/2 of the
 * general path's gradient - the gradient with respect to a symmetric matrix, and all that a
 * symmetric-graph producer (gpf_kernel.py:150-152, whose backward symmetrises) consumes.
 * The same flags value must be passed to the matching backward. */
enum { EGM_MHD_SYMMETRIC_GRAPH = 1 };
size_t egm_mhd_state_bytes(int B, int N, int D, int iters, int prec);
size_t egm_mhd_fwd_workspace(int B, int N, int D, int iters, int prec);
int egm_mhd_fwd(const float* Z, const float* G, int B, int N, int D, int iters, float eps, int flags,
                void* x_planes, float* u, float* vecs, float* mu, float* scal, void* state, int prec,
                void* ws, size_t ws_bytes, egm_stream_t stream);
size_t egm_mhd_bwd_workspace(int B, int N, int D, int iters, int prec);
int egm_mhd_bwd(const float* dv, const float* dotO, const float* du, const float* Z, const float* G,
                const float* u, const float* vecs, const float* mu, const float* scal, const void* state,
                int B, int N, int D, int iters, float eps, int flags, float* dZ, float* dG, int prec,
                void* ws, size_t ws_bytes, egm_stream_t stream);
/* out[b] = sum_n dy[b,n] * (y[b,n] - bias[n])   (bias may be NULL) */
int egm_rowdot_bias(const float* dy, const float* y, const float* bias, int B, int n, float* out,
                    egm_stream_t stream);

/* ---- Linear layers of second_net / third_net (nn.Linear at moment_head.py:187,196) -------------
 * y [M,N] = x [M,K] W^T [N,K] + bias [N] on the tcgen05 engine with split-K; state keeps the
 * operand planes for the backward (dx = dy W, dW = dy^T x, dbias = colsum dy; any may be null).
 * x == NULL: the x operand planes were already written into `state` by egm_mhd_fwd / egm_mlr_fwd.
 * Tensor-core precision modes only (the strict fp32 mode leaves this GEMM to the caller). */
size_t egm_linear_state_bytes(int M, int N, int K, int prec);
size_t egm_linear_fwd_workspace(int M, int N, int K, int prec);
int egm_linear_fwd(const float* x, const float* W, const float* bias, int M, int N, int K, float* y,
                   void* state, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);
size_t egm_linear_bwd_workspace(int M, int N, int K, int prec);
int egm_linear_bwd(const float* dy, const void* state, int M, int N, int K, float* dx, float* dW,
                   float* dbias, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);

/* ---- half-vectorisation (row-major upper triangle incl. diagonal) -------------------------- */
int egm_triu_pack(const float* O, int B, int D, float* v, egm_stream_t stream);
int egm_triu_unpack(const float* dv, int B, int D, float* dO, egm_stream_t stream);

/* ---- third-order count-sketch product --------------------------------------------------------
 * x [B,D] -> out [B,S]; cs [3,B,S] saved. off [3,S+1], idx [3,D] (int32) and sgn [3,D] (fp32)
 * are the CSR inverse of the hash buffers; hash, sign [3,D] int64 are the state_dict buffers. */
int egm_sketch_fwd(const float* x, int B, int D, int S, const int* off, const int* idx,
                   const float* sgn, float* cs, float* out, egm_stream_t stream);
int egm_sketch_bwd(const float* dout, const float* cs, int B, int D, int S, const long long* hash,
                   const long long* sign, float* dx, egm_stream_t stream);

/* ---- graph alignment loss (EGOMomentCLEViT._graph_alignment_loss, ego_moment_clevit.py:278-316) ----
 * g[b] = mean(G[b]); loss = mean_ij (sigmoid(g_i g_j) - [labels_i == labels_j])^2 - the reference fills
 * the B x B matrix with a Python double loop (B^2 autograd in-place writes). G [B,N,N], labels [B] int64.
 * Outputs: g [B], dg [B] = d loss / d g (saved for the backward), rowloss [B] scratch, loss [1].
 * Backward: dG[b,:,:] = dloss[0] * dg[b] / N^2 (dloss: device scalar). */
int egm_align_fwd(const float* G, const long long* labels, int B, int N, float* g, float* dg,
                  float* rowloss, float* loss, egm_stream_t stream);
int egm_align_bwd(const float* dg, const float* dloss, int B, int N, float* dG, egm_stream_t stream);

/* ---- stand-alone matrix helpers (src/utils/ops.py) -------------------------------------------- */
size_t egm_gram_workspace(int B, int N, int D, int prec);
int egm_gram_fwd(const float* x, int B, int N, int D, int cosine, float eps, float* R, float* nrm,
                 int prec, void* ws, size_t ws_bytes, egm_stream_t stream);
int egm_gram_bwd(const float* dR, const float* x, const float* nrm, int B, int N, int D, int cosine,
                 float eps, float* dx, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);
/* method 0: D^-1/2 G D^-1/2 ; method 1: D^-1 G.  deg [B,N] receives the clamped degrees. */
int egm_normalize_graph(const float* G, int B, int N, int method, float eps, float* out, float* deg,
                        egm_stream_t stream);
int egm_batch_trace(const float* M, int B, int D, float* tr, egm_stream_t stream);

/* ---- feature-net tail: BatchNorm1d -> GELU -> Dropout ---------------------------------------
 * The layers that follow the Linear of second_net / third_net (moment_head.py:186-191, 195-200) in one
 * launch each way. y [M,N] = the Linear's output. training != 0: batch statistics (biased variance),
 * running_mean / running_var (optional) updated in place with `momentum` and the unbiased variance
 * exactly like nn.BatchNorm1d, inverted dropout with keep-mask = hash(seed, element) >= drop_p (the
 * same seed must be passed to the backward; the mask is NOT torch's Philox stream). training == 0:
 * running statistics, no dropout. GELU is the exact erf form. save_mean / save_rstd [N] are outputs. */
int egm_feature_tail_fwd(const float* y, int M, int N, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, int training, float momentum, float bn_eps,
                         float drop_p, unsigned long long seed, float* out, float* save_mean, float* save_rstd,
                         egm_stream_t stream);
int egm_feature_tail_bwd(const float* dout, const float* y, const float* gamma, const float* beta,
                         const float* save_mean, const float* save_rstd, int M, int N, int training,
                         float drop_p, unsigned long long seed, float* dy, float* dgamma, float* dbeta,
                         egm_stream_t stream);

/* Plain batched product C = alpha * op(A) op(B) through the active engine (used by the
 * native tests and the benchmark's tensor-pipe probe). A [B, M|K, K|M], B [B, K|N, N|K]. */
size_t egm_bmm_workspace(int B, int M, int N, int K, int prec);
int egm_bmm(const float* A, int transA, const float* Bm, int transB, int B, int M, int N, int K,
            float alpha, float* C, int prec, void* ws, size_t ws_bytes, egm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EGM_B200_H_ */
