// Internal description of one batched GEMM-with-epilogue, shared by the two CUDA back
// ends of the library:
//   * egm_gemm_tc.cu   - tcgen05/TMEM tensor-core engine fed by TMA (bf16 hi/lo planes)
//   * egm_gemm_simt.cu - fp32 FFMA engine for shapes TMA cannot address and for the
//                        strict-fp32 mode
//
//   C[b] = alpha * alpha_b[b] * sum_t op(A_t[b]) * op(B_t[b])  +  beta_eye * beta_b[b] * I
//          +  gamma * gamma_b[b] * E[b]
//
// Every dense contraction on the moment-pooling path (Gram matrices, W*Zc, Zc^T*U, the
// Newton-Schulz chain and all of their backward products) is an instance of this.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace egm {

// Precision modes of the path (public: the `precision_mode` argument of the C ABI).
enum Precision : int {
  PREC_FP32_SIMT = 0,  // fp32 FFMA on CUDA cores (exact fp32 products)
  PREC_BF16X3 = 1,     // fp32 emulated on tcgen05: x = hi + lo (bf16 each), 3 MMAs per product
  PREC_BF16 = 2,       // single bf16 MMA, fp32 accumulate
};

// Batched row-major matrix view [batch][rows][cols].
// fp32 storage : p0 = float*,  p1 unused.
// plane storage: p0 = bf16 "hi" plane, p1 = bf16 "lo" plane (x ~= hi + lo), same ld/bstride.
struct Mat {
  void* p0 = nullptr;
  void* p1 = nullptr;
  int rows = 0, cols = 0;
  long long ld = 0;       // elements between consecutive rows
  long long bstride = 0;  // elements between consecutive batch items
};

struct GemmTerm {
  Mat A;       // op(A) is M x K : stored [M,K] (transA = 0) or [K,M] (transA = 1)
  int transA = 0;
  Mat B;       // op(B) is K x N : stored [K,N] (transB = 0) or [N,K] (transB = 1)
  int transB = 0;
  int K = 0;
  // tcgen05 engine only: the operand is a SYMMETRIC square matrix of which only the upper
  // 256 x 256 blocks (block column >= block row, diagonal blocks complete) are stored - the
  // layout a `sym_out` product writes. Reads that fall into an absent block are redirected to
  // the mirrored block with the operand's major-ness flipped in the UMMA descriptor, so no
  // transpose pass and no mirrored store ever happens. transA / transB are ignored.
  int symA = 0;
  int symB = 0;
};

struct GemmProblem {
  int M = 0, N = 0, batch = 1;
  int nterms = 1;
  int split_k = 1;   // tcgen05 engine only: slice K into split_k partial products, written to
                     // Cf[s] (s = 0..split_k-1, batch stride Cf.bstride); needs batch == 1
  GemmTerm t[2];
  // epilogue
  float alpha = 1.f;
  const float* alpha_b = nullptr;  // optional per-batch multiplier (device pointer)
  float beta_eye = 0.f;
  const float* beta_b = nullptr;   // optional per-batch multiplier of beta_eye
  float gamma = 0.f;
  const float* gamma_b = nullptr;  // optional per-batch multiplier of gamma
  Mat E;             // optional addend, storage given by e_planes
  int e_planes = 0;
  Mat Cp;            // optional output as bf16 planes (p0 == nullptr: absent)
  Mat Cf;            // optional output as fp32       (p0 == nullptr: absent)
  // ---- tcgen05 engine only -------------------------------------------------------------
  // second plane output of the same accumulator: Cp2 = c2_scale * C + c2_eye * I
  // (e.g. A = M/tau and T_0 = 1.5 I - 0.5 A from one product)
  Mat Cp2;
  float c2_scale = 0.f, c2_eye = 0.f;
  // dot_out[b] = <C[b], F[b]> (deterministic: per-warp partials in dot_ws, then a fixed-order
  // reduction). dot_ws needs gemm_tc_dot_ws_floats(g) floats.
  Mat F;
  int f_planes = 0;
  float* dot_out = nullptr;
  float* dot_ws = nullptr;
  // packed upper triangle (row-major, diagonal included) of the square result as bf16 planes:
  // X.p0/p1[b * X.ld + i*N - i(i-1)/2 + j - i] = C[b][i][j], j >= i. Tiles strictly below the
  // diagonal are not computed at all. Excludes Cp / Cp2.
  Mat X;
  // the result is symmetric (M == N): only tiles that touch the upper 256 x 256 blocks are
  // computed and stored (6 of 9 at D = 768); consumers read it back through symA / symB.
  // With dot_out, off-diagonal blocks count twice. E and F are read in the same upper blocks.
  int sym_out = 0;
};

// floats of workspace `dot_ws` must provide for problem g (0 when no dot is requested)
size_t gemm_tc_dot_ws_floats(const GemmProblem& g);

// Returns cudaSuccess or the launch error. No allocation, no synchronisation.
cudaError_t gemm_tc(const GemmProblem& g, int npass /*1 or 3*/, cudaStream_t stream);
cudaError_t gemm_simt(const GemmProblem& g, cudaStream_t stream);  // all Mats fp32
// True when every operand of `g` satisfies the TMA addressing rules of the tcgen05 engine.
bool gemm_tc_supported(const GemmProblem& g, int npass);

const char* last_error();
void note_launch();                 // every kernel launch of the library calls this
unsigned long long launch_count();
void set_error(const char* fmt, ...);

// optional per-launch timing of the tcgen05 engine (egm_error.cu; C ABI egm_prof_*)
bool prof_enabled();
int prof_level();                   // 0 off, 1 every launch, 2 chain groups only
void prof_enable(int level);
// level 2: one record for all engine launches between begin and end (same thread, same stream)
bool prof_group_open();
void prof_group_begin(cudaStream_t st);
void prof_group_note(double flops, const int dims[6]);
void prof_group_end(cudaStream_t st);
void prof_reset();
int prof_begin(cudaStream_t st, double flops, const int dims[6]);   // dims: M, N, K0, K1, batch, passes
void prof_end(int id, cudaStream_t st);
int prof_count();
int prof_read(int i, float* ms, double* flops, int* dims);

}  // namespace egm
