// fp32 FFMA batched GEMM with the same contract as the tcgen05 engine (egm_gemm.h).
// Used (a) in the strict-fp32 precision mode, and (b) for operands TMA cannot address
// (rows that are not 16-byte aligned, e.g. a caller-supplied [B,197,197] fp32 graph).
// Plain shared-memory tiling: 64x64 output tile, K step 16, 4x4 outputs per thread.
#include <cuda_runtime.h>

#include "egm_gemm.h"

namespace egm {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtParams {
  const float* A[2];
  const float* B[2];
  long long ldA[2], bsA[2], ldB[2], bsB[2];
  int tA[2], tB[2], K[2];
  int nterms, M, N, batch;
  float alpha, beta_eye, gamma;
  const float* alpha_b;
  const float* beta_b;
  const float* gamma_b;
  const float* E;
  long long ldE, bsE;
  float* C;
  long long ldC, bsC;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int t = 0; t < p.nterms; ++t) {
    const float* A = p.A[t] + (long long)b * p.bsA[t];
    const float* B = p.B[t] + (long long)b * p.bsB[t];
    const long long ldA = p.ldA[t], ldB = p.ldB[t];
    const int K = p.K[t];
    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + i * 256;
        int m, k;
        if (p.tA[t]) { m = e & 63; k = e >> 6; } else { k = e & 15; m = e >> 4; }
        const int gm = m0 + m, gk = k0 + k;
        float v = 0.f;
        if (gm < p.M && gk < K) v = p.tA[t] ? A[(long long)gk * ldA + gm] : A[(long long)gm * ldA + gk];
        As[k][m] = v;
        int n, k2;
        if (p.tB[t]) { k2 = e & 15; n = e >> 4; } else { n = e & 63; k2 = e >> 6; }
        const int gn = n0 + n, gk2 = k0 + k2;
        float w = 0.f;
        if (gn < p.N && gk2 < K) w = p.tB[t] ? B[(long long)gn * ldB + gk2] : B[(long long)gk2 * ldB + gn];
        Bs[k2][n] = w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        float a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  const float a_eff = p.alpha * (p.alpha_b ? p.alpha_b[b] : 1.f);
  const float beta = p.beta_eye * (p.beta_b ? p.beta_b[b] : 1.f);
  const float gamma = p.gamma * (p.gamma_b ? p.gamma_b[b] : 1.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float o = a_eff * acc[i][j];
      if (m == n) o += beta;
      if (p.E) o += gamma * p.E[(long long)b * p.bsE + (long long)m * p.ldE + n];
      p.C[(long long)b * p.bsC + (long long)m * p.ldC + n] = o;
    }
  }
}

}  // namespace

cudaError_t gemm_simt(const GemmProblem& g, cudaStream_t stream) {
  if (!g.Cf.p0 || g.nterms < 1 || g.nterms > 2) {
    set_error("gemm_simt: needs an fp32 output and 1..2 terms");
    return cudaErrorInvalidValue;
  }
  if (g.sym_out || g.t[0].symA || g.t[0].symB || (g.nterms > 1 && (g.t[1].symA || g.t[1].symB))) {
    set_error("gemm_simt: symmetric block storage is a tcgen05-engine format");
    return cudaErrorInvalidValue;
  }
  SimtParams p = {};
  for (int t = 0; t < g.nterms; ++t) {
    p.A[t] = static_cast<const float*>(g.t[t].A.p0);
    p.B[t] = static_cast<const float*>(g.t[t].B.p0);
    p.ldA[t] = g.t[t].A.ld; p.bsA[t] = g.t[t].A.bstride;
    p.ldB[t] = g.t[t].B.ld; p.bsB[t] = g.t[t].B.bstride;
    p.tA[t] = g.t[t].transA; p.tB[t] = g.t[t].transB; p.K[t] = g.t[t].K;
  }
  p.nterms = g.nterms; p.M = g.M; p.N = g.N; p.batch = g.batch;
  p.alpha = g.alpha; p.beta_eye = g.beta_eye; p.gamma = g.gamma; p.alpha_b = g.alpha_b;
  p.beta_b = g.beta_b; p.gamma_b = g.gamma_b;
  if (g.E.p0 && g.gamma != 0.f) {
    if (g.e_planes) {
      set_error("gemm_simt: addend must be fp32");
      return cudaErrorInvalidValue;
    }
    p.E = static_cast<const float*>(g.E.p0); p.ldE = g.E.ld; p.bsE = g.E.bstride;
  }
  p.C = static_cast<float*>(g.Cf.p0); p.ldC = g.Cf.ld; p.bsC = g.Cf.bstride;
  dim3 grid((g.N + TN - 1) / TN, (g.M + TM - 1) / TM, g.batch);
  if (grid.z > 65535) {
    set_error("gemm_simt: batch %d exceeds grid.z", g.batch);
    return cudaErrorInvalidValue;
  }
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
  note_launch();
  return cudaGetLastError();
}

}  // namespace egm
