// Stand-alone correctness + throughput check of the tcgen05 GEMM engine (egm_gemm_tc.cu)
// against a double-precision host product. Runs on the GPU box only:
//   make -C tests/native && tests/native/test_gemm_tc [quick]
// Every case prints max|err|/max|ref|; the process exits non-zero if any case fails.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../ego-moment-cle-vit_b200/csrc/egm_gemm.h"

using egm::GemmProblem;
using egm::Mat;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, \
             egm::last_error());                                                    \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

static uint32_t g_seed = 12345;
static float frand() {
  g_seed = g_seed * 1664525u + 1013904223u;
  return ((g_seed >> 8) & 0xFFFF) / 32768.0f - 1.0f;
}

struct HostMat {
  int rows, cols, batch;
  long long ld, bs;
  std::vector<float> f;               // the fp32 values
  std::vector<__nv_bfloat16> hi, lo;  // their planes
  __nv_bfloat16 *d_hi = nullptr, *d_lo = nullptr;
  float* d_f = nullptr;
  void init(int r, int c, int b, int pad, bool random) {
    rows = r; cols = c; batch = b;
    ld = ((c + 7) / 8) * 8 + pad;
    bs = (long long)rows * ld + 8 * pad;
    f.assign((size_t)bs * batch, 0.f);
    hi.resize(f.size()); lo.resize(f.size());
    for (int bb = 0; bb < batch; ++bb)
      for (int i = 0; i < rows; ++i)
        for (long long j = 0; j < ld; ++j) {
          // padding columns get garbage on purpose: TMA bounds must hide them
          float x = random ? frand() : 0.f;
          if (j >= cols) x = 1000.f;
          f[(size_t)bb * bs + i * ld + j] = x;
        }
    for (size_t i = 0; i < f.size(); ++i) {
      hi[i] = __float2bfloat16_rn(f[i]);
      lo[i] = __float2bfloat16_rn(f[i] - __bfloat162float(hi[i]));
    }
    CK(cudaMalloc(&d_hi, hi.size() * 2));
    CK(cudaMalloc(&d_lo, lo.size() * 2));
    CK(cudaMalloc(&d_f, f.size() * 4));
    CK(cudaMemcpy(d_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_f, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
  }
  Mat planes() const {
    Mat m; m.p0 = d_hi; m.p1 = d_lo; m.rows = rows; m.cols = cols; m.ld = ld; m.bstride = bs;
    return m;
  }
  Mat f32() const {
    Mat m; m.p0 = d_f; m.rows = rows; m.cols = cols; m.ld = ld; m.bstride = bs;
    return m;
  }
  double val(int b, int i, int j, int npass) const {
    size_t o = (size_t)b * bs + (size_t)i * ld + j;
    double v = __bfloat162float(hi[o]);
    if (npass == 3) v += __bfloat162float(lo[o]);
    return v;
  }
  void free_() { cudaFree(d_hi); cudaFree(d_lo); cudaFree(d_f); }
};

struct Case {
  const char* name;
  int M, N, batch, nterms;
  int K[2], tA[2], tB[2];
  int npass;
  float alpha, beta, gamma;
  int e_mode;      // 0 none 1 planes 2 f32
  int use_alpha_b;
  int out_planes, out_f32;
  int pad;         // extra leading-dimension padding (elements, multiple of 8)
  int f32_unaligned;  // fp32 output with ld == N (exercises the scalar store path)
  int sample;      // verify only `sample` random entries per image (0 = all)
  int c2;          // also request the secondary plane output  -0.5*C + 1.5*I
  int dot;         // also request <C, F> per image (F: planes)
  int triu;        // output = packed upper triangle planes (excludes out_planes)
};

static int run_case(const Case& c) {
  HostMat A[2], B[2], E;
  GemmProblem g;
  g.M = c.M; g.N = c.N; g.batch = c.batch; g.nterms = c.nterms;
  for (int t = 0; t < c.nterms; ++t) {
    if (c.tA[t]) A[t].init(c.K[t], c.M, c.batch, c.pad, true); else A[t].init(c.M, c.K[t], c.batch, c.pad, true);
    if (c.tB[t]) B[t].init(c.N, c.K[t], c.batch, c.pad, true); else B[t].init(c.K[t], c.N, c.batch, c.pad, true);
    g.t[t].A = A[t].planes(); g.t[t].transA = c.tA[t];
    g.t[t].B = B[t].planes(); g.t[t].transB = c.tB[t];
    g.t[t].K = c.K[t];
  }
  g.alpha = c.alpha; g.beta_eye = c.beta; g.gamma = c.gamma;
  std::vector<float> alpha_b(c.batch, 1.f);
  float* d_alpha_b = nullptr;
  if (c.use_alpha_b) {
    for (int b = 0; b < c.batch; ++b) alpha_b[b] = 0.5f + 0.25f * b;
    CK(cudaMalloc(&d_alpha_b, c.batch * 4));
    CK(cudaMemcpy(d_alpha_b, alpha_b.data(), c.batch * 4, cudaMemcpyHostToDevice));
    g.alpha_b = d_alpha_b;
  }
  if (c.e_mode) {
    E.init(c.M, c.N, c.batch, c.pad, true);
    g.E = c.e_mode == 1 ? E.planes() : E.f32();
    g.e_planes = c.e_mode == 1;
    if (c.e_mode == 1 && c.npass != 3) g.E.p1 = nullptr;   // single-pass mode carries no lo plane
  }
  // outputs
  const long long ldp = ((c.N + 7) / 8) * 8 + c.pad;
  const long long bsp = (long long)c.M * ldp;
  const long long ldf = c.f32_unaligned ? c.N : ((c.N + 3) / 4) * 4 + c.pad;
  const long long bsf = (long long)c.M * ldf;
  __nv_bfloat16 *d_ch = nullptr, *d_cl = nullptr; float* d_cf = nullptr;
  if (c.out_planes) {
    CK(cudaMalloc(&d_ch, bsp * c.batch * 2)); CK(cudaMalloc(&d_cl, bsp * c.batch * 2));
    CK(cudaMemset(d_ch, 0xFF, bsp * c.batch * 2)); CK(cudaMemset(d_cl, 0xFF, bsp * c.batch * 2));
    g.Cp.p0 = d_ch; g.Cp.p1 = (c.npass == 3) ? d_cl : nullptr; g.Cp.rows = c.M; g.Cp.cols = c.N;
    g.Cp.ld = ldp; g.Cp.bstride = bsp;
  }
  if (c.out_f32) {
    CK(cudaMalloc(&d_cf, bsf * c.batch * 4));
    CK(cudaMemset(d_cf, 0xFF, bsf * c.batch * 4));
    g.Cf.p0 = d_cf; g.Cf.rows = c.M; g.Cf.cols = c.N; g.Cf.ld = ldf; g.Cf.bstride = bsf;
  }
  __nv_bfloat16 *d_c2h = nullptr, *d_c2l = nullptr;
  if (c.c2) {
    CK(cudaMalloc(&d_c2h, bsp * c.batch * 2)); CK(cudaMalloc(&d_c2l, bsp * c.batch * 2));
    CK(cudaMemset(d_c2h, 0xFF, bsp * c.batch * 2)); CK(cudaMemset(d_c2l, 0xFF, bsp * c.batch * 2));
    g.Cp2.p0 = d_c2h; g.Cp2.p1 = (c.npass == 3) ? d_c2l : nullptr; g.Cp2.rows = c.M; g.Cp2.cols = c.N;
    g.Cp2.ld = ldp; g.Cp2.bstride = bsp; g.c2_scale = -0.5f; g.c2_eye = 1.5f;
  }
  HostMat F;
  float *d_dot = nullptr, *d_dotws = nullptr;
  if (c.dot) {
    F.init(c.M, c.N, c.batch, c.pad, true);
    g.F = F.planes(); g.f_planes = 1;
    if (c.npass != 3) g.F.p1 = nullptr;
    CK(cudaMalloc(&d_dot, c.batch * 4));
    g.dot_out = d_dot;
    CK(cudaMalloc(&d_dotws, egm::gemm_tc_dot_ws_floats(g) * 4 + 16));
    g.dot_ws = d_dotws;
  }
  const long long Lx = (long long)c.N * (c.N + 1) / 2, ldx = (Lx + 7) / 8 * 8;
  __nv_bfloat16 *d_xh = nullptr, *d_xl = nullptr;
  if (c.triu) {
    CK(cudaMalloc(&d_xh, ldx * c.batch * 2)); CK(cudaMalloc(&d_xl, ldx * c.batch * 2));
    CK(cudaMemset(d_xh, 0xFF, ldx * c.batch * 2)); CK(cudaMemset(d_xl, 0xFF, ldx * c.batch * 2));
    g.X.p0 = d_xh; g.X.p1 = (c.npass == 3) ? d_xl : nullptr; g.X.ld = ldx;
  }
  cudaError_t le = egm::gemm_tc(g, c.npass, 0);
  if (le != cudaSuccess) {
    printf("[FAIL] %-44s launch: %s (%s)\n", c.name, cudaGetErrorString(le), egm::last_error());
    return 1;
  }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) {
    printf("[FAIL] %-44s kernel: %s\n", c.name, cudaGetErrorString(se));
    exit(3);  // context is gone
  }
  std::vector<__nv_bfloat16> ch, cl; std::vector<float> cf;
  if (c.out_planes) {
    ch.resize(bsp * c.batch); cl.resize(bsp * c.batch);
    CK(cudaMemcpy(ch.data(), d_ch, ch.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cl.data(), d_cl, cl.size() * 2, cudaMemcpyDeviceToHost));
  }
  if (c.out_f32) { cf.resize(bsf * c.batch); CK(cudaMemcpy(cf.data(), d_cf, cf.size() * 4, cudaMemcpyDeviceToHost)); }

  std::vector<__nv_bfloat16> c2h, c2l, xh, xl;
  std::vector<float> dots;
  if (c.c2) {
    c2h.resize(bsp * c.batch); c2l.resize(bsp * c.batch);
    CK(cudaMemcpy(c2h.data(), d_c2h, c2h.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c2l.data(), d_c2l, c2l.size() * 2, cudaMemcpyDeviceToHost));
  }
  if (c.triu) {
    xh.resize(ldx * c.batch); xl.resize(ldx * c.batch);
    CK(cudaMemcpy(xh.data(), d_xh, xh.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(xl.data(), d_xl, xl.size() * 2, cudaMemcpyDeviceToHost));
  }
  if (c.dot) { dots.resize(c.batch); CK(cudaMemcpy(dots.data(), d_dot, c.batch * 4, cudaMemcpyDeviceToHost)); }
  double max_err_c2 = 0, max_err_x = 0, max_err_dot = 0;
  std::vector<double> dot_ref(c.batch, 0.0), dot_abs(c.batch, 0.0);
  double max_ref = 0, max_err_f = 0, max_err_p = 0;
  int wb = 0, wm = 0, wn = 0; double wgot = 0, wref = 0;
  uint32_t s = 777;
  for (int b = 0; b < c.batch; ++b) {
    long long count = c.sample ? c.sample : (long long)c.M * c.N;
    for (long long idx = 0; idx < count; ++idx) {
      int m, n;
      if (c.sample) { s = s * 1664525u + 1013904223u; m = (s >> 8) % c.M; s = s * 1664525u + 1013904223u; n = (s >> 8) % c.N;
                      if (idx == 0) { m = c.M - 1; n = c.N - 1; } if (idx == 1) { m = 0; n = 0; } }
      else { m = (int)(idx / c.N); n = (int)(idx % c.N); }
      double acc = 0;
      for (int t = 0; t < c.nterms; ++t)
        for (int k = 0; k < c.K[t]; ++k) {
          double a = c.tA[t] ? A[t].val(b, k, m, c.npass) : A[t].val(b, m, k, c.npass);
          double bb = c.tB[t] ? B[t].val(b, n, k, c.npass) : B[t].val(b, k, n, c.npass);
          acc += a * bb;
        }
      double ref = (double)c.alpha * alpha_b[b] * acc + (m == n ? c.beta : 0.0);
      if (c.e_mode == 1) ref += c.gamma * E.val(b, m, n, c.npass == 3 ? 3 : 1);
      if (c.e_mode == 2) ref += c.gamma * (double)E.f[(size_t)b * E.bs + (size_t)m * E.ld + n];
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
      if (c.c2) {
        size_t o = (size_t)b * bsp + (size_t)m * ldp + n;
        double got = __bfloat162float(c2h[o]);
        if (c.npass == 3) got += __bfloat162float(c2l[o]);
        double e = fabs(got - (-0.5 * ref + (m == n ? 1.5 : 0.0)));
        if (e > max_err_c2) max_err_c2 = e;
      }
      if (c.dot) {
        double f = F.val(b, m, n, c.npass == 3 ? 3 : 1);
        dot_ref[b] += ref * f; dot_abs[b] += fabs(ref * f);
      }
      if (c.triu && n >= m) {
        size_t o = (size_t)b * ldx + (size_t)m * c.N - (size_t)m * (m - 1) / 2 + (n - m);
        double got = __bfloat162float(xh[o]);
        if (c.npass == 3) got += __bfloat162float(xl[o]);
        double e = fabs(got - ref);
        if (!(e <= max_err_x)) { max_err_x = e; wb = b; wm = m; wn = n; wgot = got; wref = ref; }
      }
      if (c.out_f32) {
        double got = cf[(size_t)b * bsf + (size_t)m * ldf + n];
        double e = fabs(got - ref);
        if (!(e <= max_err_f)) { max_err_f = e; wb = b; wm = m; wn = n; wgot = got; wref = ref; }
      }
      if (c.out_planes) {
        size_t o = (size_t)b * bsp + (size_t)m * ldp + n;
        double got = __bfloat162float(ch[o]);
        if (c.npass == 3) got += __bfloat162float(cl[o]);
        double e = fabs(got - ref);
        if (!(e <= max_err_p)) { max_err_p = e; if (!c.out_f32) { wb = b; wm = m; wn = n; wgot = got; wref = ref; } }
      }
    }
  }
  const double tol_f = (c.npass == 3) ? 5e-5 : 2e-5;
  const double tol_p = (c.npass == 3) ? 8e-5 : 6e-3;  // single plane output: bf16 rounding
  const double rf = max_err_f / (max_ref + 1e-30), rp = max_err_p / (max_ref + 1e-30);
  if (c.dot)
    for (int b = 0; b < c.batch; ++b) {
      double e = fabs(dots[b] - dot_ref[b]) / (dot_abs[b] + 1e-30);
      if (!(e <= max_err_dot)) max_err_dot = e;
    }
  const double rc2 = max_err_c2 / (max_ref + 1e-30), rx = max_err_x / (max_ref + 1e-30);
  // padding of the packed rows must stay untouched (0xFFFF) and nothing may be written twice wrong
  bool pad_ok = true;
  if (c.triu)
    for (int b = 0; b < c.batch && pad_ok; ++b)
      for (long long j = Lx; j < ldx; ++j)
        if (__bfloat16_as_ushort(xh[(size_t)b * ldx + j]) != 0xFFFF) pad_ok = false;
  const bool ok = (!c.out_f32 || rf < tol_f) && (!c.out_planes || rp < tol_p) && max_ref > 0 &&
                  (!c.c2 || rc2 < tol_p) && (!c.triu || (rx < tol_p && pad_ok)) &&
                  (!c.dot || max_err_dot < (c.sample ? 1e30 : 2e-5));
  printf("[%s] %-44s rel_err f32=%.2e planes=%.2e (max|ref|=%.3g)", ok ? " ok " : "FAIL", c.name, rf, rp, max_ref);
  if (c.c2) printf(" c2=%.2e", rc2);
  if (c.triu) printf(" triu=%.2e%s", rx, pad_ok ? "" : " PAD-CLOBBERED");
  if (c.dot) printf(" dot=%.2e", max_err_dot);
  if (!ok) printf("  worst b=%d m=%d n=%d got=%.6g ref=%.6g", wb, wm, wn, wgot, wref);
  printf("\n");
  fflush(stdout);
  for (int t = 0; t < c.nterms; ++t) { A[t].free_(); B[t].free_(); }
  if (c.e_mode) E.free_();
  cudaFree(d_ch); cudaFree(d_cl); cudaFree(d_cf); cudaFree(d_alpha_b);
  cudaFree(d_c2h); cudaFree(d_c2l); cudaFree(d_xh); cudaFree(d_xl); cudaFree(d_dot); cudaFree(d_dotws);
  if (c.dot) F.free_();
  return ok ? 0 : 1;
}


// Symmetric block storage: operands are symmetric D x D matrices of which only the upper
// 256 x 256 blocks are valid on the device (the absent blocks hold NaN); the product is
// requested with sym_out, so only the upper blocks of C = S0 S1 (+ S2 S3) may be written.
static int run_sym_case(const char* name, int D, int batch, int nterms, int npass, int with_e, int with_dot) {
  HostMat S[4], E, F;
  auto symmetrise = [&](HostMat& m, bool poison) {
    for (int b = 0; b < m.batch; ++b)
      for (int i = 0; i < D; ++i)
        for (int j = 0; j < i; ++j) m.f[(size_t)b * m.bs + (size_t)i * m.ld + j] = m.f[(size_t)b * m.bs + (size_t)j * m.ld + i];
    for (size_t i = 0; i < m.f.size(); ++i) {
      m.hi[i] = __float2bfloat16_rn(m.f[i]);
      m.lo[i] = __float2bfloat16_rn(m.f[i] - __bfloat162float(m.hi[i]));
    }
    std::vector<__nv_bfloat16> dh = m.hi, dl = m.lo;
    if (poison)
      for (int b = 0; b < m.batch; ++b)
        for (int i = 0; i < D; ++i)
          for (int j = 0; j < D; ++j)
            if ((j >> 8) < (i >> 8)) {
              dh[(size_t)b * m.bs + (size_t)i * m.ld + j] = __ushort_as_bfloat16(0x7FC0);
              dl[(size_t)b * m.bs + (size_t)i * m.ld + j] = __ushort_as_bfloat16(0x7FC0);
            }
    CK(cudaMemcpy(m.d_hi, dh.data(), dh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m.d_lo, dl.data(), dl.size() * 2, cudaMemcpyHostToDevice));
  };
  GemmProblem g;
  g.M = D; g.N = D; g.batch = batch; g.nterms = nterms; g.sym_out = 1;
  for (int t = 0; t < 2 * nterms; ++t) { S[t].init(D, D, batch, 0, true); symmetrise(S[t], true); }
  for (int t = 0; t < nterms; ++t) {
    g.t[t].A = S[2 * t].planes(); g.t[t].symA = 1;
    g.t[t].B = S[2 * t + 1].planes(); g.t[t].symB = 1;
    if (npass != 3) { g.t[t].A.p1 = nullptr; g.t[t].B.p1 = nullptr; }
    g.t[t].K = D;
  }
  g.alpha = -0.5f; g.beta_eye = 1.5f;
  if (with_e) {
    E.init(D, D, batch, 0, true); symmetrise(E, true);
    g.E = E.planes(); g.e_planes = 1; g.gamma = -3.f;
    if (npass != 3) g.E.p1 = nullptr;
  }
  const long long ldp = ((D + 7) / 8) * 8, bsp = (long long)D * ldp;
  __nv_bfloat16 *d_ch, *d_cl;
  CK(cudaMalloc(&d_ch, bsp * batch * 2)); CK(cudaMalloc(&d_cl, bsp * batch * 2));
  CK(cudaMemset(d_ch, 0xFF, bsp * batch * 2)); CK(cudaMemset(d_cl, 0xFF, bsp * batch * 2));
  g.Cp.p0 = d_ch; g.Cp.p1 = npass == 3 ? d_cl : nullptr; g.Cp.rows = D; g.Cp.cols = D; g.Cp.ld = ldp; g.Cp.bstride = bsp;
  float *d_dot = nullptr, *d_dotws = nullptr;
  if (with_dot) {
    F.init(D, D, batch, 0, true); symmetrise(F, true);
    g.F = F.planes(); g.f_planes = 1;
    if (npass != 3) g.F.p1 = nullptr;
    CK(cudaMalloc(&d_dot, batch * 4));
    g.dot_out = d_dot;
    CK(cudaMalloc(&d_dotws, egm::gemm_tc_dot_ws_floats(g) * 4 + 16));
    g.dot_ws = d_dotws;
  }
  cudaError_t le = egm::gemm_tc(g, npass, 0);
  if (le != cudaSuccess) {
    printf("[FAIL] %-44s launch: %s (%s)\n", name, cudaGetErrorString(le), egm::last_error());
    return 1;
  }
  cudaError_t se = cudaDeviceSynchronize();
  if (se != cudaSuccess) { printf("[FAIL] %-44s kernel: %s\n", name, cudaGetErrorString(se)); exit(3); }
  std::vector<__nv_bfloat16> ch(bsp * batch), cl(bsp * batch);
  CK(cudaMemcpy(ch.data(), d_ch, ch.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cl.data(), d_cl, cl.size() * 2, cudaMemcpyDeviceToHost));
  std::vector<float> dots(batch, 0.f);
  if (with_dot) CK(cudaMemcpy(dots.data(), d_dot, batch * 4, cudaMemcpyDeviceToHost));
  double max_ref = 0, max_err = 0, max_err_dot = 0;
  bool untouched = true;
  for (int b = 0; b < batch; ++b) {
    double dref = 0, dabs = 0;
    for (int m = 0; m < D; ++m)
      for (int n = 0; n < D; ++n) {
        const size_t o = (size_t)b * bsp + (size_t)m * ldp + n;
        if ((n >> 8) < (m >> 8)) {   // absent block: must not have been written
          if (__bfloat16_as_ushort(ch[o]) != 0xFFFF) untouched = false;
          continue;
        }
        double acc = 0;
        for (int t = 0; t < nterms; ++t)
          for (int k = 0; k < D; ++k) acc += S[2 * t].val(b, m, k, npass) * S[2 * t + 1].val(b, k, n, npass);
        double ref = -0.5 * acc + (m == n ? 1.5 : 0.0);
        if (with_e) ref += -3.0 * E.val(b, m, n, npass);
        if (fabs(ref) > max_ref) max_ref = fabs(ref);
        double got = __bfloat162float(ch[o]);
        if (npass == 3) got += __bfloat162float(cl[o]);
        if (!(fabs(got - ref) <= max_err)) max_err = fabs(got - ref);
        if (with_dot) {
          const double w = ((n >> 8) > (m >> 8)) ? 2.0 : 1.0;
          dref += w * ref * F.val(b, m, n, npass); dabs += fabs(w * ref * F.val(b, m, n, npass));
        }
      }
    if (with_dot) { const double e = fabs(dots[b] - dref) / (dabs + 1e-30); if (!(e <= max_err_dot)) max_err_dot = e; }
  }
  const double tol = (npass == 3) ? 8e-5 : 6e-3;
  const double r = max_err / (max_ref + 1e-30);
  const bool ok = r < tol && untouched && max_ref > 0 && (!with_dot || max_err_dot < 2e-5);
  printf("[%s] %-44s rel_err planes=%.2e (max|ref|=%.3g)%s", ok ? " ok " : "FAIL", name, r, max_ref,
         untouched ? "" : " LOWER-BLOCK-WRITTEN");
  if (with_dot) printf(" dot=%.2e", max_err_dot);
  printf("\n");
  fflush(stdout);
  for (int t = 0; t < 2 * nterms; ++t) S[t].free_();
  if (with_e) E.free_();
  if (with_dot) F.free_();
  cudaFree(d_ch); cudaFree(d_cl); cudaFree(d_dot); cudaFree(d_dotws);
  return ok ? 0 : 1;
}

static void bench(int D, int batch, int npass, int tA, int tB, int sym = 0, int nterms = 1) {
  HostMat A, B;
  A.init(D, D, 1, 0, true); B.init(D, D, 1, 0, true);
  // replicate one image `batch` times on the device
  size_t per = (size_t)D * D;
  __nv_bfloat16 *ah, *al, *bh, *bl, *ch, *cl;
  CK(cudaMalloc(&ah, per * batch * 2)); CK(cudaMalloc(&al, per * batch * 2));
  CK(cudaMalloc(&bh, per * batch * 2)); CK(cudaMalloc(&bl, per * batch * 2));
  CK(cudaMalloc(&ch, per * batch * 2)); CK(cudaMalloc(&cl, per * batch * 2));
  for (int b = 0; b < batch; ++b) {
    CK(cudaMemcpy(ah + per * b, A.d_hi, per * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(al + per * b, A.d_lo, per * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(bh + per * b, B.d_hi, per * 2, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(bl + per * b, B.d_lo, per * 2, cudaMemcpyDeviceToDevice));
  }
  GemmProblem g;
  g.M = D; g.N = D; g.batch = batch; g.nterms = 1;
  g.t[0].A.p0 = ah; g.t[0].A.p1 = al; g.t[0].A.rows = D; g.t[0].A.cols = D; g.t[0].A.ld = D; g.t[0].A.bstride = per;
  g.t[0].B = g.t[0].A; g.t[0].B.p0 = bh; g.t[0].B.p1 = bl;
  g.t[0].transA = tA; g.t[0].transB = tB; g.t[0].K = D;
  if (sym) { g.t[0].symA = g.t[0].symB = 1; g.sym_out = 1; }
  if (nterms == 2) { g.nterms = 2; g.t[1] = g.t[0]; }
  g.alpha = -0.5f; g.beta_eye = 1.5f;
  g.Cp.p0 = ch; g.Cp.p1 = npass == 3 ? cl : nullptr; g.Cp.rows = D; g.Cp.cols = D; g.Cp.ld = D; g.Cp.bstride = per;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) CK(egm::gemm_tc(g, npass, 0));
  CK(cudaDeviceSynchronize());
  const int iters = 10;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) CK(egm::gemm_tc(g, npass, 0));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  double flops = 2.0 * D * D * D * batch * nterms;
  if (sym) {   // only the upper 256-block tiles are evaluated
    const int nb = (D + 255) / 256;
    flops *= (double)(nb * (nb + 1) / 2) / (nb * nb);
  }
  printf("[perf] D=%d batch=%d npass=%d tA=%d tB=%d sym=%d terms=%d : %.3f ms  %.1f TFLOP/s evaluated (%.1f executed)\n",
         D, batch, npass, tA, tB, sym, nterms, ms, flops / ms * 1e-9, flops * npass / ms * 1e-9);
  fflush(stdout);
  cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl); cudaFree(ch); cudaFree(cl);
  A.free_(); B.free_();
}

// time vs K at fixed output shape: slope = per-k-block time, intercept = per-tile overhead
static void sweep(int D, int batch, int npass) {
  const int Ks[] = {128, 256, 512, 768, 1024, 1536, 2048};
  for (int K : Ks) {
    size_t pa = (size_t)D * K, pc = (size_t)D * D;
    __nv_bfloat16 *ah, *al, *bh, *bl, *ch, *cl;
    CK(cudaMalloc(&ah, pa * batch * 2)); CK(cudaMalloc(&al, pa * batch * 2));
    CK(cudaMalloc(&bh, pa * batch * 2)); CK(cudaMalloc(&bl, pa * batch * 2));
    CK(cudaMalloc(&ch, pc * batch * 2)); CK(cudaMalloc(&cl, pc * batch * 2));
    CK(cudaMemset(ah, 0, pa * batch * 2)); CK(cudaMemset(al, 0, pa * batch * 2));
    CK(cudaMemset(bh, 0, pa * batch * 2)); CK(cudaMemset(bl, 0, pa * batch * 2));
    GemmProblem g;
    g.M = D; g.N = D; g.batch = batch; g.nterms = 1;
    g.t[0].A.p0 = ah; g.t[0].A.p1 = al; g.t[0].A.rows = D; g.t[0].A.cols = K; g.t[0].A.ld = K; g.t[0].A.bstride = pa;
    g.t[0].B.p0 = bh; g.t[0].B.p1 = bl; g.t[0].B.rows = K; g.t[0].B.cols = D; g.t[0].B.ld = D; g.t[0].B.bstride = pa;
    g.t[0].transA = 0; g.t[0].transB = 0; g.t[0].K = K;
    g.Cp.p0 = ch; g.Cp.p1 = npass == 3 ? cl : nullptr; g.Cp.rows = D; g.Cp.cols = D; g.Cp.ld = D; g.Cp.bstride = pc;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) CK(egm::gemm_tc(g, npass, 0));
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) CK(egm::gemm_tc(g, npass, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    const double tiles = (double)((D + 255) / 256) * ((D + 255) / 256) * batch;
    const double us_tile = ms * 1e3 / (tiles / 74.0);
    printf("[sweep] D=%d batch=%d npass=%d K=%4d : %.3f ms  %.2f us/tile/pair  %.1f TFLOP/s executed\n", D, batch,
           npass, K, ms, us_tile, 2.0 * D * D * K * batch * npass / ms * 1e-9);
    fflush(stdout);
    cudaFree(ah); cudaFree(al); cudaFree(bh); cudaFree(bl); cudaFree(ch); cudaFree(cl);
  }
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "quick");
  if (argc > 1 && !strcmp(argv[1], "sweep")) {
    sweep(768, 128, 1);
    sweep(768, 128, 3);
    return 0;
  }
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  int fails = 0;
  //            name                              M    N   b nt  K        tA     tB   np alpha beta gamma e ab op of pad unal sample
  Case cases[] = {
      {"K/K  128x256x64   x1",                  128, 256, 1, 1, {64, 0},  {0, 0}, {1, 0}, 1, 1.f, 0.f, 0.f, 0, 0, 0, 1, 0, 0, 0},
      {"K/K  128x256x256  x1 (ring wraps)",     128, 256, 1, 1, {256, 0}, {0, 0}, {1, 0}, 1, 1.f, 0.f, 0.f, 0, 0, 0, 1, 0, 0, 0},
      {"K/K  128x256x512  x3 (ring wraps)",     128, 256, 1, 1, {512, 0}, {0, 0}, {1, 0}, 3, 1.f, 0.f, 0.f, 0, 0, 0, 1, 0, 0, 0},
      {"K/N  128x256x128  x1 (B row-major)",    128, 256, 1, 1, {128, 0}, {0, 0}, {0, 0}, 1, 1.f, 0.f, 0.f, 0, 0, 0, 1, 0, 0, 0},
      {"M/K  128x256x128  x1 (A transposed)",   128, 256, 1, 1, {128, 0}, {1, 0}, {1, 0}, 1, 1.f, 0.f, 0.f, 0, 0, 0, 1, 0, 0, 0},
      {"M/N  128x256x128  x1 (A^T * B)",        128, 256, 1, 1, {128, 0}, {1, 0}, {0, 0}, 1, 1.f, 0.f, 0.f, 0, 0, 0, 1, 0, 0, 0},
      {"K/N  256x512x192  x3 b2 multi-tile",    256, 512, 2, 1, {192, 0}, {0, 0}, {0, 0}, 3, 1.f, 0.f, 0.f, 0, 0, 1, 1, 0, 0, 0},
      {"NS-style 768^3 b3 x3 1.5I-0.5P",        768, 768, 3, 1, {768, 0}, {0, 0}, {0, 0}, 3, -0.5f, 1.5f, 0.f, 0, 0, 1, 1, 0, 0, 4000},
      {"NS-style 768^3 b3 x1",                  768, 768, 3, 1, {768, 0}, {0, 0}, {0, 0}, 1, -0.5f, 1.5f, 0.f, 0, 0, 1, 1, 0, 0, 4000},
      {"Gram 197x197x768 b3 x3 ragged, ld=N",   197, 197, 3, 1, {768, 0}, {0, 0}, {1, 0}, 3, 1.f, 0.f, 0.f, 0, 0, 1, 1, 0, 1, 0},
      {"Z^T U 768x768x197 b2 x3 ragged K",      768, 768, 2, 1, {197, 0}, {1, 0}, {0, 0}, 3, 1.f, 0.f, 0.f, 0, 1, 0, 1, 8, 0, 4000},
      {"W Zc 197x768x197 b2 x3 padded ld",      197, 768, 2, 1, {197, 0}, {0, 0}, {0, 0}, 3, 1.f, 0.f, 0.f, 0, 0, 1, 0, 8, 0, 4000},
      {"two-term dY T^T + Z^T dP 384 b2 x3",    384, 384, 2, 2, {384, 384}, {0, 1}, {1, 0}, 3, 1.f, 0.f, 0.f, 0, 0, 1, 1, 0, 0, 4000},
      {"two-term + planes addend + alpha_b",    200, 264, 3, 2, {72, 136}, {1, 0}, {0, 1}, 3, 0.75f, 0.f, -0.5f, 1, 1, 1, 1, 8, 0, 0},
      {"f32 addend, x1, small 64x64x64",        64, 64, 2, 1, {64, 0},   {0, 0}, {0, 0}, 1, 2.f, 3.f, 1.f, 2, 0, 1, 1, 0, 0, 0},
      // secondary plane output + <C,F> + alpha_b (A = M/tau and T_0 from one product)
      {"Cp2 + dot, 300x300x197 b3 x3 alpha_b",  300, 300, 3, 1, {197, 0}, {1, 0}, {0, 0}, 3, 1.f, 0.f, 0.f, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0},
      {"Cp2 + dot, 200x264 b2 x1 ragged, f32",  200, 264, 2, 1, {72, 0},  {0, 0}, {1, 0}, 1, 0.5f, 0.25f, -0.5f, 1, 0, 1, 1, 8, 0, 0, 1, 1, 0},
      {"planes addend, x1, 200x264x72 ragged", 200, 264, 2, 1, {72, 0},  {0, 0}, {1, 0}, 1, 0.5f, 0.25f, -0.5f, 1, 0, 1, 1, 8, 0, 0, 0, 0, 0},
      {"dot only U=W Zc 197x768x197 b2 x3",     197, 768, 2, 1, {197, 0}, {0, 0}, {0, 0}, 3, 1.f, 0.f, 0.f, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0},
      // packed upper triangle (+ fp32 copy of the full matrix is excluded where tiles are skipped)
      {"triu packed 768^3 b2 x3 post-scaled",   768, 768, 2, 1, {768, 0}, {0, 0}, {0, 0}, 3, 1.f, 0.f, 0.f, 0, 1, 0, 0, 0, 0, 0, 0, 0, 1},
      {"triu packed 300x300x96 b3 x1 ragged",   300, 300, 3, 1, {96, 0},  {0, 0}, {1, 0}, 1, -0.5f, 1.5f, 0.f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1},
      {"triu packed 197x197x64 b2 x3 1 tile",   197, 197, 2, 1, {64, 0},  {1, 0}, {0, 0}, 3, 1.f, 0.f, 0.5f, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1},
  };
  const int ncases = sizeof(cases) / sizeof(cases[0]);
  for (int i = 0; i < ncases; ++i) fails += run_case(cases[i]);
  fails += run_sym_case("sym 768 b2 x3 1 term", 768, 2, 1, 3, 0, 0);
  fails += run_sym_case("sym 768 b2 x3 2 terms + E + dot", 768, 2, 2, 3, 1, 1);
  fails += run_sym_case("sym 768 b2 x1 2 terms + dot", 768, 2, 2, 1, 0, 1);
  fails += run_sym_case("sym 1024 b1 x3 1 term + E", 1024, 1, 1, 3, 1, 0);
  fails += run_sym_case("sym 600 b2 x3 ragged 2 terms + dot", 600, 2, 2, 3, 0, 1);
  fails += run_sym_case("sym 200 b3 x3 single block", 200, 3, 1, 3, 1, 1);
  if (!quick) {
    bench(768, 256, 3, 0, 0, 1, 1);
    bench(768, 256, 3, 0, 0, 1, 2);
    bench(768, 256, 1, 0, 0, 1, 1);
    bench(768, 256, 1, 0, 0, 1, 2);
    bench(768, 256, 3, 0, 0, 0, 2);
    bench(768, 64, 1, 0, 0);
    bench(768, 64, 3, 0, 0);
    bench(768, 256, 1, 0, 0);
    bench(768, 256, 3, 0, 0);
    bench(768, 256, 3, 0, 1);
    bench(768, 256, 3, 1, 0);
    bench(1024, 128, 3, 0, 0);
    bench(1024, 128, 1, 0, 0);
  }
  printf("%s: %d failing case(s)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
