// Definitions of the bandwidth-bound kernels declared in egm_kernels.cuh.
// Reference arithmetic being reproduced (all fp32):
//   src/models/gpf_kernel.py:75-115,133-157   normalise / Hadamard powers / symmetrise / clamp
//   src/models/moment_head.py:222-266         degree normalisation, graph-weighted mean
//   src/models/moment_head.py:42-45,67-68     trace pre-normalisation / post-compensation
//   src/models/moment_head.py:100-131         count sketch x3 and their product
//   src/models/moment_head.py:215-218         triu half-vectorisation (row-major order)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "egm_kernels.cuh"

namespace egm {
namespace k {
namespace {

constexpr int kMaxDeg = 15;  // Hadamard powers 0..15 per view

struct WPtr {
  float* f;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  long long ld, bs;
};
WPtr wptr(const W& w, int prec) {
  WPtr p = {};
  p.ld = w.ld;
  p.bs = (long long)w.rows * w.ld;
  if (prec == PREC_FP32_SIMT) {
    p.f = static_cast<float*>(w.base);
  } else {
    p.hi = static_cast<__nv_bfloat16*>(w.base);
    p.lo = (prec == PREC_BF16X3) ? p.hi + (long long)w.batch * w.rows * w.ld : nullptr;
  }
  return p;
}
__device__ __forceinline__ void wstore(const WPtr& p, long long off, float x) {
  if (p.f) {
    p.f[off] = x;
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    p.hi[off] = h;
    if (p.lo) p.lo[off] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}
__device__ __forceinline__ void wstore4(const WPtr& p, long long off, float4 x) {
  if (p.f) {
    *reinterpret_cast<float4*>(p.f + off) = x;
  } else {
    const float v[4] = {x.x, x.y, x.z, x.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = __float2bfloat16_rn(v[i]);
      l[i] = __float2bfloat16_rn(v[i] - __bfloat162float(h[i]));
    }
    uint2 hh, ll;
    hh.x = __bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    hh.y = __bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    *reinterpret_cast<uint2*>(p.hi + off) = hh;
    if (p.lo) {
      ll.x = __bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
      ll.y = __bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
      *reinterpret_cast<uint2*>(p.lo + off) = ll;
    }
  }
}
__device__ __forceinline__ void wstore2(const WPtr& p, long long off, float x0, float x1) {
  if (p.f) {
    *reinterpret_cast<float2*>(p.f + off) = make_float2(x0, x1);
  } else {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
    *reinterpret_cast<uint32_t*>(p.hi + off) =
        __bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    if (p.lo) {
      const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
      *reinterpret_cast<uint32_t*>(p.lo + off) =
          __bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
}
__device__ __forceinline__ float wload(const WPtr& p, long long off) {
  if (p.f) return p.f[off];
  float x = __bfloat162float(p.hi[off]);
  if (p.lo) x += __bfloat162float(p.lo[off]);
  return x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum, result valid in every thread; `sh` needs 32 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (wid == 0) r = warp_sum(r);
  if (threadIdx.x == 0) sh[0] = r;
  __syncthreads();
  return sh[0];
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------- affine
template <int V>
__global__ void affine_kernel(const float* __restrict__ x, long long ld, long long bs, int rows,
                              int cols, const float* __restrict__ s, float a1, float b1, WPtr o1,
                              float a2, float b2, WPtr o2, int has2) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (j >= cols) return;
  const float sc = s ? s[b] : 1.f;
  const float* src = x + (long long)b * bs + (long long)i * ld + j;
  if (V == 4) {
    const float4 v = *reinterpret_cast<const float4*>(src);
    const float in[4] = {v.x, v.y, v.z, v.w};
    float r1[4], r2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float e = (i == j + q) ? 1.f : 0.f;
      r1[q] = a1 * sc * in[q] + b1 * e;
      r2[q] = a2 * sc * in[q] + b2 * e;
    }
    wstore4(o1, (long long)b * o1.bs + (long long)i * o1.ld + j, make_float4(r1[0], r1[1], r1[2], r1[3]));
    if (has2) wstore4(o2, (long long)b * o2.bs + (long long)i * o2.ld + j, make_float4(r2[0], r2[1], r2[2], r2[3]));
  } else {
    const float e = (i == j) ? 1.f : 0.f;
    const float v = *src;
    wstore(o1, (long long)b * o1.bs + (long long)i * o1.ld + j, a1 * sc * v + b1 * e);
    if (has2) wstore(o2, (long long)b * o2.bs + (long long)i * o2.ld + j, a2 * sc * v + b2 * e);
  }
}

template <int V>
__global__ void export_kernel(WPtr in, int rows, int cols, float* __restrict__ out, long long ld,
                              long long bs) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  out[(long long)b * bs + (long long)i * ld + j] = wload(in, (long long)b * in.bs + (long long)i * in.ld + j);
}

// ------------------------------------------------------------------ rownorm
__global__ void rownorm_kernel(const float* __restrict__ x, int n, int d, float eps, int cosine,
                               float* __restrict__ nrm, WPtr xn) {
  // one warp per token row
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* src = x + ((long long)b * n + row) * d;
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) { const float v = src[j]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float nr = sqrtf(ss);
  if (lane == 0) nrm[(long long)b * n + row] = nr;
  const float inv = cosine ? 1.f / fmaxf(nr, eps) : 1.f;
  const long long o = (long long)b * xn.bs + (long long)row * xn.ld;
  for (int j = lane; j < d; j += 32) wstore(xn, o + j, src[j] * inv);
}

__global__ void rownorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ nrm,
                                   const float* __restrict__ dxn, int n, int d, float eps,
                                   float* __restrict__ dx) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const long long o = ((long long)b * n + row) * d;
  const float nr = nrm[(long long)b * n + row];
  if (nr >= eps) {
    // F.normalize: x / clamp_min(||x||, eps); the clamp passes gradient when ||x|| >= eps
    const float inv = 1.f / nr;
    float dot = 0.f;
    for (int j = lane; j < d; j += 32) dot = fmaf(x[o + j] * inv, dxn[o + j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < d; j += 32) dx[o + j] = (dxn[o + j] - x[o + j] * inv * dot) * inv;
  } else {
    const float inv = 1.f / eps;
    for (int j = lane; j < d; j += 32) dx[o + j] = dxn[o + j] * inv;
  }
}

// --------------------------------------------------------------- polynomial
template <int MD>
__device__ __forceinline__ void had_powers(float x, int deg, float (&pw)[MD + 1]) {
  // f_0 = 1, f_1 = x (NOT clamped), f_k = max(x,0)^k   (gpf_kernel.py:107-115)
  pw[0] = 1.f;
  if (MD >= 1) pw[1] = x;
  const float c = fmaxf(x, 0.f);
  float acc = c;
#pragma unroll
  for (int k = 2; k <= MD; ++k) {
    acc *= c;
    pw[k] = (k <= deg) ? acc : 0.f;
  }
}
template <int MD>
__device__ __forceinline__ void had_dpowers(float x, int deg, float (&dp)[MD + 1]) {
  // f'_0 = 0, f'_1 = 1, f'_k = k max(x,0)^(k-1)
  dp[0] = 0.f;
  if (MD >= 1) dp[1] = 1.f;
  const float c = fmaxf(x, 0.f);
  float acc = 1.f;
#pragma unroll
  for (int k = 2; k <= MD; ++k) {
    acc *= c;
    dp[k] = (k <= deg) ? k * acc : 0.f;
  }
}
// sum_{p<=P, q<=Q} coef[p,q] pa[p] pb[q]; coef is stored with row stride (Q+1)
template <int MD>
__device__ __forceinline__ float poly_eval(const float (&pa)[MD + 1], const float (&pb)[MD + 1],
                                           const float* coef, int P, int Q) {
  float f = 0.f;
#pragma unroll
  for (int p = 0; p <= MD; ++p) {
    if (p <= P) {
      float inner = 0.f;
#pragma unroll
      for (int q = 0; q <= MD; ++q)
        if (q <= Q) inner = fmaf(coef[p * (Q + 1) + q], pb[q], inner);
      f = fmaf(pa[p], inner, f);
    }
  }
  return f;
}

// f = sum_pq c_pq pa_p pb_q together with df/da-side = sum_pq c_pq da_p pb_q and df/db-side =
// sum_pq c_pq pa_p db_q, sharing the inner sums. `cp` is the coefficient table padded with zeros to
// (MD+1) x (MD+1), so no degree tests are needed (powers above the degree are zero as well).
template <int MD>
__device__ __forceinline__ void poly_eval_grad(const float (&pa)[MD + 1], const float (&pb)[MD + 1],
                                               const float (&da)[MD + 1], const float (&db)[MD + 1],
                                               const float* cp, float& f, float& dfa, float& dfb) {
  float colsum[MD + 1];   // sum_p c_pq pa_p
#pragma unroll
  for (int q = 0; q <= MD; ++q) colsum[q] = 0.f;
  f = 0.f; dfa = 0.f;
#pragma unroll
  for (int p = 0; p <= MD; ++p) {
    float inner = 0.f;    // sum_q c_pq pb_q
#pragma unroll
    for (int q = 0; q <= MD; ++q) {
      const float cpq = cp[p * (MD + 1) + q];
      inner = fmaf(cpq, pb[q], inner);
      colsum[q] = fmaf(cpq, pa[p], colsum[q]);
    }
    f = fmaf(pa[p], inner, f);
    dfa = fmaf(da[p], inner, dfa);
  }
  dfb = 0.f;
#pragma unroll
  for (int q = 0; q <= MD; ++q) dfb = fmaf(db[q], colsum[q], dfb);
}

// The polynomial kernels work on PAIRS of 32 x 32 tiles, (ti,tj) and its mirror (tj,ti), ti <= tj:
// the mirror tile is staged in shared memory with coalesced loads and read transposed, so a thread
// owns the element pair (i,j)/(j,i), evaluates it once, writes (i,j) directly and hands (j,i) back
// through shared memory for a coalesced store. No strided global access anywhere.
constexpr int kPT = 32;
__host__ __device__ inline int poly_tiles(int n) { return (n + kPT - 1) / kPT; }
__host__ __device__ inline int poly_pairs(int n) { return poly_tiles(n) * (poly_tiles(n) + 1) / 2; }
__device__ __forceinline__ void poly_pair_coords(int pair, int nt, int& ti, int& tj) {
  ti = 0;
  while (pair >= nt - ti) { pair -= nt - ti; ++ti; }
  tj = ti + pair;
}
// A pair whose column tile is the ragged last one (n % 32 columns) is walked from its mirror instead:
// the direct tile then has n % 32 full-width rows (whole warps skip the rest) rather than 32 rows with
// n % 32 active lanes each. The element-pair arithmetic is symmetric in the two roles.
__device__ __forceinline__ void poly_prefer_full_rows(int n, int& ti, int& tj) {
  if (tj != ti && tj == poly_tiles(n) - 1 && (n % kPT) != 0) { const int t = ti; ti = tj; tj = t; }
}

template <int MD>
__global__ void __launch_bounds__(256)
gpf_poly_fwd_kernel(const float* __restrict__ Ra, const float* __restrict__ Rp, long long ldR,
                    const float* __restrict__ coef, int P, int Q, int symmetric, int n,
                    float* __restrict__ G) {
  __shared__ float c[(kMaxDeg + 1) * (kMaxDeg + 1)];
  __shared__ float sa[kPT][kPT + 1], sp[kPT][kPT + 1];
  for (int t = threadIdx.x; t < (P + 1) * (Q + 1); t += blockDim.x) c[t] = coef[t];
  const int b = blockIdx.y;
  int ti, tj;
  poly_pair_coords(blockIdx.x, poly_tiles(n), ti, tj);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long rb = (long long)b * n * ldR;
  const long long gb = (long long)b * n * n;
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // mirror tile (tj, ti)
    const int r = ty + 8 * k, i2 = tj * kPT + r, j2 = ti * kPT + tx;
    const bool ok = i2 < n && j2 < n;
    sa[r][tx] = ok ? Ra[rb + (long long)i2 * ldR + j2] : 0.f;
    sp[r][tx] = ok ? Rp[rb + (long long)i2 * ldR + j2] : 0.f;
  }
  __syncthreads();
  float g2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k, i = ti * kPT + r, j = tj * kPT + tx;
    g2[k] = 0.f;
    if (i < n && j < n) {
      float pa[MD + 1], pb[MD + 1];
      had_powers<MD>(Ra[rb + (long long)i * ldR + j], P, pa);
      had_powers<MD>(Rp[rb + (long long)i * ldR + j], Q, pb);
      const float f = poly_eval<MD>(pa, pb, c, P, Q);
      had_powers<MD>(sa[tx][r], P, pa);
      had_powers<MD>(sp[tx][r], Q, pb);
      const float ft = poly_eval<MD>(pa, pb, c, P, Q);
      const float g1 = fmaxf(symmetric ? 0.5f * (f + ft) : f, 0.f);
      g2[k] = symmetric ? g1 : fmaxf(ft, 0.f);
      G[gb + (long long)i * n + j] = g1;
    }
  }
  if (ti == tj) return;   // block-uniform: a diagonal tile is its own mirror
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) sa[tx][ty + 8 * k] = g2[k];   // element (j,i): mirror-tile row tx, column r
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k, i2 = tj * kPT + r, j2 = ti * kPT + tx;
    if (i2 < n && j2 < n) G[gb + (long long)i2 * n + j2] = sa[r][tx];
  }
}

// one block per tile pair of one image; deterministic two-stage dcoef reduction
template <int MD>
__global__ void __launch_bounds__(256, (MD <= 3) ? 4 : 1)
gpf_poly_bwd_kernel(const float* __restrict__ dG, const float* __restrict__ Ra,
                    const float* __restrict__ Rp, long long ldR, const float* __restrict__ coef,
                    int P, int Q, int symmetric, int n, WPtr Ea, WPtr Ep,
                    float* __restrict__ partial) {
  __shared__ float cp[(MD + 1) * (MD + 1)];    // coefficients, zero-padded to (MD+1)^2
  __shared__ float red[(MD <= 3) ? 1 : (MD + 1) * (MD + 1)];
  __shared__ float sa[kPT][kPT + 1], sp[kPT][kPT + 1], sg[kPT][kPT + 1];
  __shared__ float wacc[8][16];
  const int nterm = (P + 1) * (Q + 1);
  for (int t = threadIdx.x; t < (MD + 1) * (MD + 1); t += blockDim.x) {
    const int pp = t / (MD + 1), qq = t % (MD + 1);
    cp[t] = (pp <= P && qq <= Q) ? coef[pp * (Q + 1) + qq] : 0.f;
    if (MD > 3) red[t] = 0.f;
  }
  const int b = blockIdx.y;
  int ti, tj;
  poly_pair_coords(blockIdx.x, poly_tiles(n), ti, tj);
  const bool offdiag = ti != tj;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long rb = (long long)b * n * ldR;
  const long long gb = (long long)b * n * n;
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // mirror tile (tj, ti)
    const int r = ty + 8 * k, i2 = tj * kPT + r, j2 = ti * kPT + tx;
    const bool ok = i2 < n && j2 < n;
    sa[r][tx] = ok ? Ra[rb + (long long)i2 * ldR + j2] : 0.f;
    sp[r][tx] = ok ? Rp[rb + (long long)i2 * ldR + j2] : 0.f;
    sg[r][tx] = ok ? dG[gb + (long long)i2 * n + j2] : 0.f;
  }
  // this thread's own elements: issue the loads before the barrier
  float ra_k[4], rp_k[4], g_k[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = ti * kPT + ty + 8 * k, j = tj * kPT + tx;
    const bool ok = i < n && j < n;
    ra_k[k] = ok ? Ra[rb + (long long)i * ldR + j] : 0.f;
    rp_k[k] = ok ? Rp[rb + (long long)i * ldR + j] : 0.f;
    g_k[k] = ok ? dG[gb + (long long)i * n + j] : 0.f;
  }
  __syncthreads();
  // small-degree fast path keeps the per-term sums in registers
  float acc[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t] = 0.f;
  float ea_k[4], ep_k[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k, i = ti * kPT + r, j = tj * kPT + tx;
    ea_k[k] = ep_k[k] = 0.f;
    if (i < n && j < n) {
      float pa[MD + 1], pb[MD + 1], da[MD + 1], db[MD + 1];
      float pat[MD + 1], pbt[MD + 1], dat[MD + 1], dbt[MD + 1];
      const float ra = ra_k[k], rp = rp_k[k];
      const float rat = sa[tx][r], rpt = sp[tx][r];
      had_powers<MD>(ra, P, pa); had_powers<MD>(rp, Q, pb);
      had_dpowers<MD>(ra, P, da); had_dpowers<MD>(rp, Q, db);
      had_powers<MD>(rat, P, pat); had_powers<MD>(rpt, Q, pbt);
      had_dpowers<MD>(rat, P, dat); had_dpowers<MD>(rpt, Q, dbt);
      float f, fa, fb, ft, fat, fbt;
      poly_eval_grad<MD>(pa, pb, da, db, cp, f, fa, fb);
      poly_eval_grad<MD>(pat, pbt, dat, dbt, cp, ft, fat, fbt);
      const float g_ij = g_k[k], g_ji = sg[tx][r];
      float dF_ij, dF_ji;
      if (symmetric) {
        const float sgate = (0.5f * (f + ft) >= 0.f) ? 1.f : 0.f;  // clamp(min=0) passes grad at 0
        dF_ij = dF_ji = 0.5f * (g_ij + g_ji) * sgate;
      } else {
        dF_ij = (f >= 0.f) ? g_ij : 0.f;
        dF_ji = (ft >= 0.f) ? g_ji : 0.f;
      }
      // E = dR + dR^T is symmetric: the same value goes to (i,j) and (j,i)
      const float ea = dF_ij * fa + dF_ji * fat;
      const float ep = dF_ij * fb + dF_ji * fbt;
      ea_k[k] = ea; ep_k[k] = ep;
      wstore(Ea, (long long)b * Ea.bs + (long long)i * Ea.ld + j, ea);
      wstore(Ep, (long long)b * Ep.bs + (long long)i * Ep.ld + j, ep);
      // dcoef: this thread's (i,j) term, plus the (j,i) term when the mirror tile is not this tile
      const float dF_m = offdiag ? dF_ji : 0.f;
      if (MD == 3) {
        // acc[4p+q]: static register indices; re-packed to the (Q+1)-strided order on output
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const float u = dF_ij * pa[pp], v = dF_m * pat[pp];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) acc[pp * 4 + qq] = fmaf(u, pb[qq], fmaf(v, pbt[qq], acc[pp * 4 + qq]));
        }
      } else {
        for (int pp = 0; pp <= P; ++pp)
          for (int qq = 0; qq <= Q; ++qq)
            atomicAdd(&red[pp * (MD + 1) + qq], dF_ij * pa[pp] * pb[qq] + dF_m * pat[pp] * pbt[qq]);
      }
    }
  }
  if (offdiag) {   // block-uniform
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) { sa[tx][ty + 8 * k] = ea_k[k]; sp[tx][ty + 8 * k] = ep_k[k]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ty + 8 * k, i2 = tj * kPT + r, j2 = ti * kPT + tx;
      if (i2 < n && j2 < n) {
        wstore(Ea, (long long)b * Ea.bs + (long long)i2 * Ea.ld + j2, sa[r][tx]);
        wstore(Ep, (long long)b * Ep.bs + (long long)i2 * Ep.ld + j2, sp[r][tx]);
      }
    }
  }
  float* out = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * nterm;
  if (MD == 3) {
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const float v = warp_sum(acc[t]);
      if (tx == 0) wacc[ty][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      const int pp = threadIdx.x >> 2, qq = threadIdx.x & 3;
      if (pp <= P && qq <= Q) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += wacc[w][threadIdx.x];
        out[pp * (Q + 1) + qq] = v;
      }
    }
  } else {
    __syncthreads();
    for (int t = threadIdx.x; t < (MD + 1) * (MD + 1); t += blockDim.x) {
      const int pp = t / (MD + 1), qq = t % (MD + 1);
      if (pp <= P && qq <= Q) out[pp * (Q + 1) + qq] = red[t];
    }
  }
}
// ---- degree <= 3 (the configurations of the reference: configs/ufg_base.yaml sweeps (1,1)..(3,3)):
// the 4 x 4 zero-padded coefficient table lives in registers and the polynomial is two nested Horner
// forms, ~20 FMA per evaluation instead of a shared-memory read per term.
struct Poly3 {
  float c[16];
  __device__ __forceinline__ void load(const float* __restrict__ coef, int P, int Q) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[p * 4 + q] = (p <= P && q <= Q) ? coef[p * (Q + 1) + q] : 0.f;
  }
  // inner_p = sum_q c_pq f_q(y)
  __device__ __forceinline__ void inner(float y, float y2, float y3, float (&in)[4]) const {
#pragma unroll
    for (int p = 0; p < 4; ++p) in[p] = fmaf(c[p * 4 + 3], y3, fmaf(c[p * 4 + 2], y2, fmaf(c[p * 4 + 1], y, c[p * 4])));
  }
  // f = sum_pq c_pq f_p(x) f_q(y);  f_0 = 1, f_1 = t, f_k = max(t,0)^k
  __device__ __forceinline__ float eval(float x, float y) const {
    const float cx = fmaxf(x, 0.f), cy = fmaxf(y, 0.f);
    const float x2 = cx * cx, x3 = x2 * cx, y2 = cy * cy, y3 = y2 * cy;
    float in[4];
    inner(y, y2, y3, in);
    return fmaf(in[3], x3, fmaf(in[2], x2, fmaf(in[1], x, in[0])));
  }
  // f, df/dx, df/dy and the power tables pa = f_p(x), pb = f_q(y)
  __device__ __forceinline__ void eval_grad(float x, float y, float& f, float& fx, float& fy,
                                            float (&pa)[4], float (&pb)[4]) const {
    const float cx = fmaxf(x, 0.f), cy = fmaxf(y, 0.f);
    pa[0] = 1.f; pa[1] = x; pa[2] = cx * cx; pa[3] = pa[2] * cx;
    pb[0] = 1.f; pb[1] = y; pb[2] = cy * cy; pb[3] = pb[2] * cy;
    float in[4];
    inner(y, pb[2], pb[3], in);
    f = fmaf(in[3], pa[3], fmaf(in[2], pa[2], fmaf(in[1], x, in[0])));
    fx = fmaf(3.f * pa[2], in[3], fmaf(2.f * cx, in[2], in[1]));      // f'_1 = 1, f'_k = k max(t,0)^(k-1)
    float col[3];                                                     // col_q = sum_p c_pq f_p(x), q = 1..3
#pragma unroll
    for (int q = 1; q < 4; ++q) col[q - 1] = fmaf(c[12 + q], pa[3], fmaf(c[8 + q], pa[2], fmaf(c[4 + q], x, c[q])));
    fy = fmaf(3.f * pb[2], col[2], fmaf(2.f * cy, col[1], col[0]));
  }
};

__global__ void __launch_bounds__(256)
gpf_poly3_fwd_kernel(const float* __restrict__ Ra, const float* __restrict__ Rp, int ldR,
                     const float* __restrict__ coef, int P, int Q, int symmetric, int n,
                     float* __restrict__ G) {
  __shared__ float sa[kPT][kPT + 1], sp[kPT][kPT + 1];
  Poly3 poly;
  poly.load(coef, P, Q);
  int ti, tj;
  poly_pair_coords(blockIdx.x, poly_tiles(n), ti, tj);
  poly_prefer_full_rows(n, ti, tj);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* ra_b = Ra + (long long)blockIdx.y * n * ldR;
  const float* rp_b = Rp + (long long)blockIdx.y * n * ldR;
  float* g_b = G + (long long)blockIdx.y * n * n;
  const int i0 = ti * kPT, j0 = tj * kPT;
  float xa[4], xp[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    const bool okm = (j0 + r < n) && (i0 + tx < n);     // mirror tile (tj, ti)
    sa[r][tx] = okm ? ra_b[(j0 + r) * ldR + i0 + tx] : 0.f;
    sp[r][tx] = okm ? rp_b[(j0 + r) * ldR + i0 + tx] : 0.f;
    const bool ok = (i0 + r < n) && (j0 + tx < n);
    xa[k] = ok ? ra_b[(i0 + r) * ldR + j0 + tx] : 0.f;
    xp[k] = ok ? rp_b[(i0 + r) * ldR + j0 + tx] : 0.f;
  }
  __syncthreads();
  float g2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    g2[k] = 0.f;
    if (i0 + r >= n) continue;   // warp-uniform: a warp owns one row of the tile
    const float f = poly.eval(xa[k], xp[k]);
    const float ft = poly.eval(sa[tx][r], sp[tx][r]);
    const float g1 = fmaxf(symmetric ? 0.5f * (f + ft) : f, 0.f);
    g2[k] = symmetric ? g1 : fmaxf(ft, 0.f);
    if ((i0 + r < n) && (j0 + tx < n)) g_b[(i0 + r) * n + j0 + tx] = g1;
  }
  if (ti == tj) return;   // block-uniform: a diagonal tile is its own mirror
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) sa[tx][ty + 8 * k] = g2[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    if ((j0 + r < n) && (i0 + tx < n)) g_b[(j0 + r) * n + i0 + tx] = sa[r][tx];
  }
}

// RSYM: R_a and R_p are symmetric bit for bit (written by the fused forward, which evaluates each pair
// once): the mirror element has the same polynomial value, derivatives and power tables, so one
// evaluation per pair suffices and only dG is read at the mirrored position.
//
// FOLD (cosine similarity): the backward of F.normalize is folded into E, so that the product with the
// RAW token planes is the token gradient itself and no rownorm_bwd pass over [B,N,D] is needed:
//   dx_i = (dxh_i - xh_i <xh_i, dxh_i>) / n_i,  dxh = E xh,  <xh_i, dxh_i> = sum_j E_ij R_ij =: s_i
//   =>  dx = E'' X  with  E''_ij = (E_ij - [n_i >= eps] s_i delta_ij) / (m_i m_j),  m = max(n, eps)
// (a row below the clamp has xh_i = x_i / eps and no projection term, gpf_kernel.py:87 / F.normalize).
// This kernel scales E by 1/(m_i m_j) and writes the per-column-tile partial sums of s; gpf_diag_fix_kernel
// adds them up in a fixed order and patches the diagonal of the planes.
template <bool RSYM, bool FOLD>
__global__ void __launch_bounds__(256, 3)
gpf_poly3_bwd_kernel(const float* __restrict__ dG, const float* __restrict__ Ra,
                     const float* __restrict__ Rp, int ldR, const float* __restrict__ coef, int P, int Q,
                     int symmetric, int n, WPtr Ea, WPtr Ep, float* __restrict__ partial,
                     const float* __restrict__ nrm_a, const float* __restrict__ nrm_p, float eps,
                     float* __restrict__ rowpart) {
  __shared__ float sa[kPT][kPT + 1], sp[kPT][kPT + 1], sg[kPT][kPT + 1];
  __shared__ float wacc[8][16];
  Poly3 poly;
  poly.load(coef, P, Q);
  int ti, tj;
  poly_pair_coords(blockIdx.x, poly_tiles(n), ti, tj);
  poly_prefer_full_rows(n, ti, tj);
  const bool offdiag = ti != tj;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* ra_b = Ra + (long long)blockIdx.y * n * ldR;
  const float* rp_b = Rp + (long long)blockIdx.y * n * ldR;
  const float* dg_b = dG + (long long)blockIdx.y * n * n;
  const int i0 = ti * kPT, j0 = tj * kPT;
  float xa[4], xp[4], xg[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    const bool okm = (j0 + r < n) && (i0 + tx < n);     // mirror tile (tj, ti)
    if (!RSYM) {
      sa[r][tx] = okm ? ra_b[(j0 + r) * ldR + i0 + tx] : 0.f;
      sp[r][tx] = okm ? rp_b[(j0 + r) * ldR + i0 + tx] : 0.f;
    }
    sg[r][tx] = okm ? dg_b[(j0 + r) * n + i0 + tx] : 0.f;
    const bool ok = (i0 + r < n) && (j0 + tx < n);
    xa[k] = ok ? ra_b[(i0 + r) * ldR + j0 + tx] : 0.f;
    xp[k] = ok ? rp_b[(i0 + r) * ldR + j0 + tx] : 0.f;
    xg[k] = ok ? dg_b[(i0 + r) * n + j0 + tx] : 0.f;
  }
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) acc[t] = 0.f;
  float ea_k[4], ep_k[4];
  float ma_k[4], mp_k[4];     // FOLD: R at the mirrored position (column sums of E * R)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = ty + 8 * k;
    ea_k[k] = ep_k[k] = 0.f;
    ma_k[k] = xa[k]; mp_k[k] = xp[k];
    if (i0 + r >= n) continue;   // warp-uniform: a warp owns one row of the tile
    float f, fa, fb, pa[4], pb[4];
    poly.eval_grad(xa[k], xp[k], f, fa, fb, pa, pb);
    const float g_ij = xg[k], g_ji = sg[tx][r];       // both zero outside the matrix
    if (RSYM) {
      float dF_ij, dF_ji;
      if (symmetric) {
        dF_ij = dF_ji = (f >= 0.f) ? 0.5f * (g_ij + g_ji) : 0.f;
      } else {
        dF_ij = (f >= 0.f) ? g_ij : 0.f;
        dF_ji = (f >= 0.f) ? g_ji : 0.f;
      }
      const float dsum = dF_ij + dF_ji;
      ea_k[k] = dsum * fa;
      ep_k[k] = dsum * fb;
      const float dc = dF_ij + (offdiag ? dF_ji : 0.f);
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const float u = dc * pa[pp];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) acc[pp * 4 + qq] = fmaf(u, pb[qq], acc[pp * 4 + qq]);
      }
      continue;
    }
    float ft, fat, fbt, pat[4], pbt[4];
    poly.eval_grad(sa[tx][r], sp[tx][r], ft, fat, fbt, pat, pbt);
    ma_k[k] = sa[tx][r]; mp_k[k] = sp[tx][r];
    float dF_ij, dF_ji;
    if (symmetric) {
      const float sgate = (0.5f * (f + ft) >= 0.f) ? 1.f : 0.f;  // clamp(min=0) passes grad at 0
      dF_ij = dF_ji = 0.5f * (g_ij + g_ji) * sgate;
    } else {
      dF_ij = (f >= 0.f) ? g_ij : 0.f;
      dF_ji = (ft >= 0.f) ? g_ji : 0.f;
    }
    // E = dR + dR^T is symmetric: the same value goes to (i,j) and (j,i)
    ea_k[k] = dF_ij * fa + dF_ji * fat;
    ep_k[k] = dF_ij * fb + dF_ji * fbt;
    // dcoef: this thread's (i,j) term, plus the (j,i) term when the mirror tile is not this tile
    const float dF_m = offdiag ? dF_ji : 0.f;
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      const float u = dF_ij * pa[pp], v = dF_m * pat[pp];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) acc[pp * 4 + qq] = fmaf(u, pb[qq], fmaf(v, pbt[qq], acc[pp * 4 + qq]));
    }
  }
  // stage the tile of E in shared memory; rows leave as 4-column groups (one 8-byte store per plane),
  // the mirror tile as the transposed read of the same staging tile
  __syncthreads();
  if (FOLD) {
    const int nt = poly_tiles(n);
    const long long bn = (long long)blockIdx.y * n;
    float* rp_a = rowpart + ((long long)blockIdx.y * 2 + 0) * nt * n;
    float* rp_p = rowpart + ((long long)blockIdx.y * 2 + 1) * nt * n;
    const int j = j0 + tx;
    const float ima_c = j < n ? 1.f / fmaxf(nrm_a[bn + j], eps) : 0.f;
    const float imp_c = j < n ? 1.f / fmaxf(nrm_p[bn + j], eps) : 0.f;
    float ca = 0.f, cp = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + ty + 8 * k;
      // row sums over this tile's columns (a warp owns the row), column sums over its rows
      const float ra = warp_sum(ea_k[k] * xa[k]), rpv = warp_sum(ep_k[k] * xp[k]);
      if (tx == 0 && i < n) { rp_a[(long long)tj * n + i] = ra; rp_p[(long long)tj * n + i] = rpv; }
      ca = fmaf(ea_k[k], ma_k[k], ca);
      cp = fmaf(ep_k[k], mp_k[k], cp);
      const float ima_r = i < n ? 1.f / fmaxf(nrm_a[bn + i], eps) : 0.f;
      const float imp_r = i < n ? 1.f / fmaxf(nrm_p[bn + i], eps) : 0.f;
      ea_k[k] *= ima_r * ima_c;
      ep_k[k] *= imp_r * imp_c;
    }
    if (offdiag) {   // block-uniform; sg is free again: rows 0-7 / 8-15 hold the per-warp column sums
      sg[ty][tx] = ca;
      sg[8 + ty][tx] = cp;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { sa[ty + 8 * k][tx] = ea_k[k]; sp[ty + 8 * k][tx] = ep_k[k]; }
  __syncthreads();
  if (FOLD && offdiag && ty == 0 && j0 + tx < n) {
    const int nt = poly_tiles(n);
    float a = 0.f, b2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += sg[w][tx]; b2 += sg[8 + w][tx]; }
    rowpart[(((long long)blockIdx.y * 2 + 0) * nt + ti) * n + j0 + tx] = a;
    rowpart[(((long long)blockIdx.y * 2 + 1) * nt + ti) * n + j0 + tx] = b2;
  }
  {
    const int r = threadIdx.x >> 3, c = (threadIdx.x & 7) * 4;
    const long long eb = (long long)blockIdx.y * Ea.bs;
    if (i0 + r < n && j0 + c < n) {
      const long long o = eb + (long long)(i0 + r) * Ea.ld + j0 + c;
      wstore4(Ea, o, make_float4(sa[r][c], sa[r][c + 1], sa[r][c + 2], sa[r][c + 3]));
      wstore4(Ep, o, make_float4(sp[r][c], sp[r][c + 1], sp[r][c + 2], sp[r][c + 3]));
    }
    if (offdiag && j0 + r < n && i0 + c < n) {
      const long long o = eb + (long long)(j0 + r) * Ea.ld + i0 + c;
      wstore4(Ea, o, make_float4(sa[c][r], sa[c + 1][r], sa[c + 2][r], sa[c + 3][r]));
      wstore4(Ep, o, make_float4(sp[c][r], sp[c + 1][r], sp[c + 2][r], sp[c + 3][r]));
    }
  }
  const int nterm = (P + 1) * (Q + 1);
  float* out = partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * nterm;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float v = warp_sum(acc[t]);
    if (tx == 0) wacc[ty][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int pp = threadIdx.x >> 2, qq = threadIdx.x & 3;
    if (pp <= P && qq <= Q) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += wacc[w][threadIdx.x];
      out[pp * (Q + 1) + qq] = v;
    }
  }
}

// E''_ii -= [n_i >= eps] * s_i / m_i^2 with s_i = sum over column tiles of the partial row sums (fixed order)
__global__ void gpf_diag_fix_kernel(const float* __restrict__ rowpart, const float* __restrict__ nrm_a,
                                    const float* __restrict__ nrm_p, int n, int nt, float eps, WPtr Ea, WPtr Ep) {
  const int b = blockIdx.y, v = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float nr = (v ? nrm_p : nrm_a)[(long long)b * n + i];
  if (nr < eps) return;                       // below the clamp: no projection term
  const float* rp = rowpart + ((long long)b * 2 + v) * nt * n;
  float s = 0.f;
  for (int t = 0; t < nt; ++t) s += rp[(long long)t * n + i];
  const WPtr& E = v ? Ep : Ea;
  const long long o = (long long)b * E.bs + (long long)i * E.ld + i;
  wstore(E, o, wload(E, o) - s / (nr * nr));
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nblocks, int nt,
                                       float* __restrict__ out) {
  __shared__ float sh[32];
  const int t = blockIdx.x;
  float a = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) a += partial[(long long)i * nt + t];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) out[t] = a;
}

// ---------------------------------------------------------- degree / weight
__global__ void degree_kernel(const float* __restrict__ G, int n, float eps, float* __restrict__ deg,
                              float* __restrict__ s) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* g = G + ((long long)b * n + row) * n;
  float a = 0.f;
  for (int j = lane; j < n; j += 32) a += g[j];
  a = warp_sum(a);
  if (lane == 0) {
    deg[(long long)b * n + row] = a;
    s[(long long)b * n + row] = rsqrtf(fmaxf(a, eps));
  }
}
__global__ void weight_kernel(const float* __restrict__ G, const float* __restrict__ s, int n,
                              WPtr Wn, float* __restrict__ w, float* __restrict__ wdiag) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* g = G + ((long long)b * n + row) * n;
  const float* sb = s + (long long)b * n;
  const float si = sb[row];
  float a = 0.f;
  const long long o = (long long)b * Wn.bs + (long long)row * Wn.ld;
  for (int j = lane; j < n; j += 32) {
    const float v = g[j] * si * sb[j];   // graph * s_i * s_j, same association as the reference
    a += v;
    wstore(Wn, o + j, v);
    if (j == row) wdiag[(long long)b * n + row] = v;
  }
  a = warp_sum(a);
  if (lane == 0) w[(long long)b * n + row] = a;
}

// ------------------------------------------------------------- mean / centre
// Block = 64 columns x 8 row groups (512 threads): row group g walks rows g, g+8, ...; the weighted
// column sums are combined through shared memory, then every group centres and stores its own rows
// (second read of the block's 64-column slab hits L2). 8x the loads in flight of a thread-per-column
// layout, which is what an HBM-bound pass over a [197 x 768] image needs.
constexpr int kMcCols = 64, kMcGroups = 8;
__global__ void __launch_bounds__(kMcCols * kMcGroups)
mean_center_kernel(const float* __restrict__ Z, const float* __restrict__ w,
                   const float* __restrict__ wdiag, int n, int d, float eps, float* __restrict__ t_out,
                   float* __restrict__ sw_out, float* __restrict__ mu, float* __restrict__ u, WPtr Zc) {
  extern __shared__ float shw[];  // n weights
  __shared__ float sh[32];
  __shared__ float part[kMcGroups][kMcCols];
  const int b = blockIdx.y;
  float tl = 0.f, sl = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float wi = w[(long long)b * n + i];
    shw[i] = wi;
    sl += wi;
    tl += wdiag[(long long)b * n + i];
  }
  const float t = block_sum(tl, sh);
  const float sw = block_sum(sl, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) { t_out[b] = t; sw_out[b] = sw; }
  const int tx = threadIdx.x % kMcCols, g = threadIdx.x / kMcCols;
  const int j = blockIdx.x * kMcCols + tx;
  const bool ok = j < d;
  const float* z = Z + (long long)b * n * d + j;
  float a = 0.f;
  if (ok)
    for (int i = g; i < n; i += kMcGroups) a = fmaf(shw[i], z[(long long)i * d], a);
  part[g][tx] = a;
  __syncthreads();
  a = 0.f;
#pragma unroll
  for (int q = 0; q < kMcGroups; ++q) a += part[q][tx];   // same order in every group: one value of mu
  __syncthreads();
  const float inv = 1.f / (t + eps);
  const float m = a * inv;
  float ua = 0.f;
  if (ok) {
    if (g == 0) mu[(long long)b * d + j] = m;
    const long long o = (long long)b * Zc.bs + j;
    for (int i = g; i < n; i += kMcGroups) {
      const float zc = z[(long long)i * d] - m;
      ua = fmaf(zc, shw[i], ua);
      wstore(Zc, o + (long long)i * Zc.ld, zc);
    }
  }
  if (u) {   // kernel-uniform
    part[g][tx] = ua;
    __syncthreads();
    if (g == 0 && ok) {
      float us = 0.f;
#pragma unroll
      for (int q = 0; q < kMcGroups; ++q) us += part[q][tx];
      u[(long long)b * d + j] = us * inv;
    }
  }
}

// same, two adjacent columns per thread (8-byte loads, packed bf16x2 stores): d and Zc.ld even
__global__ void __launch_bounds__(kMcCols * kMcGroups)
mean_center2_kernel(const float* __restrict__ Z, const float* __restrict__ w,
                    const float* __restrict__ wdiag, int n, int d, float eps, float* __restrict__ t_out,
                    float* __restrict__ sw_out, float* __restrict__ mu, float* __restrict__ u, WPtr Zc) {
  extern __shared__ float shw[];  // n weights
  __shared__ float sh[32];
  __shared__ float2 part[kMcGroups][kMcCols];
  const int b = blockIdx.y;
  float tl = 0.f, sl = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float wi = w[(long long)b * n + i];
    shw[i] = wi;
    sl += wi;
    tl += wdiag[(long long)b * n + i];
  }
  const float t = block_sum(tl, sh);
  const float sw = block_sum(sl, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) { t_out[b] = t; sw_out[b] = sw; }
  const int tx = threadIdx.x % kMcCols, g = threadIdx.x / kMcCols;
  const int j = (blockIdx.x * kMcCols + tx) * 2;
  const bool ok = j < d;
  const float* z = Z + (long long)b * n * d + j;
  float2 a = make_float2(0.f, 0.f);
  if (ok)
    for (int i = g; i < n; i += kMcGroups) {
      const float2 v = *reinterpret_cast<const float2*>(z + (long long)i * d);
      a.x = fmaf(shw[i], v.x, a.x);
      a.y = fmaf(shw[i], v.y, a.y);
    }
  part[g][tx] = a;
  __syncthreads();
  a = make_float2(0.f, 0.f);
#pragma unroll
  for (int q = 0; q < kMcGroups; ++q) { a.x += part[q][tx].x; a.y += part[q][tx].y; }
  __syncthreads();
  const float inv = 1.f / (t + eps);
  const float m0 = a.x * inv, m1 = a.y * inv;
  float2 ua = make_float2(0.f, 0.f);
  if (ok) {
    if (g == 0) *reinterpret_cast<float2*>(mu + (long long)b * d + j) = make_float2(m0, m1);
    const long long o = (long long)b * Zc.bs + j;
    for (int i = g; i < n; i += kMcGroups) {
      const float2 v = *reinterpret_cast<const float2*>(z + (long long)i * d);
      const float c0 = v.x - m0, c1 = v.y - m1;
      ua.x = fmaf(c0, shw[i], ua.x);
      ua.y = fmaf(c1, shw[i], ua.y);
      wstore2(Zc, o + (long long)i * Zc.ld, c0, c1);
    }
  }
  if (u) {   // kernel-uniform
    part[g][tx] = ua;
    __syncthreads();
    if (g == 0 && ok) {
      float2 us = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kMcGroups; ++q) { us.x += part[q][tx].x; us.y += part[q][tx].y; }
      *reinterpret_cast<float2*>(u + (long long)b * d + j) = make_float2(us.x * inv, us.y * inv);
    }
  }
}

// same, four adjacent columns per thread (16-byte loads, 8-byte packed plane stores) and 16 row groups:
// d % 4 == 0. More bytes in flight per thread for the same 1536 blocks of 512 threads.
constexpr int kMc4Cols = 32, kMc4Groups = 16;
__global__ void __launch_bounds__(kMc4Cols * kMc4Groups)
mean_center4_kernel(const float* __restrict__ Z, const float* __restrict__ w,
                    const float* __restrict__ wdiag, int n, int d, float eps, float* __restrict__ t_out,
                    float* __restrict__ sw_out, float* __restrict__ mu, float* __restrict__ u, WPtr Zc) {
  extern __shared__ float shw[];  // n weights
  __shared__ float sh[32];
  __shared__ float4 part[kMc4Groups][kMc4Cols];
  const int b = blockIdx.y;
  float tl = 0.f, sl = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float wi = w[(long long)b * n + i];
    shw[i] = wi;
    sl += wi;
    tl += wdiag[(long long)b * n + i];
  }
  const float t = block_sum(tl, sh);
  const float sw = block_sum(sl, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) { t_out[b] = t; sw_out[b] = sw; }
  const int tx = threadIdx.x % kMc4Cols, g = threadIdx.x / kMc4Cols;
  const int j = (blockIdx.x * kMc4Cols + tx) * 4;
  const bool ok = j < d;
  const float* z = Z + (long long)b * n * d + j;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok)
    for (int i = g; i < n; i += kMc4Groups) {
      const float4 v = *reinterpret_cast<const float4*>(z + (long long)i * d);
      const float wi = shw[i];
      a.x = fmaf(wi, v.x, a.x); a.y = fmaf(wi, v.y, a.y); a.z = fmaf(wi, v.z, a.z); a.w = fmaf(wi, v.w, a.w);
    }
  part[g][tx] = a;
  __syncthreads();
  a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < kMc4Groups; ++q) {     // same order in every group: one value of mu
    const float4 v = part[q][tx];
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  __syncthreads();
  const float inv = 1.f / (t + eps);
  const float4 m = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  float4 ua = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
    if (g == 0) *reinterpret_cast<float4*>(mu + (long long)b * d + j) = m;
    const long long o = (long long)b * Zc.bs + j;
    for (int i = g; i < n; i += kMc4Groups) {
      const float4 v = *reinterpret_cast<const float4*>(z + (long long)i * d);
      const float4 c = make_float4(v.x - m.x, v.y - m.y, v.z - m.z, v.w - m.w);
      const float wi = shw[i];
      ua.x = fmaf(c.x, wi, ua.x); ua.y = fmaf(c.y, wi, ua.y); ua.z = fmaf(c.z, wi, ua.z); ua.w = fmaf(c.w, wi, ua.w);
      wstore4(Zc, o + (long long)i * Zc.ld, c);
    }
  }
  if (u) {   // kernel-uniform
    part[g][tx] = ua;
    __syncthreads();
    if (g == 0 && ok) {
      float4 us = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kMc4Groups; ++q) {
        const float4 v = part[q][tx];
        us.x += v.x; us.y += v.y; us.z += v.z; us.w += v.w;
      }
      *reinterpret_cast<float4*>(u + (long long)b * d + j) = make_float4(us.x * inv, us.y * inv, us.z * inv, us.w * inv);
    }
  }
}

// ------------------------------------------------------------ trace, dots
__global__ void trace_scales_kernel(const float* __restrict__ M, int d, float eps, int post_mode,
                                    float* __restrict__ tr, float* __restrict__ inv,
                                    float* __restrict__ post) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* m = M + (long long)b * d * d;
  float a = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) a += m[(long long)i * d + i];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) {
    tr[b] = a;
    inv[b] = 1.f / (a + eps);
    const float r = sqrtf(a + eps);
    post[b] = post_mode == 0 ? 1.f / r : r;
  }
}
__global__ void batch_dot_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                 long long n_per, float* __restrict__ partial) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const float* x = X + (long long)b * n_per;
  const float* y = Y + (long long)b * n_per;
  float a = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_per;
       i += (long long)gridDim.x * blockDim.x)
    a = fmaf(x[i], y[i], a);
  a = block_sum(a, sh);
  if (threadIdx.x == 0) partial[(long long)b * gridDim.x + blockIdx.x] = a;
}
__global__ void ns_bwd_finish_kernel(const float* __restrict__ dA, const float* __restrict__ inv,
                                     const float* __restrict__ dotO, const float* __restrict__ dotA,
                                     float coef_tau, int d, float* __restrict__ dM) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const long long o = ((long long)b * d + i) * d + j;
  float v = dA[o] * inv[b];
  if (i == j) v += (coef_tau * dotO[b] - dotA[b] * inv[b]) * inv[b];
  dM[o] = v;
}

__global__ void clamped_degree_kernel(const float* __restrict__ G, int n, float eps,
                                      float* __restrict__ deg) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* g = G + ((long long)b * n + row) * n;
  float a = 0.f;
  for (int j = lane; j < n; j += 32) a += g[j];
  a = warp_sum(a);
  if (lane == 0) deg[(long long)b * n + row] = fmaxf(a, eps);
}
__global__ void normalize_graph_kernel(const float* __restrict__ G, const float* __restrict__ deg,
                                       int n, int method, float* __restrict__ out) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const long long o = ((long long)b * n + i) * n + j;
  const float di = deg[(long long)b * n + i];
  if (method == 0) {
    const float ii = 1.f / sqrtf(di), ij = 1.f / sqrtf(deg[(long long)b * n + j]);
    out[o] = G[o] * ii * ij;
  } else {
    out[o] = G[o] * (1.f / di);
  }
}
__global__ void batch_trace_kernel(const float* __restrict__ M, int d, float* __restrict__ tr) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  const float* m = M + (long long)b * d * d;
  float a = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) a += m[(long long)i * d + i];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) tr[b] = a;
}

// ------------------------------------------------------------------- triu
constexpr int kTriuRows = 8;
__global__ void __launch_bounds__(256) triu_pack_kernel(const float* __restrict__ O, int d,
                                                        float* __restrict__ v) {
  const int b = blockIdx.y;
  const long long L = (long long)d * (d + 1) / 2;
  const int i0 = blockIdx.x * kTriuRows;
  for (int r = 0; r < kTriuRows; ++r) {
    const int i = i0 + r;
    if (i >= d) break;
    const float* src = O + ((long long)b * d + i) * d;
    float* dst = v + (long long)b * L + (long long)i * d - (long long)i * (i - 1) / 2 - i;
    for (int j = i + threadIdx.x; j < d; j += blockDim.x) dst[j] = src[j];
  }
}
__global__ void __launch_bounds__(256) triu_unpack_kernel(const float* __restrict__ dv, int d,
                                                          float* __restrict__ dO) {
  const int b = blockIdx.y;
  const long long L = (long long)d * (d + 1) / 2;
  const int i0 = blockIdx.x * kTriuRows;
  for (int r = 0; r < kTriuRows; ++r) {
    const int i = i0 + r;
    if (i >= d) break;
    float* dst = dO + ((long long)b * d + i) * d;
    const float* src = dv + (long long)b * L + (long long)i * d - (long long)i * (i - 1) / 2 - i;
    for (int j = threadIdx.x; j < d; j += blockDim.x) dst[j] = (j >= i) ? src[j] : 0.f;
  }
}

// packed upper triangle [B][L] (fp32 gradient of the half-vector) -> full D x D working matrix
// out = s_b[b] * unpack(dv) with a zero lower triangle; optional per-block partial of
// trace(unpack(dv)) (fixed order: tr_partial[b * gridDim.x + blockIdx.x]).
constexpr int kUnpackRows = 8;
__global__ void __launch_bounds__(256)
triu_unpack_planes_kernel(const float* __restrict__ dv, long long ld_dv, int d,
                          const float* __restrict__ s_b, WPtr out, float* __restrict__ tr_partial) {
  __shared__ float sh[32];
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * kUnpackRows;
  const float sc = s_b ? s_b[b] : 1.f;
  const int groups = (int)(out.ld / 8);        // 8-column groups per row (ld % 8 == 0)
  const float* src_b = dv + (long long)b * ld_dv;
  float tr = 0.f;
  for (int u = threadIdx.x; u < kUnpackRows * groups; u += blockDim.x) {
    const int i = i0 + u / groups;
    if (i >= d) break;
    const int j0 = (u % groups) * 8;
    const float* src = src_b + (long long)i * d - (long long)i * (i - 1) / 2 - i;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q;
      const float x = (j >= i && j < d) ? src[j] : 0.f;
      if (j == i) tr += x;
      v[q] = sc * x;
    }
    const long long o = (long long)b * out.bs + (long long)i * out.ld + j0;
    if (out.f) {
      *reinterpret_cast<float4*>(out.f + o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out.f + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * q]), h1 = __float2bfloat16_rn(v[2 * q + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * q] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * q + 1] - __bfloat162float(h1));
        hw[q] = __bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lw[q] = __bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint4*>(out.hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      if (out.lo) *reinterpret_cast<uint4*>(out.lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
  if (tr_partial) {
    tr = block_sum(tr, sh);
    if (threadIdx.x == 0) tr_partial[(long long)b * gridDim.x + blockIdx.x] = tr;
  }
}
// out = scale * s_b[b] * sym(unpack(dv[b])) with sym(X) = (X + X^T)/2, written in the symmetric
// block storage of the GEMM engine: only 8-column groups inside the upper 256 x 256 blocks
// (block column >= block row) are stored; the blocks below are never read by the engine.
__global__ void __launch_bounds__(256)
triu_unpack_sym_planes_kernel(const float* __restrict__ dv, long long ld_dv, int d,
                              const float* __restrict__ s_b, float scale, WPtr out) {
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * kUnpackRows;
  const float sc = scale * (s_b ? s_b[b] : 1.f);
  const int groups = (int)(out.ld / 8);
  const float* src_b = dv + (long long)b * ld_dv;
  for (int u = threadIdx.x; u < kUnpackRows * groups; u += blockDim.x) {
    const int i = i0 + u / groups;
    if (i >= d) break;
    const int j0 = (u % groups) * 8;
    if ((j0 >> 8) < (i >> 8)) continue;          // absent block
    const float* src = src_b + (long long)i * d - (long long)i * (i - 1) / 2 - i;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q;
      float x = 0.f;
      if (j < d) {
        if (j >= i) x = src[j];
        else x = src_b[(long long)j * d - (long long)j * (j - 1) / 2 + (i - j)];   // mirror (diagonal blocks)
      }
      v[q] = (j == i ? sc : 0.5f * sc) * x;
    }
    const long long o = (long long)b * out.bs + (long long)i * out.ld + j0;
    if (out.f) {
      *reinterpret_cast<float4*>(out.f + o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out.f + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * q]), h1 = __float2bfloat16_rn(v[2 * q + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * q] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * q + 1] - __bfloat162float(h1));
        hw[q] = __bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lw[q] = __bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint4*>(out.hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      if (out.lo) *reinterpret_cast<uint4*>(out.lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
  }
}
// fp32 matrix -> packed upper triangle as a working-matrix row (the Linear's operand)
__global__ void __launch_bounds__(256)
triu_pack_planes_kernel(const float* __restrict__ O, int d, WPtr out) {
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * kUnpackRows;
  for (int r = 0; r < kUnpackRows; ++r) {
    const int i = i0 + r;
    if (i >= d) break;
    const float* src = O + ((long long)b * d + i) * d;
    const long long o = (long long)b * out.ld + (long long)i * d - (long long)i * (i - 1) / 2 - i;
    for (int j = i + threadIdx.x; j < d; j += blockDim.x) wstore(out, o + j, src[j]);
  }
}
// out[b] = sum_n dy[b,n] * (y[b,n] - bias[n])   ( = <dx_b, x_b> for y = x W^T + bias, dx = dy W )
__global__ void rowdot_bias_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                   const float* __restrict__ bias, int batch, int n,
                                   float* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const int lane = threadIdx.x & 31;
  float a = 0.f;
  for (int j = lane; j < n; j += 32)
    a = fmaf(dy[(long long)b * n + j], y[(long long)b * n + j] - (bias ? bias[j] : 0.f), a);
  a = warp_sum(a);
  if (lane == 0) out[b] = a;
}
// tau -> tr, inv = 1/(tau+eps), post = (tau+eps)^-1/2        (scal rows 0,1,2 of [3,B])
__global__ void mh_scalars_fwd_kernel(const float* __restrict__ tau, int batch, float eps,
                                      float* __restrict__ scal) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float t = tau[b];
  scal[b] = t;
  scal[batch + b] = 1.f / (t + eps);
  scal[2 * batch + b] = 1.f / sqrtf(t + eps);
}
// dM = inv dA + dtau I  with  dtau = (-1/2 <dO,O> - <dA,A>) inv
__global__ void mh_dtau_kernel(const float* __restrict__ scal, int batch, const float* __restrict__ dotO,
                               const float* __restrict__ dotA, float* __restrict__ dtau) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  dtau[b] = (-0.5f * dotO[b] - dotA[b]) * scal[batch + b];
}
// out[b] = sum of nper consecutive partials
__global__ void sum_partials_kernel(const float* __restrict__ partial, int nper, int batch,
                                    float* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const int lane = threadIdx.x & 31;
  float a = 0.f;
  for (int i = lane; i < nper; i += 32) a += partial[(long long)b * nper + i];
  a = warp_sum(a);
  if (lane == 0) out[b] = a;
}

// ----------------------------------------------------------------- sketch
__global__ void sketch_fwd_kernel(const float* __restrict__ x, int batch, int d, int S,
                                  const int* __restrict__ off, const int* __restrict__ idx,
                                  const float* __restrict__ sgn, float* __restrict__ cs,
                                  float* __restrict__ out) {
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float* xb = x + (long long)b * d;
  float prod = 1.f;
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    const int* o = off + (long long)h * (S + 1);
    float a = 0.f;
    for (int e = o[s]; e < o[s + 1]; ++e) a += sgn[h * d + e] * xb[idx[h * d + e]];
    cs[((long long)h * batch + b) * S + s] = a;
    prod *= a;
  }
  out[(long long)b * S + s] = prod;
}
__global__ void sketch_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ cs,
                                  int batch, int d, int S, const long long* __restrict__ hash,
                                  const long long* __restrict__ sign, float* __restrict__ dx) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  float a = 0.f;
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    const long long s = hash[(long long)h * d + j];
    float others = 1.f;
#pragma unroll
    for (int g = 0; g < 3; ++g)
      if (g != h) others *= cs[((long long)g * batch + b) * S + s];
    a += static_cast<float>(sign[(long long)h * d + j]) * dout[(long long)b * S + s] * others;
  }
  dx[(long long)b * d + j] = a;
}

// -------------------------------------------------- low-rank chain helpers
__global__ void w_fill_eye_kernel(WPtr out, int n, float c) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  wstore(out, (long long)b * out.bs + (long long)i * out.ld + j, i == j ? c : 0.f);
}
__global__ void w_lincomb_kernel(WPtr out, int rows, int cols, float a, WPtr X, float bb, WPtr Y,
                                 int hasY, float c, WPtr Z, int hasZ, const float* __restrict__ s_b) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  float v = a * wload(X, (long long)b * X.bs + (long long)i * X.ld + j);
  if (hasY) v += bb * wload(Y, (long long)b * Y.bs + (long long)i * Y.ld + j);
  if (hasZ) v += c * wload(Z, (long long)b * Z.bs + (long long)i * Z.ld + j);
  if (s_b) v *= s_b[b];
  wstore(out, (long long)b * out.bs + (long long)i * out.ld + j, v);
}
__global__ void w_dot_kernel(WPtr X, WPtr Y, int rows, int cols, float* __restrict__ out) {
  __shared__ float sh[32];
  const int b = blockIdx.x;
  float a = 0.f;
  for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
    const int i = e / cols, j = e - i * cols;
    a = fmaf(wload(X, (long long)b * X.bs + (long long)i * X.ld + j),
             wload(Y, (long long)b * Y.bs + (long long)i * Y.ld + j), a);
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0) out[b] = a;
}
__global__ void mlr_scalars_fwd_kernel(const float* __restrict__ tau, int batch, float eps, float aK,
                                       float* __restrict__ scal) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float taup = tau[b] + eps;
  const float inv = 1.f / taup, post = rsqrtf(taup);
  scal[b] = taup;
  scal[batch + b] = inv;
  scal[2 * batch + b] = post;
  scal[3 * batch + b] = post * inv;
  scal[4 * batch + b] = post * aK;
}
__global__ void mlr_scalars_bwd_kernel(const float* __restrict__ scal, int batch, float aK,
                                       const float* __restrict__ dotOO, const float* __restrict__ trdO,
                                       const float* __restrict__ dotHs, float* __restrict__ dtaup) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float taup = scal[b], inv = scal[batch + b], post = scal[2 * batch + b];
  const float c1 = scal[3 * batch + b], betaK = scal[4 * batch + b];
  const float dc1 = (dotOO[b] - betaK * trdO[b]) / c1;
  const float dpost = aK * trdO[b] + inv * dc1;
  // dHs = inv * dH  =>  <dH, H> * taup = <dHs, H> * taup^2
  const float dinv = post * dc1 + (dotHs ? dotHs[b] * taup * taup : 0.f);
  dtaup[b] = -inv * inv * dinv - 0.5f * post * inv * dpost;   // taup^-1.5 = post * inv
}

__global__ void reduce_splits_kernel(const float* __restrict__ partial, int splits, long long mn, int n,
                                     const float* __restrict__ bias, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float a = bias ? bias[i % n] : 0.f;
  for (int s = 0; s < splits; ++s) a += partial[(long long)s * mn + i];
  y[i] = a;
}
__global__ void colsum_kernel(const float* __restrict__ x, int m, int n, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float a = 0.f;
  for (int i = 0; i < m; ++i) a += x[(long long)i * n + j];
  out[j] = a;
}

// ------------------------------------------------- BatchNorm1d -> GELU -> Dropout (feature nets)
// second_net / third_net after their Linear (moment_head.py:186-191,195-200): BatchNorm1d (train: batch
// statistics, biased variance for the normalisation, unbiased for the running update, momentum 0.1;
// eval: running statistics), exact (erf) GELU, inverted Dropout. One block owns 32 features for all
// rows, so the per-feature reductions never leave the block: deterministic, one launch each way.
constexpr int kFtCols = 32, kFtRows = 32;
__device__ __forceinline__ float gelu_f(float z) { return 0.5f * z * (1.f + erff(z * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_f(float z) {
  return 0.5f * (1.f + erff(z * 0.70710678118654752f)) + z * __expf(-0.5f * z * z) * 0.39894228040143268f;
}
// keep-mask of inverted dropout: a counter-based hash of (seed, element index) -> uniform [0,1)
__device__ __forceinline__ float drop_scale(unsigned long long seed, long long idx, float p, float inv_keep) {
  if (p <= 0.f) return 1.f;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.f / 16777216.f);
  return u >= p ? inv_keep : 0.f;
}
// column sums over the block's rows: sh[kFtRows][kFtCols]; result in every thread of the column
__device__ __forceinline__ float ft_colsum(float v, float (*sh)[kFtCols], int tx, int ty) {
  __syncthreads();
  sh[ty][tx] = v;
  __syncthreads();
  float a = 0.f;
#pragma unroll
  for (int r = 0; r < kFtRows; ++r) a += sh[r][tx];
  return a;
}
__global__ void __launch_bounds__(kFtCols * kFtRows)
feature_tail_fwd_kernel(const float* __restrict__ y, int m, int n, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ run_mean,
                        float* __restrict__ run_var, int training, float momentum, float bn_eps, float drop_p,
                        unsigned long long seed, float* __restrict__ out, float* __restrict__ save_mean,
                        float* __restrict__ save_rstd) {
  __shared__ float sh[kFtRows][kFtCols];
  const int tx = threadIdx.x & (kFtCols - 1), ty = threadIdx.x / kFtCols;
  const int j = blockIdx.x * kFtCols + tx;
  const bool ok = j < n;
  float mean, rstd;
  if (training) {
    float a = 0.f;
    for (int i = ty; i < m; i += kFtRows) a += ok ? y[(long long)i * n + j] : 0.f;
    mean = ft_colsum(a, sh, tx, ty) / (float)m;
    float v = 0.f;
    for (int i = ty; i < m; i += kFtRows) {
      const float d = ok ? y[(long long)i * n + j] - mean : 0.f;
      v = fmaf(d, d, v);
    }
    const float var = ft_colsum(v, sh, tx, ty) / (float)m;
    rstd = rsqrtf(var + bn_eps);
    if (ok && ty == 0 && run_mean) {
      run_mean[j] = (1.f - momentum) * run_mean[j] + momentum * mean;
      run_var[j] = (1.f - momentum) * run_var[j] + momentum * var * ((float)m / (float)(m > 1 ? m - 1 : 1));
    }
  } else {
    mean = ok ? run_mean[j] : 0.f;
    rstd = ok ? rsqrtf(run_var[j] + bn_eps) : 0.f;
  }
  if (ok && ty == 0) { save_mean[j] = mean; save_rstd[j] = rstd; }
  if (!ok) return;
  const float g = gamma ? gamma[j] : 1.f, bt = beta ? beta[j] : 0.f;
  const float inv_keep = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  for (int i = ty; i < m; i += kFtRows) {
    const long long o = (long long)i * n + j;
    const float z = (y[o] - mean) * rstd * g + bt;
    out[o] = gelu_f(z) * (training ? drop_scale(seed, o, drop_p, inv_keep) : 1.f);
  }
}
__global__ void __launch_bounds__(kFtCols * kFtRows)
feature_tail_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y, int m, int n,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                        int training, float drop_p, unsigned long long seed, float* __restrict__ dy,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sh[kFtRows][kFtCols];
  const int tx = threadIdx.x & (kFtCols - 1), ty = threadIdx.x / kFtCols;
  const int j = blockIdx.x * kFtCols + tx;
  const bool ok = j < n;
  const float mean = ok ? save_mean[j] : 0.f, rstd = ok ? save_rstd[j] : 0.f;
  const float g = (ok && gamma) ? gamma[j] : 1.f, bt = (ok && beta) ? beta[j] : 0.f;
  const float inv_keep = drop_p < 1.f ? 1.f / (1.f - drop_p) : 0.f;
  // dz = dout * mask * gelu'(z);  dbeta = sum dz;  dgamma = sum dz * xhat
  float sb = 0.f, sg = 0.f;
  for (int i = ty; i < m; i += kFtRows) {
    if (!ok) break;
    const long long o = (long long)i * n + j;
    const float xh = (y[o] - mean) * rstd;
    const float dz = dout[o] * (training ? drop_scale(seed, o, drop_p, inv_keep) : 1.f) * dgelu_f(xh * g + bt);
    sb += dz;
    sg = fmaf(dz, xh, sg);
  }
  const float db = ft_colsum(sb, sh, tx, ty);
  const float dg = ft_colsum(sg, sh, tx, ty);
  if (ok && ty == 0) {
    if (dbeta) dbeta[j] = db;
    if (dgamma) dgamma[j] = dg;
  }
  if (!ok) return;
  const float inv_m = 1.f / (float)m;
  for (int i = ty; i < m; i += kFtRows) {
    const long long o = (long long)i * n + j;
    const float xh = (y[o] - mean) * rstd;
    const float dz = dout[o] * (training ? drop_scale(seed, o, drop_p, inv_keep) : 1.f) * dgelu_f(xh * g + bt);
    // train: d xhat = g dz; dy = rstd (d xhat - mean(d xhat) - xhat mean(d xhat xhat)); eval: constants
    dy[o] = training ? rstd * g * (dz - db * inv_m - xh * dg * inv_m) : dz * g * rstd;
  }
}

// ---------------------------------------------------------- pooling backward
__global__ void __launch_bounds__(128)
pool_bwd_dmu_kernel(const float* __restrict__ dZc, const float* __restrict__ du,
                    const float* __restrict__ sw, const float* __restrict__ t, int n, int d,
                    float eps, float* __restrict__ dmu) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const float* p = dZc + (long long)b * n * d + j;
  float a = 0.f;
  for (int i = 0; i < n; ++i) a += p[(long long)i * d];
  if (du) a += sw[b] * du[(long long)b * d + j] / (t[b] + eps);
  dmu[(long long)b * d + j] = -a;
}
__global__ void pool_bwd_rows_kernel(const float* __restrict__ dZc, const float* __restrict__ Z,
                                     WPtr Zc, const float* __restrict__ w, const float* __restrict__ t,
                                     const float* __restrict__ mu, const float* __restrict__ u,
                                     const float* __restrict__ du, const float* __restrict__ dmu,
                                     int n, int d, float eps, float* __restrict__ dZ,
                                     float* __restrict__ dw, float* __restrict__ dt) {
  // one warp per token row; block (0, b) additionally reduces dt
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const float inv = 1.f / (t[b] + eps);
  const float* dmub = dmu + (long long)b * d;
  const float* dub = du ? du + (long long)b * d : nullptr;
  if (row < n) {
    const float wi = w[(long long)b * n + row] * inv;
    const long long o = ((long long)b * n + row) * d;
    const long long oc = (long long)b * Zc.bs + (long long)row * Zc.ld;
    float a = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float dmj = dmub[j];
      const float duj = dub ? dub[j] : 0.f;
      dZ[o + j] = dZc[o + j] + wi * (duj + dmj);
      a = fmaf(Z[o + j], dmj, a);
      if (dub) a = fmaf(wload(Zc, oc + j), duj, a);
    }
    a = warp_sum(a);
    if (lane == 0) dw[(long long)b * n + row] = a * inv;
  }
  if (blockIdx.x == 0 && (threadIdx.x >> 5) == 0) {
    float a = 0.f;
    for (int j = lane; j < d; j += 32) {
      a = fmaf(mu[(long long)b * d + j], dmub[j], a);
      if (dub) a = fmaf(u[(long long)b * d + j], dub[j], a);
    }
    a = warp_sum(a);
    if (lane == 0) dt[b] = -a * inv;
  }
}
__global__ void pool_bwd_ds_kernel(const float* __restrict__ dW, long long ldW,
                                   const float* __restrict__ dw, const float* __restrict__ dt,
                                   const float* __restrict__ G, const float* __restrict__ s, int n,
                                   int sym, float* __restrict__ ds) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* dWb = dW + (long long)b * n * ldW;
  const float* Gb = G + (long long)b * n * n;
  const float* sb = s + (long long)b * n;
  const float* dwb = dw + (long long)b * n;
  const float dtb = dt[b];
  const int i = row;
  float a = 0.f;
  if (sym) {
    // G and dW symmetric (EGM_MHD_SYMMETRIC_GRAPH): the mirrored elements are the row's own
    for (int j = lane; j < n; j += 32) {
      const float d = dWb[(long long)i * ldW + j] + (i == j ? dtb : 0.f);
      a += (2.f * d + dwb[i] + dwb[j]) * Gb[(long long)i * n + j] * sb[j];
    }
  } else {
    for (int j = lane; j < n; j += 32) {
      const float dij = dWb[(long long)i * ldW + j] + dwb[i] + (i == j ? dtb : 0.f);
      const float dji = dWb[(long long)j * ldW + i] + dwb[j] + (i == j ? dtb : 0.f);
      a += (dij * Gb[(long long)i * n + j] + dji * Gb[(long long)j * n + i]) * sb[j];
    }
  }
  a = warp_sum(a);
  if (lane == 0) ds[(long long)b * n + i] = a;
}
// A block walks 32 rows of one image (a warp takes 4 of them); the per-token vectors s, dw and
// ddeg = -0.5 s^3 ds [deg >= eps] are staged in shared memory once per block, so an element costs one
// global load (dW) and one store.
constexpr int kDgRows = 32;
__global__ void __launch_bounds__(256)
pool_bwd_dG_kernel(const float* __restrict__ dW, long long ldW, const float* __restrict__ dw,
                   const float* __restrict__ dt, const float* __restrict__ s,
                   const float* __restrict__ deg, const float* __restrict__ ds, int n, float eps,
                   int sym, float* __restrict__ dG) {
  extern __shared__ float shv[];          // s[n], dw[n], ddeg[n]
  float* ss = shv;
  float* sdw = shv + n;
  float* sdd = shv + 2 * n;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float si = s[(long long)b * n + i];
    ss[i] = si;
    sdw[i] = dw[(long long)b * n + i];
    // s = max(deg,eps)^(-1/2): d s/d deg = -0.5 s^3 where the clamp is inactive (deg >= eps)
    sdd[i] = (deg[(long long)b * n + i] >= eps) ? -0.5f * si * si * si * ds[(long long)b * n + i] : 0.f;
  }
  __syncthreads();
  const float dtb = dt[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < kDgRows; r += 8) {
    const int i = blockIdx.x * kDgRows + r;
    if (i >= n) break;
    const float* dWr = dW + ((long long)b * n + i) * ldW;
    float* out = dG + ((long long)b * n + i) * n;
    const float si = ss[i], dwi = sdw[i], ddeg = sdd[i];
    if (sym) {
      // symmetric graph: return the symmetric part (dG + dG^T)/2 of the reference's gradient, i.e. the
      // gradient with respect to a symmetric matrix (dW is symmetric here; row terms are averaged)
      for (int j = lane; j < n; j += 32) {
        const float dwf = dWr[j] + 0.5f * (dwi + sdw[j]) + (i == j ? dtb : 0.f);
        out[j] = si * dwf * ss[j] + 0.5f * (ddeg + sdd[j]);
      }
    } else {
      for (int j = lane; j < n; j += 32) {
        const float dwf = dWr[j] + dwi + (i == j ? dtb : 0.f);
        out[j] = si * dwf * ss[j] + ddeg;
      }
    }
  }
}

// --------------------------------------------------------- graph alignment loss
// g[b] = mean(G[b]): one block per image, fixed summation order
__global__ void __launch_bounds__(1024)
graph_mean_kernel(const float* __restrict__ G, long long per, float* __restrict__ g) {
  __shared__ float sh[32];
  const float* src = G + (long long)blockIdx.x * per;
  float a = 0.f;
  for (long long i = threadIdx.x; i < per; i += blockDim.x) a += src[i];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) g[blockIdx.x] = a / (float)per;
}
// row i of S = sigmoid(g g^T) against L_ij = [labels_i == labels_j]:
//   rowloss_i = sum_j (S_ij - L_ij)^2 / B^2,  dg_i = (4/B^2) sum_j (S_ij - L_ij) S_ij (1 - S_ij) g_j
// (S and L are symmetric, so the (i,j) and (j,i) terms of d/dg_i coincide)
__global__ void __launch_bounds__(256)
align_rows_kernel(const float* __restrict__ g, const long long* __restrict__ labels, int batch,
                  float* __restrict__ rowloss, float* __restrict__ dg) {
  __shared__ float sh[32];
  const int i = blockIdx.x;
  const float gi = g[i];
  const long long li = labels[i];
  float l = 0.f, d = 0.f;
  for (int j = threadIdx.x; j < batch; j += blockDim.x) {
    const float gj = g[j];
    const float s = 1.f / (1.f + __expf(-gi * gj));
    const float e = s - (labels[j] == li ? 1.f : 0.f);
    l = fmaf(e, e, l);
    d = fmaf(e * s * (1.f - s), gj, d);
  }
  l = block_sum(l, sh);
  d = block_sum(d, sh);
  if (threadIdx.x == 0) {
    const float inv = 1.f / ((float)batch * (float)batch);
    rowloss[i] = l * inv;
    dg[i] = 4.f * d * inv;
  }
}
// dG[b, :, :] = dloss * dg[b] / per
__global__ void align_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ dloss,
                                 long long per, float* __restrict__ dG) {
  const float v = dloss[0] * dg[blockIdx.y] / (float)per;
  float* dst = dG + (long long)blockIdx.y * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = v;
}

}  // namespace

// ================================================================ launchers
void affine(const float* x, long long ld, long long bs, int batch, int rows, int cols,
            const float* s, float a1, float b1, const W& out1, float a2, float b2, const W* out2,
            int prec, cudaStream_t st) {
  WPtr o1 = wptr(out1, prec), o2 = out2 ? wptr(*out2, prec) : WPtr{};
  const bool vec = (cols % 4 == 0) && (ld % 4 == 0) && (bs % 4 == 0) && aligned16(x) && aligned16(out1.base) &&
                   (!out2 || aligned16(out2->base));
  if (vec) {
    dim3 grid((cols / 4 + 127) / 128, rows, batch);
    affine_kernel<4><<<grid, 128, 0, st>>>(x, ld, bs, rows, cols, s, a1, b1, o1, a2, b2, o2, out2 != nullptr);
  note_launch();
  } else {
    dim3 grid((cols + 127) / 128, rows, batch);
    affine_kernel<1><<<grid, 128, 0, st>>>(x, ld, bs, rows, cols, s, a1, b1, o1, a2, b2, o2, out2 != nullptr);
  note_launch();
  }
}
void export_f32(const W& in, float* out, long long ld, long long bs, int prec, cudaStream_t st) {
  dim3 grid((in.cols + 127) / 128, in.rows, in.batch);
  export_kernel<1><<<grid, 128, 0, st>>>(wptr(in, prec), in.rows, in.cols, out, ld, bs);
  note_launch();
}
void rownorm(const float* x, int batch, int n, int d, float eps, int cosine, float* nrm, const W& xn,
             int prec, cudaStream_t st) {
  dim3 grid((n + 7) / 8, batch);
  rownorm_kernel<<<grid, 256, 0, st>>>(x, n, d, eps, cosine, nrm, wptr(xn, prec));
  note_launch();
}
void rownorm_bwd(const float* x, const float* nrm, const float* dxn, int batch, int n, int d,
                 float eps, float* dx, cudaStream_t st) {
  dim3 grid((n + 7) / 8, batch);
  rownorm_bwd_kernel<<<grid, 256, 0, st>>>(x, nrm, dxn, n, d, eps, dx);
  note_launch();
}
void gpf_poly_fwd(const float* Ra, const float* Rp, long long ldR, const float* coef, int P, int Q,
                  int symmetric, int batch, int n, float* G, cudaStream_t st) {
  dim3 grid(poly_pairs(n), batch);
  if (P <= 3 && Q <= 3 && (long long)n * ldR < (1ll << 31))
    gpf_poly3_fwd_kernel<<<grid, 256, 0, st>>>(Ra, Rp, (int)ldR, coef, P, Q, symmetric, n, G);
  else
    gpf_poly_fwd_kernel<kMaxDeg><<<grid, 256, 0, st>>>(Ra, Rp, ldR, coef, P, Q, symmetric, n, G);
  note_launch();
}
int gpf_poly_bwd_blocks(int batch, int n) { return batch * poly_pairs(n); }
bool gpf_poly_bwd_can_fold(int P, int Q, int n, long long ldR, const W& Ea) {
  return P <= 3 && Q <= 3 && (long long)n * ldR < (1ll << 31) && Ea.ld % 4 == 0;
}
size_t gpf_poly_bwd_rowpart_floats(int batch, int n) { return (size_t)batch * 2 * poly_tiles(n) * n; }
void gpf_poly_bwd(const float* dG, const float* Ra, const float* Rp, long long ldR,
                  const float* coef, int P, int Q, int symmetric, int batch, int n, const W& Ea,
                  const W& Ep, float* partial, int nblocks, float* dcoef, int prec,
                  cudaStream_t st, const float* nrm_a, const float* nrm_p, float eps, float* rowpart) {
  dim3 grid(poly_pairs(n), batch);
  if (rowpart && gpf_poly_bwd_can_fold(P, Q, n, ldR, Ea)) {
    // rows of the last (ragged) tile column that no block writes must read as zero
    cudaMemsetAsync(rowpart, 0, gpf_poly_bwd_rowpart_floats(batch, n) * sizeof(float), st);
    if (symmetric & 2)
      gpf_poly3_bwd_kernel<true, true><<<grid, 256, 0, st>>>(dG, Ra, Rp, (int)ldR, coef, P, Q, symmetric & 1, n,
                                                             wptr(Ea, prec), wptr(Ep, prec), partial, nrm_a, nrm_p,
                                                             eps, rowpart);
    else
      gpf_poly3_bwd_kernel<false, true><<<grid, 256, 0, st>>>(dG, Ra, Rp, (int)ldR, coef, P, Q, symmetric & 1, n,
                                                              wptr(Ea, prec), wptr(Ep, prec), partial, nrm_a, nrm_p,
                                                              eps, rowpart);
    note_launch();
    gpf_diag_fix_kernel<<<dim3((n + 127) / 128, batch, 2), 128, 0, st>>>(rowpart, nrm_a, nrm_p, n, poly_tiles(n), eps,
                                                                          wptr(Ea, prec), wptr(Ep, prec));
    note_launch();
    const int nt = (P + 1) * (Q + 1);
    reduce_partials_kernel<<<nt, 256, 0, st>>>(partial, nblocks, nt, dcoef);
    note_launch();
    return;
  }
  // bit 1 of `symmetric`: R_a / R_p are exactly symmetric (the fused forward wrote them)
  const bool rsym = (symmetric & 2) != 0;
  symmetric &= 1;
  if (P <= 3 && Q <= 3 && (long long)n * ldR < (1ll << 31) && Ea.ld % 4 == 0) {
    if (rsym)
      gpf_poly3_bwd_kernel<true, false><<<grid, 256, 0, st>>>(dG, Ra, Rp, (int)ldR, coef, P, Q, symmetric, n,
                                                              wptr(Ea, prec), wptr(Ep, prec), partial, nullptr,
                                                              nullptr, 0.f, nullptr);
    else
      gpf_poly3_bwd_kernel<false, false><<<grid, 256, 0, st>>>(dG, Ra, Rp, (int)ldR, coef, P, Q, symmetric, n,
                                                               wptr(Ea, prec), wptr(Ep, prec), partial, nullptr,
                                                               nullptr, 0.f, nullptr);
  } else {
    gpf_poly_bwd_kernel<kMaxDeg><<<grid, 256, 0, st>>>(dG, Ra, Rp, ldR, coef, P, Q, symmetric, n,
                                                       wptr(Ea, prec), wptr(Ep, prec), partial);
  }
  note_launch();
  const int nt = (P + 1) * (Q + 1);
  reduce_partials_kernel<<<nt, 256, 0, st>>>(partial, nblocks, nt, dcoef);
  note_launch();
}
void degree(const float* G, int batch, int n, float eps, float* deg, float* s, cudaStream_t st) {
  dim3 grid((n + 7) / 8, batch);
  degree_kernel<<<grid, 256, 0, st>>>(G, n, eps, deg, s);
  note_launch();
}
void weight(const float* G, const float* s, int batch, int n, const W& Wn, float* w, float* wdiag,
            int prec, cudaStream_t st) {
  dim3 grid((n + 7) / 8, batch);
  weight_kernel<<<grid, 256, 0, st>>>(G, s, n, wptr(Wn, prec), w, wdiag);
  note_launch();
}
void mean_center(const float* Z, const float* w, const float* wdiag, int batch, int n, int d,
                 float eps, float* t, float* sw, float* mu, float* u, const W& Zc, int prec,
                 cudaStream_t st) {
  const bool even = d % 2 == 0 && Zc.ld % 2 == 0 && (reinterpret_cast<uintptr_t>(Z) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(mu) & 7) == 0 && (!u || (reinterpret_cast<uintptr_t>(u) & 7) == 0) &&
                    (reinterpret_cast<uintptr_t>(Zc.base) & 7) == 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool quad = d % 4 == 0 && Zc.ld % 4 == 0 && al16(Z) && al16(mu) && (!u || al16(u)) && al16(Zc.base);
  if (quad) {
    dim3 grid((d / 4 + kMc4Cols - 1) / kMc4Cols, batch);
    mean_center4_kernel<<<grid, kMc4Cols * kMc4Groups, n * sizeof(float), st>>>(Z, w, wdiag, n, d, eps, t, sw, mu,
                                                                               u, wptr(Zc, prec));
  } else if (even) {
    dim3 grid((d / 2 + kMcCols - 1) / kMcCols, batch);
    mean_center2_kernel<<<grid, kMcCols * kMcGroups, n * sizeof(float), st>>>(Z, w, wdiag, n, d, eps, t, sw, mu,
                                                                              u, wptr(Zc, prec));
  } else {
    dim3 grid((d + kMcCols - 1) / kMcCols, batch);
    mean_center_kernel<<<grid, kMcCols * kMcGroups, n * sizeof(float), st>>>(Z, w, wdiag, n, d, eps, t, sw, mu,
                                                                             u, wptr(Zc, prec));
  }
  note_launch();
}
void trace_scales(const float* M, int batch, int d, float eps, int post_mode, float* tr, float* inv,
                  float* post, cudaStream_t st) {
  trace_scales_kernel<<<batch, 256, 0, st>>>(M, d, eps, post_mode, tr, inv, post);
  note_launch();
}
void batch_dot(const float* X, const float* Y, int batch, long long n_per, float* out,
               cudaStream_t st) {
  // single deterministic block per image (grid.x = 1): n_per is at most ~1M elements
  dim3 grid(1, batch);
  batch_dot_kernel<<<grid, 1024, 0, st>>>(X, Y, n_per, out);
  note_launch();
}
void ns_bwd_finish(const float* dA, const float* inv, const float* dotO, const float* dotA,
                   float coef_tau, int batch, int d, float* dM, cudaStream_t st) {
  dim3 grid((d + 127) / 128, d, batch);
  ns_bwd_finish_kernel<<<grid, 128, 0, st>>>(dA, inv, dotO, dotA, coef_tau, d, dM);
  note_launch();
}
void normalize_graph(const float* G, int batch, int n, int method, float eps, float* out,
                     float* deg, cudaStream_t st) {
  dim3 g1((n + 7) / 8, batch);
  clamped_degree_kernel<<<g1, 256, 0, st>>>(G, n, eps, deg);
  note_launch();
  dim3 g2((n + 127) / 128, n, batch);
  normalize_graph_kernel<<<g2, 128, 0, st>>>(G, deg, n, method, out);
  note_launch();
}
void batch_trace(const float* M, int batch, int d, float* tr, cudaStream_t st) {
  batch_trace_kernel<<<batch, 256, 0, st>>>(M, d, tr);
  note_launch();
}
void triu_pack(const float* O, int batch, int d, float* v, cudaStream_t st) {
  dim3 grid((d + kTriuRows - 1) / kTriuRows, batch);
  triu_pack_kernel<<<grid, 256, 0, st>>>(O, d, v);
  note_launch();
}
void triu_unpack(const float* dv, int batch, int d, float* dO, cudaStream_t st) {
  dim3 grid((d + kTriuRows - 1) / kTriuRows, batch);
  triu_unpack_kernel<<<grid, 256, 0, st>>>(dv, d, dO);
  note_launch();
}
void sketch_fwd(const float* x, int batch, int d, int S, const int* off, const int* idx,
                const float* sgn, float* cs, float* out, cudaStream_t st) {
  dim3 grid((S + 255) / 256, batch);
  sketch_fwd_kernel<<<grid, 256, 0, st>>>(x, batch, d, S, off, idx, sgn, cs, out);
  note_launch();
}
void sketch_bwd(const float* dout, const float* cs, int batch, int d, int S, const long long* hash,
                const long long* sign, float* dx, cudaStream_t st) {
  dim3 grid((d + 255) / 256, batch);
  sketch_bwd_kernel<<<grid, 256, 0, st>>>(dout, cs, batch, d, S, hash, sign, dx);
  note_launch();
}
void w_fill_eye(const W& out, float c, int prec, cudaStream_t st) {
  dim3 grid((out.cols + 127) / 128, out.rows, out.batch);
  w_fill_eye_kernel<<<grid, 128, 0, st>>>(wptr(out, prec), out.cols, c);
  note_launch();
}
void w_lincomb(const W& out, float a, const W& X, float b, const W* Y, float c, const W* Z,
               const float* s_b, int prec, cudaStream_t st) {
  dim3 grid((out.cols + 127) / 128, out.rows, out.batch);
  w_lincomb_kernel<<<grid, 128, 0, st>>>(wptr(out, prec), out.rows, out.cols, a, wptr(X, prec), b,
                                         Y ? wptr(*Y, prec) : WPtr{}, Y != nullptr, c,
                                         Z ? wptr(*Z, prec) : WPtr{}, Z != nullptr, s_b);
  note_launch();
}
void w_dot(const W& X, const W& Y, float* out, int prec, cudaStream_t st) {
  w_dot_kernel<<<X.batch, 512, 0, st>>>(wptr(X, prec), wptr(Y, prec), X.rows, X.cols, out);
  note_launch();
}
void mlr_scalars_fwd(const float* tau, int batch, float eps, float aK, float* scal, cudaStream_t st) {
  mlr_scalars_fwd_kernel<<<(batch + 127) / 128, 128, 0, st>>>(tau, batch, eps, aK, scal);
  note_launch();
}
void mlr_scalars_bwd(const float* scal, int batch, float aK, const float* dotOO, const float* trdO,
                     const float* dotHs, float* dtaup, cudaStream_t st) {
  mlr_scalars_bwd_kernel<<<(batch + 127) / 128, 128, 0, st>>>(scal, batch, aK, dotOO, trdO, dotHs, dtaup);
  note_launch();
}
void reduce_splits(const float* partial, int splits, int m, int n, const float* bias, float* y,
                   cudaStream_t st) {
  const long long mn = (long long)m * n;
  reduce_splits_kernel<<<(unsigned)((mn + 255) / 256), 256, 0, st>>>(partial, splits, mn, n, bias, y);
  note_launch();
}
void feature_tail_fwd(const float* y, int m, int n, const float* gamma, const float* beta, float* run_mean,
                      float* run_var, int training, float momentum, float bn_eps, float drop_p,
                      unsigned long long seed, float* out, float* save_mean, float* save_rstd, cudaStream_t st) {
  feature_tail_fwd_kernel<<<(n + kFtCols - 1) / kFtCols, kFtCols * kFtRows, 0, st>>>(
      y, m, n, gamma, beta, run_mean, run_var, training, momentum, bn_eps, drop_p, seed, out, save_mean, save_rstd);
  note_launch();
}
void feature_tail_bwd(const float* dout, const float* y, int m, int n, const float* gamma, const float* beta,
                      const float* save_mean, const float* save_rstd, int training, float drop_p,
                      unsigned long long seed, float* dy, float* dgamma, float* dbeta, cudaStream_t st) {
  feature_tail_bwd_kernel<<<(n + kFtCols - 1) / kFtCols, kFtCols * kFtRows, 0, st>>>(
      dout, y, m, n, gamma, beta, save_mean, save_rstd, training, drop_p, seed, dy, dgamma, dbeta);
  note_launch();
}
void colsum(const float* x, int m, int n, float* out, cudaStream_t st) {
  colsum_kernel<<<(n + 127) / 128, 128, 0, st>>>(x, m, n, out);
  note_launch();
}
void pool_bwd_dmu(const float* dZc, const float* du, const float* sw, const float* t, int batch,
                  int n, int d, float eps, float* dmu, cudaStream_t st) {
  dim3 grid((d + 127) / 128, batch);
  pool_bwd_dmu_kernel<<<grid, 128, 0, st>>>(dZc, du, sw, t, n, d, eps, dmu);
  note_launch();
}
void pool_bwd_rows(const float* dZc, const float* Z, const W& Zc, const float* w, const float* t,
                   const float* mu, const float* u, const float* du, const float* dmu, int batch,
                   int n, int d, float eps, float* dZ, float* dw, float* dt, int prec,
                   cudaStream_t st) {
  dim3 grid((n + 7) / 8, batch);
  pool_bwd_rows_kernel<<<grid, 256, 0, st>>>(dZc, Z, wptr(Zc, prec), w, t, mu, u, du, dmu, n, d, eps,
                                             dZ, dw, dt);
  note_launch();
}
void pool_bwd_ds(const float* dW, long long ldW, const float* dw, const float* dt, const float* G,
                 const float* s, int batch, int n, int sym, float* ds, cudaStream_t st) {
  dim3 grid((n + 7) / 8, batch);
  pool_bwd_ds_kernel<<<grid, 256, 0, st>>>(dW, ldW, dw, dt, G, s, n, sym, ds);
  note_launch();
}
void pool_bwd_dG(const float* dW, long long ldW, const float* dw, const float* dt, const float* s,
                 const float* deg, const float* ds, int batch, int n, float eps, int sym, float* dG,
                 cudaStream_t st) {
  dim3 grid((n + kDgRows - 1) / kDgRows, batch);
  const size_t smem = 3 * (size_t)n * sizeof(float);
  if (smem > 48 * 1024)   // more than 4096 tokens: opt in to the large shared-memory carve-out
    cudaFuncSetAttribute(pool_bwd_dG_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pool_bwd_dG_kernel<<<grid, 256, smem, st>>>(dW, ldW, dw, dt, s, deg, ds, n, eps, sym, dG);
  note_launch();
}

int triu_unpack_blocks(int d) { return (d + kUnpackRows - 1) / kUnpackRows; }
void triu_unpack_planes(const float* dv, long long ld_dv, int batch, int d, const float* s_b,
                        const W& out, float* tr_partial, int prec, cudaStream_t st) {
  dim3 grid(triu_unpack_blocks(d), batch);
  triu_unpack_planes_kernel<<<grid, 256, 0, st>>>(dv, ld_dv, d, s_b, wptr(out, prec), tr_partial);
  note_launch();
}
void triu_unpack_sym_planes(const float* dv, long long ld_dv, int batch, int d, const float* s_b,
                            float scale, const W& out, int prec, cudaStream_t st) {
  dim3 grid(triu_unpack_blocks(d), batch);
  triu_unpack_sym_planes_kernel<<<grid, 256, 0, st>>>(dv, ld_dv, d, s_b, scale, wptr(out, prec));
  note_launch();
}
void triu_pack_planes(const float* O, int batch, int d, const W& out, int prec, cudaStream_t st) {
  dim3 grid((d + kUnpackRows - 1) / kUnpackRows, batch);
  triu_pack_planes_kernel<<<grid, 256, 0, st>>>(O, d, wptr(out, prec));
  note_launch();
}
void rowdot_bias(const float* dy, const float* y, const float* bias, int batch, int n, float* out,
                 cudaStream_t st) {
  rowdot_bias_kernel<<<(batch + 7) / 8, 256, 0, st>>>(dy, y, bias, batch, n, out);
  note_launch();
}
void mh_scalars_fwd(const float* tau, int batch, float eps, float* scal, cudaStream_t st) {
  mh_scalars_fwd_kernel<<<(batch + 127) / 128, 128, 0, st>>>(tau, batch, eps, scal);
  note_launch();
}
void mh_dtau(const float* scal, int batch, const float* dotO, const float* dotA, float* dtau,
             cudaStream_t st) {
  mh_dtau_kernel<<<(batch + 127) / 128, 128, 0, st>>>(scal, batch, dotO, dotA, dtau);
  note_launch();
}
void graph_mean(const float* G, int batch, long long per, float* g, cudaStream_t st) {
  graph_mean_kernel<<<batch, 1024, 0, st>>>(G, per, g);
  note_launch();
}
void align_rows(const float* g, const long long* labels, int batch, float* rowloss, float* dg,
                cudaStream_t st) {
  align_rows_kernel<<<batch, 256, 0, st>>>(g, labels, batch, rowloss, dg);
  note_launch();
}
void align_bwd(const float* dg, const float* dloss, int batch, long long per, float* dG, cudaStream_t st) {
  dim3 grid((unsigned)((per + 1023) / 1024 < 64 ? (per + 1023) / 1024 : 64), batch);
  align_bwd_kernel<<<grid, 256, 0, st>>>(dg, dloss, per, dG);
  note_launch();
}
void sum_partials(const float* partial, int nper, int batch, float* out, cudaStream_t st) {
  sum_partials_kernel<<<(batch + 7) / 8, 256, 0, st>>>(partial, nper, batch, out);
  note_launch();
}

}  // namespace k
}  // namespace egm
