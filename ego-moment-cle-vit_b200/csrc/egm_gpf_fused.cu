// GraphPolynomialFusion.forward as ONE pass over the tokens (north-star kernel 1; reference:
// src/models/gpf_kernel.py:117-159 = _compute_similarity x2 (:75-94), the (P+1)(Q+1)-term Hadamard
// polynomial (:133-150), symmetrise (:153-154), clamp (:157) - ~120 ATen launches and ~100 passes
// over 40 MB tensors in the reference; 5 launches and ~1 GB of traffic in round 1 of this library).
//
// One persistent CTA per SM walks the images. For an image it streams both token matrices from HBM
// exactly once:
//   warp 0      TMA producer: fp32 boxes {32 features x NP/2 tokens} x 2 of one view into a 3-slot
//               staging ring (128B swizzle), mbarrier complete_tx
//   warps 2..9  converters: one token row per thread - read the fp32 row slice, add it to the row's
//               running sum of squares (the norm of F.normalize, exact fp32), split into bf16 hi/lo
//               and write both planes into the tcgen05 operand ring in the canonical 128B-swizzled
//               K-major layout (no normalised copy of the tokens ever exists: the cosine scaling
//               1/(n_i n_j) is applied to the accumulator in the epilogue)
//   warp 1      MMA issuer: R_view += X X^T on the tensor cores - A and B operands are the SAME
//               shared-memory rows (a Gram matrix), 3 bf16 MMAs per product in fp32 mode - into two
//               TMEM accumulators (R_anchor at column 0, R_positive at column 256)
//   warps 2..9  epilogue: tcgen05.ld both accumulators, scale, evaluate sum_pq c_pq f_p(Ra) f_q(Rp)
//               (Horner), clamp, and write G - only elements on or above the diagonal are evaluated;
//               each is stored at (i,j) through a per-warp transpose buffer and at (j,i) straight
//               from registers (TMEM lanes are rows, so the mirrored store is the coalesced one).
//               G is therefore symmetric bit for bit, which MomentHead's fast path relies on.
// Row tile 0 = token rows 0..127 against all columns; row tile 1 (N > 128) = rows 128.. against columns
// 128.. only, re-reading those rows (L2 hits: the image was streamed microseconds earlier).
// Algorithmic bytes per image: 2 N D 4 read + N N 4 written (1.366 MB at N=197, D=768).
// In training mode the cosine matrices R_a, R_p are written as well (the backward's polynomial needs
// them) and the normalised operand planes are re-derived in the backward.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "egm_gemm.h"
#include "egm_kernels.cuh"
#include "egm_ptx.cuh"

namespace egm {
namespace k {
namespace {

constexpr int kRowsMax = 208;                 // token rows per operand slot (N <= 208)
constexpr int kPlane = kRowsMax * 128;        // one bf16 plane of a slot: rows x 64 bf16 (26 KiB)
constexpr int kStgSlots = 3;                  // fp32 staging ring: boxes of 32 features
constexpr int kStgBytes = kRowsMax * 128;     // rows x 32 fp32
constexpr int kOpSlots = 2;                   // operand ring: 2 slots x 2 halves of 32 K-elements
constexpr int kOpBytes = 2 * kPlane;          // hi + lo
constexpr int kWorkers = 8;                   // converter / epilogue warps
constexpr int kThreadsF = 32 * (2 + kWorkers);
constexpr int kTbuf = 32 * 33 * 4;            // per-warp transpose buffer
constexpr int kMaxCoef = 256;
constexpr int kSmemF = 1024 + kStgSlots * kStgBytes + kOpSlots * kOpBytes + kWorkers * kTbuf +
                       2 * kRowsMax * 4 + kMaxCoef * 4 + 256;

struct FusedParams {
  CUtensorMap tm[2];      // anchor / positive tokens, fp32 [B][N][D], box {32, NP/2, 1}, 128B swizzle
  int B, N, D, NP;
  int P, Q, cosine, npass;
  float eps;
  const float* coef;      // [(P+1)(Q+1)], row stride Q+1
  float* G;               // [B][N][N]
  float* Ra;              // optional [B][N][ldR]
  float* Rp;
  long long ldR;
  float* nrm[2];          // [B][N]
};

template <int MD>
__device__ __forceinline__ void powers(float x, int deg, float (&pw)[MD + 1]) {
  // f_0 = 1, f_1 = x (NOT clamped), f_k = max(x,0)^k   (gpf_kernel.py:107-115); powers above the
  // degree are forced to zero (their coefficients are zero padding; inf * 0 must not appear)
  pw[0] = 1.f;
  pw[1] = x;
  const float c = fmaxf(x, 0.f);
  float acc = c;
#pragma unroll
  for (int q = 2; q <= MD; ++q) {
    acc *= c;
    pw[q] = (q <= deg) ? acc : 0.f;
  }
}
// sum_{p,q} c[p][q] pa[p] pb[q]; c is zero-padded to (MD+1) x (MD+1)
template <int MD>
__device__ __forceinline__ float poly(const float (&pa)[MD + 1], const float (&pb)[MD + 1], const float* c) {
  float f = 0.f;
#pragma unroll
  for (int p = MD; p >= 0; --p) {
    float inner = 0.f;
#pragma unroll
    for (int q = MD; q >= 0; --q) inner = fmaf(c[p * (MD + 1) + q], pb[q], inner);
    f = fmaf(pa[p], inner, f);
  }
  return f;
}

template <int MD>
__global__ void __launch_bounds__(kThreadsF, 1) gpf_fused_fwd_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t stg = base;
  const uint32_t ops = stg + kStgSlots * kStgBytes;
  const uint32_t tbuf0 = ops + kOpSlots * kOpBytes;
  float* tbuf_gen = reinterpret_cast<float*>(gen + kStgSlots * kStgBytes + kOpSlots * kOpBytes);
  float* inv_gen = tbuf_gen + kWorkers * 32 * 33;            // [2][kRowsMax] 1/max(norm, eps)
  float* coef_gen = inv_gen + 2 * kRowsMax;                  // (MD+1)^2, zero padded (MD=15: 256)
  const uint32_t bars = tbuf0 + kWorkers * kTbuf + 2 * kRowsMax * 4 + kMaxCoef * 4;
  auto stg_full = [&](int s) { return bars + 8u * s; };
  auto stg_empty = [&](int s) { return bars + 8u * (kStgSlots + s); };
  auto op_full = [&](int h) { return bars + 8u * (2 * kStgSlots + h); };
  auto op_empty = [&](int h) { return bars + 8u * (2 * kStgSlots + 4 + h); };
  const uint32_t tfull = bars + 8u * (2 * kStgSlots + 8);
  const uint32_t tempty = bars + 8u * (2 * kStgSlots + 9);
  const uint32_t tmem_slot = bars + 8u * (2 * kStgSlots + 10);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      reinterpret_cast<uint8_t*>(coef_gen + kMaxCoef) + 8 * (2 * kStgSlots + 10));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, NP = p.NP, D = p.D;
  const int nkb = (D + 31) / 32;
  const int ntile = N > 128 ? 2 : 1;
  const int units = 2 * nkb;                    // (k-block, view) units per row tile
  const int box_rows = NP / 2;

  if (threadIdx.x == 0) {
    ptx::prefetch_tensormap(&p.tm[0]);
    ptx::prefetch_tensormap(&p.tm[1]);
    for (int s = 0; s < kStgSlots; ++s) { ptx::mbar_init(stg_full(s), 1); ptx::mbar_init(stg_empty(s), kWorkers); }
    for (int h = 0; h < 4; ++h) { ptx::mbar_init(op_full(h), kWorkers); ptx::mbar_init(op_empty(h), 1); }
    ptx::mbar_init(tfull, 1);
    ptx::mbar_init(tempty, kWorkers);
    ptx::fence_barrier_init();
  }
  for (int t = threadIdx.x; t < (MD + 1) * (MD + 1); t += blockDim.x) {
    const int pp = t / (MD + 1), qq = t % (MD + 1);
    coef_gen[t] = (pp <= p.P && qq <= p.Q) ? p.coef[pp * (p.Q + 1) + qq] : 0.f;
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool issuer = ptx::elect_one();
    uint32_t n = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x)
      for (int r = 0; r < ntile; ++r)
        for (int u = 0; u < units; ++u, ++n) {
          const int s = n % kStgSlots;
          ptx::mbar_wait(stg_empty(s), ((n / kStgSlots) & 1u) ^ 1u);
          if (issuer) {
            ptx::mbar_arrive_expect_tx(stg_full(s), 2u * box_rows * 128u);
            const CUtensorMap* tm = &p.tm[u & 1];
            const int k0 = (u >> 1) * 32, row0 = r * 128;
            ptx::tma_load_3d(tm, stg_full(s), stg + s * kStgBytes, k0, row0, b);
            ptx::tma_load_3d(tm, stg_full(s), stg + s * kStgBytes + box_rows * 128, k0, row0 + box_rows, b);
          }
          __syncwarp();
        }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    const bool issuer = ptx::elect_one();
    uint32_t n = 0, tiles = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x)
      for (int r = 0; r < ntile; ++r, ++tiles) {
        const int ncols = r == 0 ? NP : NP - 128;
        const uint32_t idesc = ptx::idesc_bf16_f32(128, ncols, 0, 0);
        ptx::mbar_wait(tempty, (tiles & 1u) ^ 1u);          // the previous tile's epilogue has drained TMEM
        ptx::tc_fence_after();
        for (int u = 0; u < units; ++u, ++n) {
          const int hs = n & 3;
          ptx::mbar_wait(op_full(hs), (n >> 2) & 1u);
          ptx::tc_fence_after();
          if (issuer) {
            const uint32_t hi = ops + (hs >> 1) * kOpBytes + (hs & 1) * 64;
            const uint32_t d = tmem + (u & 1) * 256;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t dh = ptx::smem_desc_sw128(hi + kk * 32, 0, 1024);
              const uint32_t acc0 = (u < 2 && kk == 0) ? 0u : 1u;
              if (p.npass == 3) {
                const uint64_t dl = ptx::smem_desc_sw128(hi + kPlane + kk * 32, 0, 1024);
                ptx::mma_bf16_ss(d, dl, dh, idesc, acc0);
                ptx::mma_bf16_ss(d, dh, dh, idesc, 1u);
                ptx::mma_bf16_ss(d, dh, dl, idesc, 1u);
              } else {
                ptx::mma_bf16_ss(d, dh, dh, idesc, acc0);
              }
            }
            ptx::tc_commit(op_empty(hs));
            if (u == units - 1) ptx::tc_commit(tfull);
          }
          __syncwarp();
        }
      }
  } else {
    // ------------------------------------------------- converters, then the tile's epilogue
    const int w = warp - 2;
    const int t = w * 32 + lane;                 // the slot row this thread converts
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int half = w >> 2;                     // which half of the column chunks it takes
    float* tb = tbuf_gen + w * 32 * 33;
    uint32_t n = 0, tiles = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x)
      for (int r = 0; r < ntile; ++r, ++tiles) {
        const int row_base = r * 128;            // global row of slot row 0
        const int rows = r == 0 ? NP : 128;      // slot rows that feed an operand
        const bool mine = t < rows && t < kRowsMax;
        float ssq[2] = {0.f, 0.f};
        for (int u = 0; u < units; ++u, ++n) {
          const int s = n % kStgSlots, hs = n & 3;
          ptx::mbar_wait(stg_full(s), (n / kStgSlots) & 1u);
          ptx::mbar_wait(op_empty(hs), ((n >> 2) & 1u) ^ 1u);
          if (mine) {
            const uint32_t src = stg + s * kStgBytes + t * 128;
            const uint32_t dst = ops + (hs >> 1) * kOpBytes + t * 128;
            const uint32_t sw = t & 7;
            float acc = ssq[u & 1];
#pragma unroll
            for (int c = 0; c < 4; ++c) {        // 8 features -> one 16-byte bf16 chunk per plane
              float v[8];
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                uint32_t x0, x1, x2, x3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                             : "r"(src + (((2 * c + h2) ^ sw) << 4)));
                v[4 * h2] = __uint_as_float(x0); v[4 * h2 + 1] = __uint_as_float(x1);
                v[4 * h2 + 2] = __uint_as_float(x2); v[4 * h2 + 3] = __uint_as_float(x3);
              }
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x0 = v[2 * e], x1 = v[2 * e + 1];
                acc = fmaf(x0, x0, acc);
                acc = fmaf(x1, x1, acc);
                uint32_t h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
                hw[e] = h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;"
                    : "=r"(lw[e])
                    : "f"(x1 - __uint_as_float(h & 0xFFFF0000u)), "f"(x0 - __uint_as_float(h << 16)));
              }
              const uint32_t off = (((hs & 1) * 4 + c) ^ sw) << 4;
              ptx::sts128(dst + off, hw[0], hw[1], hw[2], hw[3]);
              if (p.npass == 3) ptx::sts128(dst + kPlane + off, lw[0], lw[1], lw[2], lw[3]);
            }
            ssq[u & 1] = acc;
          }
          ptx::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(op_full(hs));
            ptx::mbar_arrive(stg_empty(s));
          }
        }
        // norms of this tile's rows -> shared (scaling of the columns) and global (saved for backward)
        if (mine) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const float nr = sqrtf(ssq[v]);
            inv_gen[v * kRowsMax + t] = p.cosine ? 1.f / fmaxf(nr, p.eps) : 1.f;
            const int gr = row_base + t;
            if (gr < N) p.nrm[v][(long long)b * N + gr] = nr;      // tile 1 rewrites rows 128.. with the same value
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");
        // ------------------------------------------------------------------- epilogue
        ptx::mbar_wait(tfull, tiles & 1u);
        ptx::tc_fence_after();
        const int li = q * 32 + lane;            // row within the tile = TMEM lane
        const int i = row_base + li;             // global row
        const int ncols = r == 0 ? NP : NP - 128;
        const int nch = (ncols + 31) / 32;
        const int c_lo = half * ((nch + 1) / 2), c_hi = half ? nch : (nch + 1) / 2;
        const float ia_i = inv_gen[li], ip_i = inv_gen[kRowsMax + li];   // slot row li is global row i
        const uint32_t t_row = tmem + (static_cast<uint32_t>(q * 32) << 16);
        float* Gb = p.G + (long long)b * N * N;
        for (int c = c_lo; c < c_hi; ++c) {
          const int lj0 = c * 32;                // column within the tile == slot row of that token
          const int j0 = row_base + lj0;         // global column
          if (j0 + 31 < row_base + q * 32) continue;      // chunk entirely below the diagonal (warp-uniform)
          uint32_t va[32], vp[32];
          ptx::tmem_ld_32x32(t_row + lj0, va);
          ptx::tmem_ld_32x32(t_row + 256 + lj0, vp);
          ptx::tmem_ld_wait();
          float g[32];
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int ljc = min(lj0 + jj, kRowsMax - 1);
            const float ra = __uint_as_float(va[jj]) * ia_i * inv_gen[ljc];
            const float rp = __uint_as_float(vp[jj]) * ip_i * inv_gen[kRowsMax + ljc];
            va[jj] = __float_as_uint(ra);
            vp[jj] = __float_as_uint(rp);
            float pa[MD + 1], pb[MD + 1];
            powers<MD>(ra, p.P, pa);
            powers<MD>(rp, p.Q, pb);
            g[jj] = fmaxf(poly<MD>(pa, pb, coef_gen), 0.f);
          }
          const int row0 = row_base + q * 32;    // first global row of this warp
          // three outputs share the two store patterns: G always, R_a / R_p when the backward will run
          for (int o = 0; o < 3; ++o) {
            float* dstb;
            long long ld;
            if (o == 0) { dstb = Gb; ld = N; }
            else {
              float* R = o == 1 ? p.Ra : p.Rp;
              if (!R) break;
              dstb = R + (long long)b * N * p.ldR; ld = p.ldR;
            }
            // mirrored store (j, i), j > i: lanes are consecutive i - coalesced as is
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const int j = j0 + jj;
              const float val = o == 0 ? g[jj] : __uint_as_float(o == 1 ? va[jj] : vp[jj]);
              if (j > i && j < N && i < N) dstb[(long long)j * ld + i] = val;
            }
            // direct store (i, j), j >= i: through the warp's transpose buffer
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              tb[lane * 33 + jj] = o == 0 ? g[jj] : __uint_as_float(o == 1 ? va[jj] : vp[jj]);
            __syncwarp();
            const int j = j0 + lane;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
              const int i2 = row0 + rr;
              if (i2 < N && j >= i2 && j < N) dstb[(long long)i2 * ld + j] = tb[rr * 33 + lane];
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty);
      }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

bool token_map(CUtensorMap* tm, const float* x, int B, int N, int D, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)N * D * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(tokens) failed (%d): B=%d N=%d D=%d", (int)r, B, N, D);
    return false;
  }
  return true;
}

}  // namespace

bool gpf_fused_supported(int n, int d, int P, int Q, const float* a, const float* p) {
  static const bool off = []() { const char* e = getenv("EGM_GPF_FUSED"); return e && e[0] == '0'; }();
  if (off) return false;
  return n >= 1 && n <= kRowsMax && d >= 1 && d % 4 == 0 && P <= 15 && Q <= 15 &&
         (reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

cudaError_t gpf_fused_fwd(const float* a, const float* p, const float* coef, int batch, int n, int d, int P,
                          int Q, int cosine, float eps, float* G, float* Ra, float* Rp, long long ldR,
                          float* nrm_a, float* nrm_p, int npass, cudaStream_t st) {
  FusedParams fp;
  memset(&fp, 0, sizeof(fp));
  fp.B = batch; fp.N = n; fp.D = d;
  fp.NP = (n + 15) / 16 * 16;
  fp.P = P; fp.Q = Q; fp.cosine = cosine; fp.npass = npass; fp.eps = eps;
  fp.coef = coef; fp.G = G; fp.Ra = Ra; fp.Rp = Rp; fp.ldR = ldR;
  fp.nrm[0] = nrm_a; fp.nrm[1] = nrm_p;
  if (!token_map(&fp.tm[0], a, batch, n, d, fp.NP / 2) || !token_map(&fp.tm[1], p, batch, n, d, fp.NP / 2))
    return cudaErrorInvalidValue;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const bool small = P <= 3 && Q <= 3;
  static thread_local int configured[2] = {-1, -1};
  if (configured[small] != dev) {
    e = small ? cudaFuncSetAttribute(gpf_fused_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF)
              : cudaFuncSetAttribute(gpf_fused_fwd_kernel<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemF);
    if (e != cudaSuccess) return e;
    configured[small] = dev;
  }
  const int grid = batch < sms ? batch : sms;
  if (small) gpf_fused_fwd_kernel<3><<<grid, kThreadsF, kSmemF, st>>>(fp);
  else gpf_fused_fwd_kernel<15><<<grid, kThreadsF, kSmemF, st>>>(fp);
  note_launch();
  return cudaGetLastError();
}

}  // namespace k
}  // namespace egm
