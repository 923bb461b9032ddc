// GraphPolynomialFusion.forward as ONE pass over the tokens (north-star kernel 1; reference:
// src/models/gpf_kernel.py:117-159 = _compute_similarity x2 (:75-94), the (P+1)(Q+1)-term Hadamard
// polynomial (:133-150), symmetrise (:153-154), clamp (:157) - ~120 ATen launches and ~100 passes
// over 40 MB tensors in the reference; 5 launches and ~1 GB of traffic in round 1 of this library).
//
// One persistent CTA per SM walks the images. For an image it streams both token matrices from HBM
// exactly once:
//   warp 0      TMA producer: fp32 boxes {32 features x NP/2 tokens} x 2 of one view into a 5-slot
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
// The staging ring is what keeps HBM busy: a slot is occupied from the TMA issue until the converters
// release it (~2 us of memory latency + transfer, then ~0.3 us of conversion), so the bytes in flight
// per SM are (slots - 1) x 26 KiB. Three slots measured 2.1 TB/s (latency-bound, ncu: the workers wait
// on stg_full); five slots cover the per-SM share of HBM bandwidth at that latency.
constexpr int kStgSlots = 5;                  // fp32 staging ring: boxes of 32 features
constexpr int kStgBytes = kRowsMax * 128;     // rows x 32 fp32
constexpr int kOpUnits = 2;                   // operand ring: the two 64-byte halves of one 128-byte-row slot
constexpr int kOpBytes = 2 * kPlane;          // hi + lo
constexpr int kWorkers = 8;                   // converter / epilogue warps
constexpr int kThreadsF = 32 * (2 + kWorkers);
constexpr int kTbuf = 32 * 17 * 4;            // per-warp transpose buffer (16 columns at a time)
constexpr int kMaxCoef = 256;
constexpr int kSmemF = 1024 + kStgSlots * kStgBytes + kOpBytes + kWorkers * kTbuf +
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
  // optional (training): the RAW tokens as bf16 hi/lo operand planes [B][N][ld_pl] for the backward's
  // dx = E'' X product - the converter threads hold exactly these values in registers
  __nv_bfloat16* xpl[2];
  long long pl_lo_off;    // elements from a hi plane to its lo plane
  int ld_pl;
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
template <int MD, typename C>
__device__ __forceinline__ float poly(const float (&pa)[MD + 1], const float (&pb)[MD + 1], const C& c) {
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
  const uint32_t tbuf0 = ops + kOpBytes;
  float* tbuf_gen = reinterpret_cast<float*>(gen + kStgSlots * kStgBytes + kOpBytes);
  float* inv_gen = tbuf_gen + kWorkers * 32 * 17;            // [2][kRowsMax] 1/max(norm, eps)
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
    for (int h = 0; h < kOpUnits; ++h) { ptx::mbar_init(op_full(h), kWorkers); ptx::mbar_init(op_empty(h), 1); }
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
            // row tile 1 needs token rows 128..N-1 only: one box when they fit (rows past it feed
            // accumulator lanes of tokens that do not exist)
            const int row0 = r * 128;
            const bool two = row0 + box_rows < N;
            ptx::mbar_arrive_expect_tx(stg_full(s), (two ? 2u : 1u) * box_rows * 128u);
            const CUtensorMap* tm = &p.tm[u & 1];
            const int k0 = (u >> 1) * 32;
            ptx::tma_load_3d(tm, stg_full(s), stg + s * kStgBytes, k0, row0, b);
            if (two) ptx::tma_load_3d(tm, stg_full(s), stg + s * kStgBytes + box_rows * 128, k0, row0 + box_rows, b);
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
          const int hs = n & 1;
          ptx::mbar_wait(op_full(hs), (n >> 1) & 1u);
          ptx::tc_fence_after();
          if (issuer) {
            const uint32_t hi = ops + hs * 64;
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
    // 8 worker warps, one slot row per thread. (16 workers with half a row each measured SLOWER, 220 vs
    // 166 us: 18 warps cap the kernel at 96 registers and every extra warp adds its own barrier round
    // trip per unit; the unit's latency chain - wait, load, convert, store, proxy fence, arrive, MMA -
    // not the issue rate, is what bounds this kernel.)
    const int w = warp - 2;
    const int t = w * 32 + lane;                              // slot row
#ifdef EGM_GPF_PROFILE
    long long pf_wait_stg = 0, pf_wait_op = 0, pf_conv = 0, pf_wait_t = 0, pf_epi = 0, pf_bar = 0, pf_t0 = clock64();
#define PF(acc) do { const long long c_ = clock64(); acc += c_ - pf_t0; pf_t0 = c_; } while (0)
#else
#define PF(acc) do {} while (0)
#endif
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int cg = w >> 2;                       // its share of the 16-column chunks: c16 = cg, cg + 4, ...
    float creg[MD <= 3 ? (MD + 1) * (MD + 1) : 1];
    if (MD <= 3) {
#pragma unroll
      for (int k = 0; k < (MD + 1) * (MD + 1); ++k) creg[MD <= 3 ? k : 0] = coef_gen[k];
    }
    float* tb = tbuf_gen + w * 32 * 17;
    uint32_t n = 0, tiles = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x)
      for (int r = 0; r < ntile; ++r, ++tiles) {
        const int row_base = r * 128;            // global row of slot row 0
        // slot rows that must hold real operands: every column of the tile (B operand) and every row
        // whose token exists (A operand); the rest feeds accumulator lanes nobody reads
        const int rows = r == 0 ? NP : NP - 128;
        const bool mine = t < rows;
        float ssq[2] = {0.f, 0.f};
        for (int u = 0; u < units; ++u, ++n) {
          const int s = n % kStgSlots, hs = n & 1;
          ptx::mbar_wait(stg_full(s), (n / kStgSlots) & 1u);
          PF(pf_wait_stg);
          ptx::mbar_wait(op_empty(hs), ((n >> 1) & 1u) ^ 1u);
          PF(pf_wait_op);
          if (mine) {
            const uint32_t src = stg + s * kStgBytes + t * 128;
            const uint32_t dst = ops + t * 128;
            const uint32_t sw = t & 7;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;    // independent partial sums of squares
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {     // 16 features at a time: loads in flight before the math
              uint32_t x[16];
#pragma unroll
              for (int c = 0; c < 4; ++c)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(x[4 * c]), "=r"(x[4 * c + 1]), "=r"(x[4 * c + 2]), "=r"(x[4 * c + 3])
                             : "r"(src + (((4 * hh + c) ^ sw) << 4)));
              uint32_t hw[8], lw[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float x0 = __uint_as_float(x[2 * e]), x1 = __uint_as_float(x[2 * e + 1]);
                if (e & 1) { a2 = fmaf(x0, x0, a2); a3 = fmaf(x1, x1, a3); }
                else { a0 = fmaf(x0, x0, a0); a1 = fmaf(x1, x1, a1); }
                uint32_t h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
                hw[e] = h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;"
                    : "=r"(lw[e])
                    : "f"(x1 - __uint_as_float(h & 0xFFFF0000u)), "f"(x0 - __uint_as_float(h << 16)));
              }
#pragma unroll
              for (int c = 0; c < 2; ++c) {      // 8 features -> one 16-byte bf16 chunk per plane
                const uint32_t off = ((hs * 4 + 2 * hh + c) ^ sw) << 4;
                ptx::sts128(dst + off, hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
                if (p.npass == 3)
                  ptx::sts128(dst + kPlane + off, lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
              }
              // training: the same values leave as the backward's operand planes (row tile 0 covers every row)
              if (p.xpl[u & 1] && r == 0 && t < N) {
                const int kcol = (u >> 1) * 32 + 16 * hh;
                __nv_bfloat16* gp = p.xpl[u & 1] + ((long long)b * N + t) * p.ld_pl + kcol;
#pragma unroll
                for (int c = 0; c < 2; ++c)
                  if (kcol + 8 * c < p.ld_pl) {
                    *reinterpret_cast<uint4*>(gp + 8 * c) = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
                    if (p.npass == 3)
                      *reinterpret_cast<uint4*>(gp + p.pl_lo_off + 8 * c) =
                          make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
                  }
              }
            }
            ssq[u & 1] += (a0 + a1) + (a2 + a3);
          }
          ptx::fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(op_full(hs));
            ptx::mbar_arrive(stg_empty(s));
          }
          PF(pf_conv);
        }
        // norms of this tile's rows -> shared (scaling of the columns) and global (saved for backward)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (mine) {
            const float nr = sqrtf(ssq[v]);
            inv_gen[v * kRowsMax + t] = p.cosine ? 1.f / fmaxf(nr, p.eps) : 1.f;
            const int gr = row_base + t;
            if (gr < N) p.nrm[v][(long long)b * N + gr] = nr;      // tile 1 rewrites rows 128.. with the same value
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers * 32) : "memory");
        PF(pf_bar);
        // ------------------------------------------------------------------- epilogue
        ptx::mbar_wait(tfull, tiles & 1u);
        ptx::tc_fence_after();
        PF(pf_wait_t);
        const int li = q * 32 + lane;            // row within the tile = TMEM lane
        const int i = row_base + li;             // global row
        const int ncols = r == 0 ? NP : NP - 128;
        const int nch = ncols / 16;              // NP is a multiple of 16
        const int lic = min(li, kRowsMax - 1);
        const float ia_i = inv_gen[lic], ip_i = inv_gen[kRowsMax + lic];   // slot row li is global row i
        const uint32_t t_row = tmem + (static_cast<uint32_t>(q * 32) << 16);
        const int row0 = row_base + q * 32;      // first global row of this warp
        float* Gb = p.G + (long long)b * N * N;
        for (int c = cg; c < nch; c += kWorkers / 4) {
          const int lj0 = c * 16;                // column within the tile == slot row of that token
          const int j0 = row_base + lj0;         // global column
          if (j0 + 15 < row0 || row0 >= N) continue;      // chunk entirely below the diagonal / no rows (warp-uniform)
          uint32_t va[16], vp[16];
          ptx::tmem_ld_32x16(t_row + lj0, va);
          ptx::tmem_ld_32x16(t_row + 256 + lj0, vp);
          // the 16 column scales are the same for every lane: four 16-byte broadcast loads per view
          float sa[16], sp[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 x = *reinterpret_cast<const float4*>(inv_gen + lj0 + 4 * k4);
            const float4 y = *reinterpret_cast<const float4*>(inv_gen + kRowsMax + lj0 + 4 * k4);
            sa[4 * k4] = x.x * ia_i; sa[4 * k4 + 1] = x.y * ia_i; sa[4 * k4 + 2] = x.z * ia_i; sa[4 * k4 + 3] = x.w * ia_i;
            sp[4 * k4] = y.x * ip_i; sp[4 * k4 + 1] = y.y * ip_i; sp[4 * k4 + 2] = y.z * ip_i; sp[4 * k4 + 3] = y.w * ip_i;
          }
          ptx::tmem_ld_wait();
          float g[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float ra = __uint_as_float(va[jj]) * sa[jj];
            const float rp = __uint_as_float(vp[jj]) * sp[jj];
            va[jj] = __float_as_uint(ra);
            vp[jj] = __float_as_uint(rp);
            float pa[MD + 1], pb[MD + 1];
            powers<MD>(ra, p.P, pa);
            powers<MD>(rp, p.Q, pb);
            g[jj] = fmaxf(MD <= 3 ? poly<MD>(pa, pb, creg) : poly<MD>(pa, pb, coef_gen), 0.f);
          }
          // interior chunk: every element is strictly above the diagonal and inside the matrix
          const bool interior = j0 >= row0 + 32 && j0 + 16 <= N && row0 + 32 <= N;
          // three outputs share the two store patterns: G always, R_a / R_p when the backward will run
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            float* dstb;
            int ld;
            if (o == 0) { dstb = Gb; ld = N; }
            else {
              float* R = o == 1 ? p.Ra : p.Rp;
              if (!R) break;
              dstb = R + (long long)b * N * p.ldR; ld = (int)p.ldR;
            }
            float val[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) val[jj] = o == 0 ? g[jj] : __uint_as_float(o == 1 ? va[jj] : vp[jj]);
            // mirrored store (j, i), j > i: lanes are consecutive i - coalesced as is
            float* col = dstb + (long long)j0 * ld + i;
            if (interior) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) { *col = val[jj]; col += ld; }
            } else if (i < N) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                if (j0 + jj > i && j0 + jj < N) *col = val[jj];
                col += ld;
              }
            }
            // direct store (i, j), j >= i: through the warp's transpose buffer (two rows per pass:
            // lanes 0-15 row rr, lanes 16-31 row rr + 1)
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) tb[lane * 17 + jj] = val[jj];
            __syncwarp();
            const int jc = lane & 15, rh = lane >> 4;
            const int j = j0 + jc;
            float* rowp = dstb + (long long)(row0 + rh) * ld + j;
            const float* tbr = tb + rh * 17 + jc;
            if (interior) {
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) { *rowp = tbr[rr * 17]; rowp += 2 * ld; }
            } else if (j < N) {
#pragma unroll
              for (int rr = 0; rr < 32; rr += 2) {
                const int i2 = row0 + rr + rh;
                if (i2 < N && j >= i2) *rowp = tbr[rr * 17];
                rowp += 2 * ld;
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty);
        PF(pf_epi);
      }
#ifdef EGM_GPF_PROFILE
    if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 120))
      printf("gpf_prof blk %d warp %d: wait_stg %lld wait_op %lld conv %lld bar %lld wait_tfull %lld epi %lld\n",
             (int)blockIdx.x, warp, pf_wait_stg, pf_wait_op, pf_conv, pf_bar, pf_wait_t, pf_epi);
#endif
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
                          float* nrm_a, float* nrm_p, int npass, cudaStream_t st, const W* xa_raw,
                          const W* xp_raw) {
  FusedParams fp;
  memset(&fp, 0, sizeof(fp));
  if (xa_raw && xp_raw) {
    fp.xpl[0] = static_cast<__nv_bfloat16*>(xa_raw->base);
    fp.xpl[1] = static_cast<__nv_bfloat16*>(xp_raw->base);
    fp.pl_lo_off = (long long)xa_raw->batch * xa_raw->rows * xa_raw->ld;
    fp.ld_pl = (int)xa_raw->ld;
  }
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
