// tcgen05 / TMEM batched GEMM engine of the moment-pooling path (sm_100a only).
//
// Replaces, for every dense contraction of the path, what the reference leaves to
// torch.bmm -> cuBLAS (reference: src/models/moment_head.py:53-64 Newton-Schulz loop,
// :292-293 pooling products, src/models/gpf_kernel.py:88 Gram) and the elementwise passes
// around them (`3*I - ZY`, `0.5 *`, `/ sqrt(trace)`), which are folded into the epilogue.
//
// Shape of the kernel (one persistent CTA per SM, 320 threads; CTAs work in pairs, see below):
//   warp 0      TMA producer: cp.async.bulk.tensor 3-D boxes {64 x rows x 1 image} of the
//               bf16 operand planes into 128B-swizzled shared memory, mbarrier-signalled
//   warp 1      MMA issuer: one lane issues tcgen05.mma.kind::f16 (128 x 256 x 16, or 256 x 256 x 16
//               over a CTA pair) into one of two 256-column fp32 accumulators in TMEM
//   warps 2..9  epilogue, two per TMEM lane quarter (each takes half of the columns): tcgen05.ld the
//               finished accumulator (32 lanes x 32 columns per load, the next load in flight),
//               apply  alpha*acc + beta*I + gamma*E, split to bf16 hi/lo planes and/or write fp32
//               through swizzled staging + TMA stores, while the MMA warp fills the other accumulator
//
// fp32 fidelity ("bf16x3"): every fp32 operand x is carried as two bf16 planes
// hi = bf16(x), lo = bf16(x - hi).  A*B ~= Ah*Bh + Ah*Bl + Al*Bh  (3 MMAs, fp32 accumulate,
// relative error ~1e-5).  The single-pass mode issues Ah*Bh only.
//
// Up to two products are accumulated into the same TMEM tile (C = A0*B0 + A1*B1), which
// is what the Newton-Schulz backward needs (e.g. Y'_{k+1} = Y'_k*T_k + Y_k*T'_k) without a
// read-modify-write epilogue.
//
// Symmetric matrices (the whole Newton-Schulz chain when the graph is symmetric) are kept as their
// upper 256 x 256 blocks only: a `sym_out` product computes just the tiles that touch those blocks,
// and a `sym` operand whose block is absent is loaded from the mirrored block with the operand's
// major-ness (K-major <-> MN-major) flipped in the UMMA descriptors - see sym_a_mn / sym_b_mn.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "egm_gemm.h"
#include "egm_ptx.cuh"

namespace egm {

namespace {

constexpr int BM = 128;        // tile rows  (= UMMA M, one TMEM lane per row)
constexpr int BN = 256;        // tile cols  (= UMMA N, one TMEM column per col)
constexpr int BK = 64;         // K per pipeline stage: 64 bf16 = one 128-byte swizzle line
constexpr int UK = 16;         // K per tcgen05.mma (bf16)
constexpr int kTileA = BM * BK * 2;   // 16 KiB per plane
constexpr int kChunk = BK * 128;      // one MN-major TMA box: 64 K-rows x 128 B = 8 KiB
constexpr int kEpiWarps = 8;          // two per TMEM lane quarter, each takes half of the columns
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kTmemCols = 512;        // two 256-column accumulators
constexpr int kEpiWarpBytes = 4096;   // per epilogue warp: one staging slot (32 rows x 128 B)
constexpr int kEpiBytes = kEpiWarps * kEpiWarpBytes;

struct alignas(64) TcParams {
  CUtensorMap tm[2][4];  // [term][A_hi, A_lo, B_hi, B_lo]
  CUtensorMap tmC[3];    // TMA-store maps: C_hi, C_lo (bf16, 64B swizzle), C_f32 (128B swizzle)
  CUtensorMap tmC2[2];   // TMA-store maps of the secondary plane output
  int tma_cp, tma_cf, tma_c2;  // which outputs leave through TMA stores
  int epi_split;         // hi and lo plane stores of a chunk as separate bulk groups
  unsigned epi_sleep_ns; // back-off of the epilogue warps while they wait for an accumulator
  unsigned* sched;       // dynamic tile scheduler: [0] next tile, [1] clusters done (self-resetting)
  int splits;            // split-K: `batch` counts K-slices of ONE problem (A/B batch index 0)
  int M, N, batch, nterms;
  int K[2], a_mn[2], b_mn[2];
  int a_sym[2], b_sym[2];  // operand stored as the upper 256-blocks of a symmetric matrix
  int sym_out;           // symmetric result: upper tiles only, off-diagonal dot partials x2
  int tiles_m, tiles_n;
  int tiles_per_img;     // tiles_m * tiles_n, or the upper-triangular subset (triu_tiles)
  int triu_tiles;        // enumerate only tiles that touch j >= i
  int tile_rows;         // rows of one tile: BM (1-CTA kernel) or 2*BM (pair kernel)
  float alpha, beta_eye, gamma;
  const float* alpha_b;
  const float* beta_b;
  const float* gamma_b;
  const void* E0;
  const void* E1;
  long long ldE, bsE;
  int e_mode;  // 0 none, 1 bf16 planes, 2 fp32
  __nv_bfloat16* Cp_hi;
  __nv_bfloat16* Cp_lo;
  long long ldCp, bsCp;
  float* Cf;
  long long ldCf, bsCf;
  // secondary plane output  c2_scale * C + c2_eye * I
  __nv_bfloat16* C2_hi;
  __nv_bfloat16* C2_lo;
  long long ldC2, bsC2;
  float c2_scale, c2_eye;
  // <C, F> partials
  const void* F0;
  const void* F1;
  long long ldF, bsF;
  int f_mode;  // 0 none, 1 bf16 planes, 2 fp32
  float* dot_ws;
  // packed upper triangle as planes
  __nv_bfloat16* X_hi;
  __nv_bfloat16* X_lo;
  long long ldX;
};

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// tile index within one image -> (row tile, column tile). With triu_tiles only the tiles that
// contain at least one element on or above the diagonal are enumerated (row-tile major).
__device__ __forceinline__ void tile_coords(const TcParams& p, int r, int& tm, int& tn) {
  if (!p.triu_tiles) {
    tm = r / p.tiles_n;
    tn = r - tm * p.tiles_n;
    return;
  }
  tm = 0;
  for (;;) {
    const int first = (tm * p.tile_rows) / BN;
    const int cnt = p.tiles_n - first;
    if (r < cnt) { tn = first + r; return; }
    r -= cnt;
    ++tm;
  }
}

// Symmetric operand kept as its upper 256 x 256 blocks: major-ness of the load that fetches the
// logical block (rows r0.., K columns k0..) of an A operand / (K rows k0.., columns c0..) of a B
// operand. A block that is absent is read from its mirror image, i.e. with the other major-ness.
__device__ __forceinline__ int sym_a_mn(int r0, int k0) { return (k0 >> 8) < (r0 >> 8) ? 1 : 0; }
__device__ __forceinline__ int sym_b_mn(int k0, int c0) { return (c0 >> 8) >= (k0 >> 8) ? 1 : 0; }

// 32 consecutive elements of row `row` (columns col0..col0+31) of a batched matrix stored as
// bf16 hi(/lo) planes (mode 1) or fp32 (mode 2); columns >= ncols read as zero.
__device__ __forceinline__ void load_chunk(const void* P0, const void* P1, int mode, long long ld,
                                           long long bs, int b, int row, int col0, int ncols,
                                           float (&e)[32]) {
  const bool full = (col0 + 32 <= ncols);
  if (mode == 1) {
    const __nv_bfloat16* eh = static_cast<const __nv_bfloat16*>(P0) + b * bs + (long long)row * ld + col0;
    const __nv_bfloat16* el =
        P1 ? static_cast<const __nv_bfloat16*>(P1) + b * bs + (long long)row * ld + col0 : nullptr;
    if (full && (ld & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(eh) + j);
        const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          e[j * 8 + 2 * w] = __uint_as_float(hw[w] << 16);
          e[j * 8 + 2 * w + 1] = __uint_as_float(hw[w] & 0xFFFF0000u);
        }
        if (el) {
          const uint4 l = __ldg(reinterpret_cast<const uint4*>(el) + j);
          const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            e[j * 8 + 2 * w] += __uint_as_float(lw[w] << 16);
            e[j * 8 + 2 * w + 1] += __uint_as_float(lw[w] & 0xFFFF0000u);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = 0.f;
        if (col0 + j < ncols) {
          x = __bfloat162float(eh[j]);
          if (el) x += __bfloat162float(el[j]);
        }
        e[j] = x;
      }
    }
  } else {
    const float* ef = static_cast<const float*>(P0) + b * bs + (long long)row * ld + col0;
    if (full && (ld & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(ef) + j);
        e[4 * j] = x.x; e[4 * j + 1] = x.y; e[4 * j + 2] = x.z; e[4 * j + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = (col0 + j < ncols) ? ef[j] : 0.f;
    }
  }
}

// two floats -> packed bf16x2 (round to nearest even); `lo` lands in bits [0,16)
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// one lane's row of a 32 x 32 chunk -> a 2 KiB plane staging area (rows of 64 B, 64B swizzle): the
// bf16 hi plane (LO = false) or the bf16 plane of the residuals x - hi (LO = true). Four words at a
// time, so no 16-word arrays stay live next to the chunk and the TMEM load in flight.
template <bool LO>
__device__ __forceinline__ void stage_plane(uint32_t buf, int lane, const float (&o)[32]) {
  const uint32_t sw = (lane >> 1) & 3;       // 64B swizzle: 16B chunk ^= addr bits [7,9)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x1 = o[j * 8 + 2 * q + 1], x0 = o[j * 8 + 2 * q];
      const uint32_t h = cvt_bf16x2(x1, x0);
      w[q] = LO ? cvt_bf16x2(x1 - __uint_as_float(h & 0xFFFF0000u), x0 - __uint_as_float(h << 16)) : h;
    }
    ptx::sts128(buf + lane * 64 + ((j ^ sw) << 4), w[0], w[1], w[2], w[3]);
  }
}
// one lane's row of a 32 x 32 chunk -> bf16 hi/lo staging slot (hi at +0, lo at +2048)
__device__ __forceinline__ void stage_planes(uint32_t buf, int lane, const float (&o)[32], bool has_lo) {
  stage_plane<false>(buf, lane, o);
  if (has_lo) stage_plane<true>(buf + 2048, lane, o);
}

// Plane output of one chunk: TMA store through the warp's staging slot when the layout allows
// it (full-line writes, no LSU traffic, ragged edges clipped by the tensor map), else scalar.
// The slot is two 2 KiB halves and every half-store is its own bulk group, so a half is rewritten
// only after the store issued two groups earlier has read it (wait_group.read 1): with hi + lo
// planes the halves are the two planes, with a single plane they alternate (`half_slot`).
__device__ __forceinline__ void store_planes(const float (&o)[32], bool tma, const CUtensorMap* tmh,
                                             const CUtensorMap* tml, __nv_bfloat16* hi,
                                             __nv_bfloat16* lo, long long ld, long long bs, int b,
                                             int row, int row0, int col0, int ncols, bool row_ok,
                                             int lane, uint32_t buf, int& half_slot, bool split) {
  if (tma) {
    if (lo && !split) {
      // one bulk group per chunk: the whole slot must have drained
      if (lane == 0) ptx::bulk_wait_read<0>();
      __syncwarp();
      stage_planes(buf, lane, o, true);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_3d(tmh, buf, col0, row0, b);
        ptx::tma_store_3d(tml, buf + 2048, col0, row0, b);
        ptx::bulk_commit();
      }
      return;
    }
    const uint32_t hbuf = lo ? buf : buf + half_slot * 2048;
    if (!lo) half_slot ^= 1;
    if (lane == 0) ptx::bulk_wait_read<1>();   // the store that last used this half has drained
    __syncwarp();
    stage_plane<false>(hbuf, lane, o);
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_3d(tmh, hbuf, col0, row0, b);
      ptx::bulk_commit();
    }
    if (lo) {
      if (lane == 0) ptx::bulk_wait_read<1>();
      __syncwarp();
      stage_plane<true>(buf + 2048, lane, o);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_3d(tml, buf + 2048, col0, row0, b);
        ptx::bulk_commit();
      }
    }
  } else if (row_ok) {
    __nv_bfloat16* ch = hi + b * bs + (long long)row * ld + col0;
    __nv_bfloat16* cl = lo ? lo + b * bs + (long long)row * ld + col0 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j < ncols) {
        __nv_bfloat16 h, l;
        split_bf16(o[j], h, l);
        ch[j] = h;
        if (cl) cl[j] = l;
      }
  }
}

// Epilogue of one accumulator tile for one warp: TMEM -> registers -> alpha*acc + beta*I +
// gamma*E -> bf16 hi/lo planes and/or fp32 (+ optional secondary planes, <C,F> partial, packed
// upper triangle). `t_addr` addresses this warp's 32 TMEM lanes, `row0` is the warp's first
// output row (a multiple of 32), `n0` the first output column of the 256-wide tile; the warp
// handles the 32-column chunks [c_begin, c_end). The TMEM load of chunk c+1 is in flight while
// chunk c is processed; one 4 KiB staging slot per warp (`buf`) feeds the TMA stores
// (`half_slot`: which 2 KiB half the next hi-only store uses, carried across tiles).
// Returns this lane's share of <C, F>.
__device__ __forceinline__ float epilogue_tile(const TcParams& p, uint32_t t_addr, int b, int row0,
                                               int lane, int n0, uint32_t buf, int c_begin, int c_end,
                                               int& half_slot) {
  const int row = row0 + lane;
  const bool row_ok = row < p.M;
  const float a_eff = p.alpha * (p.alpha_b ? __ldg(p.alpha_b + b) : 1.f);
  const float beta = p.beta_eye * (p.beta_b ? __ldg(p.beta_b + b) : 1.f);
  const float gamma = p.gamma * (p.gamma_b ? __ldg(p.gamma_b + b) : 1.f);
  float dsum = 0.f;
  // warp-uniform chunk range: stop at N, and skip chunks entirely below the diagonal for X
  while (c_end > c_begin && n0 + (c_end - 1) * 32 >= p.N) --c_end;
  if (p.X_hi)
    while (c_begin < c_end && n0 + c_begin * 32 + 31 < row0) ++c_begin;
  if (c_begin >= c_end) return 0.f;
  // The addend E / dot operand F of a chunk is a dependent global load in the middle of the chunk's
  // work (a DRAM miss: the matrix was written by an earlier launch and is larger than L2). A cache
  // prefetch of the NEXT chunk's lines, issued before this chunk is converted and stored, hides that
  // latency at no register cost (holding the data itself would take 32 registers of a 168 budget).
  auto prefetch_ef = [&](int col0) {
    if (!row_ok || col0 >= p.N) return;
    if (p.e_mode) {
      const int es = p.e_mode == 1 ? 2 : 4;
      const char* e0 = static_cast<const char*>(p.E0) + ((long long)b * p.bsE + (long long)row * p.ldE + col0) * es;
      ptx::prefetch_l2(e0);
      if (p.e_mode == 1 && p.E1)
        ptx::prefetch_l2(static_cast<const char*>(p.E1) + ((long long)b * p.bsE + (long long)row * p.ldE + col0) * es);
    }
    if (p.f_mode) {
      const int fs = p.f_mode == 1 ? 2 : 4;
      const char* f0 = static_cast<const char*>(p.F0) + ((long long)b * p.bsF + (long long)row * p.ldF + col0) * fs;
      ptx::prefetch_l2(f0);
      if (p.f_mode == 1 && p.F1)
        ptx::prefetch_l2(static_cast<const char*>(p.F1) + ((long long)b * p.bsF + (long long)row * p.ldF + col0) * fs);
    }
  };
  const bool has_ef = p.e_mode || p.f_mode;
  if (has_ef) prefetch_ef(n0 + c_begin * 32);
  uint32_t v[32];
  ptx::tmem_ld_32x32(t_addr + c_begin * 32, v);
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int col0 = n0 + c * 32;
    if (has_ef && c + 1 < c_end) prefetch_ef(col0 + 32);
    ptx::tmem_ld_wait();
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = a_eff * __uint_as_float(v[j]);
    if (c + 1 < c_end) ptx::tmem_ld_32x32(t_addr + (c + 1) * 32, v);   // overlaps the work below
    // rows and columns of a chunk are both 32-aligned: the diagonal crosses it iff col0 == row0
    const bool diag = (col0 == row0);
    if (beta != 0.f && diag) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == lane) o[j] += beta;
    }
    if (p.e_mode && row_ok) {
      float e[32];
      load_chunk(p.E0, p.E1, p.e_mode, p.ldE, p.bsE, b, row, col0, p.N, e);
#pragma unroll
      for (int j = 0; j < 32; ++j) o[j] = fmaf(gamma, e[j], o[j]);
    }
    if (p.f_mode && row_ok) {
      float f[32];
      load_chunk(p.F0, p.F1, p.f_mode, p.ldF, p.bsF, b, row, col0, p.N, f);
#pragma unroll
      for (int j = 0; j < 32; ++j) dsum = fmaf(o[j], f[j], dsum);
    }
    if (p.Cp_hi)
      store_planes(o, p.tma_cp != 0, &p.tmC[0], &p.tmC[1], p.Cp_hi, p.Cp_lo, p.ldCp, p.bsCp, b, row, row0,
                   col0, p.N, row_ok, lane, buf, half_slot, p.epi_split != 0);
    if (p.C2_hi) {
      float o2[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) o2[j] = p.c2_scale * o[j];
      if (p.c2_eye != 0.f && diag) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j == lane) o2[j] += p.c2_eye;
      }
      store_planes(o2, p.tma_c2 != 0, &p.tmC2[0], &p.tmC2[1], p.C2_hi, p.C2_lo, p.ldC2, p.bsC2, b, row,
                   row0, col0, p.N, row_ok, lane, buf, half_slot, p.epi_split != 0);
    }
    if (p.X_hi) {
      // packed upper triangle: stage the chunk, then every row leaves as one contiguous run
      __syncwarp();
      stage_planes(buf, lane, o, p.X_lo != nullptr);
      __syncwarp();
      const int j = col0 + lane;
      const int rows = min(32, p.M - row0);
      // packed offset of (i, j) is base_i + j with base_{i+1} = base_i + N - i - 1: one add per row
      long long base = (long long)b * p.ldX + (long long)row0 * p.N - (long long)row0 * (row0 - 1) / 2 - row0;
      const uint32_t lane_off = (lane & 7) * 2;
      const uint32_t lane_sw = lane >> 3;
      const bool has_lo = p.X_lo != nullptr;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (r < rows && j >= row0 + r && j < p.N) {
          const uint32_t off = r * 64 + ((lane_sw ^ ((r >> 1) & 3)) << 4) + lane_off;
          p.X_hi[base + j] = __ushort_as_bfloat16(ptx::lds16(buf + off));
          if (has_lo) p.X_lo[base + j] = __ushort_as_bfloat16(ptx::lds16(buf + 2048 + off));
        }
        base += p.N - (row0 + r) - 1;
      }
    }
    if (p.Cf) {
      if (p.tma_cf) {
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
        const uint32_t sw = lane & 7;              // 128B swizzle: 16B chunk ^= addr bits [7,10)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ptx::sts128(buf + lane * 128 + ((j ^ sw) << 4), __float_as_uint(o[4 * j]),
                      __float_as_uint(o[4 * j + 1]), __float_as_uint(o[4 * j + 2]),
                      __float_as_uint(o[4 * j + 3]));
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_3d(&p.tmC[2], buf, col0, row0, b);
          ptx::bulk_commit();
        }
      } else if (row_ok) {
        float* cf = p.Cf + b * p.bsCf + (long long)row * p.ldCf + col0;
        if (col0 + 32 <= p.N && (p.ldCf & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(cf)[j] =
                make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) cf[j] = o[j];
        }
      }
    }
  }
  return dsum;
}

// this warp's <C,F> partial of one tile -> dot_ws[(b * tiles_per_img + r) * nwarps + w]
__device__ __forceinline__ void write_dot_partial(const TcParams& p, float dsum, int lane, long long slot_idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
  if (lane == 0) p.dot_ws[slot_idx] = dsum;
}

// dot_out[b] = sum of the image's partials, fixed order (one warp per image)
__global__ void dot_reduce_kernel(const float* __restrict__ ws, int per_img, int batch,
                                  float* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const int lane = threadIdx.x & 31;
  float a = 0.f;
  for (int i = lane; i < per_img; i += 32) a += ws[(long long)b * per_img + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[b] = a;
}

// ------------------------------------------------------------------ the kernel
// A CTA pair (cluster of 2, one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2:
// each CTA stages its own 128 rows of A and HALF of B's 256 columns, so the shared-memory
// fill per MMA cycle is a third below a single CTA's 128 x 256 tile and the freed space buys a
// deeper ring (3 x 64 KiB stages for bf16x3, 6 x 32 KiB for single pass). Only the leader CTA
// issues MMAs; its commits are multicast to the mbarriers of both CTAs. (A single-CTA variant of
// this kernel measured 1155 TFLOP/s executed against 1334 for the pair in round 1 and was retired.)
constexpr int kTileBh = (BN / 2) * BK * 2;  // 16 KiB: this CTA's half of the B tile

// Dynamic tile scheduler. Tiles are handed out by an atomic counter in global memory instead of the
// static `tile += nclusters` walk, so a CTA pair that starts late (its SMs were held by another
// kernel - NCCL's all-reduce CTAs under the Newton-Schulz backward in data-parallel training) simply
// takes fewer tiles instead of becoming the tail of every launch. The leader CTA's producer warp
// fetches the next tile index and publishes it through a small ring in the shared memory of BOTH
// CTAs (plain stores + cluster-scope release/acquire mbarriers); the peer's producer, the MMA warp
// and the 16 epilogue warps each read it and free the slot. -1 ends every role's loop.
// The counter pair of a launch is one slot of a device-global pool picked round-robin on the host;
// the last pair to finish zeroes it, so a slot is clean long before it comes round again.
constexpr int kSched = 4;                          // ring depth (tiles the producer may run ahead)
constexpr int kSchedReaders = 2 * kEpiWarps + 2;   // leader: MMA + 8 epilogue; peer: producer + 8 epilogue
constexpr int kSchedPool = 4096;
__device__ unsigned g_sched_pool[kSchedPool][2];

template <int NPASS>
struct Cfg2 {
  static constexpr int kPlanes = (NPASS == 3) ? 2 : 1;
  static constexpr int kStageBytes = kPlanes * (kTileA + kTileBh);     // 64 KiB / 32 KiB
  static constexpr int kStages = (NPASS == 3) ? 3 : 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 + 256;
};

template <int NPASS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ TcParams p) {
  using C = Cfg2<NPASS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = smem_base + C::kStages * C::kStageBytes;
  const uint32_t bar_base = epi_base + kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::kStages + 4);
  auto sfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 5 + s); };
  auto sempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 5 + kSched + s); };
  const uint32_t stile = bar_base + 8u * (2 * C::kStages + 5 + 2 * kSched);   // kSched x int
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + C::kStages * C::kStageBytes + kEpiBytes + 8 * (2 * C::kStages + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);

  // The launch's counter slot is untouched since the kernel that last used it zeroed it (kSchedPool
  // launches ago), so the first fetch may run before griddep_wait(), under the previous kernel's tail.
  int next_fetch = 0;
  if (warp == 0 && lane == 0) {
    if (leader) next_fetch = p.sched ? (int)atomicAdd(p.sched, 1u) : (int)(blockIdx.x >> 1);
    for (int t = 0; t < p.nterms; ++t)
      for (int i = 0; i < 4; ++i)
        if (i % 2 == 0 || NPASS == 3) ptx::prefetch_tensormap(&p.tm[t][i]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kSched; ++s) {
        ptx::mbar_init(sfull_bar(s), 1);                 // the leader's producer publishes
        ptx::mbar_init(sempty_bar(s), kSchedReaders);    // used in the leader: every reader of both CTAs
      }
      for (int s = 0; s < C::kStages; ++s) {
        ptx::mbar_init(full_bar(s), 1);    // used in the leader: its own arrive.expect_tx
        ptx::mbar_init(empty_bar(s), 1);   // multicast commit of the leader's MMA warp
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);   // multicast commit
        ptx::mbar_init(tempty_bar(a), 2 * kEpiWarps);  // used in the leader: epilogue warps x 2 CTAs
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc_2sm(tmem_slot, kTmemCols);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // programmatic dependent launch: the set-up above may overlap the tail of the previous kernel of the
  // stream; nothing below touches global memory before that kernel has completed
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();

  const int tiles_per_img = p.tiles_per_img;             // tiles_m counts 256-row tiles here
  const int ntiles = tiles_per_img * p.batch;
  const int nclusters = gridDim.x >> 1;
  // a reader's side of the ring: wait for slot `it`, read the tile index, free the slot (in the leader)
  auto next_tile = [&](int it, uint32_t sleep_ns) -> int {
    const int s = it % kSched;
    ptx::mbar_wait_cluster(sfull_bar(s), (uint32_t)(it / kSched) & 1u, sleep_ns);
    const int t = (int)ptx::ld_shared_u32(stile + 4u * s);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_release_cluster(ptx::mapa(sempty_bar(s), 0));
    return t;
  };

  if (warp == 0) {
    // --------------------------------------------- TMA producer (both CTAs of the pair)
    // warp-uniform loop; one elected lane issues the copies (see the 1-CTA kernel)
    {
      const bool issuer = ptx::elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0;; ++it) {
        int tile;
        if (leader) {
          // the scheduler: fetch, publish to both CTAs' rings
          const int s = it % kSched;
          ptx::mbar_wait_cluster(sempty_bar(s), ((uint32_t)(it / kSched) & 1u) ^ 1u);
          tile = 0;
          if (lane == 0) {
            // the index fetched one tile ago; the fetch for the tile after this one goes out now and its
            // latency hides under this tile's loads (p.sched == nullptr: the static walk, for A/B runs)
            tile = next_fetch;
            if (tile >= ntiles) tile = -1;
            else next_fetch = p.sched ? (int)atomicAdd(p.sched, 1u) : tile + nclusters;
            ptx::st_shared_cluster_u32(ptx::mapa(stile + 4u * s, 0), (uint32_t)tile);
            ptx::st_shared_cluster_u32(ptx::mapa(stile + 4u * s, 1), (uint32_t)tile);
            ptx::mbar_arrive_release_cluster(ptx::mapa(sfull_bar(s), 0));
            ptx::mbar_arrive_release_cluster(ptx::mapa(sfull_bar(s), 1));
          }
          tile = __shfl_sync(0xffffffffu, tile, 0);
        } else {
          tile = next_tile(it, 0);
        }
        if (tile < 0) break;
        const int b = tile / tiles_per_img;
        const int r = tile - b * tiles_per_img;
        int tm, tn;
        tile_coords(p, r, tm, tn);
        const int m0 = tm * (2 * BM) + rank * BM;        // this CTA's 128 rows
        const int n0 = tn * BN + rank * (BN / 2);        // this CTA's 128 columns of B
        const int bl = p.splits > 1 ? 0 : b;   // batch coordinate of the operand loads
        for (int t = 0; t < p.nterms; ++t) {
          const int a_sym = p.a_sym[t], b_sym = p.b_sym[t];
          const int a_kboxes = a_sym ? BM / 64 : 1;        // symmetric operands: 64-row boxes in
          const int b_kboxes = b_sym ? BN / 128 : 1;       // both major-nesses (one tensor map)
          const int nkb_all = (p.K[t] + BK - 1) / BK;
          const int kb_lo = p.splits > 1 ? (int)((long long)b * nkb_all / p.splits) : 0;
          const int kb_hi = p.splits > 1 ? (int)((long long)(b + 1) * nkb_all / p.splits) : nkb_all;
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
            ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t fb = ptx::mapa(full_bar(stage), 0);         // the leader's barrier
            const uint32_t sA = smem_base + stage * C::kStageBytes;
            const uint32_t sB = sA + C::kPlanes * kTileA;
            const int k0 = kb * BK;
            // both CTAs of the pair sit in the same 256-block of rows / columns: same decision
            const int a_mn = a_sym ? sym_a_mn(m0, k0) : p.a_mn[t];
            const int b_mn = b_sym ? sym_b_mn(k0, n0) : p.b_mn[t];
            if (issuer) {
              if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C::kStageBytes);
#pragma unroll
              for (int pl = 0; pl < C::kPlanes; ++pl) {
                if (!a_mn) {
                  for (int j = 0; j < a_kboxes; ++j)
                    ptx::tma_load_3d_2sm(&p.tm[t][pl], fb, sA + pl * kTileA + j * kChunk, k0, m0 + 64 * j, bl);
                } else {
#pragma unroll
                  for (int j = 0; j < BM / 64; ++j)
                    ptx::tma_load_3d_2sm(&p.tm[t][pl], fb, sA + pl * kTileA + j * kChunk,
                                         m0 + 64 * j, k0, bl);
                }
                if (!b_mn) {
                  for (int j = 0; j < b_kboxes; ++j)
                    ptx::tma_load_3d_2sm(&p.tm[t][2 + pl], fb, sB + pl * kTileBh + j * kChunk, k0, n0 + 64 * j, bl);
                } else {
#pragma unroll
                  for (int j = 0; j < BN / 128; ++j)
                    ptx::tma_load_3d_2sm(&p.tm[t][2 + pl], fb, sB + pl * kTileBh + j * kChunk,
                                         n0 + 64 * j, k0, bl);
                }
              }
            }
            __syncwarp();
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      const bool issuer = ptx::elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = 0;; ++it) {
        const int tile = next_tile(it, 0);
        if (tile < 0) break;
        const int bt = tile / tiles_per_img;
        int tm, tn;
        tile_coords(p, tile - bt * tiles_per_img, tm, tn);
        const int m0 = tm * (2 * BM), n0 = tn * BN;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int t = 0; t < p.nterms; ++t) {
          const int a_sym = p.a_sym[t], b_sym = p.b_sym[t];
          const int nkb_all = (p.K[t] + BK - 1) / BK;
          const int kb_lo = p.splits > 1 ? (int)((long long)bt * nkb_all / p.splits) : 0;
          const int kb_hi = p.splits > 1 ? (int)((long long)(bt + 1) * nkb_all / p.splits) : nkb_all;
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
            const int a_mn = a_sym ? sym_a_mn(m0, kb * BK) : p.a_mn[t];
            const int b_mn = b_sym ? sym_b_mn(kb * BK, n0) : p.b_mn[t];
            const uint32_t idesc = ptx::idesc_bf16_f32(2 * BM, BN, a_mn, b_mn);
            const uint32_t a_step = a_mn ? 2048u : 32u;
            const uint32_t b_step = b_mn ? 2048u : 32u;
            const uint32_t a_lbo = a_mn ? kChunk : 0u;
            const uint32_t b_lbo = b_mn ? kChunk : 0u;
            ptx::mbar_wait(full_bar(stage), phase);
            ptx::tc_fence_after();
            const uint32_t sA = smem_base + stage * C::kStageBytes;
            const uint32_t sB = sA + C::kPlanes * kTileA;
            if (issuer) {
#pragma unroll
              for (int kk = 0; kk < BK / UK; ++kk) {
                const uint64_t ah = ptx::smem_desc_sw128(sA + kk * a_step, a_lbo, 1024);
                const uint64_t bh = ptx::smem_desc_sw128(sB + kk * b_step, b_lbo, 1024);
                if (NPASS == 3) {
                  const uint64_t al = ptx::smem_desc_sw128(sA + kTileA + kk * a_step, a_lbo, 1024);
                  const uint64_t bl = ptx::smem_desc_sw128(sB + kTileBh + kk * b_step, b_lbo, 1024);
                  // neighbouring MMAs share an operand (bh, then ah); A/B on one box: -1.5 % on the
                  // one-term products against the (al bh, ah bl, ah bh) order, nothing on two-term ones
                  ptx::mma_bf16_ss_2sm(d_tmem, al, bh, idesc, (kk == 0) ? accumulate : 1u);
                  ptx::mma_bf16_ss_2sm(d_tmem, ah, bh, idesc, 1u);
                  ptx::mma_bf16_ss_2sm(d_tmem, ah, bl, idesc, 1u);
                } else {
                  ptx::mma_bf16_ss_2sm(d_tmem, ah, bh, idesc, (kk == 0) ? accumulate : 1u);
                }
              }
              ptx::tc_commit_2sm(empty_bar(stage), 0x3);   // frees the slot in both CTAs
            }
            __syncwarp();
            accumulate = 1u;
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
        }
        if (issuer) ptx::tc_commit_2sm(tfull_bar(acc), 0x3);   // both CTAs' epilogues may start
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ----------------------------------------------------- epilogue (both CTAs, own rows)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int cb = half * (BN / 64), ce = cb + BN / 64;
    int half_slot = 0;                  // persists across tiles: a store may still be in flight
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = 0;; ++it) {
      const int tile = next_tile(it, p.epi_sleep_ns);
      if (tile < 0) break;
      const int b = tile / tiles_per_img;
      const int r = tile - b * tiles_per_img;
      int tm, tn;
      tile_coords(p, r, tm, tn);
      const int m0 = tm * (2 * BM) + rank * BM;
      const int n0 = tn * BN;
      ptx::mbar_wait_relaxed(tfull_bar(acc), acc_phase, p.epi_sleep_ns);
      ptx::tc_fence_after();
      const float ds = epilogue_tile(p, tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16), b,
                                     m0 + q * 32, lane, n0, epi_base + (warp - 2) * kEpiWarpBytes, cb, ce, half_slot);
      if (p.dot_ws)
        write_dot_partial(p, (p.sym_out && (n0 >> 8) > (m0 >> 8)) ? 2.f * ds : ds, lane,
                          (long long)tile * (2 * kEpiWarps) + rank * kEpiWarps + (warp - 2));
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(tempty_bar(acc));
        else ptx::mbar_arrive_cluster(ptx::mapa(tempty_bar(acc), 0));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) ptx::bulk_wait_all();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();   // neither CTA may retire while its peer can still touch it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
  if (leader && threadIdx.x == 0 && p.sched) {
    // this pair has seen the end of the tile list; the last pair leaves the slot clean for its next user
    if (atomicAdd(p.sched + 1, 1u) == (unsigned)nclusters - 1u) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  // libcuda is not linked (the library is built on a GPU-less box): resolve at run time.
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// bf16 plane [batch][rows][cols] -> 3-D tiled map with a {64, box_rows, 1} box, 128B swizzle.
// Out-of-range box elements read as zero, so ragged M/N/K edges need no special casing.
bool make_plane_map(CUtensorMap* tm, const void* base, const Mat& m, int batch, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return false;
  }
  const long long bs = (batch > 1) ? m.bstride : (long long)m.rows * m.ld;
  cuuint64_t dims[3] = {(cuuint64_t)m.cols, (cuuint64_t)m.rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)m.ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rows=%d cols=%d ld=%lld bstride=%lld batch=%d",
              (int)r, m.rows, m.cols, m.ld, bs, batch);
    return false;
  }
  return true;
}

// Output map for the epilogue's TMA stores: {32 cols, 32 rows, 1 image} boxes.
bool make_store_map(CUtensorMap* tm, void* base, int rows, int cols, long long ld, long long bstride,
                    int batch, bool f32) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  const int es = f32 ? 4 : 2;
  const long long bs = (batch > 1) ? bstride : (long long)rows * ld;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * es, (cuuint64_t)bs * es};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base,
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(store) failed (%d): rows=%d cols=%d ld=%lld", (int)r, rows, cols, ld);
    return false;
  }
  return true;
}

bool plane_ok(const Mat& m, int batch, bool need_lo) {
  if (!m.p0 || (need_lo && !m.p1)) return false;
  if ((reinterpret_cast<uintptr_t>(m.p0) & 15) || (need_lo && (reinterpret_cast<uintptr_t>(m.p1) & 15)))
    return false;
  if (m.ld % 8 != 0 || m.ld < m.cols) return false;
  if (batch > 1 && (m.bstride % 8 != 0 || m.bstride <= 0)) return false;
  return true;
}

// EGM_PDL=0 turns programmatic dependent launch off (A/B switch)
bool pdl_enabled() {
  static bool on = []() { const char* e = getenv("EGM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

template <int NPASS>
cudaError_t launch2(const TcParams& p, cudaStream_t stream) {
  using C = Cfg2<NPASS>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (configured_dev != dev) {
    e = cudaFuncSetAttribute(gemm_tc2_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             C::kSmemBytes);
    if (e != cudaSuccess) return e;
    configured_dev = dev;
  }
  // one counter pair of the device-global pool per launch, round-robin (see g_sched_pool)
  // (the sequence number is process-wide: two host threads launching on the same device must never
  // share a slot, or their kernels would split one tile list between them)
  static thread_local unsigned* pool = nullptr;
  static thread_local int pool_dev = -1;
  static std::atomic<unsigned> seq{0};
  if (pool_dev != dev) {
    void* sym = nullptr;
    e = cudaGetSymbolAddress(&sym, g_sched_pool);
    if (e != cudaSuccess) return e;
    pool = static_cast<unsigned*>(sym);
    pool_dev = dev;
  }
  // EGM_SCHED=0: static round-robin tiles (A/B switch)
  static const bool dynamic = []() { const char* e = getenv("EGM_SCHED"); return !(e && e[0] == '0'); }();
  TcParams q = p;
  q.sched = dynamic ? pool + 2u * (seq.fetch_add(1u, std::memory_order_relaxed) % (unsigned)kSchedPool) : nullptr;
  int sms = 0;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const long long ntiles = (long long)p.tiles_per_img * p.batch;
  const long long pairs = sms / 2;
  const int grid = 2 * (int)(ntiles < pairs ? ntiles : pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<NPASS>, q);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

// hi and lo plane stores of a chunk as one bulk group or two. Two groups keep a store in flight while
// the other half is restaged: -2 % on the long-K Newton-Schulz products (mainloop-bound, the epilogue
// only has to stay out of the way), +5 % on short-K products whose time IS the epilogue (measured A/B
// on one box). EGM_EPI_SPLIT=0/1 forces either.
int epi_split_for(int total_k, int m) {
  static int forced = []() {
    const char* e = getenv("EGM_EPI_SPLIT");
    return !e ? -1 : (e[0] == '0' ? 0 : 1);
  }();
  // long K and more than one row tile per image: the D x D x D chain products, not V = Zc dA (M = N_tokens)
  return forced >= 0 ? forced : ((total_k >= 512 && m > 256) ? 1 : 0);
}

constexpr int kCtas = 2;   // CTAs per tile (cta_group::2)

}  // namespace

bool gemm_tc_supported(const GemmProblem& g, int npass) {
  const bool lo = (npass == 3);
  for (int t = 0; t < g.nterms; ++t) {
    if (!plane_ok(g.t[t].A, g.batch, lo) || !plane_ok(g.t[t].B, g.batch, lo)) return false;
  }
  if (g.Cp.p0 && (g.Cp.ld < g.N)) return false;
  return g.M > 0 && g.N > 0 && g.batch > 0;
}

// number of tiles per image and (for triu_tiles) the upper-triangular subset
static int count_tiles(int tiles_m, int tiles_n, int tile_rows, bool triu) {
  if (!triu) return tiles_m * tiles_n;
  int n = 0;
  for (int tm = 0; tm < tiles_m; ++tm) {
    const int first = (tm * tile_rows) / BN;
    if (first < tiles_n) n += tiles_n - first;
  }
  return n;
}

size_t gemm_tc_dot_ws_floats(const GemmProblem& g) {
  if (!g.dot_out) return 0;
  const int ctas = kCtas;
  const int tiles_m = (g.M + ctas * BM - 1) / (ctas * BM);
  const int tiles_n = (g.N + BN - 1) / BN;
  return (size_t)g.batch * tiles_m * tiles_n * kEpiWarps * ctas;
}

cudaError_t gemm_tc(const GemmProblem& g, int npass, cudaStream_t stream) {
  if (npass != 1 && npass != 3) {
    set_error("gemm_tc: npass must be 1 or 3");
    return cudaErrorInvalidValue;
  }
  if (g.nterms < 1 || g.nterms > 2 || !gemm_tc_supported(g, npass)) {
    set_error("gemm_tc: operand layout not addressable by TMA (need 16B-aligned planes, ld%%8==0)");
    return cudaErrorInvalidValue;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M;
  p.N = g.N;
  p.batch = g.batch;
  p.nterms = g.nterms;
  p.splits = 1;
  const int ctas = kCtas;
  p.tiles_m = (g.M + ctas * BM - 1) / (ctas * BM);
  p.tiles_n = (g.N + BN - 1) / BN;
  p.tile_rows = ctas * BM;
  p.triu_tiles = (g.X.p0 || g.sym_out) ? 1 : 0;
  p.sym_out = g.sym_out ? 1 : 0;
  if (g.sym_out && (g.M != g.N || g.split_k > 1)) {
    set_error("gemm_tc: sym_out needs a square result and excludes split-K");
    return cudaErrorInvalidValue;
  }
  p.tiles_per_img = count_tiles(p.tiles_m, p.tiles_n, p.tile_rows, p.triu_tiles != 0);
  if (g.X.p0) {
    if (g.M != g.N || g.Cp.p0 || g.Cp2.p0 || g.split_k > 1 || g.dot_out) {
      set_error("gemm_tc: the packed-triangle output needs a square result and excludes Cp/Cp2/split-K/dot");
      return cudaErrorInvalidValue;
    }
    p.X_hi = static_cast<__nv_bfloat16*>(g.X.p0);
    p.X_lo = static_cast<__nv_bfloat16*>(g.X.p1);
    p.ldX = g.X.ld;
  }
  for (int t = 0; t < g.nterms; ++t) {
    const GemmTerm& gt = g.t[t];
    p.K[t] = gt.K;
    p.a_mn[t] = gt.transA ? 1 : 0;   // stored [K,M]  -> M-major
    p.b_mn[t] = gt.transB ? 0 : 1;   // stored [K,N]  -> N-major ; [N,K] -> K-major
    p.a_sym[t] = gt.symA ? 1 : 0;
    p.b_sym[t] = gt.symB ? 1 : 0;
    if ((gt.symA && g.M != gt.K) || (gt.symB && g.N != gt.K) || ((gt.symA || gt.symB) && g.split_k > 1)) {
      set_error("gemm_tc: a symmetric operand must be square (M == K / N == K) and excludes split-K");
      return cudaErrorInvalidValue;
    }
    const int a_rows = (gt.transA || gt.symA) ? BK : BM;
    const int b_rows = gt.symB ? BK : (gt.transB ? BN / ctas : BK);
    // logical extents double as the TMA bounds: whatever lies outside reads as zero
    Mat A = gt.A, B = gt.B;
    A.rows = (gt.transA && !gt.symA) ? gt.K : g.M;
    A.cols = (gt.transA && !gt.symA) ? g.M : gt.K;
    B.rows = (gt.transB && !gt.symB) ? g.N : gt.K;
    B.cols = (gt.transB && !gt.symB) ? gt.K : g.N;
    const int ab_batch = g.split_k > 1 ? 1 : g.batch;
    if (!make_plane_map(&p.tm[t][0], A.p0, A, ab_batch, a_rows)) return cudaErrorInvalidValue;
    if (!make_plane_map(&p.tm[t][2], B.p0, B, ab_batch, b_rows)) return cudaErrorInvalidValue;
    if (npass == 3) {
      if (!make_plane_map(&p.tm[t][1], A.p1, A, ab_batch, a_rows)) return cudaErrorInvalidValue;
      if (!make_plane_map(&p.tm[t][3], B.p1, B, ab_batch, b_rows)) return cudaErrorInvalidValue;
    }
  }
  p.epi_split = epi_split_for(g.t[0].K + (g.nterms > 1 ? g.t[1].K : 0), g.M);
  {
    // 256 ns: -0.5 % Newton-Schulz time against pure spinning (A/B on one box); EGM_EPI_SLEEP_NS overrides
    static unsigned ns = []() { const char* e = getenv("EGM_EPI_SLEEP_NS"); return e ? (unsigned)atoi(e) : 256u; }();
    p.epi_sleep_ns = ns;
  }
  if (g.split_k > 1) {
    // K-slices of one product; slice s writes partial sums to output batch index s
    const int nkb = (g.t[0].K + BK - 1) / BK;
    if (g.nterms != 1 || g.batch != 1 || g.split_k > nkb || g.Cp.p0 || g.Cp2.p0 || g.dot_out || !g.Cf.p0) {
      set_error("gemm_tc: split-K needs one term, batch 1, an fp32 output only and split_k <= K/64");
      return cudaErrorInvalidValue;
    }
    p.splits = g.split_k;
    p.batch = g.split_k;
  }
  p.alpha = g.alpha;
  p.alpha_b = g.alpha_b;
  p.beta_b = g.beta_b;
  p.gamma_b = g.gamma_b;
  p.beta_eye = g.beta_eye;
  p.gamma = g.gamma;
  if (g.E.p0 && g.gamma != 0.f) {
    p.E0 = g.E.p0;
    p.E1 = g.e_planes ? g.E.p1 : nullptr;
    p.ldE = g.E.ld;
    p.bsE = g.E.bstride;
    p.e_mode = g.e_planes ? 1 : 2;
  }
  if (g.dot_out) {
    if (!g.F.p0 || !g.dot_ws) {
      set_error("gemm_tc: dot_out needs F and dot_ws");
      return cudaErrorInvalidValue;
    }
    p.F0 = g.F.p0;
    p.F1 = g.f_planes ? g.F.p1 : nullptr;
    p.ldF = g.F.ld;
    p.bsF = g.F.bstride;
    p.f_mode = g.f_planes ? 1 : 2;
    p.dot_ws = g.dot_ws;
  }
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  auto planes_tma_ok = [&](const Mat& m) {
    return m.ld % 8 == 0 && (g.batch == 1 || m.bstride % 8 == 0) && al16(m.p0) && (!m.p1 || al16(m.p1));
  };
  if (g.Cp.p0) {
    p.Cp_hi = static_cast<__nv_bfloat16*>(g.Cp.p0);
    p.Cp_lo = static_cast<__nv_bfloat16*>(g.Cp.p1);
    p.ldCp = g.Cp.ld;
    p.bsCp = g.Cp.bstride;
    if (planes_tma_ok(g.Cp)) {
      if (!make_store_map(&p.tmC[0], g.Cp.p0, g.M, g.N, g.Cp.ld, g.Cp.bstride, g.batch, false))
        return cudaErrorInvalidValue;
      if (g.Cp.p1 && !make_store_map(&p.tmC[1], g.Cp.p1, g.M, g.N, g.Cp.ld, g.Cp.bstride, g.batch, false))
        return cudaErrorInvalidValue;
      p.tma_cp = 1;
    }
  }
  if (g.Cp2.p0) {
    if (g.Cp2.ld < g.N) {
      set_error("gemm_tc: Cp2.ld < N");
      return cudaErrorInvalidValue;
    }
    p.C2_hi = static_cast<__nv_bfloat16*>(g.Cp2.p0);
    p.C2_lo = static_cast<__nv_bfloat16*>(g.Cp2.p1);
    p.ldC2 = g.Cp2.ld;
    p.bsC2 = g.Cp2.bstride;
    p.c2_scale = g.c2_scale;
    p.c2_eye = g.c2_eye;
    if (planes_tma_ok(g.Cp2)) {
      if (!make_store_map(&p.tmC2[0], g.Cp2.p0, g.M, g.N, g.Cp2.ld, g.Cp2.bstride, g.batch, false))
        return cudaErrorInvalidValue;
      if (g.Cp2.p1 && !make_store_map(&p.tmC2[1], g.Cp2.p1, g.M, g.N, g.Cp2.ld, g.Cp2.bstride, g.batch, false))
        return cudaErrorInvalidValue;
      p.tma_c2 = 1;
    }
  }
  if (g.Cf.p0) {
    p.Cf = static_cast<float*>(g.Cf.p0);
    p.ldCf = g.Cf.ld;
    p.bsCf = g.Cf.bstride;
    // one staging slot per chunk: fp32 takes the TMA path only when no plane output shares it
    if (!g.Cp.p0 && !g.Cp2.p0 && !g.X.p0 && g.Cf.ld % 4 == 0 && (p.batch == 1 || g.Cf.bstride % 4 == 0) &&
        al16(g.Cf.p0)) {
      if (!make_store_map(&p.tmC[2], g.Cf.p0, g.M, g.N, g.Cf.ld, g.Cf.bstride, p.batch, true))
        return cudaErrorInvalidValue;
      p.tma_cf = 1;
    }
  }
  int prof_id = -1;
  if (prof_enabled()) {
    // algorithmic flops of what is actually evaluated (skipped lower tiles are not counted)
    double flops = 0;
    for (int t = 0; t < g.nterms; ++t) flops += 2.0 * g.M * g.N * (double)g.t[t].K * g.batch;
    if (p.triu_tiles) flops *= (double)p.tiles_per_img / (p.tiles_m * p.tiles_n);
    const int dims[6] = {g.M, g.N, g.t[0].K, g.nterms > 1 ? g.t[1].K : 0, g.batch, npass};
    if (prof_level() == 1) prof_id = prof_begin(stream, flops, dims);
    else prof_group_note(flops, dims);
  }
  cudaError_t le;
  le = npass == 3 ? launch2<3>(p, stream) : launch2<1>(p, stream);
  if (prof_id >= 0) prof_end(prof_id, stream);
  if (le != cudaSuccess) return le;
  if (g.dot_out) {
    const int per_img = p.tiles_per_img * kEpiWarps * ctas;
    dot_reduce_kernel<<<(g.batch + 7) / 8, 256, 0, stream>>>(g.dot_ws, per_img, g.batch, g.dot_out);
    note_launch();
    le = cudaGetLastError();
  }
  return le;
}

}  // namespace egm
