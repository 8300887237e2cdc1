// Nearest-code search for resident codebooks, flat (N, 64) rows: the streaming engine of vqs_vq_assign.
//
// Replaces /root/reference/src/models/vector_quantizer_ema.py:109-119,143-150 (distances, argmin, one-hot statistics)
// for the row stream.  The row tiles are never touched by a thread on their way in:
//   * a TMA tensor map delivers 128-row x 64-column fp32 tiles straight into the UMMA K-major SWIZZLE_128B layout
//     (cp.async.bulk.tensor, 4-stage mbarrier ring, one elected thread);
//   * the RAW fp32 words are the tensor-core operands: tcgen05.mma kind::tf32 reads the upper 19 bits, so ONE pass of 8
//     MMAs (128 x Kpad x 8) gives every score s_k = |e_k|^2 - 2 x.e_k to ~2e-3 (|x|^2 + |e_k|^2);
//   * scan warps keep the three smallest lower bounds per row (code index packed into the low mantissa bits: pure
//     min/max, no compares).  A row whose runner-up is further than twice the error bound is settled.  Otherwise (~5 % of
//     random rows) its two candidates are re-scored in fp32 by the whole warp; only if those land within 4e-5 of each
//     other (or a third candidate is close) the row is settled with the CUDA-core kernel's own formula and summation
//     order -- so the indices are IDENTICAL to the fp32 search (torch.argmin semantics, lowest index on ties);
//   * per-code statistics (counts, dw = encodings^T x) without atomics: the tile's rows are counting-sorted by code
//     (ranks from match.any in the scan warps, offsets from one warp scan), eight warps sum runs of equal codes in
//     registers from the same shared-memory tile and add each run to the CTA's private bins in a fixed order
//     (run continuations across warp ranges go through a carry slot) -> deterministic.
// Warp roles (576 threads, one persistent CTA per SM): warps 0-7 scan (two groups alternating tiles, TMEM lane quarter =
// warp % 4), warps 8-15 statistics, warp 16 TMA producer, warp 17 TMEM allocator + MMA issuer.
#include <cuda.h>
#include <math.h>

#include "tc_common.cuh"

namespace vqs {

struct AssignTmaParams {
  const float* cb;
  int64_t* idx;
  float* partials;  // [grid][K*65]
  long long N;
  int K, Kpad, ntiles;
};

namespace {

constexpr int TR = 128;                  // rows per tile
constexpr int XT = TR * 128;             // bytes of one k-block image (32 columns)
constexpr int TILE_BYTES = 2 * XT;       // 32 KB
constexpr int NSTAGE = 4;
constexpr int SCAN_WARPS = 8, STAT_WARPS = 8;
constexpr int TMA_WARP = SCAN_WARPS + STAT_WARPS, MMA_WARP = TMA_WARP + 1;
constexpr int NT = (MMA_WARP + 1) * 32;
constexpr int KMAX = 64;
// Tensor-core filter.  kind::tf32 drops the low 13 mantissa bits of both operands: |x_j e_j - tf32(x_j) tf32(e_j)| <
// (2 * 2^-10 + 2^-20) |x_j e_j|, so the score error is < 2 * 1.955e-3 |x||e_k| <= 1.955e-3 (|x|^2 + |e_k|^2); packing the
// code index into 6 mantissa bits adds < 1.6e-5, the fp32 rounding of the formula itself < 1e-5 of the same scale.
constexpr float EPS1 = 2.2e-3f;
// fp32 filter: two fp32 evaluations of (|x|^2 + |e|^2) - 2 x.e in ANY summation order differ by < 2e-5 (|x|^2 + |e|^2)
constexpr float EPS2 = 2.0e-5f;
constexpr float BIG = 3.0e38f;

struct Sh {
  uint64_t full[NSTAGE], empty[NSTAGE], tmem_full[2], tmem_empty[2], idx_ready[2], idx_free[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// byte offset of element (row r, column j) of a 64-column tile stored as two k-block images of `rows` x 128 B
__device__ __forceinline__ uint32_t elem_off64(int r, int j, int rows) {
  return (uint32_t)((j >> 5) * rows * 128) + sw128_off(r, j & 31);
}

// sum_j x^2 in the summation order of the CUDA-core kernel (vq_kernels.cu: 8 partials over j = 4p + 32s + e, then a tree)
__device__ __forceinline__ float row_sumsq_canon(const uint8_t* xt, int r) {
  float part[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float s = 0.f;
#pragma unroll
    for (int j = p * 4; j < 64; j += 32) {
      const float4 v = *reinterpret_cast<const float4*>(xt + elem_off64(r, j, TR));
      s = fmaf(v.x, v.x, s);
      s = fmaf(v.y, v.y, s);
      s = fmaf(v.z, v.z, s);
      s = fmaf(v.w, v.w, s);
    }
    part[p] = s;
  }
  const float a0 = part[0] + part[1], a1 = part[2] + part[3], a2 = part[4] + part[5], a3 = part[6] + part[7];
  const float b0 = a0 + a1, b1 = a2 + a3;
  return b0 + b1;
}

__global__ void __launch_bounds__(NT, 1) vq_assign_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const AssignTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, Kpad = p.Kpad;
  uint8_t* xs = smem;                                               // [NSTAGE][2 k-blocks][128 rows][128 B], raw fp32
  uint8_t* cbs = xs + NSTAGE * TILE_BYTES;                          // [2 k-blocks][Kpad codes][128 B], raw fp32
  float* dw_s = reinterpret_cast<float*>(cbs + 2 * Kpad * 128);     // [Kpad][64] this CTA's dw bins
  float* se = dw_s + Kpad * 64;                                     // [64]  |e_k|^2
  float* sea = se + KMAX;                                           // [64]  |e_k|^2 (1 - EPS1), BIG beyond K
  int* cnt_s = reinterpret_cast<int*>(sea + KMAX);                  // [64]  this CTA's counts
  int* cw = cnt_s + KMAX;                                           // [2][4][64] rows per (tile parity, scan warp, code)
  int* qoff = cw + 2 * 4 * KMAX;                                    // [4][64] first sorted position of (scan warp, code)
  int* start = qoff + 4 * KMAX;                                     // [64] first sorted position of a code
  int* sidx = start + KMAX;                                         // [2][128] code | rank << 8 (-1: no row)
  int* order = sidx + 2 * TR;                                       // [128] sorted: row * 128 | code << 16
  float* carry = reinterpret_cast<float*>(order + TR);              // [8][64] run continuations
  int* carryk = reinterpret_cast<int*>(carry + STAT_WARPS * 64);    // [8]
  Sh* sh = reinterpret_cast<Sh*>(carryk + STAT_WARPS);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sh->full[s], 1);
      mbar_init(&sh->empty[s], 1 + 4 + STAT_WARPS);   // score MMAs (commit) + the tile's scan group + statistics warps
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full[a], 1);
      mbar_init(&sh->tmem_empty[a], 4);
      mbar_init(&sh->idx_ready[a], 4);
      mbar_init(&sh->idx_free[a], STAT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * Kpad) tmem_cols <<= 1;
  if (warp == MMA_WARP) tmem_alloc(&sh->tmem_base, tmem_cols);
  for (int e = tid; e < Kpad * 64; e += NT) {
    const int k = e >> 6, j = e & 63;
    const float v = (k < K) ? __ldg(p.cb + (size_t)k * 64 + j) : 0.f;
    *reinterpret_cast<float*>(cbs + elem_off64(k, j, Kpad)) = v;
    dw_s[e] = 0.f;
  }
  for (int k = tid; k < KMAX; k += NT) {
    float s = 0.f;
    if (k < K)
      for (int j = 0; j < 64; ++j) {
        const float v = __ldg(p.cb + (size_t)k * 64 + j);
        s = fmaf(v, v, s);
      }
    se[k] = (k < K) ? s : BIG;
    sea[k] = (k < K) ? s * (1.f - EPS1) : BIG;
    cnt_s[k] = 0;
  }
  for (int i = tid; i < 2 * 4 * KMAX; i += NT) cw[i] = 0;
  if (tid < STAT_WARPS) carryk[tid] = -1;
  fence_proxy_async();   // the codebook image was written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == TMA_WARP) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % NSTAGE;
        const long long r0 = (long long)(blockIdx.x + it * gridDim.x) * TR;
        mbar_wait(&sh->empty[s], ((uint32_t)(it / NSTAGE) & 1u) ^ 1u);
        mbar_expect_tx(&sh->full[s], TILE_BYTES);
        const uint32_t dst = smem_u32(xs + s * TILE_BYTES);
        tma_load_2d(dst, &tmap, 0, (int)r0, &sh->full[s]);         // rows beyond N arrive as zeros
        tma_load_2d(dst + XT, &tmap, 32, (int)r0, &sh->full[s]);
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer: 8 single-pass tf32 MMAs per tile on the raw fp32 words =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(Kpad);
      const uint32_t cb_a = smem_u32(cbs);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % NSTAGE, a = it & 1;
        mbar_wait(&sh->full[s], (uint32_t)(it / NSTAGE) & 1u);
        mbar_wait(&sh->tmem_empty[a], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t x_a = smem_u32(xs + s * TILE_BYTES);
        const uint32_t dst = tmem_base + (uint32_t)(a * Kpad);
        uint32_t acc = 0u;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t ad = make_desc_sw128(x_a + kb * XT), bd = make_desc_sw128(cb_a + kb * Kpad * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_tf32(dst, ad + (uint64_t)((k * 32) >> 4), bd + (uint64_t)((k * 32) >> 4), idesc, acc);
            acc = 1u;
          }
        }
        umma_commit(&sh->tmem_full[a]);
        umma_commit(&sh->empty[s]);
      }
    }
    __syncwarp();
  } else if (warp < SCAN_WARPS) {
    // ================= scan warps: scores -> index, rank of the row inside its code =================
    const int g = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    for (int it = g; it < my_tiles; it += 2) {
      const int s = it % NSTAGE;
      const uint32_t ph2 = (uint32_t)(it >> 1) & 1u;
      const long long r0 = (long long)(blockIdx.x + it * gridDim.x) * TR;
      const long long left = p.N - r0;
      const int rows = left < TR ? (int)left : TR;
      const uint8_t* xt = xs + s * TILE_BYTES;
      mbar_wait(&sh->full[s], (uint32_t)(it / NSTAGE) & 1u);   // acquire the TMA writes for this thread's own reads
      mbar_wait(&sh->tmem_full[g], ph2);
      tc_fence_after();
      // ---- lower bounds adj_k = |e_k|^2 (1 - EPS1) - 2 x.e_k with k in the low 6 mantissa bits: three smallest ----
      float b = INFINITY, s2 = INFINITY, t3 = INFINITY;
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * Kpad);
      for (int c0 = 0; c0 < Kpad; c0 += 16) {
        float v[16];
        tmem_ld16(ta + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float adj = fmaf(-2.f, v[j], sea[c0 + j]);
          const float key = __uint_as_float((__float_as_uint(adj) & 0xFFFFFFC0u) | (uint32_t)(c0 + j));
          const float hi1 = fmaxf(b, key);
          b = fminf(b, key);
          const float hi2 = fmaxf(s2, hi1);
          s2 = fminf(s2, hi1);
          t3 = fminf(t3, hi2);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty[g]);   // the accumulator can be overwritten
      const int k1 = (int)(__float_as_uint(b) & 63u);
      const int k2 = (int)(__float_as_uint(s2) & 63u);
      int bk = k1;
      // |x|^2 (any order: it only scales the bounds); lanes walk the 16-byte chunks of their row in a rotated order so
      // that every quarter-warp touches 8 distinct bank groups
      float sx = 0.f;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pc = (r + i) & 7;
          const float4 h = *reinterpret_cast<const float4*>(xt + kb * XT + r * 128 + (pc << 4));
          sx = fmaf(h.x, h.x, sx);
          sx = fmaf(h.y, h.y, sx);
          sx = fmaf(h.z, h.z, sx);
          sx = fmaf(h.w, h.w, sx);
        }
      }
      sx *= 1.0001f;
      const float tol2 = 2.f * EPS1 * (sx + se[k1]);
      const bool close2 = !((s2 - b) > tol2);
      const bool close3 = !((t3 - b) > tol2);
      // ---- exactly two candidates: fp32 re-score by the whole warp (lane = column pair) ----
      unsigned m2 = __ballot_sync(0xffffffffu, close2 && !close3);
      unsigned canon = __ballot_sync(0xffffffffu, close2 && close3);
      while (m2) {
        const int rr = __ffs(m2) - 1;
        m2 &= m2 - 1;
        const int R = q * 32 + rr;
        const int c1 = __shfl_sync(0xffffffffu, k1, rr), c2 = __shfl_sync(0xffffffffu, k2, rr);
        const float sxr = __shfl_sync(0xffffffffu, sx, rr);
        const float2 xv = *reinterpret_cast<const float2*>(xt + elem_off64(R, 2 * lane, TR));
        const float2 e1 = *reinterpret_cast<const float2*>(cbs + elem_off64(c1, 2 * lane, Kpad));
        const float2 e2 = *reinterpret_cast<const float2*>(cbs + elem_off64(c2, 2 * lane, Kpad));
        float d1 = fmaf(xv.y, e1.y, xv.x * e1.x), d2 = fmaf(xv.y, e2.y, xv.x * e2.x);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          d1 += __shfl_xor_sync(0xffffffffu, d1, o);
          d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        const float se1 = se[c1], se2 = se[c2];
        const float dd1 = (sxr + se1) - 2.f * d1, dd2 = (sxr + se2) - 2.f * d2;
        const float diff = dd1 - dd2;
        if (fabsf(diff) > 2.f * EPS2 * (sxr + fmaxf(se1, se2))) {
          if (lane == rr) bk = diff < 0.f ? c1 : c2;
        } else {
          canon |= 1u << rr;   // too close for an order-independent decision
        }
      }
      // ---- canonical settlement: same fp32 formula and summation order as the CUDA-core kernel, lanes = codes ----
      while (canon) {
        const int rr = __ffs(canon) - 1;
        canon &= canon - 1;
        const int R = q * 32 + rr;
        const float sxr = row_sumsq_canon(xt, R);
        float bd = INFINITY;
        int bkk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
          float dot = 0.f;
          for (int j = 0; j < 64; ++j)
            dot = fmaf(*reinterpret_cast<const float*>(xt + elem_off64(R, j, TR)),
                       *reinterpret_cast<const float*>(cbs + elem_off64(k, j, Kpad)), dot);
          const float dd = __fsub_rn(__fadd_rn(sxr, se[k]), __fmul_rn(2.0f, dot));
          if (dd < bd) {
            bd = dd;
            bkk = k;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, bd, o);
          const int ok = __shfl_xor_sync(0xffffffffu, bkk, o);
          if (od < bd || (od == bd && ok < bkk)) {
            bd = od;
            bkk = ok;
          }
        }
        if (lane == rr) bk = bkk < K ? bkk : K - 1;
      }
      const bool valid = r < rows;
      if (valid) p.idx[r0 + r] = (int64_t)bk;
      // ---- rank of this row among the rows of its scan warp with the same code (stable counting sort, part 1) ----
      const int kk = valid ? bk : 255;
      const unsigned same = __match_any_sync(0xffffffffu, kk);
      const int rank = __popc(same & ((1u << lane) - 1u));
      mbar_wait(&sh->idx_free[g], ph2 ^ 1u);            // the statistics warps are done with tile it - 2
      if (valid && rank == 0) cw[(g * 4 + q) * KMAX + bk] = __popc(same);
      sidx[g * TR + r] = valid ? (bk | (rank << 8)) : -1;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sh->idx_ready[g]);
        mbar_arrive(&sh->empty[s]);                     // this warp no longer reads the row tile
      }
    }
  } else {
    // ================= statistics warps: counting sort by code, run sums in registers, ordered bin updates =================
    const int sw = warp - SCAN_WARPS, st = tid - SCAN_WARPS * 32;
    const uint32_t dw_a = smem_u32(dw_s), carry_a = smem_u32(carry), order_a = smem_u32(order);
    // this lane's column pair (2 lane, 2 lane + 1): k-block image, 16-byte chunk, byte inside the chunk
    const uint32_t lane_base = (uint32_t)((lane >> 4) * XT + ((lane & 1) << 3));
    const uint32_t lane_cx = (uint32_t)(((lane & 15) >> 1) << 4);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NSTAGE, a = it & 1;
      const long long r0 = (long long)(blockIdx.x + it * gridDim.x) * TR;
      const long long left = p.N - r0;
      const int rows = left < TR ? (int)left : TR;
      mbar_wait(&sh->full[s], (uint32_t)(it / NSTAGE) & 1u);
      mbar_wait(&sh->idx_ready[a], (uint32_t)(it >> 1) & 1u);
      const int* cwa = cw + a * 4 * KMAX;
      if (sw == 0) {
        // offsets: lane owns codes lane and lane + 32
        int c0[4], c1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          c0[q] = cwa[q * KMAX + lane];
          c1[q] = cwa[q * KMAX + 32 + lane];
        }
        const int t0 = c0[0] + c0[1] + c0[2] + c0[3], t1 = c1[0] + c1[1] + c1[2] + c1[3];
        int i0 = t0, i1 = t1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u0 = __shfl_up_sync(0xffffffffu, i0, o), u1 = __shfl_up_sync(0xffffffffu, i1, o);
          if (lane >= o) {
            i0 += u0;
            i1 += u1;
          }
        }
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31);
        int o0 = i0 - t0, o1 = tot0 + i1 - t1;
        start[lane] = o0;
        start[32 + lane] = o1;
        cnt_s[lane] += t0;
        cnt_s[32 + lane] += t1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          qoff[q * KMAX + lane] = o0;
          qoff[q * KMAX + 32 + lane] = o1;
          o0 += c0[q];
          o1 += c1[q];
        }
      }
      bar_sync(1, STAT_WARPS * 32);
      if (st < TR) {
        const int v = sidx[a * TR + st];
        if (v >= 0) {
          const int k = v & 255, rank = v >> 8;
          order[qoff[(st >> 5) * KMAX + k] + rank] = (st << 7) | (k << 16);
        }
      } else {
        cw[a * 4 * KMAX + (st - TR)] = 0;             // the next tile of this parity starts from zero counts
        cw[a * 4 * KMAX + st] = 0;
      }
      bar_sync(1, STAT_WARPS * 32);
      // ---- this warp's 16 sorted positions: all loads first, then the run sums ----
      const int p0 = sw * 16;
      const int n = rows - p0 < 16 ? rows - p0 : 16;
      if (lane == 0) carryk[sw] = -1;
      if (n > 0) {
        uint32_t rk[16];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const uint4 o4 = lds_u4(order_a + (uint32_t)((p0 + i4 * 4) * 4));
          rk[i4 * 4 + 0] = o4.x;
          rk[i4 * 4 + 1] = o4.y;
          rk[i4 * 4 + 2] = o4.z;
          rk[i4 * 4 + 3] = o4.w;
        }
        const uint32_t x_a = smem_u32(xs + s * TILE_BYTES) + lane_base;
        float2 v[16];
        uint32_t lastrk = rk[0];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i < n) lastrk = rk[i];
          else rk[i] = lastrk;                         // padding: same code, value zero
          const uint32_t r128 = rk[i] & 0xFFFFu;
          v[i] = lds_v2(x_a + r128 + (((r128 >> 3) & 0x70u) ^ lane_cx));
          if (i >= n) v[i] = make_float2(0.f, 0.f);
        }
        int kprev = (int)(rk[0] >> 16);
        const bool cont = start[kprev] < p0;           // the first run continues a code that began in an earlier warp's range
        bool first = true;
        float2 acc = v[0];
#pragma unroll
        for (int i = 1; i <= 16; ++i) {
          const int k = (i < 16) ? (int)(rk[i] >> 16) : -1;
          if (k != kprev) {                             // warp-uniform
            if (first && cont) {
              sts_v2(carry_a + (uint32_t)((sw * 64 + 2 * lane) * 4), acc);
              if (lane == 0) carryk[sw] = kprev;
            } else {
              const uint32_t ba = dw_a + (uint32_t)((kprev * 64 + 2 * lane) * 4);
              float2 bin = lds_v2(ba);
              bin.x += acc.x;
              bin.y += acc.y;
              sts_v2(ba, bin);
            }
            first = false;
            kprev = k;
            acc = make_float2(0.f, 0.f);
          }
          if (i < 16) {
            acc.x += v[i].x;
            acc.y += v[i].y;
          }
        }
      }
      bar_sync(1, STAT_WARPS * 32);
      if (lane == 0) {
        mbar_arrive(&sh->empty[s]);
        mbar_arrive(&sh->idx_free[a]);
      }
      if (sw == 0) {
        // run continuations, in warp order
        for (int w = 1; w < STAT_WARPS; ++w) {
          const int kc = carryk[w];
          if (kc >= 0) {
            const float2 c = lds_v2(carry_a + (uint32_t)((w * 64 + 2 * lane) * 4));
            const uint32_t ba = dw_a + (uint32_t)((kc * 64 + 2 * lane) * 4);
            float2 bin = lds_v2(ba);
            bin.x += c.x;
            bin.y += c.y;
            sts_v2(ba, bin);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  // ---- publish this CTA's partial statistics ----
  {
    float* out = p.partials + (size_t)blockIdx.x * K * 65;
    for (int i = tid; i < K; i += NT) out[i] = (float)cnt_s[i];
    for (int i = tid; i < K * 64; i += NT) out[K + i] = dw_s[i];
  }
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

size_t smem_bytes_tma(int Kpad) {
  size_t b = (size_t)NSTAGE * TILE_BYTES + (size_t)2 * Kpad * 128 + (size_t)Kpad * 64 * 4;
  b += (size_t)(2 * KMAX) * 4 + (size_t)KMAX * 4 + (size_t)(2 * 4 * KMAX) * 4 + (size_t)(4 * KMAX) * 4 + (size_t)KMAX * 4;
  b += (size_t)(2 * TR) * 4 + (size_t)TR * 4 + (size_t)STAT_WARPS * 64 * 4 + (size_t)STAT_WARPS * 4 + sizeof(Sh);
  return b + 1024 + 64;
}

}  // namespace

// flat (N, 64) rows, codebook resident in shared memory, 16-byte aligned z
bool assign_tma_supported(int layout, int K, int D, long long N) {
  return layout == VQS_LAYOUT_FLAT_ND && D == 64 && K >= 1 && K <= KMAX && N >= 1 && N < (1ll << 31) - TR;
}

// `partials` must hold grid * K*(D+1) floats; the grid size comes back through *grid_out.
int launch_assign_tma(const float* z, long long N, const float* cb, int K, int64_t* idx, float* partials, int max_grid,
                      int* grid_out, cudaStream_t st) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_error("vq_assign: cuTensorMapEncodeTiled is not available from the driver");
    return VQS_ERR_ARG;
  }
  CUtensorMap map;
  const cuuint64_t dims[2] = {64, (cuuint64_t)N};
  const cuuint64_t strides[1] = {64 * sizeof(float)};
  const cuuint32_t box[2] = {32, TR};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(z), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("vq_assign: cuTensorMapEncodeTiled failed (%d) for N=%lld", (int)cr, N);
    return VQS_ERR_ARG;
  }
  AssignTmaParams p;
  p.cb = cb; p.idx = idx; p.partials = partials; p.N = N; p.K = K;
  p.Kpad = (K + 15) / 16 * 16;
  p.ntiles = (int)((N + TR - 1) / TR);
  const size_t smem = smem_bytes_tma(p.Kpad);
  static size_t configured = 0;
  if (smem > configured) {
    VQS_CUDA(cudaFuncSetAttribute(vq_assign_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid > max_grid) grid = max_grid;
  *grid_out = grid;
  vq_assign_tma_kernel<<<grid, NT, smem, st>>>(map, p);
  VQS_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqs
