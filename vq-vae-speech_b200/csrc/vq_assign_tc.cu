// Nearest-code search on the tensor cores (tcgen05) with an exact fp32 re-check -- the fast engine of vqs_vq_assign.
//
// The distance search d[n,k] = (|x_n|^2 + |e_k|^2) - 2 x_n.e_k is a dense N x D . D x K contraction.  On CUDA cores the
// K = 44 search costs 5632 FLOP per 256-byte row, twice the FMA time the HBM stream allows; here the dot products run as
// tcgen05.mma kind::tf32 into TMEM and the CUDA cores only scan K scores per row.  Indices must nevertheless equal the
// fp32 search bit for bit (torch.argmin semantics, lowest index on ties), so:
//   1. operands are split x = xh + xl, e = eh + el (h = top 19 bits) and D = xl.eh + xh.el + xh.eh (3xTF32): the score
//      s_k = |e_k|^2 - 2 x.e_k is accurate to ~2^-21 |x||e_k|;
//   2. each row keeps the best and second-best score; if their gap exceeds 2*tol (tol: a rigorous bound on the score
//      error) the best code is the fp32 argmin;
//   3. otherwise (rare: ~1e-4 of random rows) the row is re-scored EXACTLY by the whole warp with the same fp32 formula
//      and summation order as the CUDA-core kernel (vq_kernels.cu), so both engines return identical indices.
// Per-code statistics (counts, dw) are accumulated from the same shared-memory tile exactly as in the CUDA-core kernel
// (warp-owned codes, run-length register accumulation, deterministic two-stage reduction).
//
// Warp roles (416 threads): warps 0-3 epilogue (TMEM lane quarter = warp), warps 4-11 producers (global -> hi/lo split ->
// UMMA K-major SWIZZLE_128B tiles; the (B, D, T) layout is transposed on the fly; every load of a tile is issued before
// its first store and the next tile's loads fly during the empty-slot wait), warp 12 TMEM allocator + MMA issuer.
// Pipelines: full[stage] / empty[stage] over 2 row-tile stages, tmem_full[2] / tmem_empty[2] over 2 TMEM accumulators.
#include <math.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace vqs {

struct AssignTcParams {
  const float* z;
  const float* cb;
  int64_t* idx;
  float* partials;  // [grid][K*(D+1)]
  long long N;
  int layout, B, D, T, K;
  int Kpad;    // codes padded to a multiple of 16 (UMMA N)
  int nkb;     // D / 32
  int ntiles;
  int G;       // private copies of the statistics bins (row groups of the statistics pass)
  int debug;   // profiling aid (env VQS_TC_DEBUG): bit 0 skips the statistics pass, bit 1 the tolerance / re-check
  FastDiv divD;
};

namespace {

constexpr int TROWS = 128;
constexpr int XT_BYTES = TROWS * 128;     // one [128 rows x 32 floats] k-block tile
constexpr int STAGES = 2;
constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int MMA_WARP = 4 + PROD_WARPS;
constexpr int NTHREADS = (MMA_WARP + 1) * 32;

struct Shared {
  uint64_t full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], idx_ready[2];
  uint32_t tmem_base;
  float emax;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
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

// element (row r, column j) of a row tile / codebook tile stored as nkb k-block tiles of `rows` x 128 B, hi then lo copy
__device__ __forceinline__ uint32_t elem_off(int r, int j, int rows) {
  return (uint32_t)((j >> 5) * rows * 128) + sw128_off(r, j & 31);
}

// exact fp32 value of a split element
__device__ __forceinline__ float ld_exact(const uint8_t* hi, const uint8_t* lo, uint32_t off) {
  return *reinterpret_cast<const float*>(hi + off) + *reinterpret_cast<const float*>(lo + off);
}

// sum_j x^2 in the summation order of the CUDA-core kernel (8 partials over j = 4p + 32 s + e, then a butterfly tree)
__device__ __forceinline__ float row_sumsq(const uint8_t* xh, const uint8_t* xl, int r, int D) {
  float part[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float s = 0.f;
    for (int j = p * 4; j < D; j += 32) {
      const uint32_t off = elem_off(r, j, TROWS);
      float4 h = *reinterpret_cast<const float4*>(xh + off);
      float4 l = *reinterpret_cast<const float4*>(xl + off);
      float a = h.x + l.x, b = h.y + l.y, c = h.z + l.z, d = h.w + l.w;
      s = fmaf(a, a, s);
      s = fmaf(b, b, s);
      s = fmaf(c, c, s);
      s = fmaf(d, d, s);
    }
    part[p] = s;
  }
  // butterfly over the 8 "code lanes" of the CUDA-core kernel: xor 1, 2, 4
  float a0 = part[0] + part[1], a1 = part[2] + part[3], a2 = part[4] + part[5], a3 = part[6] + part[7];
  float b0 = a0 + a1, b1 = a2 + a3;
  return b0 + b1;
}

__global__ void __launch_bounds__(NTHREADS, 1) vq_assign_tc_kernel(const AssignTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = p.D, K = p.K, Kpad = p.Kpad, nkb = p.nkb;
  const int x_tile = nkb * XT_BYTES;            // one copy (hi or lo) of a row tile
  const int stage_bytes = 2 * x_tile;
  const int cb_tile = nkb * Kpad * 128;         // one copy of the codebook
  uint8_t* xs = smem;                            // [STAGES][hi | lo]
  uint8_t* cbh = smem + STAGES * stage_bytes;    // codebook hi, lo
  uint8_t* cbl = cbh + cb_tile;
  float* se = reinterpret_cast<float*>(cbl + cb_tile);   // [Kpad]
  int* sidx = reinterpret_cast<int*>(se + Kpad);         // [2][128]
  float* dws = reinterpret_cast<float*>(sidx + 2 * TROWS);  // [G][K*D] then cnt[G][K]
  float* cnts = dws + p.G * K * D;
  Shared* sh = reinterpret_cast<Shared*>(cnts + ((p.G * K + 3) & ~3));

  // ---- one-time setup: barriers, TMEM, codebook tiles (exact split), |e|^2 ----
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sh->full[s], PROD_WARPS);
      mbar_init(&sh->empty[s], 4 + PROD_WARPS);   // scan warps + statistics (= producer) warps
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sh->tmem_full[a], 1);
      mbar_init(&sh->tmem_empty[a], 4);
      mbar_init(&sh->idx_ready[a], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * Kpad) tmem_cols <<= 1;
  if (warp == MMA_WARP) tmem_alloc(&sh->tmem_base, tmem_cols);
  for (int e = tid; e < Kpad * D; e += NTHREADS) {
    uint32_t k, j;
    p.divD.divmod((uint32_t)e, k, j);
    const float v = ((int)k < K) ? __ldg(p.cb + (size_t)k * D + j) : 0.f;
    const float h = tf32_hi(v);
    const uint32_t off = elem_off((int)k, (int)j, Kpad);
    *reinterpret_cast<float*>(cbh + off) = h;
    *reinterpret_cast<float*>(cbl + off) = v - h;
  }
  for (int i = tid; i < p.G * K * (D + 1); i += NTHREADS) dws[i] = 0.f;
  __syncthreads();
  for (int k = tid; k < Kpad; k += NTHREADS) {
    float s = 0.f;
    if (k < K) {
      for (int j = 0; j < D; ++j) {
        float v = ld_exact(cbh, cbl, elem_off(k, j, Kpad));
        s = fmaf(v, v, s);
      }
    }
    se[k] = (k < K) ? s : INFINITY;
  }
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int k = 0; k < K; ++k) m = fmaxf(m, se[k]);
    sh->emax = sqrtf(m);
  }
  fence_proxy_async();   // codebook tiles were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const float emax = sh->emax;

  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 4 && warp < MMA_WARP) {
    // ================= producers =================
    const int ptid = tid - 128;
    const uint32_t xs_a = smem_u32(xs);
    const bool flat = p.layout == VQS_LAYOUT_FLAT_ND;
    constexpr int FC = TROWS * 16 / PROD_THREADS;   // flat: 16-byte chunks per thread per tile at D = 64 (8)
    const int cpr = D >> 2;                         // chunks per row
    const int nchunk = (TROWS * cpr + PROD_THREADS - 1) / PROD_THREADS;   // <= FC for D <= 64
    float4 cur[FC], nxt[FC];
    auto load_flat = [&](int it, float4 (&rg)[FC]) {
      const long long r0 = (long long)(blockIdx.x + it * gridDim.x) * TROWS;
      const long long left = p.N - r0;
      const int total = (left < TROWS ? (int)left : TROWS) * cpr;
      const float4* src = reinterpret_cast<const float4*>(p.z + r0 * D);
#pragma unroll
      for (int u = 0; u < FC; ++u) {
        const int c = ptid + u * PROD_THREADS;
        rg[u] = (u < nchunk && c < total) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // Per-code statistics of a finished tile, taken from its shared-memory stage by the SAME 8 warps.  Thread = (column j,
    // row group g): it walks its 128/G rows in ascending order and adds x[row][j] into its PRIVATE bin dws[g][idx[row]][j]
    // -- no atomics, no conflicts (a warp reads one row's 32 consecutive columns; all its lanes hit the same code), fixed
    // order -> deterministic.  (History: doing this in the 4 scan warps cost 20k cycles per tile; a warp-owned-codes
    // variant with generic loads 10k; profiles/r01h, vq_phase_probe.py.)
    const uint32_t dws_a = smem_u32(dws), cnts_a = smem_u32(cnts), sidx_a = smem_u32(sidx);
    const int sj = ptid % D, sg = ptid / D;              // column, row group
    const int rpg = TROWS / p.G;                         // rows per group
    const uint32_t xcol = (uint32_t)((sj >> 5) * XT_BYTES + ((sj & 3) << 2));   // k-block tile + byte within the chunk
    const int xchunk = (sj & 31) >> 2;
    auto tile_stats = [&](int jt) {
      const int s2 = jt % STAGES;
      const long long r02 = (long long)(blockIdx.x + jt * gridDim.x) * TROWS;
      const long long left2 = p.N - r02;
      const int rows2 = left2 < TROWS ? (int)left2 : TROWS;
      const uint32_t xh2 = xs_a + (uint32_t)(s2 * stage_bytes), xl2 = xh2 + (uint32_t)x_tile;
      const uint32_t si = sidx_a + (uint32_t)((jt & 1) * TROWS * 4);
      mbar_wait(&sh->idx_ready[jt & 1], (uint32_t)(jt >> 1) & 1u);
      if (sg < p.G && !(p.debug & 1)) {
        const int rbeg = sg * rpg;
        int rend = rbeg + rpg;
        if (rend > rows2) rend = rows2;
        const uint32_t bins = dws_a + (uint32_t)((sg * K * D + sj) * 4);
        const uint32_t cbin = cnts_a + (uint32_t)(sg * K * 4);
        // 4 rows per step: their index and value loads are issued together, the bin updates stay in row order
        for (int R0 = rbeg; R0 < rend; R0 += 4) {
          int k[4];
          float x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int R = R0 + u < rend ? R0 + u : rend - 1;
            k[u] = lds_s32(si + (uint32_t)(R * 4));
            const uint32_t xo = xcol + (uint32_t)(R * 128 + ((xchunk ^ (R & 7)) << 4));
            x[u] = lds_f32(xh2 + xo) + lds_f32(xl2 + xo);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (R0 + u < rend) {
              const uint32_t ba = bins + (uint32_t)(k[u] * D * 4);
              sts_f32(ba, lds_f32(ba) + x[u]);
              if (sj == 0) sts_f32(cbin + (uint32_t)(k[u] * 4), lds_f32(cbin + (uint32_t)(k[u] * 4)) + 1.f);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[s2]);   // this warp is done with the row tile
    };
    if (flat && my_tiles > 0) load_flat(0, cur);
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      const long long r0 = (long long)tile * TROWS;
      const long long left = p.N - r0;
      const int rows = left < TROWS ? (int)left : TROWS;
      if (flat && it + 1 < my_tiles) load_flat(it + 1, nxt);   // next tile's loads fly during the stats + wait + stores
      if (it >= STAGES) tile_stats(it - STAGES);               // the tile that last used this stage
      mbar_wait(&sh->empty[s], ph ^ 1u);
      const uint32_t xh = xs_a + (uint32_t)(s * stage_bytes);
      const uint32_t xl = xh + (uint32_t)x_tile;
      if (flat) {
        // 16-byte chunks: 8 consecutive threads cover one 128-byte k-block row (coalesced reads, conflict-free STS);
        // rows beyond N arrive as zeros from load_flat
#pragma unroll
        for (int u = 0; u < FC; ++u) {
          const int c = ptid + u * PROD_THREADS;
          if (u < nchunk && c < TROWS * cpr) {
            const uint32_t row = p.divD.div((uint32_t)c << 2);
            const int c16 = c - (int)row * cpr;
            const uint32_t off = (uint32_t)((c16 >> 3) * XT_BYTES + row * 128 + (((c16 & 7) ^ (row & 7)) << 4));
            const float4 v = cur[u];
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            sts_v4(xh + off, h);
            sts_v4(xl + off, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
          }
        }
#pragma unroll
        for (int u = 0; u < FC; ++u) cur[u] = nxt[u];
      } else {
        if (rows < TROWS) {   // zero the tail rows
          for (int e = ptid; e < (TROWS - rows) * D; e += PROD_THREADS) {
            uint32_t r, j;
            p.divD.divmod((uint32_t)e, r, j);
            const uint32_t off = elem_off(rows + (int)r, (int)j, TROWS);
            sts_f32(xh + off, 0.f);
            sts_f32(xl + off, 0.f);
          }
        }
        // (B, D, T) read in the reference's (D, T, B) row order: flat f = d*P + t*B + b  <-  z[(b*D + d)*T + t];
        // loads are issued 8 at a time before the first store
        const int B = p.B, T = p.T;
        const long long P = (long long)T * B;
        const long long f0 = r0 * D;
        const long long f1 = f0 + (long long)rows * D;
        const long long d_first = f0 / P, d_last = (f1 - 1) / P;
        for (long long d = d_first; d <= d_last; ++d) {
          const long long base = d * P;
          const int p0 = (int)((f0 > base ? f0 : base) - base);
          const int p1 = (int)((f1 < base + P ? f1 : base + P) - base);
          const int t0 = p0 / B, t1 = (p1 + B - 1) / B;
          const int span = t1 - t0, count = span * B;
          const int qs = PROD_THREADS / span, rs = PROD_THREADS - qs * span;
          int b = ptid / span, tt = ptid - b * span;
          const int foff = (int)(base - f0);
          const float* zd = p.z + (size_t)d * T;
          for (int e = ptid; e < count; e += PROD_THREADS * 8) {
            float v[8];
            uint32_t off[8];
            bool ok[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int t = t0 + tt;
              const int pp = t * B + b;
              ok[u] = (e + u * PROD_THREADS < count) && pp >= p0 && pp < p1;
              v[u] = ok[u] ? __ldg(zd + (size_t)b * D * T + t) : 0.f;
              uint32_t row, j;
              p.divD.divmod((uint32_t)(foff + pp), row, j);
              off[u] = elem_off((int)row, (int)j, TROWS);
              tt += rs;
              b += qs;
              if (tt >= span) {
                tt -= span;
                ++b;
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (ok[u]) {
                const float h = tf32_hi(v[u]);
                sts_f32(xh + off[u], h);
                sts_f32(xl + off[u], v[u] - h);
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full[s]);
    }
    for (int jt = (my_tiles > STAGES ? my_tiles - STAGES : 0); jt < my_tiles; ++jt) tile_stats(jt);
    named_bar_sync(2, PROD_THREADS);
    // fold the G private copies in a fixed order and publish this CTA's partial
    float* out = p.partials + (size_t)blockIdx.x * K * (D + 1);
    for (int i = ptid; i < K; i += PROD_THREADS) {
      float a = 0.f;
      for (int g = 0; g < p.G; ++g) a += cnts[g * K + i];
      out[i] = a;
    }
    for (int i = ptid; i < K * D; i += PROD_THREADS) {
      float a = 0.f;
      for (int g = 0; g < p.G; ++g) a += dws[g * K * D + i];
      out[K + i] = a;
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(Kpad);
      const uint32_t cbh_a = smem_u32(cbh), cbl_a = smem_u32(cbl);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        const int a = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&sh->full[s], ph);
        mbar_wait(&sh->tmem_empty[a], aph ^ 1u);
        tc_fence_after();
        const uint32_t xh_a = smem_u32(xs + s * stage_bytes), xl_a = xh_a + (uint32_t)x_tile;
        const uint32_t dst = tmem_base + (uint32_t)(a * Kpad);
        uint32_t acc = 0u;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint64_t ah = make_desc_sw128(xh_a + kb * XT_BYTES), al = make_desc_sw128(xl_a + kb * XT_BYTES);
          const uint64_t bh = make_desc_sw128(cbh_a + kb * Kpad * 128), bl = make_desc_sw128(cbl_a + kb * Kpad * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            umma_tf32(dst, al + adv, bh + adv, idesc, acc);
            umma_tf32(dst, ah + adv, bl + adv, idesc, 1u);
            umma_tf32(dst, ah + adv, bh + adv, idesc, 1u);
            acc = 1u;
          }
        }
        umma_commit(&sh->tmem_full[a]);
      }
    }
    __syncwarp();
  } else {
    // ================= scan warps: scores -> indices (+ exact re-check of near-ties) =================
    const int r = warp * 32 + lane;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = it % STAGES;
      const int a = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const long long r0 = (long long)tile * TROWS;
      const long long left = p.N - r0;
      const int rows = left < TROWS ? (int)left : TROWS;
      const uint8_t* xh = xs + s * stage_bytes;
      const uint8_t* xl = xh + x_tile;
      mbar_wait(&sh->tmem_full[a], aph);
      tc_fence_after();
      // ---- approximate scores s_k = |e_k|^2 - 2 x.e_k: best / second best, two interleaved dependency chains ----
      float b0 = INFINITY, s0 = INFINITY, b1 = INFINITY, s1 = INFINITY;
      int k0 = 0, k1 = 0;
      const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * Kpad);
      for (int c0 = 0; c0 < Kpad; c0 += 16) {
        float v[16];
        tmem_ld16(ta + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float sa = fmaf(-2.f, v[j], se[c0 + j]);
          const float sb = fmaf(-2.f, v[j + 1], se[c0 + j + 1]);
          if (sa < b0) { s0 = b0; b0 = sa; k0 = c0 + j; } else if (sa < s0) { s0 = sa; }
          if (sb < b1) { s1 = b1; b1 = sb; k1 = c0 + j + 1; } else if (sb < s1) { s1 = sb; }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty[a]);   // the accumulator can be overwritten
      // merge the chains: lowest score, lowest index on equal scores
      const bool take1 = (b1 < b0) || (b1 == b0 && k1 < k0);
      const float best = take1 ? b1 : b0;
      const int bkm = take1 ? k1 : k0;
      const float second = fminf(fminf(s0, s1), take1 ? b0 : b1);
      int bk = bkm;
      // |x|^2 only feeds the tolerance here, so the hi copy suffices (2^-10 relative); lanes walk the 16-byte chunks of
      // their row in a rotated order so that every quarter-warp touches 8 distinct bank groups
      float sx = 0.f;
      for (int kb = 0; kb < ((p.debug & 2) ? 0 : nkb); ++kb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pc = (r + i) & 7;   // physical chunk position
          const float4 h = *reinterpret_cast<const float4*>(xh + kb * XT_BYTES + r * 128 + (pc << 4));
          sx = fmaf(h.x, h.x, sx);
          sx = fmaf(h.y, h.y, sx);
          sx = fmaf(h.z, h.z, sx);
          sx = fmaf(h.w, h.w, sx);
        }
      }
      sx *= 1.01f;
      // ---- bound on |score - fp32 distance|:  3xTF32 operand error 2 * 3 * 2^-20 |x||e| + truncating accumulation over
      // 24 MMAs 2 * 24 * 2^-23 |x||e|  (together < 1.2e-5 |x||e|), plus the rounding of the fp32 formula itself
      // (~66 * 2^-24 (|x|^2 + |e|^2)); both with margin.  Scores closer than 2*tol are re-checked exactly. ----
      const float tol = sqrtf(sx) * emax * 1.5e-5f + (sx + emax * emax) * 6e-6f;
      const bool flagged = (p.debug & 2) ? false : !((second - best) > 2.f * tol);
      // ---- exact re-check by the whole warp, same fp32 formula and order as the CUDA-core kernel ----
      unsigned m = __ballot_sync(0xffffffffu, flagged);
      while (m) {
        const int rr = __ffs(m) - 1;
        m &= m - 1;
        const int R = warp * 32 + rr;
        const float sxr = row_sumsq(xh, xl, R, D);   // the canonical fp32 sum (all lanes compute the same value)
        float bd = INFINITY;
        int bkk = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
          float dot = 0.f;
          for (int j = 0; j < D; ++j)
            dot = fmaf(ld_exact(xh, xl, elem_off(R, j, TROWS)), ld_exact(cbh, cbl, elem_off(k, j, Kpad)), dot);
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
      sidx[(it & 1) * TROWS + r] = bk;
      if (r < rows) p.idx[r0 + r] = (int64_t)bk;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sh->idx_ready[it & 1]);   // release: this warp's 32 indices are in shared memory
        mbar_arrive(&sh->empty[s]);            // and it no longer reads the row tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

static size_t smem_bytes_g(int K, int D, int G) {
  const int Kpad = (K + 15) / 16 * 16;
  const int nkb = D / 32;
  size_t x = (size_t)STAGES * 2 * nkb * XT_BYTES;
  size_t cb = (size_t)2 * nkb * Kpad * 128;
  size_t rest = (size_t)Kpad * 4 + 2 * TROWS * 4 + (size_t)G * K * D * 4 + (size_t)((G * K + 3) & ~3) * 4 + sizeof(Shared);
  return x + cb + rest + 1024 + 64;
}
// private statistics copies: as many row groups as the 256 statistics threads cover (256 / D), fewer if smem is short
static int stats_groups(int K, int D) {
  int G = PROD_THREADS / D;
  while (G > 1 && smem_bytes_g(K, D, G) > 220 * 1024) G >>= 1;
  return G;
}
size_t assign_tc_smem_bytes(int K, int D) { return smem_bytes_g(K, D, stats_groups(K, D)); }

bool assign_tc_supported(int K, int D) {
  if (D % 32 != 0 || D > 64 || K < 1 || K > 256) return false;
  return assign_tc_smem_bytes(K, D) <= 220 * 1024;
}

// Launches the tensor-core search; `partials` must hold grid * K*(D+1) floats; returns the grid size through *grid_out.
int launch_assign_tc(const float* z, int layout, int B, int D, int T, const float* cb, int K, int64_t* idx,
                     float* partials, int max_grid, int* grid_out, cudaStream_t st) {
  AssignTcParams p;
  p.z = z; p.cb = cb; p.idx = idx; p.partials = partials;
  p.N = (long long)B * T;
  p.layout = layout; p.B = B; p.D = D; p.T = T; p.K = K;
  p.Kpad = (K + 15) / 16 * 16;
  p.nkb = D / 32;
  p.ntiles = (int)((p.N + TROWS - 1) / TROWS);
  p.divD = FastDiv((uint32_t)D);
  p.G = stats_groups(K, D);
  {
    const char* dbg = getenv("VQS_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  const size_t smem = assign_tc_smem_bytes(K, D);
  static size_t configured = 0;
  if (smem > configured) {
    VQS_CUDA(cudaFuncSetAttribute(vq_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int grid = num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid > max_grid) grid = max_grid;
  *grid_out = grid;
  vq_assign_tc_kernel<<<grid, NTHREADS, smem, st>>>(p);
  VQS_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqs
