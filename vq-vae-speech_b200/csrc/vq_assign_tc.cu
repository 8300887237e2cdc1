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
// Score error bound.  With tau_k = EPS_SCORE * (|x|^2 + |e_k|^2) the tensor-core score s_k = |e_k|^2 - 2 x.e_k satisfies
// |s_k + |x|^2 - d_k| <= tau_k, d_k being the fp32 distance of the exact path: 3xTF32 operand error 2 * 3 * 2^-20 |x||e_k|
// plus truncating accumulation over 24 MMAs 2 * 24 * 2^-23 |x||e_k| (together < 1.2e-5 |x||e_k| <= 0.6e-5 (|x|^2+|e_k|^2))
// plus the rounding of the fp32 formula itself (~66 * 2^-24 (|x|^2 + |e_k|^2) < 0.4e-5 ...), with margin.  The scan ranks
// the LOWER bounds adj_k = s_k - EPS_SCORE |e_k|^2; the best code b is certainly the fp32 argmin when
// adj_second - adj_b > 2 EPS_SCORE (|x|^2 + |e_b|^2).  The bound is per code, so a few huge (e.g. never-used EMA) codes
// do not inflate it for everybody.
constexpr float EPS_SCORE = 1.6e-5f;
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
    // ================= MMA issuer (whole converged warp, MMAs guarded by elect_one(): see gemm_tc.cu::issue_mmas) ==========
    {
      const uint32_t idesc = make_idesc_tf32(Kpad);
      const bool elected = elect_one();
      const uint64_t cbh_d = make_desc_sw128(smem_u32(cbh)), cbl_d = make_desc_sw128(smem_u32(cbl));
      const uint64_t xs_d = make_desc_sw128(smem_u32(xs));
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        const int a = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&sh->full[s], ph);
        mbar_wait(&sh->tmem_empty[a], aph ^ 1u);
        tc_fence_after();
        const uint64_t xh_d = xs_d + (uint64_t)((s * stage_bytes) >> 4), xl_d = xh_d + (uint64_t)(x_tile >> 4);
        const uint32_t dst = tmem_base + (uint32_t)(a * Kpad);
        if (elected) {
          uint32_t acc = 0u;
          for (int kb = 0; kb < nkb; ++kb) {
            const uint64_t ao = (uint64_t)((kb * XT_BYTES) >> 4), bo = (uint64_t)((kb * Kpad * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              umma_tf32(dst, xl_d + ao + adv, cbh_d + bo + adv, idesc, acc);
              umma_tf32(dst, xh_d + ao + adv, cbl_d + bo + adv, idesc, 1u);
              umma_tf32(dst, xh_d + ao + adv, cbh_d + bo + adv, idesc, 1u);
              acc = 1u;
            }
          }
          umma_commit(&sh->tmem_full[a]);
        }
        __syncwarp();
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
          const float sa = fmaf(-2.f, v[j], se[c0 + j] * (1.f - EPS_SCORE));
          const float sb = fmaf(-2.f, v[j + 1], se[c0 + j + 1] * (1.f - EPS_SCORE));
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
      const bool flagged = (p.debug & 2) ? false : !((second - best) > 2.f * EPS_SCORE * (sx + se[bkm]));
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
    p.debug = 0;
#ifdef VQS_DEBUG   /* profiling builds only (-DVQS_DEBUG): these bits switch phases off and give WRONG results */
    const char* dbg = getenv("VQS_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
#endif
  }
  const size_t smem = assign_tc_smem_bytes(K, D);
  static DevCache configured;
  if (dev_needs(configured, smem))
    VQS_CUDA(cudaFuncSetAttribute(vq_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid > max_grid) grid = max_grid;
  *grid_out = grid;
  vq_assign_tc_kernel<<<grid, NTHREADS, smem, st>>>(p);
  VQS_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqs

// =====================================================================================================================
// Large codebooks (K > what fits shared memory, e.g. the 512 / 4096-code sweep): the distance search is a real GEMM
// N x 64 . 64 x K.  PERSISTENT kernel, one CTA per SM walking over 128-row tiles: a tile stays in shared memory (hi/lo split)
// while the codebook is streamed past it in 128-code chunks (pre-built shared-memory images in global memory, one
// cp.async.bulk per chunk into a 2-stage ring), 24 + 1 tf32 MMAs per chunk into one of four TMEM accumulators; 8 scan warps keep
// a running top-3 of the scores per row.  Rows whose best and second-best scores are closer than 2*tol are settled exactly
// in fp32, so indices equal the fp32 search.
//
// Why persistent (round 2, profiles/r04b_k4096_phases.txt): one CTA per tile spent 18.7 us per tile with NOTHING to do (launch,
// barrier / TMEM set-up, the row tile's HBM round trip before the first MMA, merge + settlement + 8192 statistics atomics after the
// last one) against 27.8 us of MMA work per tile at K = 4096, and a 220 KB CTA cannot share its SM with a second one.  Now the next
// tile's rows are prefetched into registers while the MMAs of the current tile run and stored the moment the last MMA has read
// the tile (a_empty: tcgen05.commit), the chunk ring and the two accumulators run on across tile boundaries, and the
// statistics of tile i (global atomics, low contention at large K) are issued by the otherwise idle producer warps from
// the L2-hot rows while the scan warps are already on tile i + 1.
// Warp roles (576 threads): 0-7 scan (TMEM lane quarter = warp & 3, column half = warp >> 2), 8-15 row-tile producers +
// settlement + statistics, 16 MMA issuer (owns the TMEM allocation), 17 chunk streamer (one lane).
// =====================================================================================================================
namespace vqs {

struct SearchLargeParams {
  const float* z;
  const float* cb;      // exact codebook (K, D)
  const float* img;     // per-chunk shared-memory images of the split codebook (cb_prep_kernel)
  const float* se;      // |e|^2, nchunks*128 entries (+inf beyond K)
  int64_t* idx;
  float* stats;         // [counts | dw], zeroed by the launcher
  long long N;
  int layout, B, D, T, K;
  int nkb, nchunks, ntiles;
  int debug;            // phase probes of -DVQS_DEBUG builds (env VQS_TC_DEBUG), see L_DBG
  float* dbg;           // [3] counters: rows settled by the top-2 check / a 64-code group scan / a full exact scan
  FastDiv divD;
};

namespace {

#ifdef VQS_DEBUG   /* phase probes of profiling builds: bit 3 no score scan, bit 4 no MMAs, bit 5 no chunk loads (WRONG results) */
#define L_DBG(bit) (p.debug & (bit))
#else
#define L_DBG(bit) false
#endif
constexpr int L_SCAN_WARPS = 8, L_PROD_WARPS = 8;
constexpr int L_PROD_THREADS = L_PROD_WARPS * 32;
constexpr int L_MMA_WARP = L_SCAN_WARPS + L_PROD_WARPS;
constexpr int CHUNK = 128;    // codes per streamed chunk (UMMA N)
constexpr int BSTAGES = 2;    // chunk ring depth (a probe build with FOUR barrier stages aliasing two buffers, chunk loads off, ran
                              // no faster: ring depth is not what holds the MMAs back, profiles/r04g_k4096_bbars4.txt)
#ifndef VQS_LARGE_ISSUERS
#define VQS_LARGE_ISSUERS 1
#endif
constexpr int ISSUERS = VQS_LARGE_ISSUERS;   // MMA issuer warps (see the issuer loop)
static_assert(ISSUERS == 1 || ISSUERS == 2, "one or two issuer warps");
constexpr int L_LOAD_WARP = L_MMA_WARP + ISSUERS;
constexpr int L_THREADS = (L_LOAD_WARP + 1) * 32;
#ifndef VQS_LARGE_TBUF
#define VQS_LARGE_TBUF 4
#endif
constexpr int TBUF = VQS_LARGE_TBUF;   // TMEM accumulators (CHUNK columns each; 4 x 128 = all 512 columns)
constexpr int TBUF_LOG2 = TBUF == 4 ? 2 : 1;
static_assert(TBUF == 2 || TBUF == 4, "accumulator ring of 2 or 4");
// (Sharing every chunk between the CTAs of a cluster through a multicast bulk copy was built and measured: 2 CTAs 3.76 ms
// against 3.70 ms alone, 4 CTAs 4.28 ms at N = 2^20 -- the L2 -> SM stream of 74 KB per chunk per SM is not the limit.)

constexpr int M3_MAX = 7;     // whole-codebook re-scores per tile done by all producer warps together (more: one warp each)
struct LShared {
  uint64_t a_full, a_empty, b_full[BSTAGES], b_empty[BSTAGES], tmem_full[TBUF], tmem_empty[TBUF], idx_ready[2];
  uint32_t tmem_base;
};

// Pre-splits the codebook and lays every 128-code chunk out as the exact shared-memory image the MMAs read
// ([hi | lo] x [k-block] x [128 rows x 128 B, SWIZZLE_128B]), so that a chunk is ONE contiguous block in global memory
// and a single thread can stream it with cp.async.bulk (TMA bulk copy, completion on an mbarrier) -- 4096 16-byte
// cp.async per chunk gave 10 B/clk/SM.  Also |e_k|^2 (fp32, sequential order = the exact path's) and +inf padding.
// The |e_k|^2 term of the score rides in the GEMM as ONE EXTRA k-step (round 2): the row tile gets two constant columns 1, the chunk
// the columns hi and lo of g_k = -(1 - EPS_LARGE) |e_k|^2 / 2 (one MMA: 25 per chunk), so the accumulator holds
// v = x.e_k + g_k and the lower-bound score is just -2 v: the scan warps work on raw accumulator words.  The extra operands are
// [128 rows x 8 floats] tiles in the un-swizzled K-major canonical layout (8-row x 16-byte core matrices: LBO = 128 B between
// the two k-halves, SBO = 256 B between row groups).
// Score tolerance of this kernel: EPS_SCORE plus the 6 mantissa bits the scan overwrites with the column number
// (|pv - v| < 64 ulp <= 2^-17 |v|, |v| <= |x||e| + |e|^2 / 2 <= |x|^2 + |e|^2, score = -2 v: 2^-16 (|x|^2 + |e|^2)).
constexpr float EPS_LARGE = EPS_SCORE + 1.53e-5f;
constexpr int AUG_FLOATS = CHUNK * 8;                 // the extra k-step of a chunk / of the row tile
__host__ __device__ __forceinline__ int aug_off(int row, int k) {   // float index inside such a tile
  return (row >> 3) * 64 + (k >> 2) * 32 + (row & 7) * 4 + (k & 3);
}
__device__ __forceinline__ uint64_t make_desc_kmajor_plain(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;              // leading byte offset: next 16-byte k-chunk
  d |= (uint64_t)(256 >> 4) << 32;              // stride byte offset: next group of 8 rows
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}

__global__ void cb_prep_kernel(const float* __restrict__ cb, int K, int D, int Kpad, float* __restrict__ img,
                               float* __restrict__ se) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Kpad) return;
  const int nkb = D / 32;
  const int copy_floats = nkb * CHUNK * 32;            // floats per copy of one chunk
  float* chunk = img + (size_t)(k / CHUNK) * (2 * copy_floats + AUG_FLOATS);
  const int row = k % CHUNK;
  float s = 0.f;
  for (int j = 0; j < D; ++j) {
    const float v = (k < K) ? cb[(size_t)k * D + j] : 0.f;
    const float h = tf32_hi(v);
    const uint32_t off = elem_off(row, j, CHUNK) >> 2;
    chunk[off] = h;
    chunk[copy_floats + off] = v - h;
    s = fmaf(v, v, s);
  }
  se[k] = (k < K) ? s : INFINITY;
  // extra k-step: columns 0 / 1 = hi / lo part of g_k (padding codes: a huge negative value, never among the best), 2..7 = 0
  const float g = (k < K) ? -0.5f * (s * (1.f - EPS_LARGE)) : -1e30f;
  const float gh = tf32_hi(g);
  float* aug = chunk + 2 * copy_floats;
  for (int j = 0; j < 8; ++j) aug[aug_off(row, j)] = j == 0 ? gh : (j == 1 ? g - gh : 0.f);
}

struct Top3 {
  float b, s, t;
  int kb, ks;
};
// Almost every candidate is NOT among the three best seen so far, so the common path is one compare and a branch the whole
// warp skips; the update itself is branch-free selects.
__device__ __forceinline__ void top3_push(Top3& a, float sc, int k) {
  if (sc < a.t) {
    const bool lb = sc < a.b, ls = sc < a.s;
    a.t = ls ? a.s : sc;
    a.s = lb ? a.b : (ls ? sc : a.s);
    a.ks = lb ? a.kb : (ls ? k : a.ks);
    a.b = lb ? sc : a.b;
    a.kb = lb ? k : a.kb;
  }
}
// the same for the three LARGEST values (the scan runs on v = x.e + g: largest v = smallest score)
__device__ __forceinline__ void top3_push_max(Top3& a, float v, int k) {
  if (v > a.t) {
    const bool lb = v > a.b, ls = v > a.s;
    a.t = ls ? a.s : v;
    a.s = lb ? a.b : (ls ? v : a.s);
    a.ks = lb ? a.kb : (ls ? k : a.ks);
    a.b = lb ? v : a.b;
    a.kb = lb ? k : a.kb;
  }
}

// fire-and-forget fp32 add to GLOBAL memory (atomicAdd on a pointer the compiler cannot prove global also emits the shared-memory
// CAS loop: 64 unrolled ATOMS.CAST.SPIN blocks in the statistics loop)
__device__ __forceinline__ void red_add_global(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// Element (row, column j) of the row space, from global memory (the settlement paths and the statistics run after the
// tile's shared-memory copy may already hold the next tile; the rows are L2-hot).  (B, D, T) input is the reference's
// (D, T, B) row space cut into rows of D (vector_quantizer_ema.py:104-107).
__device__ __forceinline__ float z_elem(const SearchLargeParams& p, long long row, int j) {
  const long long f = row * p.D + j;
  if (p.layout == VQS_LAYOUT_FLAT_ND) return __ldg(p.z + f);
  const long long P = (long long)p.T * p.B;
  const long long d = f / P, pp = f - d * P;
  const long long t = pp / p.B, b = pp - t * p.B;
  return __ldg(p.z + ((size_t)b * p.D + (size_t)d) * p.T + t);
}
// Exact settlement of one row (the rare rows whose tensor-core scores cannot certify the argmin): candidates are re-scored
// with the fp32 formula AND summation order of the CUDA-core kernel (vq_kernels.cu: |x|^2 as 8 partials over
// j = 4p + 32 s + e and a butterfly tree, the dot product as one ascending fmaf chain), lowest index on equal distances.
// The row lives distributed over the lanes of a warp (x[lane], x[lane + 32]) and is broadcast by shuffles; one lane scores
// one candidate per pass, its code row fetched as 16-byte pieces BEFORE the dependent fmaf chain starts.
struct RowX {
  float x0, x1, sx;
};
__device__ __noinline__ RowX settle_load_row(const SearchLargeParams& p, long long row, int lane) {
  const int D = p.D;
  RowX rx;
  rx.x0 = z_elem(p, row, lane);
  rx.x1 = D > 32 ? z_elem(p, row, lane + 32) : 0.f;
  float part[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float s = 0.f;
    for (int j = q * 4; j < D; j += 32) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = __shfl_sync(0xffffffffu, (j + e) < 32 ? rx.x0 : rx.x1, (j + e) & 31);
        s = fmaf(a, a, s);
      }
    }
    part[q] = s;
  }
  const float a0 = part[0] + part[1], a1 = part[2] + part[3], a2 = part[4] + part[5], a3 = part[6] + part[7];
  const float b0 = a0 + a1, b1 = a2 + a3;
  rx.sx = b0 + b1;
  return rx;
}
// candidates: list == 1: the two codes (ka, kb2); else the n codes from k0 on.  Returns the warp's best (distance, code).
__device__ __noinline__ void settle_score(const SearchLargeParams& p, const RowX& rx, int list, int ka, int kb2, int k0, int n,
                                          int lane, float& bd_out, int& bk_out) {
  const int D = p.D;
  float bd = INFINITY;
  int bk = 0x7fffffff;
  for (int c0 = 0; c0 < n; c0 += 32) {
    const int c = c0 + lane;
    const bool ok = c < n;
    const int k = !ok ? 0 : (list ? (c == 0 ? ka : kb2) : k0 + c);
    const float4* e = reinterpret_cast<const float4*>(p.cb + (size_t)k * D);
    float4 ev[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) ev[u] = (u * 4 < D) ? __ldg(e + u) : make_float4(0.f, 0.f, 0.f, 0.f);
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (u * 4 < D) {
        const float xs = u < 8 ? rx.x0 : rx.x1;
        dot = fmaf(__shfl_sync(0xffffffffu, xs, (u * 4) & 31), ev[u].x, dot);
        dot = fmaf(__shfl_sync(0xffffffffu, xs, (u * 4 + 1) & 31), ev[u].y, dot);
        dot = fmaf(__shfl_sync(0xffffffffu, xs, (u * 4 + 2) & 31), ev[u].z, dot);
        dot = fmaf(__shfl_sync(0xffffffffu, xs, (u * 4 + 3) & 31), ev[u].w, dot);
      }
    }
    const float dd = __fsub_rn(__fadd_rn(rx.sx, __ldg(p.se + k)), __fmul_rn(2.0f, dot));
    if (ok && (dd < bd || (dd == bd && k < bk))) {
      bd = dd;
      bk = k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, bd, o);
    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
    if (od < bd || (od == bd && ok < bk)) {
      bd = od;
      bk = ok;
    }
  }
  bd_out = bd;
  bk_out = bk;
}

__global__ void __launch_bounds__(L_THREADS, 1) vq_search_large_kernel(const SearchLargeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = p.D, K = p.K, nkb = p.nkb;
  const int a_copy = nkb * XT_BYTES;                 // one copy (hi or lo) of the row tile / of a code chunk
  uint8_t* xa = smem;                                 // [hi | lo]
  uint8_t* aaug = smem + 2 * a_copy;                  // the row tile's extra k-step: columns 0, 1 = 1 (4 KB, un-swizzled)
  uint8_t* bring = aaug + AUG_FLOATS * 4;             // BSTAGES x [hi | lo | extra k-step] chunk images
  const int b_copy = nkb * CHUNK * 128;
  const int b_stage = 2 * b_copy + AUG_FLOATS * 4;
  float* mrg = reinterpret_cast<float*>(bring + BSTAGES * b_stage);   // [2][128][6] merge buffers of the upper column half
  int* sidx = reinterpret_cast<int*>(mrg + 2 * TROWS * 6);            // [2][128] indices of a tile (scan -> producers)
  int* saux = sidx + 2 * TROWS;                                       // [2][128] second-best code | settlement mode << 16
  int* m3 = saux + 2 * TROWS;                                         // [2][1 + M3_MAX] rows that need a whole-codebook re-score
  float* red = reinterpret_cast<float*>(m3 + 2 * (1 + M3_MAX));       // [8][2] per-warp partial results of such a re-score
  LShared* sh = reinterpret_cast<LShared*>(red + 2 * L_PROD_WARPS);

  if (tid == 0) {
    m3[0] = m3[1 + M3_MAX] = 0;
    mbar_init(&sh->a_full, L_PROD_WARPS);
    mbar_init(&sh->a_empty, ISSUERS + L_SCAN_WARPS / 2);   // every issuer's last MMA of the tile has read it, the |x|^2 readers are done
    for (int s = 0; s < BSTAGES; ++s) {
      mbar_init(&sh->b_full[s], 1);
      mbar_init(&sh->b_empty[s], 1);
    }
    for (int s = 0; s < TBUF; ++s) {
      mbar_init(&sh->tmem_full[s], 1);
      mbar_init(&sh->tmem_empty[s], L_SCAN_WARPS);
    }
    mbar_init(&sh->idx_ready[0], L_SCAN_WARPS / 2);
    mbar_init(&sh->idx_ready[1], L_SCAN_WARPS / 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == L_MMA_WARP) tmem_alloc(&sh->tmem_base, TBUF * CHUNK);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const int ntl = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
  const uint8_t* xh = xa;

  if (warp >= L_SCAN_WARPS && warp < L_MMA_WARP) {
    // ================= producers: row tiles in, statistics out =================
    const int ptid = tid - L_SCAN_WARPS * 32;
    const uint32_t xh_a = smem_u32(xa), xl_a = xh_a + (uint32_t)a_copy;
    if (ptid < TROWS) {                                                  // extra k-step: (1, 1, 0, ..., 0) per row, once
      const uint32_t o = smem_u32(aaug) + (uint32_t)((ptid >> 3) * 256 + (ptid & 7) * 16);
      sts_v4(o, make_float4(1.f, 1.f, 0.f, 0.f));
      sts_v4(o + 128, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    const int cpr = D >> 2;                      // 16-byte pieces per row
    const int cshift = D == 64 ? 4 : 3;          // log2(cpr): D is 32 or 64 (search_large_supported)
    for (int i = 0; i <= ntl; ++i) {
      if (i < ntl) {
        const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * TROWS;
        const long long left = p.N - r0;
        const int rows = left < TROWS ? (int)left : TROWS;
        if (p.layout == VQS_LAYOUT_FLAT_ND) {
          // the whole tile in registers BEFORE the wait: the HBM round trip hides under the previous tile's MMAs
          const float4* src = reinterpret_cast<const float4*>(p.z + r0 * D);
          const int total = rows * cpr;
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = ptid + u * L_PROD_THREADS;
            v[u] = (c < total) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (i > 0) mbar_wait(&sh->a_empty, (uint32_t)(i - 1) & 1u);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = ptid + u * L_PROD_THREADS;
            if (c < TROWS * cpr) {               // rows beyond the input are stored as zeros
              const uint32_t row = (uint32_t)c >> cshift;
              const int c16 = c - (int)(row << cshift);
              const uint32_t off = (uint32_t)((c16 >> 3) * XT_BYTES + row * 128 + (((c16 & 7) ^ (row & 7)) << 4));
              const float4 h = make_float4(tf32_hi(v[u].x), tf32_hi(v[u].y), tf32_hi(v[u].z), tf32_hi(v[u].w));
              sts_v4(xh_a + off, h);
              sts_v4(xl_a + off, make_float4(v[u].x - h.x, v[u].y - h.y, v[u].z - h.z, v[u].w - h.w));
            }
          }
        } else {
          if (i > 0) mbar_wait(&sh->a_empty, (uint32_t)(i - 1) & 1u);
          for (int e = ptid; e < (TROWS - rows) * D; e += L_PROD_THREADS) {   // zero tail rows
            uint32_t r, j;
            p.divD.divmod((uint32_t)e, r, j);
            const uint32_t off = elem_off(rows + (int)r, (int)j, TROWS);
            sts_f32(xh_a + off, 0.f);
            sts_f32(xl_a + off, 0.f);
          }
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
            const int qs = L_PROD_THREADS / span, rs = L_PROD_THREADS - qs * span;
            int b = ptid / span, tt = ptid - b * span;
            const int foff = (int)(base - f0);
            const float* zd = p.z + (size_t)d * T;
            for (int e = ptid; e < count; e += L_PROD_THREADS * 8) {
              float v[8];
              uint32_t off[8];
              bool ok[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int t = t0 + tt;
                const int pp = t * B + b;
                ok[u] = (e + u * L_PROD_THREADS < count) && pp >= p0 && pp < p1;
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
                  sts_f32(xh_a + off[u], h);
                  sts_f32(xl_a + off[u], v[u] - h);
                }
              }
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->a_full);
      }
      if (i > 0) {
        // ---- statistics of tile i - 1: global atomics, values re-read from the (L2-hot) rows ----
        const int it = i - 1;
        const long long r0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TROWS;
        const long long left = p.N - r0;
        const int rows = left < TROWS ? (int)left : TROWS;
        mbar_wait(&sh->idx_ready[it & 1], (uint32_t)(it >> 1) & 1u);
        int* si = sidx + (it & 1) * TROWS;
        {
          // near-ties the scan could not certify are settled HERE, by warps that would otherwise idle, so that neither the scan
          // warps nor (through the accumulators) the MMAs ever wait for an exact re-score; warp w owns rows 16 w .. 16 w + 15.
          // A whole-codebook re-score (mode 3, a handful per million rows, but ~300 us when one warp did it alone and every
          // CTA's tile list is fixed) is shared by all eight warps, an eighth of the codes each.
          const int pw = ptid >> 5;
          const int R = pw * 16 + (lane & 15);
          int bk = si[R];
          const int aux = saux[(it & 1) * TROWS + R];
          int* m3l = m3 + (it & 1) * (1 + M3_MAX);
          const int n3 = m3l[0] < M3_MAX ? m3l[0] : M3_MAX;
          for (int q3 = 0; q3 < n3; ++q3) {
            const int R3 = m3l[1 + q3];
            const RowX rx = settle_load_row(p, r0 + R3, lane);
            const int per = (K + L_PROD_WARPS - 1) / L_PROD_WARPS;
            const int k0 = pw * per, n = min(per, K - k0);
            float bd;
            int bkk;
            settle_score(p, rx, 0, 0, 0, k0, n, lane, bd, bkk);
            if (lane == 0) {
              red[2 * pw] = bd;
              red[2 * pw + 1] = __int_as_float(bkk);
              if (pw == 0) red_add_global(p.dbg + 2, 1.f);
            }
            named_bar_sync(2, L_PROD_THREADS);
            if (pw == (R3 >> 4) && (lane & 15) == (R3 & 15)) {
              float best = INFINITY;
              int bb = 0x7fffffff;
              for (int w = 0; w < L_PROD_WARPS; ++w) {
                const float od = red[2 * w];
                const int ok = __float_as_int(red[2 * w + 1]);
                if (od < best || (od == best && ok < bb)) {
                  best = od;
                  bb = ok;
                }
              }
              bk = bb < K ? bb : K - 1;
            }
            named_bar_sync(2, L_PROD_THREADS);
          }
          // two candidates (mode 1), one 64-code group (mode 2), and whole-codebook rows beyond the shared list: one warp each
          unsigned mk = __ballot_sync(0xffffffffu, lane < 16 && (aux >> 16) != 0 && ((aux >> 16) != 3 || (aux & 0xffff) >= n3));
          while (mk) {
            const int rr = __ffs(mk) - 1;
            mk &= mk - 1;
            const int ka = __shfl_sync(0xffffffffu, bk, rr), ax = __shfl_sync(0xffffffffu, aux, rr);
            const int mode = ax >> 16;
            if (lane == 0) red_add_global(p.dbg + mode - 1, 1.f);
            const RowX rx = settle_load_row(p, r0 + pw * 16 + rr, lane);
            const int k0 = mode == 3 ? 0 : (ka >> 6) << 6;
            const int n = mode == 1 ? 2 : (mode == 3 ? K : min(64, K - k0));
            float bd;
            int bkk;
            settle_score(p, rx, mode == 1, ka, ax & 0xffff, k0, n, lane, bd, bkk);
            if (lane == rr) bk = bkk < K ? bkk : K - 1;
          }
          if (lane < 16) {
            si[R] = bk;
            if (R < rows) p.idx[r0 + R] = (int64_t)bk;
          }
          named_bar_sync(2, L_PROD_THREADS);      // every row's final index is in si[], every warp has read the list
          if (ptid == 0) m3l[0] = 0;              // the list is free again (next used at the end of tile it + 2)
        }
        if (p.layout == VQS_LAYOUT_FLAT_ND) {
          const float* src = p.z + r0 * D;
#pragma unroll 4
          for (int e = ptid; e < rows * D; e += L_PROD_THREADS) {
            uint32_t R, j;
            p.divD.divmod((uint32_t)e, R, j);
            red_add_global(p.stats + K + (size_t)si[R] * D + j, __ldg(src + e));
          }
        } else {
#pragma unroll 1
          for (int e = ptid; e < rows * D; e += L_PROD_THREADS) {
            uint32_t R, j;
            p.divD.divmod((uint32_t)e, R, j);
            red_add_global(p.stats + K + (size_t)si[R] * D + j, z_elem(p, r0 + R, (int)j));
          }
        }
        for (int R = ptid; R < rows; R += L_PROD_THREADS) red_add_global(p.stats + si[R], 1.f);
      }
    }
  } else if (warp == L_LOAD_WARP) {
    // ================= chunk streamer: one thread, one TMA bulk copy per chunk image, across tile boundaries =================
    if (lane == 0) {
      const uint32_t ring_a = smem_u32(bring);
      const uint32_t chunk_bytes = (uint32_t)b_stage;
      const int total = ntl * p.nchunks;
      int c = 0;
      for (int g = 0; g < total; ++g) {
        const int s = g & 1;
        mbar_wait(&sh->b_empty[s], (((uint32_t)g >> 1) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&sh->b_full[s]);
        if (L_DBG(32)) {
          mbar_arrive(&sh->b_full[s]);
        } else {
          asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
                       "r"(chunk_bytes)
                       : "memory");
          const char* src = reinterpret_cast<const char*>(p.img) + (size_t)c * chunk_bytes;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           ring_a + (uint32_t)(s * b_stage)),
                       "l"(src), "r"(chunk_bytes), "r"(bar)
                       : "memory");
        }
        if (++c == p.nchunks) c = 0;
      }
    }
    __syncwarp();
  } else if (warp >= L_MMA_WARP && warp < L_LOAD_WARP) {
    // ================= MMA issuer(s): the whole converged warp runs the loop, the MMAs are guarded by elect_one() (bare
    // UTCHMMA instead of an ELECT / BRA.U.ANY loop per MMA, see gemm_tc.cu::issue_mmas).  With VQS_LARGE_ISSUERS = 2 two
    // warps take alternate chunks (issuer w: ring stage w, accumulators w and w + 2).  Measured (profiles/r04j_k4096_issuers.txt):
    // the empty protocol gets faster (0.81 -> 0.53 ms) but the kernel does not (2.25 vs 2.28 ms): what looked like an issue
    // bubble (2190 "cycles" per chunk of 1664) was the clock -- under this load the SMs run at 1.66 GHz, not 1.92
    // (ncu: tensor pipe active 83 % of active cycles).  One issuer is the default. =================
    const uint32_t w = (uint32_t)(warp - L_MMA_WARP);
    constexpr uint32_t idesc = make_idesc_tf32(CHUNK);
    const bool elected = elect_one();
    const uint64_t xh_d = make_desc_sw128(smem_u32(xa));
    const uint64_t xl_d = xh_d + (uint64_t)(a_copy >> 4);
    const uint64_t ring_d = make_desc_sw128(smem_u32(bring));
    const uint64_t aug_a = make_desc_kmajor_plain(smem_u32(aaug));
    const uint64_t aug_b0 = make_desc_kmajor_plain(smem_u32(bring) + (uint32_t)(2 * b_copy));
    uint32_t g = 0;
#pragma unroll 1
    for (int i = 0; i < ntl; ++i) {
      mbar_wait(&sh->a_full, (uint32_t)i & 1u);
      bool mine = false;
#pragma unroll 1
      for (int c = 0; c < p.nchunks; ++c, ++g) {
        if (ISSUERS == 2 && (g & 1u) != w) continue;
        mine = true;
        const uint32_t s = g & 1u, par = (g >> 1) & 1u;
        const uint32_t ta = g & (uint32_t)(TBUF - 1), tpar = (g >> TBUF_LOG2) & 1u;
        mbar_wait(&sh->b_full[s], par);
        mbar_wait(&sh->tmem_empty[ta], tpar ^ 1u);
        tc_fence_after();
        const uint64_t bh_d = ring_d + (uint64_t)((s * (uint32_t)b_stage) >> 4), bl_d = bh_d + (uint64_t)(b_copy >> 4);
        const uint32_t dst = tmem_base + ta * CHUNK;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          if (kb < nkb) {
            const uint64_t ao = (uint64_t)((kb * XT_BYTES) >> 4), bo = (uint64_t)((kb * CHUNK * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              if (elected && !L_DBG(16)) {
                umma_tf32(dst, xl_d + ao + adv, bh_d + bo + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
                umma_tf32(dst, xh_d + ao + adv, bl_d + bo + adv, idesc, 1u);
                umma_tf32(dst, xh_d + ao + adv, bh_d + bo + adv, idesc, 1u);
              }
            }
          }
        }
        if (elected && !L_DBG(16)) {
          // the |e|^2 term: (1, 1, 0, ...) x (g_hi, g_lo, 0, ...)
          umma_tf32(dst, aug_a, aug_b0 + (uint64_t)((s * (uint32_t)b_stage) >> 4), idesc, 1u);
        }
        if (elected) {
          umma_commit(&sh->b_empty[s]);
          umma_commit(&sh->tmem_full[ta]);
        }
        __syncwarp();
      }
      if (elected) {                                 // the row tile may be replaced once every issuer has said so
        if (mine) umma_commit(&sh->a_empty);
        else mbar_arrive(&sh->a_empty);
      }
      __syncwarp();
    }
    __syncwarp();
  } else {
    // ================= scan warps: running top-3 per row over the chunks of a tile =================
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;
    const int colmask = ~63 | (K >> 30);   // = ~63 (K <= 8192), opaque to the compiler so that it stays in a register
    uint32_t g = 0;
    for (int i = 0; i < ntl; ++i) {
      const long long r0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * TROWS;
      const long long left = p.N - r0;
      const int rows = left < TROWS ? (int)left : TROWS;
      // ---- |x|^2 for the bound, from the tile's hi copy (rotated conflict-free read) while it is certainly there ----
      float sx = 0.f;
      if (half == 0) {
        mbar_wait(&sh->a_full, (uint32_t)i & 1u);
        for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int pc = (r + u) & 7;
            const float4 h = *reinterpret_cast<const float4*>(xh + kb * XT_BYTES + r * 128 + (pc << 4));
            sx = fmaf(h.x, h.x, sx);
            sx = fmaf(h.y, h.y, sx);
            sx = fmaf(h.z, h.z, sx);
            sx = fmaf(h.w, h.w, sx);
          }
        }
        sx *= 1.01f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->a_empty);
      }
      Top3 top;
      top.b = top.s = top.t = -INFINITY;     // the three largest accumulator words v = x.e - (1 - EPS) |e|^2 / 2
      top.kb = top.ks = 0;
      for (int c = 0; c < p.nchunks; ++c, ++g) {
        const uint32_t a = g & (uint32_t)(TBUF - 1);
        mbar_wait(&sh->tmem_full[a], (g >> TBUF_LOG2) & 1u);
        tc_fence_after();
        // this warp's share of the chunk: lane quarter q, columns [half * CHUNK/2, (half + 1) * CHUNK/2)
        float v[CHUNK / 2];
#pragma unroll
        for (int h2 = 0; h2 < CHUNK / 64; ++h2) {
          float t[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (a * CHUNK + (uint32_t)(half * (CHUNK / 2) + h2 * 32)), t);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[h2 * 32 + j] = t[j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->tmem_empty[a]);      // scores are in registers: the accumulator is free again
        const int kbase = c * CHUNK + half * (CHUNK / 2);
        if (L_DBG(8)) {
          if (v[0] == 12345.f) top.b = 1.f;
          continue;
        }
        // Branch-free top-2 of this warp's 64 columns: the column number rides in the low 6 mantissa bits of every value
        // (error <= 2^-17 |v|, part of EPS_LARGE), so a maximum carries its index and one element costs LOP3 + 3 FMNMX in
        // four independent chains -- a compare-and-branch per element (BSSY / FSETP / BRA / BSYNC, ~20 cycles of control
        // latency with two warps per scheduler) made the scan, not the tensor pipe, the limiter: 3450 of 3960 cycles per
        // chunk (profiles/r04b_k4096_phases.txt).
        float m1[4], m2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) m1[u] = m2[u] = -INFINITY;
#pragma unroll
        for (int j = 0; j < CHUNK / 2; ++j) {
          float pv;   // (v & ~63) | j as ONE LOP3 (mask in a register the compiler cannot fold, column as the immediate)
          asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=f"(pv) : "f"(v[j]), "r"(colmask), "r"(j));
          const float lo = fminf(m1[j & 3], pv);
          m1[j & 3] = fmaxf(m1[j & 3], pv);
          m2[j & 3] = fmaxf(m2[j & 3], lo);
        }
#pragma unroll
        for (int w = 2; w > 0; w >>= 1) {
#pragma unroll
          for (int u = 0; u < w; ++u) {
            const float lo = fminf(m1[u], m1[u + w]);
            m1[u] = fmaxf(m1[u], m1[u + w]);
            m2[u] = fmaxf(lo, fmaxf(m2[u], m2[u + w]));
          }
        }
        // the chunk's two best enter the running top-3 (rarely, after the first chunks: the whole warp skips the updates)
        top3_push_max(top, m1[0], kbase + (__float_as_int(m1[0]) & 63));
        top3_push_max(top, m2[0], kbase + (__float_as_int(m2[0]) & 63));
      }
      // back to lower-bound scores (smaller = better): s = (1 - EPS) |e|^2 - 2 x.e = -2 v
      top.b *= -2.f;
      top.s *= -2.f;
      top.t *= -2.f;
      // ---- merge the two column halves (ties need no care here: equal scores are re-checked exactly) ----
      float* m = mrg + ((i & 1) * TROWS + r) * 6;
      if (half == 1) {
        m[0] = top.b; m[1] = top.s; m[2] = top.t;
        m[3] = __int_as_float(top.kb); m[4] = __int_as_float(top.ks);
      }
      named_bar_sync(1, L_SCAN_WARPS * 32);
      if (half == 0) {
        top3_push(top, m[0], __float_as_int(m[3]));
        top3_push(top, m[1], __float_as_int(m[4]));
        top3_push(top, m[2], 0);   // only its value matters (third place)
        const float tol2 = 2.f * EPS_LARGE * (sx + __ldg(p.se + top.kb));
        // The scan keeps the two best of every 64-column group, so a value it dropped lies below two kept values of its own
        // group: with the best and the second in DIFFERENT groups and the third kept value far, nothing else is near (mode 1:
        // two candidates); with both in the same group and the third kept value far every candidate lies in that group
        // (mode 2); otherwise the whole codebook is re-scored (mode 3).  Rows beyond the input are zeros: never settled.
        const bool live = r < rows && !L_DBG(8 | 16 | 32 | 64);   // (bit 6: settlement off)
        const bool close2 = live && !((top.s - top.b) > tol2);
        const bool close3 = live && !((top.t - top.b) > tol2);
        const bool same_group = (top.kb >> 6) == (top.ks >> 6);
        int mode = !close2 ? 0 : (close3 ? 3 : (same_group ? 2 : 1));
        if ((L_DBG(128) && mode == 3) || (L_DBG(256) && mode == 1) || (L_DBG(512) && mode == 2)) mode = 0;   // probes
        int second = top.ks;
        if (mode == 3) {                        // position in the tile's whole-codebook list (the second-best code is not needed)
          second = atomicAdd(m3 + (i & 1) * (1 + M3_MAX), 1);
          if (second < M3_MAX) m3[(i & 1) * (1 + M3_MAX) + 1 + second] = r;
        }
        sidx[(i & 1) * TROWS + r] = L_DBG(8 | 16 | 32) ? (r & 63) : top.kb;
        saux[(i & 1) * TROWS + r] = (second & 0xffff) | (mode << 16);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->idx_ready[i & 1]);   // the producers take the statistics from here
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == L_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TBUF * CHUNK);
  }
}

}  // namespace

bool search_large_supported(int K, int D) { return (D == 32 || D == 64) && K >= 1 && K <= 8192; }

size_t search_large_workspace_bytes(int K, int D) {
  const int nchunks = (K + CHUNK - 1) / CHUNK;
  return ((size_t)nchunks * (2 * CHUNK * D + AUG_FLOATS) + (size_t)nchunks * CHUNK + 8) * sizeof(float);
}

int launch_search_large(const float* z, int layout, int B, int D, int T, const float* cb, int K, int64_t* idx,
                        float* stats, void* workspace, cudaStream_t st) {
  SearchLargeParams p;
  const int nchunks = (K + CHUNK - 1) / CHUNK;
  float* img = (float*)workspace;
  float* se = img + (size_t)nchunks * (2 * CHUNK * D + AUG_FLOATS);
  p.z = z; p.cb = cb; p.img = img; p.se = se; p.idx = idx; p.stats = stats;
  p.N = (long long)B * T;
  p.layout = layout; p.B = B; p.D = D; p.T = T; p.K = K;
  p.nkb = D / 32;
  p.nchunks = nchunks;
  p.ntiles = (int)((p.N + TROWS - 1) / TROWS);
  p.divD = FastDiv((uint32_t)D);
  p.dbg = se + (size_t)nchunks * CHUNK + 2;
  {
    p.debug = 0;
#ifdef VQS_DEBUG   /* profiling builds only (-DVQS_DEBUG): these bits switch phases off and give WRONG results */
    const char* dbg = getenv("VQS_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
#endif
  }
  VQS_CUDA(cudaMemsetAsync(se + (size_t)nchunks * CHUNK, 0, 8 * sizeof(float), st));
  VQS_CUDA(cudaMemsetAsync(stats, 0, (size_t)K * (D + 1) * sizeof(float), st));
  if (p.ntiles == 0) return 0;
  cb_prep_kernel<<<(nchunks * CHUNK + 127) / 128, 128, 0, st>>>(cb, K, D, nchunks * CHUNK, img, se);
  VQS_LAUNCH_CHECK();
  const int a_copy = p.nkb * XT_BYTES;
  const size_t smem = (size_t)2 * a_copy + AUG_FLOATS * 4 + (size_t)BSTAGES * (2 * p.nkb * CHUNK * 128 + AUG_FLOATS * 4) +
                      ((size_t)2 * TROWS * 6 + 4 * TROWS + 2 * (1 + M3_MAX) + 2 * L_PROD_WARPS) * 4 + sizeof(LShared) + 1024 + 64;
  if (smem > 226 * 1024) {
    set_error("vq_search_large: codebook of %d codes needs %zu bytes of shared memory", K, smem);
    return VQS_ERR_ARG;
  }
  static DevCache configured;
  if (dev_needs(configured, smem))
    VQS_CUDA(cudaFuncSetAttribute(vq_search_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = p.ntiles < num_sms() ? p.ntiles : num_sms();
  vq_search_large_kernel<<<grid, L_THREADS, smem, st>>>(p);
  VQS_LAUNCH_CHECK();
#ifdef VQS_DEBUG
  {
    const char* dbg = getenv("VQS_TC_DEBUG");
    if (dbg && (atoi(dbg) & 4)) {   // profiling aid: how many rows needed the exact paths (synchronises!)
      float h[3] = {0.f, 0.f, 0.f};
      cudaStreamSynchronize(st);
      cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[vq_search_large] N=%lld K=%d: top-2 settlements %.0f, 64-code group scans %.0f, full exact scans %.0f\n",
              p.N, K, h[0], h[1], h[2]);
    }
  }
#endif
  return 0;
}

}  // namespace vqs
