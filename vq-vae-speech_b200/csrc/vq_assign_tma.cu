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
//   * per-code statistics (dw = encodings^T x) without floating-point atomics and without barriers: each of the four
//     statistics warps owns 32 rows of every tile and a PRIVATE copy of the bins in shared memory (lane = column pair),
//     so its read-add-write sequences are ordered and the final sum over warps / CTAs is fixed -> deterministic.  Counts
//     are integers (match.any + one shared-memory atomic per distinct code and warp).
// What the statistics pass costs and what did not make it cheaper (phase probe profiles/probe_tma_phases.py, 2^22 rows:
// 0.393 ms complete, 0.316 ms without statistics, 0.329 ms without the settlement, 0.188 ms with neither = the HBM floor):
//   * sums in REGISTERS of code-owning warps (warp sw owns k = sw mod 4, lane = column pair, rows prefetched one ahead):
//     correct, but selecting the accumulator costs an indirect branch per row (LDC + BRX + reconvergence): 1.27 - 1.87 ms;
//   * NO dedicated statistics warps (four groups of four warps scan a tile, exchange indices through a named barrier, and
//     warp q of the group adds the rows of the codes k = q mod 4 into per-group bins): correct, 0.305 ms without and
//     0.491 ms with statistics -- the read-add-write chains then sit on the critical path of every group.
//   * the dedicated warps' read-add-write steps SOFTWARE-PIPELINED (bit masks of the codes decide whether the next step's
//     bins may be fetched before this step's stores): correct, 0.50 ms -- the extra mask arithmetic costs more issue slots
//     than the hidden shared-memory latency gives back.
// The first is in the round-1 history (commit "register-resident statistics"); none of them is built.
// Warp roles (576 threads, one persistent CTA per SM): warps 0-11 scan (three groups taking tiles round-robin, TMEM lane
// quarter = warp % 4), warps 12-15 statistics, warp 16 TMA producer, warp 17 TMEM allocator + MMA issuer.
//
// (B, D, T) rows (the reference's own layout, ema.py:101-106: row r = 64 consecutive elements of the (D, T, B) flattening)
// run through the SAME pipeline when B is a multiple of 64: a row is then the 64 batch items b of one (d, t), a
// transposing gather no TMA box can express.  The BDT instantiation replaces the TMA warp by two producer warps that
// move a tile of TT frames x GB batch groups (TT * GB = 128 rows of one channel d) with 4-byte cp.async copies straight into the
// swizzled operand layout: a warp request covers 8 consecutive t of 4 batch items (four full 32-byte sectors of HBM, 32
// distinct banks of shared memory), one to three tiles per thread stay in flight, and a stage is published by cp.async.wait_group
// -> fence.proxy.async -> mbarrier.arrive of every producer thread.  Frames beyond T are zero-filled rows that are not
// counted.  Everything downstream (MMAs, scan, settlement, statistics) is shared with the flat instantiation.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace vqs {

struct AssignTmaParams {
  const float* cb;
  int64_t* idx;
  float* partials;  // [grid][K*65]
  long long N;
  int K, Kpad, ntiles;
  int debug;   // profiling aid (env VQS_TMA_DEBUG): bit 0 skips the statistics pass, bit 1 the settlement, bit 2 the scan
  // (B, D, T) instantiation only: tile = (batch block bbb, channel d, frame block tb), tile id = (bbb * 64 + d) * NTB + tb
  const float* z;
  int T, Q;          // frames, batch groups per (d, t) = B / 64
  int TT, GB;        // frames x batch groups per tile (TT * GB = 128), both powers of two, TT >= 8
  int tt_shift;      // log2(TT)
  int lag;           // tiles a producer thread keeps in flight behind the one it is issuing (1 .. 3)
  FastDiv divNTB;    // NTB = ceil(T / TT) frame blocks
};

namespace {

constexpr int TR = 128;                  // rows per tile
constexpr int XT = TR * 128;             // bytes of one k-block image (32 columns)
constexpr int TILE_BYTES = 2 * XT;       // 32 KB
constexpr int NSTAGE = 5;
constexpr int NGROUP = 3;                // scan groups (4 warps each) taking tiles round-robin, one TMEM accumulator each
constexpr int SCAN_WARPS = 4 * NGROUP, STAT_WARPS = 4;
constexpr int NIDX = 4;                  // index buffers between the scan and the statistics warps
constexpr int SROWS = TR / STAT_WARPS;   // rows of a tile per statistics warp
constexpr int TMA_WARP = SCAN_WARPS + STAT_WARPS, MMA_WARP = TMA_WARP + 1;
constexpr int NT = (MMA_WARP + 1) * 32;
// (B, D, T) instantiation: warp TMA_WARP and one more behind the MMA warp.  19 warps = at most 5 per SM sub-partition keeps
// the 96-register budget of the scan warps (21 warps: ptxas caps at 80 registers and spills 180 bytes).
constexpr int PROD_WARPS = 2;
constexpr int NT_BDT = NT + (PROD_WARPS - 1) * 32;
constexpr int KMAX = 48;              // four private bin copies of K x 64 floats must fit beside the tile ring
// Tensor-core filter.  kind::tf32 drops the low 13 mantissa bits of both operands: |x_j e_j - tf32(x_j) tf32(e_j)| <
// (2 * 2^-10 + 2^-20) |x_j e_j|, so the score s_k = |e_k|^2 - 2 x.e_k is off by < 3.91e-3 |x||e_k| (Cauchy-Schwarz); packing
// the code index into 6 mantissa bits adds < 1.6e-5 (|x|^2 + |e_k|^2), the fp32 rounding of the formula itself < 1e-5 of
// the same scale:  tau_k = EPSA |x||e_k| + EPSB (|x|^2 + |e_k|^2).  (The |x||e_k| form matters: an EMA-trained codebook
// has |e_k| << |x|, and with the looser (|x|^2 + |e_k|^2) / 2 twice as many rows went to the fp32 settlement.)
constexpr float EPSA = 4.0e-3f, EPSB = 3.2e-5f;
// fp32 filter: two fp32 evaluations of (|x|^2 + |e|^2) - 2 x.e in ANY summation order differ by < 2e-5 (|x|^2 + |e|^2)
constexpr float EPS2 = 2.0e-5f;
constexpr float BIG = 3.0e38f;

struct Sh {
  uint64_t full[NSTAGE], empty[NSTAGE], tmem_full[NGROUP], tmem_empty[NGROUP], idx_ready[NIDX], idx_free[NIDX];
  uint32_t tmem_base;
};

// issues the load only: tmem_ld_wait() before the values are used
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
// 4-byte asynchronous copy; src_bytes = 0 writes a zero instead of reading
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const float* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float2 lds_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
template <int IMM>
__device__ __forceinline__ float4 lds_v4_imm(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4 + %5];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr), "n"(IMM)
               : "memory");
  return v;
}
template <int IMM>
__device__ __forceinline__ float2 lds_v2_imm(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2 + %3];" : "=f"(v.x), "=f"(v.y) : "r"(addr), "n"(IMM) : "memory");
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

// one physical 16-byte chunk (both k-block images) of a row against the matching chunks of two code rows
__device__ __forceinline__ void settle_chunk(uint32_t xa, uint32_t pc4, uint32_t e1a, uint32_t e2a, uint32_t e1b,
                                             uint32_t e2b, float (&p)[4], float (&q)[4]) {
  {
    const float4 xv = lds_v4_imm<0>(xa);
    const float4 ea = lds_v4_imm<0>(e1a ^ pc4), eb = lds_v4_imm<0>(e2a ^ pc4);
    p[0] = fmaf(xv.x, ea.x, p[0]);
    p[1] = fmaf(xv.y, ea.y, p[1]);
    p[2] = fmaf(xv.z, ea.z, p[2]);
    p[3] = fmaf(xv.w, ea.w, p[3]);
    q[0] = fmaf(xv.x, eb.x, q[0]);
    q[1] = fmaf(xv.y, eb.y, q[1]);
    q[2] = fmaf(xv.z, eb.z, q[2]);
    q[3] = fmaf(xv.w, eb.w, q[3]);
  }
  {
    const float4 xv = lds_v4_imm<XT>(xa);
    const float4 ea = lds_v4_imm<0>(e1b ^ pc4), eb = lds_v4_imm<0>(e2b ^ pc4);
    p[0] = fmaf(xv.x, ea.x, p[0]);
    p[1] = fmaf(xv.y, ea.y, p[1]);
    p[2] = fmaf(xv.z, ea.z, p[2]);
    p[3] = fmaf(xv.w, ea.w, p[3]);
    q[0] = fmaf(xv.x, eb.x, q[0]);
    q[1] = fmaf(xv.y, eb.y, q[1]);
    q[2] = fmaf(xv.z, eb.z, q[2]);
    q[3] = fmaf(xv.w, eb.w, q[3]);
  }
}

// 16 rows of a statistics warp (H = 0: rows 0-15, 1: rows 16-31): this lane's column pair of every row
template <int H>
__device__ __forceinline__ void stats_load_half(const uint32_t (&xrow)[8], float2 (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = lds_v2_imm<H * 2048>(xrow[i]);
    v[8 + i] = lds_v2_imm<H * 2048 + 1024>(xrow[i]);
  }
}
// ... added to the warp's private bins.  Four rows per step: their bins are fetched together when all four differ (87 % of
// random steps), otherwise one row after the other; stores in row order.  Rows beyond N are zeros and carry a valid code,
// so they need no special case.  sidx holds the byte offset of each row's bin row.
__device__ __forceinline__ void stats_rmw_half(uint32_t sidx_a, uint32_t bins_a, const float2 (&v)[16], int n) {
  int kc[16];
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const uint4 o4 = lds_u4(sidx_a + (uint32_t)(i4 * 16));
    kc[i4 * 4 + 0] = (int)o4.x;
    kc[i4 * 4 + 1] = (int)o4.y;
    kc[i4 * 4 + 2] = (int)o4.z;
    kc[i4 * 4 + 3] = (int)o4.w;
  }
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    if (i < n) {                                        // warp-uniform
      const uint32_t a0 = bins_a + (uint32_t)kc[i], a1 = bins_a + (uint32_t)kc[i + 1];
      const uint32_t a2 = bins_a + (uint32_t)kc[i + 2], a3 = bins_a + (uint32_t)kc[i + 3];
      if (a0 != a1 && a0 != a2 && a0 != a3 && a1 != a2 && a1 != a3 && a2 != a3) {
        float2 b0 = lds_v2(a0), b1 = lds_v2(a1), b2 = lds_v2(a2), b3 = lds_v2(a3);
        b0.x += v[i].x;
        b0.y += v[i].y;
        b1.x += v[i + 1].x;
        b1.y += v[i + 1].y;
        b2.x += v[i + 2].x;
        b2.y += v[i + 2].y;
        b3.x += v[i + 3].x;
        b3.y += v[i + 3].y;
        sts_v2(a0, b0);
        sts_v2(a1, b1);
        sts_v2(a2, b2);
        sts_v2(a3, b3);
      } else {
        float2 b0 = lds_v2(a0);
        b0.x += v[i].x;
        b0.y += v[i].y;
        sts_v2(a0, b0);
        b0 = lds_v2(a1);
        b0.x += v[i + 1].x;
        b0.y += v[i + 1].y;
        sts_v2(a1, b0);
        b0 = lds_v2(a2);
        b0.x += v[i + 2].x;
        b0.y += v[i + 2].y;
        sts_v2(a2, b0);
        b0 = lds_v2(a3);
        b0.x += v[i + 3].x;
        b0.y += v[i + 3].y;
        sts_v2(a3, b0);
      }
    }
  }
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

// where tile `tile` of the (B, D, T) instantiation lives: channel d, first frame t0, first batch group bb0
struct BdtTile {
  int d, t0, bb0;
};
__device__ __forceinline__ BdtTile bdt_tile(const AssignTmaParams& p, int tile) {
  uint32_t rest, tb;
  p.divNTB.divmod((uint32_t)tile, rest, tb);
  BdtTile t;
  t.d = (int)(rest & 63u);
  t.t0 = (int)tb << p.tt_shift;
  t.bb0 = (int)(rest >> 6) * p.GB;
  return t;
}

template <bool BDT>
__global__ void __launch_bounds__(BDT ? NT_BDT : NT, 1) vq_assign_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                             const AssignTmaParams p) {
  constexpr int NT = BDT ? vqs::NT_BDT : vqs::NT;      // threads of this instantiation (shadows the flat constant)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, Kpad = p.Kpad;
  uint8_t* xs = smem;                                               // [NSTAGE][2 k-blocks][128 rows][128 B], raw fp32
  uint8_t* cbs = xs + NSTAGE * TILE_BYTES;                          // [2 k-blocks][Kpad codes][128 B], raw fp32
  float* bins = reinterpret_cast<float*>(cbs + 2 * Kpad * 128);     // [STAT_WARPS][K][64] private dw bins
  float* se = bins + STAT_WARPS * K * 64;                           // [KMAX]  |e_k|^2
  float* sea = se + KMAX;                                           // [KMAX]  |e_k|^2 (1 - EPSB), BIG beyond K
  float* sqa = sea + KMAX;                                          // [KMAX]  EPSA |e_k|
  int* cnt_s = reinterpret_cast<int*>(sqa + KMAX);                  // [KMAX]  this CTA's counts
  int* sidx = cnt_s + KMAX;                                         // [NIDX][128] code of every row of a tile
  Sh* sh = reinterpret_cast<Sh*>(sidx + NIDX * TR);

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&sh->full[s], BDT ? PROD_WARPS * 32 : 1);
      mbar_init(&sh->empty[s], 1 + 4 + STAT_WARPS);   // score MMAs (commit) + the tile's scan group + statistics warps
    }
    for (int a = 0; a < NGROUP; ++a) {
      mbar_init(&sh->tmem_full[a], 1);
      mbar_init(&sh->tmem_empty[a], 4);
    }
    for (int a = 0; a < NIDX; ++a) {
      mbar_init(&sh->idx_ready[a], 4);
      mbar_init(&sh->idx_free[a], STAT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < NGROUP * Kpad) tmem_cols <<= 1;
  if (warp == MMA_WARP) tmem_alloc(&sh->tmem_base, tmem_cols);
  for (int e = tid; e < Kpad * 64; e += NT) {
    const int k = e >> 6, j = e & 63;
    const float v = (k < K) ? __ldg(p.cb + (size_t)k * 64 + j) : 0.f;
    *reinterpret_cast<float*>(cbs + elem_off64(k, j, Kpad)) = v;
  }
  for (int e = tid; e < STAT_WARPS * K * 64; e += NT) bins[e] = 0.f;
  for (int k = tid; k < KMAX; k += NT) {
    float s = 0.f;
    if (k < K)
      for (int j = 0; j < 64; ++j) {
        const float v = __ldg(p.cb + (size_t)k * 64 + j);
        s = fmaf(v, v, s);
      }
    se[k] = (k < K) ? s : BIG;
    sea[k] = (k < K) ? s * (1.f - EPSB) : BIG;
    sqa[k] = (k < K) ? EPSA * sqrtf(s) * 1.0001f : 0.f;
    cnt_s[k] = 0;
  }
  fence_proxy_async();   // the codebook image was written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (BDT && (warp == TMA_WARP || warp > MMA_WARP)) {
    // ================= gather producers ((B, D, T) rows): 4-byte cp.async into the swizzled operand layout =================
    const int pw = warp == TMA_WARP ? 0 : warp - MMA_WARP;          // 0 .. PROD_WARPS-1: owns the 16-byte chunks jc = pw + PROD_WARPS u
    const int l7 = lane & 7, jl = lane >> 3;                        // frame within a group of 8, column within a chunk
    const int ntc = p.TT >> 3;                                      // groups of 8 frames per tile
    const uint32_t colT = (uint32_t)(64 * p.T);                     // elements between consecutive batch items
    const uint32_t lx = (uint32_t)((l7 ^ pw) << 4);
    const size_t stepJ = (size_t)(PROD_WARPS * 4) * colT;           // elements between this warp's consecutive chunks
    static_assert(PROD_WARPS == 2, "the chunk arithmetic below assumes jc = pw + 2u");
    const int L = p.lag;
    for (int it = 0; it < my_tiles + L + 1; ++it) {
      if (it >= L + 1) {                                            // tile it-L-1 has landed: publish its stage
        if (L == 1) cp_async_wait<1>();
        else if (L == 2) cp_async_wait<2>();
        else cp_async_wait<3>();
        fence_proxy_async();
        mbar_arrive(&sh->full[(it - L - 1) % NSTAGE]);
      }
      if (it < my_tiles) {
        const int s = it % NSTAGE;
        const BdtTile tl = bdt_tile(p, (int)blockIdx.x + it * (int)gridDim.x);
        mbar_wait_sleep(&sh->empty[s], ((uint32_t)(it / NSTAGE) & 1u) ^ 1u);
        const int tlast = p.T - 1 - tl.t0;                           // last frame of the tensor, relative to the tile
        if (p.TT >= 32 && !(p.debug & 8)) {
          // ---- whole lines: a warp request = 32 consecutive frames of ONE batch item (one 128-byte line of HBM; the 32 rows
          //      collide four ways on the shared-memory banks, which have cycles to spare).  Warp pw owns k-block image pw:
          //      columns j = 32 pw + u; column u sits at byte ((u >> 2) ^ (row & 7)) << 4 | (u & 3) << 2 of its row ----
          const uint32_t dst0 = smem_u32(xs + s * TILE_BYTES) + (uint32_t)(pw * XT + lane * 128 + ((lane & 7) << 4));
          const float* src0 = p.z + ((size_t)((tl.bb0 * 64 + pw * 32) * 64 + tl.d) * p.T + tl.t0);
          const int nt32 = p.TT >> 5;
          for (int bi = 0; bi < p.GB; ++bi) {
            const float* srcb = src0 + (size_t)((uint32_t)(bi * 64) * colT);
            for (int tc = 0; tc < nt32; ++tc) {
              const int tt = tc * 32 + lane;
              const uint32_t ok = tt <= tlast ? 4u : 0u;
              const float* src = srcb + (tt <= tlast ? tt : tlast);
              const uint32_t dst = dst0 + (uint32_t)(((bi << p.tt_shift) + tc * 32) * 128);   // tile row = that + lane
#pragma unroll
              for (int u = 0; u < 32; ++u)
                cp_async4_zfill((dst ^ (uint32_t)((u >> 2) << 4)) + (uint32_t)((u & 3) << 2), src + (size_t)u * colT, ok);
            }
          }
        } else {
        // The producers share their schedulers with the scan warps (the kernel is issue-bound), so the inner loop is kept
        // to ~3 instructions per copy: chunk jc = pw + 2u sits at byte ((jc & 7) ^ l7) << 4 = lx ^ ((u & 3) << 5) of its row
        // with lx = (l7 ^ pw) << 4 (disjoint bits), its k-block image is u >> 2, and consecutive u lie 8 batch items apart.
        const uint32_t dst0 = smem_u32(xs + s * TILE_BYTES) + (uint32_t)(l7 * 128 + jl * 4) + lx;
        // element (b, d, t) of z at ((b * 64 + d) * T + t); this lane: b = (bb0 + bi) * 64 + jc * 4 + jl, t = t0 + 8 tc + l7
        const float* src0 = p.z + ((size_t)((tl.bb0 * 64 + pw * 4 + jl) * 64 + tl.d) * p.T + tl.t0);
        for (int bi = 0; bi < p.GB; ++bi) {
          const float* srcb = src0 + (size_t)((uint32_t)(bi * 64) * colT);
          for (int tc = 0; tc < ntc; ++tc) {
            const int tt = tc * 8 + l7;
            const uint32_t ok = tt <= tlast ? 4u : 0u;               // frames beyond T: zero rows (address clamped, 0 bytes read)
            const float* src = srcb + (tt <= tlast ? tt : tlast);
            const uint32_t dst = dst0 + (uint32_t)(((bi << p.tt_shift) + tc * 8) * 128);   // tile row = that + l7
#pragma unroll
            for (int u = 0; u < 16 / PROD_WARPS; ++u)
              cp_async4_zfill((dst ^ (uint32_t)((u & 3) << 5)) + (uint32_t)((u >> 2) * XT), src + (size_t)u * stepJ, ok);
          }
        }
        }
      }
      cp_async_commit();
    }
  } else if (!BDT && warp == TMA_WARP) {
    // ================= TMA producer =================
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % NSTAGE;
        const long long r0 = (long long)(blockIdx.x + it * gridDim.x) * TR;
        mbar_wait_sleep(&sh->empty[s], ((uint32_t)(it / NSTAGE) & 1u) ^ 1u);
        mbar_expect_tx(&sh->full[s], TILE_BYTES);
        const uint32_t dst = smem_u32(xs + s * TILE_BYTES);
        tma_load_2d(dst, &tmap, 0, (int)r0, &sh->full[s]);         // rows beyond N arrive as zeros
        tma_load_2d(dst + XT, &tmap, 32, (int)r0, &sh->full[s]);
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer: 8 single-pass tf32 MMAs per tile on the raw fp32 words =================
    // (the whole converged warp runs the loop and the MMAs are guarded by elect_one(): bare UTCHMMA instructions instead of
    // an ELECT / BRA.U.ANY loop per MMA under `if (lane == 0)`, see gemm_tc.cu::issue_mmas)
    {
      const uint32_t idesc = make_idesc_tf32(Kpad);
      const bool elected = elect_one();
      const uint64_t cb_d = make_desc_sw128(smem_u32(cbs));
      const uint64_t x_d0 = make_desc_sw128(smem_u32(xs));
      const uint64_t cb_kb = (uint64_t)((Kpad * 128) >> 4);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % NSTAGE, a = it % NGROUP;
        mbar_wait_sleep(&sh->full[s], (uint32_t)(it / NSTAGE) & 1u);
        mbar_wait_sleep(&sh->tmem_empty[a], ((uint32_t)(it / NGROUP) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t x_d = x_d0 + (uint64_t)((s * TILE_BYTES) >> 4);
        const uint32_t dst = tmem_base + (uint32_t)(a * Kpad);
        if (elected) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t ad = x_d + (uint64_t)((kb * XT) >> 4), bd = cb_d + (uint64_t)kb * cb_kb;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(dst, ad + (uint64_t)((k * 32) >> 4), bd + (uint64_t)((k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&sh->tmem_full[a]);
          umma_commit(&sh->empty[s]);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else if (warp < SCAN_WARPS) {
    // ================= scan warps: scores -> index =================
    const int g = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    for (int it = g; it < my_tiles; it += NGROUP) {
      const int s = it % NSTAGE;
      const uint32_t ph2 = (uint32_t)(it / NGROUP) & 1u;
      // this thread's row: its position in idx and whether it exists
      long long grow;
      bool valid;
      if (BDT) {
        const BdtTile tl = bdt_tile(p, (int)blockIdx.x + it * (int)gridDim.x);
        const int t = tl.t0 + (r & (p.TT - 1));
        valid = t < p.T;
        grow = ((long long)tl.d * p.T + t) * p.Q + tl.bb0 + (r >> p.tt_shift);
      } else {
        grow = (long long)(blockIdx.x + it * gridDim.x) * TR + r;
        valid = grow < p.N;
      }
      const uint8_t* xt = xs + s * TILE_BYTES;
      mbar_wait_sleep(&sh->full[s], (uint32_t)(it / NSTAGE) & 1u);   // acquire the TMA writes for this thread's own reads
      // |x|^2 (any order: it only scales the bounds) while the MMAs run; lanes walk the 16-byte chunks of their row in a
      // rotated order so that every quarter-warp touches 8 distinct bank groups; four independent chains
      float sx;
      {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        // physical chunk (r & 7) ^ i: row base and chunk offset combine by XOR (the row base is 128-byte aligned)
        const uint32_t x0 = (smem_u32(xt) + (uint32_t)(r * 128)) ^ (uint32_t)((r & 7) << 4);
        if (!(p.debug & 2)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 h = lds_v4_imm<0>(x0 ^ (uint32_t)(i << 4)), h2 = lds_v4_imm<XT>(x0 ^ (uint32_t)(i << 4));
            a0 = fmaf(h.x, h.x, a0);
            a1 = fmaf(h.y, h.y, a1);
            a2 = fmaf(h.z, h.z, a2);
            a3 = fmaf(h.w, h.w, a3);
            a0 = fmaf(h2.x, h2.x, a0);
            a1 = fmaf(h2.y, h2.y, a1);
            a2 = fmaf(h2.z, h2.z, a2);
            a3 = fmaf(h2.w, h2.w, a3);
          }
        }
        sx = ((a0 + a1) + (a2 + a3)) * 1.0001f;
      }
      const float nrsx = -sqrtf(sx) * 1.0001f;             // -|x| (upper bound)
      mbar_wait_sleep(&sh->tmem_full[g], ph2);
      tc_fence_after();
      // ---- lower bounds adj_k = s_k - EPSB |e_k|^2 - EPSA |x||e_k| with k in the low 6 mantissa bits: three smallest, kept by
      //      two independent min/max chains (even / odd codes) ----
      float b = INFINITY, s2 = INFINITY, t3 = INFINITY, b1 = INFINITY, s21 = INFINITY, t31 = INFINITY;
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * Kpad);
      float v[KMAX];
#pragma unroll
      for (int c0 = 0; c0 < KMAX; c0 += 16)
        if (c0 < Kpad) tmem_ld16(ta + (uint32_t)c0, v + c0);
      tmem_ld_wait();
#pragma unroll
      for (int c0 = 0; c0 < KMAX; c0 += 16) {
        if (c0 < ((p.debug & 4) ? 16 : Kpad)) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float adj0 = fmaf(nrsx, sqa[c0 + j], fmaf(-2.f, v[c0 + j], sea[c0 + j]));
            const float adj1 = fmaf(nrsx, sqa[c0 + j + 1], fmaf(-2.f, v[c0 + j + 1], sea[c0 + j + 1]));
            const float key0 = __uint_as_float((__float_as_uint(adj0) & 0xFFFFFFC0u) | (uint32_t)(c0 + j));
            const float key1 = __uint_as_float((__float_as_uint(adj1) & 0xFFFFFFC0u) | (uint32_t)(c0 + j + 1));
            const float h0 = fmaxf(b, key0), h1 = fmaxf(b1, key1);
            b = fminf(b, key0);
            b1 = fminf(b1, key1);
            const float g0 = fmaxf(s2, h0), g1 = fmaxf(s21, h1);
            s2 = fminf(s2, h0);
            s21 = fminf(s21, h1);
            t3 = fminf(t3, g0);
            t31 = fminf(t31, g1);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty[g]);   // the accumulator can be overwritten
      // merge the chains: push (b1, s21, t31) into (b, s2, t3)
      {
        float h = fmaxf(b, b1);
        b = fminf(b, b1);
        float gg = fmaxf(s2, h);
        s2 = fminf(s2, h);
        t3 = fminf(t3, gg);
        h = fmaxf(b, s21);       // s21 >= b1 >= new b: only second / third place
        gg = fmaxf(s2, h);
        s2 = fminf(s2, h);
        t3 = fminf(t3, gg);
        t3 = fminf(t3, fmaxf(s2, t31));
      }
      const int k1 = (int)(__float_as_uint(b) & 63u);
      const int k2 = (int)(__float_as_uint(s2) & 63u);
      int bk = k1;
      const float tol2 = 2.f * (EPSB * (sx + se[k1]) - nrsx * sqa[k1]);
      const bool close2 = (p.debug & 2) ? false : !((s2 - b) > tol2);
      const bool close3 = !((t3 - b) > tol2);
      // ---- exactly two candidates (~7 % of random rows): the row's own lane re-scores both in fp32 ----
      bool unsafe = false;
      if (close2 && !close3) {
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        // walk the PHYSICAL 16-byte chunks of the row (immediate offsets); the matching chunk of a code row is an XOR away
        const uint32_t xr = smem_u32(xt) + (uint32_t)(r * 128);
        const uint32_t e1a = (smem_u32(cbs) + (uint32_t)(k1 * 128)) ^ (uint32_t)(((r ^ k1) & 7) << 4);
        const uint32_t e2a = (smem_u32(cbs) + (uint32_t)(k2 * 128)) ^ (uint32_t)(((r ^ k2) & 7) << 4);
        const uint32_t e1b = e1a + (uint32_t)(Kpad * 128), e2b = e2a + (uint32_t)(Kpad * 128);
        float pa[4] = {0.f, 0.f, 0.f, 0.f}, qa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
        for (int pc = 0; pc < 8; ++pc) settle_chunk(xr + (uint32_t)(pc << 4), (uint32_t)(pc << 4), e1a, e2a, e1b, e2b, pa, qa);
        p0 = pa[0]; p1 = pa[1]; p2 = pa[2]; p3 = pa[3];
        q0 = qa[0]; q1 = qa[1]; q2 = qa[2]; q3 = qa[3];
        const float se1 = se[k1], se2 = se[k2];
        const float dd1 = (sx + se1) - 2.f * ((p0 + p1) + (p2 + p3)), dd2 = (sx + se2) - 2.f * ((q0 + q1) + (q2 + q3));
        const float diff = dd1 - dd2;
        if (fabsf(diff) > 2.f * EPS2 * (sx + fmaxf(se1, se2))) bk = diff < 0.f ? k1 : k2;
        else unsafe = true;                              // too close for an order-independent decision
      }
      unsigned canon = __ballot_sync(0xffffffffu, close2 && close3);
      const unsigned canon2 = __ballot_sync(0xffffffffu, unsafe);
      // ---- three or more candidates: fp32 re-score of ALL codes (lanes = codes lane, lane + 32), best two ----
      {
        unsigned m3 = canon;
        canon = canon2;
        while (m3) {
          const int rr = __ffs(m3) - 1;
          m3 &= m3 - 1;
          const int R = q * 32 + rr;
          const float sxr = __shfl_sync(0xffffffffu, sx, rr);
          const int kA = lane, kB = lane + 32 < Kpad ? lane + 32 : lane;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll 2
          for (int c = 0; c < 16; ++c) {
            const int kbo = c >> 3, ch = c & 7;
            const float4 xv = *reinterpret_cast<const float4*>(xt + kbo * XT + R * 128 + ((ch ^ (R & 7)) << 4));
            const float4 ea = *reinterpret_cast<const float4*>(cbs + kbo * Kpad * 128 + kA * 128 + ((ch ^ (kA & 7)) << 4));
            const float4 eb = *reinterpret_cast<const float4*>(cbs + kbo * Kpad * 128 + kB * 128 + ((ch ^ (kB & 7)) << 4));
            a0 = fmaf(xv.x, ea.x, a0);
            a1 = fmaf(xv.y, ea.y, a1);
            a2 = fmaf(xv.z, ea.z, a2);
            a3 = fmaf(xv.w, ea.w, a3);
            c0 = fmaf(xv.x, eb.x, c0);
            c1 = fmaf(xv.y, eb.y, c1);
            c2 = fmaf(xv.z, eb.z, c2);
            c3 = fmaf(xv.w, eb.w, c3);
          }
          const float dA = kA < K ? (sxr + se[kA]) - 2.f * ((a0 + a1) + (a2 + a3)) : BIG;
          const float dB = lane + 32 < K ? (sxr + se[kB]) - 2.f * ((c0 + c1) + (c2 + c3)) : BIG;
          // (best, its code, second best, its code) of this lane, then a butterfly merge
          float m1 = dA, n1 = dB;
          int i1 = kA, j1 = lane + 32;
          if (n1 < m1) {
            const float tf = m1; m1 = n1; n1 = tf;
            const int ti = i1; i1 = j1; j1 = ti;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m1, o), on = __shfl_xor_sync(0xffffffffu, n1, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i1, o), oj = __shfl_xor_sync(0xffffffffu, j1, o);
            if (om < m1 || (om == m1 && oi < i1)) {      // partner's best wins: second = min(my best, partner's second)
              if (m1 < on || (m1 == on && i1 < oj)) { n1 = m1; j1 = i1; } else { n1 = on; j1 = oj; }
              m1 = om; i1 = oi;
            } else {                                      // mine wins: second = min(my second, partner's best)
              if (om < n1 || (om == n1 && oi < j1)) { n1 = om; j1 = oi; }
            }
          }
          const float sc2 = sxr + fmaxf(se[i1 < K ? i1 : 0], se[j1 < K ? j1 : 0]);
          if ((n1 - m1) > 2.f * EPS2 * sc2) {
            if (lane == rr) bk = i1;
          } else {
            canon |= 1u << rr;
          }
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
#pragma unroll 4
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
      if (valid) p.idx[grow] = (int64_t)bk;
      // ---- counts: one integer atomic per distinct code of this warp ----
      const int kk = valid ? bk : 255;
      const unsigned same = __match_any_sync(0xffffffffu, kk);
      if (valid && (same & ((1u << lane) - 1u)) == 0u) atomicAdd(&cnt_s[bk], __popc(same));
      const int ib = it % NIDX;
      mbar_wait_sleep(&sh->idx_free[ib], ((uint32_t)(it / NIDX) & 1u) ^ 1u);   // the statistics warps are done with tile it - NIDX
      sidx[ib * TR + r] = bk * 256;                    // byte offset of the code's bin row
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sh->idx_ready[ib]);
        mbar_arrive(&sh->empty[s]);                     // this warp no longer reads the row tile
      }
    }
  } else if (warp < TMA_WARP) {
    // ================= statistics warps: 32 rows each, private bins, ordered read-add-write =================
    const int sw = warp - SCAN_WARPS;
    const uint32_t bins_a = smem_u32(bins) + (uint32_t)((sw * K * 64 + 2 * lane) * 4);
    const uint32_t sidx_a = smem_u32(sidx);
    // this lane's column pair (2 lane, 2 lane + 1): k-block image, byte inside the chunk; 16-byte chunk index
    const uint32_t lane_base = (uint32_t)((lane >> 4) * XT + ((lane & 1) << 3) + sw * SROWS * 128);
    const uint32_t lane_c = (uint32_t)((lane & 15) >> 1);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NSTAGE, a = it % NIDX;
      int n = SROWS;                                   // (B, D, T): rows beyond T are zero rows with a valid code
      if (!BDT) {
        const long long left = p.N - (long long)(blockIdx.x + it * gridDim.x) * TR;
        n = (left < TR ? (int)left : TR) - sw * SROWS;  // rows of this warp that exist (may be <= 0 or > SROWS)
      }
      mbar_wait_sleep(&sh->full[s], (uint32_t)(it / NSTAGE) & 1u);
      // the row values do not depend on the indices: load them while the scan warps work (rows beyond N are zeros)
      const uint32_t x_a = smem_u32(xs + s * TILE_BYTES) + lane_base;
      uint32_t xrow[8];                                // rows i, i + 8, ... share the swizzle: one base per i & 7
#pragma unroll
      for (int i = 0; i < 8; ++i) xrow[i] = x_a + (uint32_t)(i * 128) + ((((uint32_t)i) ^ lane_c) << 4);
      float2 v[16];
      stats_load_half<0>(xrow, v);
      mbar_wait_sleep(&sh->idx_ready[a], (uint32_t)(it / NIDX) & 1u);
      const uint32_t sidx_t = sidx_a + (uint32_t)((a * TR + sw * SROWS) * 4);
      if (!(p.debug & 1)) stats_rmw_half(sidx_t, bins_a, v, n);
      stats_load_half<1>(xrow, v);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->empty[s]);       // the row values are in registers: the stage can be refilled
      if (!(p.debug & 1)) stats_rmw_half(sidx_t + 64u, bins_a, v, n - 16);
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->idx_free[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  // ---- publish this CTA's partial statistics (the private bins are summed in warp order) ----
  {
    float* out = p.partials + (size_t)blockIdx.x * K * 65;
    for (int i = tid; i < K; i += NT) out[i] = (float)cnt_s[i];
    for (int i = tid; i < K * 64; i += NT) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < STAT_WARPS; ++w) a += bins[w * K * 64 + i];
      out[K + i] = a;
    }
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

size_t smem_bytes_tma(int K, int Kpad) {
  size_t b = (size_t)NSTAGE * TILE_BYTES + (size_t)2 * Kpad * 128 + (size_t)STAT_WARPS * K * 64 * 4;
  b += (size_t)(4 * KMAX) * 4 + (size_t)(NIDX * TR) * 4 + sizeof(Sh);
  return b + 1024 + 64;
}

// frames x batch groups of a (B, D, T) tile: the feasible power-of-two split of 128 rows that wastes the fewest frames
bool bdt_tile_shape(int B, int T, int* TT_out, int* GB_out) {
  if (B % 64 != 0) return false;
  const int Q = B / 64;
  int best = 0;
  long long best_cost = 0;
  const int order[5] = {32, 64, 16, 128, 8};                        // preference on equal cost (32 x 4 is the measured shape)
  for (int i = 0; i < 5; ++i) {
    const int TT = order[i];
    const int GB = TR / TT;
    if (Q % GB != 0) continue;
    const long long cost = (long long)((T + TT - 1) / TT) * TT;     // frames processed per batch group
    if (cost > 2ll * T) continue;                                   // more padding than data: leave it to the other engines
    if (best == 0 || cost < best_cost) {
      best = TT;
      best_cost = cost;
    }
  }
  if (best == 0) return false;
  *TT_out = best;
  *GB_out = TR / best;
  return true;
}

}  // namespace

// codebook resident in shared memory, D = 64: flat (N, 64) rows (16-byte aligned z) or (B, D, T) rows with B % 64 == 0
bool assign_tma_supported(int layout, int B, int T, int K, int D) {
  const long long N = (long long)B * T;
  if (D != 64 || K < 1 || K > KMAX || N < 1 || N >= (1ll << 31) - TR) return false;
  if (layout == VQS_LAYOUT_FLAT_ND) return true;
  int TT, GB;
  return layout == VQS_LAYOUT_BDT_AS_DTB && bdt_tile_shape(B, T, &TT, &GB);
}

// `partials` must hold grid * K*(D+1) floats; the grid size comes back through *grid_out.
int launch_assign_tma(const float* z, int layout, int B, int T, const float* cb, int K, int64_t* idx, float* partials,
                      int max_grid, int* grid_out, cudaStream_t st) {
  const long long N = (long long)B * T;
  AssignTmaParams p;
  p.cb = cb; p.idx = idx; p.partials = partials; p.N = N; p.K = K;
  p.Kpad = (K + 15) / 16 * 16;
  p.z = z; p.T = T; p.Q = B / 64; p.TT = 0; p.GB = 0; p.tt_shift = 0;
  {
    p.debug = 0;
#ifdef VQS_DEBUG   /* profiling builds only (-DVQS_DEBUG): these bits switch phases off and give WRONG results */
    const char* dbg = getenv("VQS_TMA_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
#endif
  }
  {
    const char* lg = getenv("VQS_TMA_LAG");
    p.lag = lg ? atoi(lg) : 1;
    if (p.lag < 1) p.lag = 1;
    if (p.lag > 3) p.lag = 3;
  }
  const bool bdt = layout == VQS_LAYOUT_BDT_AS_DTB;
  CUtensorMap map;
  if (bdt) {
    if (!bdt_tile_shape(B, T, &p.TT, &p.GB)) {
      set_error("vq_assign: no (B, D, T) tile shape for B=%d T=%d", B, T);
      return VQS_ERR_ARG;
    }
    while ((1 << p.tt_shift) < p.TT) ++p.tt_shift;
    const int NTB = (T + p.TT - 1) / p.TT;
    p.divNTB = FastDiv((uint32_t)NTB);
    p.ntiles = (B / (64 * p.GB)) * 64 * NTB;
    memset(&map, 0, sizeof(map));                  // unused by this instantiation
  } else {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
      set_error("vq_assign: cuTensorMapEncodeTiled is not available from the driver");
      return VQS_ERR_ARG;
    }
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
    p.ntiles = (int)((N + TR - 1) / TR);
  }
  const size_t smem = smem_bytes_tma(K, p.Kpad);
  static DevCache configured[2];
  if (dev_needs(configured[bdt ? 1 : 0], smem)) {
    if (bdt)
      VQS_CUDA(cudaFuncSetAttribute(vq_assign_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      VQS_CUDA(cudaFuncSetAttribute(vq_assign_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int grid = num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid > max_grid) grid = max_grid;
  *grid_out = grid;
  if (bdt) vq_assign_tma_kernel<true><<<grid, NT_BDT, smem, st>>>(map, p);
  else vq_assign_tma_kernel<false><<<grid, NT, smem, st>>>(map, p);
  VQS_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqs
