// tcgen05 (5th-generation tensor core) implicit-GEMM kernels for the conv / wgrad GEMMs of the training step.
//
//   D[128 x BN] (fp32, TMEM) += A[128 x 32] (smem, K-major, 128B swizzle) * B[BN x 32]^T (smem, K-major, 128B swizzle)
//
// kind::tf32 MMAs issued by one thread, accumulators in tensor memory, operands produced by 8 "producer" warps that gather
// the implicit-GEMM operands straight from the NCL activation / weight tensors (padding, stride, tap flip, input ReLU are
// index arithmetic, exactly as in the CUDA-core kernels) and write them into the UMMA canonical shared-memory layout.
// Two precisions:
//   1  single-pass TF32: operands are used as stored (the tensor core reads the top 19 bits) -- what cuDNN does by default
//      for the reference's convs on a GPU (torch.backends.cudnn.allow_tf32 = True);
//   2  3xTF32: every operand is split x = hi + lo (hi = top 19 bits, lo = x - hi exactly), D += Al*Bh + Ah*Bl + Ah*Bh.
//      Dropped terms are <= 2^-21 relative, accumulation is fp32 -> meets the 1e-5 parity bar of the exact-fp32 path
//      while running on the tensor pipe.  fp32 operands cost 4 B per element, so one 128x128 tile needs 1 KB of operands per
//      k: at ~42 B/clk/SM of L2 bandwidth the single-pass variant is operand-bandwidth-bound at about a third of the TF32
//      peak and the three MMAs of the split ride in that shadow -- the exact mode is (almost) free.
//      The tensor core adds into its fp32 accumulator with truncation, a bias that grows with the number of accumulation
//      steps (measured: 1.8e-5 relative after 864 MMAs into one accumulator).  Mode 2 therefore spreads the main term
//      Ah*Bh round-robin over three TMEM accumulators and keeps the small correction terms in a fourth; the epilogue adds
//      the four in fp32 round-to-nearest.  All 512 TMEM columns are used (BN = 128).
//
// Warp roles (576 threads): warps 0-15 producers (all of them drain the epilogue: TMEM lane quarter = warp id % 4),
// warp 16 = TMEM allocator + MMA issuer, warp 17 = second MMA issuer of the 3xTF32 split (see issue_mmas).  Pipelines:
// full[stage] (one arrival per producer warp) / empty[stage] (one tcgen05.commit per issuer), tmem_full (one
// tcgen05.commit per issuer after the last k-block).
//
// K order of the conv-like GEMM is TAP-MAJOR: kk = j*Cred + c, A = [M][ksz][Cred] (vqs_permute_weight modes 1/2), so a
// 32-wide k-block has ONE tap j and 32 consecutive channels: the padding / stride / bounds arithmetic of the gather is
// done once per k-block per thread and the loads walk a constant channel stride (first build: (c, j) order, 815
// instructions per thread per k-block, issue-bound at 43 TFLOP/s -- profiles/r01c_ncu_full_gemm_tc_vq.txt).
#include <stdlib.h>

#include <type_traits>

#include "gemm_params.cuh"
#include "tc_common.cuh"

namespace vqs {
namespace {

constexpr int BM = 128;
constexpr int BKF = 32;                    // k-elements (floats) per stage row = one 128-byte swizzle atom
constexpr int N_PROD_WARPS = 16;
constexpr int N_PROD = N_PROD_WARPS * 32;
constexpr int TC_THREADS = N_PROD + 64;   // + warp 16 (TMEM allocator, MMA issuer) + warp 17 (second MMA issuer, 3xTF32)

// PAIR = 1: the CTA is one half of a cta_group::2 pair (cluster of two CTAs along grid.y = two neighbouring 128-row tiles of
// M sharing one BN-column tile): it stages its own 128 x 32 block of A and only BNL = BN / 2 rows of B (rows
// [rank * BNL, rank * BNL + BNL) of the tile); the MMAs (M = 256) are issued by the leader CTA and read both halves.
template <int BN, int PASSES, int PAIR = 0>
struct TcCfg {
  static constexpr int NOPS = (PASSES == 3) ? 2 : 1;                // hi (+ lo) copies per operand
  static constexpr int BNL = PAIR ? BN / 2 : BN;                    // B rows staged by this CTA
  static constexpr int A_BYTES = BM * 128, B_BYTES = BNL * 128;
  static constexpr int STAGE_BYTES = NOPS * (A_BYTES + B_BYTES);
  // (3-pass, BN = 64: a ring of 3 x 48 KB measured faster than 4 -- 584 vs 743 cycles per k-block in the protocol
  // microbenchmark profiles/mma_pipe.cu)
  static constexpr int STAGES_MAX = PAIR ? 4 : ((PASSES == 3) ? 3 : 4);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > STAGES_MAX ? STAGES_MAX : (200 * 1024) / STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers, row table*/;
  static constexpr int NACC = (PASSES == 3) ? 4 : 1;                // TMEM accumulators of BN columns each
  static constexpr int TMEM_USED = NACC * BN;                        // accumulator a lives at columns [a * BN, a * BN + BN)
  static constexpr int TMEM_COLS = TMEM_USED <= 64 ? 64 : (TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512));   // allocations are powers of two
  static constexpr int N_ISSUERS = (PASSES == 3) ? 2 : 1;           // threads that issue tcgen05.mma (see issue_mmas)
  static_assert(STAGES >= 2, "need at least a double buffer");
};

// Profiling builds only (-DVQS_GEMM_TIMING): CTA (0, 0, 0) leaves clock64() stamps of its phases in g_tc_timing, read back by
// vqs_debug_gemm_timing (profiles/probe_gemm_phases.py).  0: prologue done, 1: first stage full (issuer), 2: last MMA issued,
// 3: accumulators complete (epilogue warp 0), 4: epilogue done, 5: kernel entry, 6 / 7: globaltimer at entry / end.
#ifdef VQS_GEMM_TIMING
__device__ long long g_tc_timing[48];
#define TC_STAMP(i)                                                                        \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_timing[i] = clock64(); \
  } while (0)
__device__ __forceinline__ long long tc_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP_PEER(i)                                                                   \
  do {                                                                                     \
    if (blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_timing[i] = clock64(); \
  } while (0)
#define TC_STAMP_NS(i)                                                                           \
  do {                                                                                           \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_timing[i] = tc_globaltimer(); \
  } while (0)
#else
#define TC_STAMP(i) do {} while (0)
#define TC_STAMP_PEER(i) do {} while (0)
#define TC_STAMP_NS(i) do {} while (0)
#endif
#ifdef VQS_GEMM_TIMING
#define TC_ACC_BEGIN() const long long _t0 = clock64()
#define TC_ACC_END(v) v += clock64() - _t0
#define TC_ACC_PUT(i, v)                                                                 \
  do {                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_timing[i] = (v);     \
  } while (0)
#else
#define TC_ACC_BEGIN() do {} while (0)
#define TC_ACC_END(v) do {} while (0)
#define TC_ACC_PUT(i, v) do {} while (0)
#endif

struct TcShared {
  uint64_t full[4], empty[4], tmem_full;
  uint32_t tmem_base;
  int2 btab[64];   // wgrad pair kernel: per B row of this CTA (c * Lx, j * j_mul + off), or a very negative shift for rows beyond Nw
};

// ---------------------------------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------------------------------
template <int PASSES>
__device__ __forceinline__ void st_elem(uint32_t hi, uint32_t lo, float v) {
  if (PASSES == 3) {
    float h = tf32_hi(v);
    sts_f32(hi, h);
    sts_f32(lo, v - h);
  } else {
    sts_f32(hi, v);
  }
}
template <int PASSES>
__device__ __forceinline__ void st_vec4(uint32_t hi, uint32_t lo, float4 v) {
  if (PASSES == 3) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    sts_v4(hi, h);
    sts_v4(lo, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
  } else {
    sts_v4(hi, v);
  }
}

// MODE 0: conv-like GEMM (A = dense weights [M][Ktot], B gathered from NCL activations)
// MODE 1: wgrad GEMM     (A = Aact[b, m, l], B = X'[b, c, l*l_mul + j*j_mul + off], reduction over (b, l))
template <int MODE>
struct TcParams;
template <>
struct TcParams<0> {
  ConvParams p;
};
template <>
struct TcParams<1> {
  WgradParams p;
};

// Producer register stage: every global load of one k-block is issued before the first shared-memory store, and the
// loads of k-block i+1 are in flight while k-block i is being stored (software pipelining across the empty-slot wait).
template <int BN>
struct ConvRegs {
  static constexpr int TPR = N_PROD / BN;   // threads per B row: 4 (BN = 128) or 8 (BN = 64)
  static constexpr int CH = 8 / TPR;        // 16-byte chunks of a B row per thread: 2 or 1
  static constexpr int AI = BM * 8 / N_PROD;  // A chunks per thread: 2
  float4 a[AI];
  float b[CH * 4];
};

template <int BN>
__device__ __forceinline__ void load_conv(const ConvParams& p, int kb, ConvRegs<BN>& rg, int m0, int ptid, bool n_ok,
                                          const float* xb, int lbase, int bk0) {
  const vqs_conv_gemm_desc& d = p.d;
  // tap-major K: k-block kb = tap j, channels [c0, c0 + 32)
  const uint32_t j = p.divCpb.div((uint32_t)kb);
  const int c0 = (kb - (int)j * p.cpb) * BKF;
  if (d.a_tap_major != 2) {   // (with an operand image the A tile arrives by cp.async.bulk, see the stage code)
    const int c = ptid & 7;  // A: float4 along k; 8 consecutive threads cover one 128-byte row
    const float* arow = d.A + (size_t)kb * BKF + c * 4;
#pragma unroll
    for (int i = 0; i < ConvRegs<BN>::AI; ++i) {
      const int m = m0 + (ptid >> 3) + i * (N_PROD / 8);
      rg.a[i] = (m < d.M) ? __ldg(reinterpret_cast<const float4*>(arow + (size_t)m * p.Ktot))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // B: thread = column n (fixed (b, l)); one bounds decision per k-block, then a constant channel stride
  int pn = lbase + (int)j * d.j_mul;
  bool ok = n_ok && pn >= 0;
  if (d.l_div == 2) {
    ok = ok && ((pn & 1) == 0);
    pn >>= 1;
  }
  ok = ok && pn < d.Lin;
  const float* ptr = xb + (long long)pn * d.x_sl + (long long)(c0 + bk0 * 4) * d.x_sc;
  const int cleft = p.cred_real - (c0 + bk0 * 4);     // channels of X from this thread's first one on (padded images: < CH * 4)
#pragma unroll
  for (int e = 0; e < ConvRegs<BN>::CH * 4; ++e) rg.b[e] = (ok && e < cleft) ? __ldg(ptr + (long long)e * d.x_sc) : 0.f;
}

// offA = byte offset of this thread's first A chunk, offB[ci] of its B chunks (k-block invariant)
template <int BN, int PASSES>
__device__ __forceinline__ void store_conv(const ConvParams& p, const ConvRegs<BN>& rg, uint32_t sA_hi, uint32_t sA_lo,
                                           uint32_t sB_hi, uint32_t sB_lo, uint32_t offA,
                                           const uint32_t (&offB)[ConvRegs<BN>::CH]) {
  if (p.d.a_tap_major != 2) {
#pragma unroll
    for (int i = 0; i < ConvRegs<BN>::AI; ++i) {
      const uint32_t o = offA + (uint32_t)(i * (N_PROD / 8) * 128);   // rows step by 64: (r & 7) unchanged
      st_vec4<PASSES>(sA_hi + o, sA_lo + o, rg.a[i]);
    }
  }
  const bool rl = p.d.x_relu != 0;
#pragma unroll
  for (int ci = 0; ci < ConvRegs<BN>::CH; ++ci) {
    float4 v = make_float4(rg.b[ci * 4 + 0], rg.b[ci * 4 + 1], rg.b[ci * 4 + 2], rg.b[ci * 4 + 3]);
    if (rl) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    st_vec4<PASSES>(sB_hi + offB[ci], sB_lo + offB[ci], v);
  }
}

template <int BN>
struct WgradRegs {
  static constexpr int RA = BM / N_PROD_WARPS, RB = BN / N_PROD_WARPS;  // rows per warp
  float a[RA];
  float b[RB];
};
// per-thread row constants of the wgrad loaders (rows are fixed for the whole kernel, only (b, l) moves)
template <int BN>
struct WgradRows {
  int xoff[WgradRegs<BN>::RB];   // c * Lx
  int pj[WgradRegs<BN>::RB];     // j * j_mul + off   (very negative for rows beyond Nw)
  int na;                        // number of valid A rows of this warp
};

template <int BN>
__device__ __forceinline__ void load_wgrad(const WgradParams& p, int kb, WgradRegs<BN>& rg, const WgradRows<BN>& rows,
                                           int m0, int pwarp, int lane) {
  const vqs_wgrad_desc& d = p.d;
  const int kk = kb * BKF + lane;           // lane = k index within the block: global reads run along l (coalesced)
  const bool k_ok = kk < p.Kred;
  uint32_t b = 0, l = 0;
  if (k_ok) p.divLa.divmod((uint32_t)kk, b, l);
  const float* ab = d.Aact + ((size_t)b * d.M + m0 + pwarp) * d.La + l;
  const float* xb = d.X + ((size_t)b * d.Cred) * d.Lx;
  const int lp = (int)l * d.l_mul;
  const size_t astep = (size_t)N_PROD_WARPS * d.La;
#pragma unroll
  for (int i = 0; i < WgradRegs<BN>::RA; ++i) rg.a[i] = (k_ok && i < rows.na) ? __ldg(ab + i * astep) : 0.f;
#pragma unroll
  for (int i = 0; i < WgradRegs<BN>::RB; ++i) {
    const int pn = lp + rows.pj[i];
    rg.b[i] = (k_ok && pn >= 0 && pn < d.Lx) ? __ldg(xb + rows.xoff[i] + pn) : 0.f;
  }
}

// off0 = sw128_off(warp, lane); rows step by 16 so (r & 7) and with it the swizzled chunk stay fixed: +2048 B per row step
template <int BN, int PASSES>
__device__ __forceinline__ void store_wgrad(const WgradParams& p, const WgradRegs<BN>& rg, uint32_t sA_hi,
                                            uint32_t sA_lo, uint32_t sB_hi, uint32_t sB_lo, uint32_t off0) {
#pragma unroll
  for (int i = 0; i < WgradRegs<BN>::RA; ++i) {
    const uint32_t o = off0 + (uint32_t)(i * N_PROD_WARPS * 128);
    st_elem<PASSES>(sA_hi + o, sA_lo + o, rg.a[i]);
  }
  const bool rl = p.d.x_relu != 0;
#pragma unroll
  for (int i = 0; i < WgradRegs<BN>::RB; ++i) {
    const uint32_t o = off0 + (uint32_t)(i * N_PROD_WARPS * 128);
    st_elem<PASSES>(sB_hi + o, sB_lo + o, rl ? fmaxf(rg.b[i], 0.f) : rg.b[i]);
  }
}

__device__ __forceinline__ bool tc_mask_on(const void* m, int kind, size_t i) {
  if (kind == 1) return reinterpret_cast<const float*>(m)[i] > 0.f;
  return reinterpret_cast<const unsigned char*>(m)[i] != 0;
}

// ---------------------------------------------------------------------------------------------------
// MMA issue loop, executed by a whole converged warp with the MMAs guarded by elect_one().  role 0: single-pass (one
// accumulator); role 1: main term of the 3xTF32 split; role 2: its correction terms.  The stage ring is unrolled: k-block i0 + j lives in stage j, so all descriptors are base + constant
// and the round-robin accumulator of a k-step is a compile-time constant.
// ---------------------------------------------------------------------------------------------------
template <int PAIR>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (PAIR) umma_tf32_pair(tmem_d, a, b, idesc, acc);
  else umma_tf32(tmem_d, a, b, idesc, acc);
}
template <int PAIR>
__device__ __forceinline__ void umma_done(uint64_t* bar) {
  if (PAIR) umma_commit_pair(bar);
  else umma_commit(bar);
}

template <int BN, int PASSES, int ROLE, int PAIR>
__device__ __forceinline__ void issue_mmas(TcShared* sh, uint32_t smem_base, uint32_t tmem_base, int nkb) {
  using Cfg = TcCfg<BN, PASSES, PAIR>;
  constexpr uint32_t idesc = make_idesc_tf32(BN, PAIR ? 256 : 128);
  constexpr int S = Cfg::STAGES;
  const uint64_t d0 = make_desc_sw128(smem_base);
  const bool elected = elect_one();
  uint32_t par = 0;
  [[maybe_unused]] long long acc_full = 0;
#pragma unroll 1
  for (int i0 = 0; i0 < nkb; i0 += S) {
    const uint32_t nz = i0 > 0 ? 1u : 0u;      // accumulate flag of the first MMA into each accumulator
    // k-step t of the CTA goes to main accumulator 1 + t % 3 whatever the ring depth (S * 4 need not be a multiple of 3)
    const uint32_t rot = (uint32_t)(i0 * (BKF / 8)) % 3u;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      if (i0 + j < nkb) {
        {
          TC_ACC_BEGIN();
          mbar_wait(&sh->full[j], par);
          TC_ACC_END(acc_full);
        }
        tc_fence_after();
        if (ROLE != 2 && i0 == 0 && j == 0 && elected) TC_STAMP(1);
        if (elected && i0 + j == 40) TC_STAMP(ROLE == 2 ? 25 : 23);
        if (elected && i0 + j == 41) TC_STAMP(ROLE == 2 ? 28 : 27);
        const uint64_t a_hi = d0 + (uint64_t)((j * Cfg::STAGE_BYTES) >> 4);
        const uint64_t b_hi = a_hi + (uint64_t)(Cfg::A_BYTES >> 4);
        const uint64_t a_lo = a_hi + (uint64_t)((Cfg::A_BYTES + Cfg::B_BYTES) >> 4);
        const uint64_t b_lo = a_lo + (uint64_t)(Cfg::A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < BKF / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);  // +32 bytes per K = 8 step inside the swizzle atom
          const int g = j * (BKF / 8) + k;                    // k-step within the unrolled ring pass
          if (!elected) continue;
          if (PASSES == 3) {
            if (ROLE == 1) {
              // main term: accumulators 1..3 round-robin (the tensor core truncates when it accumulates: one accumulator
              // drifts 1.8e-5 relative over 864 MMAs, see the file header)
              uint32_t acc = (uint32_t)(g % 3) + rot;
              acc = acc >= 3u ? acc - 3u : acc;
              umma_issue<PAIR>(tmem_base + (1u + acc) * (uint32_t)BN, a_hi + adv, b_hi + adv, idesc, g < 3 ? nz : 1u);
            } else {
              umma_issue<PAIR>(tmem_base, a_lo + adv, b_hi + adv, idesc, g == 0 ? nz : 1u);
              umma_issue<PAIR>(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
            }
          } else {
            umma_issue<PAIR>(tmem_base, a_hi + adv, b_hi + adv, idesc, g == 0 ? nz : 1u);
          }
        }
        if (elected) umma_done<PAIR>(&sh->empty[j]);  // frees the smem slot (of both CTAs of a pair) once this thread's MMAs have read it
        if (elected && i0 + j == 40) TC_STAMP(ROLE == 2 ? 26 : 24);
        __syncwarp();
      }
    }
    par ^= 1u;
  }
  if (ROLE != 2 && elected) TC_STAMP(2);
  if (elected) TC_ACC_PUT(ROLE == 2 ? 15 : 14, acc_full);
  if (elected) umma_done<PAIR>(&sh->tmem_full);   // this thread's share of the accumulators is complete
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
template <int MODE, int BN, int PASSES, int KSZ, int PAIR>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const TcParams<MODE> prm) {
  using Cfg = TcCfg<BN, PASSES, PAIR>;
  constexpr int BNL = Cfg::BNL;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader (issues the MMAs), 1 = peer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  TcShared* sh = reinterpret_cast<TcShared*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    TC_STAMP(5);
    TC_STAMP_NS(6);
  }
  // pairs are the CTAs (2i, 2i + 1) of grid.x (cta_group::2 wants the pair adjacent in x): grid = (M tiles, N tiles, splits)
  const int tile_m = PAIR ? blockIdx.x : blockIdx.y, tile_n = PAIR ? blockIdx.y : blockIdx.x;
  const int m0 = tile_m * BM, n0 = tile_n * BN;

  bool a_image = false;
  if constexpr (MODE == 0) a_image = prm.p.d.a_tap_major == 2;
  int nkb, kb_begin;
  if constexpr (MODE == 0) {
    const int total = (prm.p.Ktot + BKF - 1) / BKF;
    kb_begin = prm.p.splits > 1 ? blockIdx.z * prm.p.kt_per_split : 0;
    int kb_end = prm.p.splits > 1 ? kb_begin + prm.p.kt_per_split : total;
    if (kb_end > total) kb_end = total;
    nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;
  } else {
    const int total = (prm.p.Kred + BKF - 1) / BKF;
    kb_begin = blockIdx.z * prm.p.kt_per_split;
    int kb_end = kb_begin + prm.p.kt_per_split;
    if (kb_end > total) kb_end = total;
    nkb = kb_end - kb_begin;
    if (nkb < 0) nkb = 0;
  }

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      // (pair: the leader's full[] also waits for one arrival of the peer's relay warp = "the peer's half is staged")
      mbar_init(&sh->full[s], (PAIR ? N_PROD_WARPS / Cfg::STAGES : N_PROD_WARPS) + (a_image ? 1 : 0) + ((PAIR && rank == 0) ? 1 : 0));
      mbar_init(&sh->empty[s], Cfg::N_ISSUERS);
    }
    mbar_init(&sh->tmem_full, Cfg::N_ISSUERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == N_PROD_WARPS) {
    if (PAIR) tmem_alloc_pair(&sh->tmem_base, Cfg::TMEM_COLS);
    else tmem_alloc(&sh->tmem_base, Cfg::TMEM_COLS);
  }
  if constexpr (PAIR && MODE == 1) {
    if (tid < BNL) {
      const int nrow = n0 + (int)rank * BNL + tid;
      const int c = nrow / KSZ, j = nrow - c * KSZ;
      sh->btab[tid] = make_int2(c * prm.p.d.Lx, nrow < prm.p.Nw ? j * prm.p.d.j_mul + prm.p.d.off : -(1 << 29));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  pdl_prologue_done();
  if (tid == 0) TC_STAMP(0);

  constexpr bool GROUPS = PAIR != 0;           // one group of producer warps per ring stage (CTA pairs)
  if constexpr (GROUPS) {
  if (warp < N_PROD_WARPS) {
    // ================= producers of a CTA pair: one group of warps per ring stage =================
    // A thread that fills EVERY stage spends ~1000 cycles per k-block on serially dependent work (index arithmetic, waiting
    // for an empty slot, stores, fence.proxy.async, arrive: measured with -DVQS_GEMM_TIMING) against 768 cycles of MMA
    // work -- the producers' critical path, not their throughput, paced the ring.  Here group g = warp / 4 owns stage g and
    // handles k-blocks g, g + 4, ...: four times the data per visit, a quarter of the visits, so the fixed costs are paid once
    // per four k-blocks and the loads of the group's next k-block have ~3 k-block periods to land (one register buffer).
    // Conv: a thread owns one B row (n) and 16 of the 32 channels: a warp's load covers 32 consecutive positions of one
    // channel (whole 128-byte lines), its 16-byte stores hit each bank group from exactly four lanes (the minimum).
    constexpr int S = Cfg::STAGES;
    constexpr int GW = N_PROD_WARPS / S, GT = GW * 32;
    static_assert(!PAIR || (N_PROD_WARPS % S == 0 && GW == 4), "one group of four producer warps per stage");
    const int g = warp / GW, gt = tid - g * GT;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sA_hi = smem_base + (uint32_t)(g * Cfg::STAGE_BYTES);
    const uint32_t sB_hi = sA_hi + Cfg::A_BYTES;
    const uint32_t sA_lo = sA_hi + Cfg::A_BYTES + Cfg::B_BYTES;
    const uint32_t sB_lo = sA_lo + Cfg::A_BYTES;
    [[maybe_unused]] long long acc_empty = 0, acc_store = 0, acc_fence = 0;
    auto wait_empty = [&](int i) {
      TC_ACC_BEGIN();
      mbar_wait(&sh->empty[g], ((uint32_t)(i / S) & 1u) ^ 1u);
      TC_ACC_END(acc_empty);
    };
    auto publish = [&]() {
      TC_ACC_BEGIN();
      fence_proxy_async();  // this thread's generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full[g]);          // one arrival per warp of the group
      TC_ACC_END(acc_fence);
    };
    if constexpr (MODE == 0) {
      constexpr int PARTS = GT / BNL;          // threads per B row: 2
      constexpr int CPT = BKF / PARTS;         // channels per thread: 16
      static_assert(GT % BNL == 0 && BKF % PARTS == 0 && CPT % 4 == 0, "B row split");
      const vqs_conv_gemm_desc& d = prm.p.d;
      const int brow = gt % BNL, part = gt / BNL;
      const int n = n0 + (int)rank * BNL + brow;
      const bool n_ok = n < prm.p.Ntot;
      uint32_t bb = 0, ll = 0;
      if (n_ok) prm.p.divL.divmod((uint32_t)n, bb, ll);
      const float* xb = d.X + (long long)bb * d.x_sb + (long long)(part * CPT) * d.x_sc;
      const int lbase = (int)ll * d.l_mul + d.off;
      uint32_t offB[CPT / 4];
#pragma unroll
      for (int ci = 0; ci < CPT / 4; ++ci) offB[ci] = (uint32_t)(brow * 128 + (((part * (CPT / 4) + ci) ^ (brow & 7)) << 4));
      const int nkb_all = prm.p.Ktot / BKF;
      const float* ablk = d.A + ((size_t)tile_m * nkb_all + kb_begin) * 8192;
      // without an image (a_tap_major = 1: plain fp32 matrix [M][ksz][Cred]) the group loads and splits the 128 x 32 block of A
      // itself: half the L2 -> SM bytes of the image (hi and lo copies), and with both operands at 43 KB per k-block and
      // CTA the image-fed pair sat on the L2 slices' output limit (~42 B/clk per SM with every SM asking)
      constexpr int AI = BM * 8 / GT;            // 16-byte chunks of A per thread: 8
      const int arow0 = gt >> 3, acol = gt & 7;  // 8 consecutive threads read one 128-byte row segment
      const uint32_t offA = (uint32_t)(arow0 * 128 + ((acol ^ (arow0 & 7)) << 4));   // rows step by 16: (r & 7) fixed
      const float* arow = d.A + (size_t)(m0 + arow0) * prm.p.Ktot + (size_t)kb_begin * BKF + acol * 4;
      const bool rl = d.x_relu != 0;
      float v[CPT];
      float4 va[AI];
      auto load = [&](int i) {
        if (!a_image) {
#pragma unroll
          for (int r = 0; r < AI; ++r)
            va[r] = (m0 + arow0 + r * (GT / 8) < d.M)
                        ? __ldg(reinterpret_cast<const float4*>(arow + (size_t)r * (GT / 8) * prm.p.Ktot + (size_t)i * BKF))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int kb = kb_begin + i;
        const uint32_t j = prm.p.divCpb.div((uint32_t)kb);     // tap-major K: k-block = tap j, channels [c0, c0 + 32)
        const int c0 = (kb - (int)j * prm.p.cpb) * BKF;
        int pn = lbase + (int)j * d.j_mul;
        bool ok = n_ok && pn >= 0;
        if (d.l_div == 2) {
          ok = ok && ((pn & 1) == 0);
          pn >>= 1;
        }
        ok = ok && pn < d.Lin;
        const float* ptr = xb + (long long)pn * d.x_sl + (long long)c0 * d.x_sc;
        const int cleft = prm.p.cred_real - (c0 + part * CPT);   // channels of X from this thread's first one on
#pragma unroll
        for (int e = 0; e < CPT; ++e) v[e] = (ok && e < cleft) ? __ldg(ptr + (long long)e * d.x_sc) : 0.f;
      };
      if (g < nkb) load(g);
      for (int i = g; i < nkb; i += S) {
        wait_empty(i);
        if (gt == 0 && i == 40) TC_STAMP(20);
        if (lane == 0 && i == 40) TC_STAMP(36 + (warp & 3));
        if (gt == 0 && i == 40) TC_STAMP_PEER(29);
        if (gt == 0 && i == 44) TC_STAMP(22);
        TC_ACC_BEGIN();
        if (!a_image) {
#pragma unroll
          for (int r = 0; r < AI; ++r) {
            const uint32_t o = offA + (uint32_t)(r * (GT / 8) * 128);
            st_vec4<PASSES>(sA_hi + o, sA_lo + o, va[r]);
          }
        } else if (gt == 0) {
          // A operand: one bulk copy per copy (hi, lo) of the pre-built 128 x 32 image block of this k-block; the copy
          // engine signals full[g] with complete_tx (async proxy: no fence needed), this thread adds the extra arrival
          const float* blk = ablk + (size_t)i * 8192;
          const uint32_t bar = smem_u32(&sh->full[g]);
          constexpr uint32_t bytes = (PASSES == 3) ? 2u * Cfg::A_BYTES : (uint32_t)Cfg::A_BYTES;
          asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
                       "r"(bytes)
                       : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           sA_hi),
                       "l"(blk), "r"((uint32_t)Cfg::A_BYTES), "r"(bar)
                       : "memory");
          if (PASSES == 3)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             sA_lo),
                         "l"(blk + 4096), "r"((uint32_t)Cfg::A_BYTES), "r"(bar)
                         : "memory");
        }
#pragma unroll
        for (int ci = 0; ci < CPT / 4; ++ci) {
          float4 q = make_float4(v[ci * 4 + 0], v[ci * 4 + 1], v[ci * 4 + 2], v[ci * 4 + 3]);
          if (rl) q = make_float4(fmaxf(q.x, 0.f), fmaxf(q.y, 0.f), fmaxf(q.z, 0.f), fmaxf(q.w, 0.f));
          st_vec4<PASSES>(sB_hi + offB[ci], sB_lo + offB[ci], q);
        }
        TC_ACC_END(acc_store);
        if (lane == 0 && i == 40) TC_STAMP(40 + (warp & 3));
        publish();
        if (gt == 0 && i == 40) TC_STAMP(21);
        if (lane == 0 && i == 40) TC_STAMP(32 + (warp & 3));
        if (gt == 0 && i == 40) TC_STAMP_PEER(30);
        if (i + S < nkb) load(i + S);
      }
    } else {
      // wgrad: lane = position inside the k-block (global reads run along l: whole lines), warp wg of the group owns rows
      // wg, wg + 4, ... of both tiles.  Per element: one address, one load, the split, two stores with immediate offsets.
      constexpr int RA = BM / GW, RB = BNL / GW;     // 32 rows of A, 16 of B per thread
      const vqs_wgrad_desc& d = prm.p.d;
      const int wg = warp - g * GW;
      // rows step by 4: (row & 7) alternates between two values, two rows further on the offset grows by 1024 B
      const uint32_t off_e = sw128_off(wg, lane), off_o = sw128_off(wg + 4, lane);
      const int na = (d.M - m0 - wg + GW - 1) / GW;  // A rows of this warp inside the matrix (<= 0: none)
      const uint32_t btab = smem_u32(&sh->btab[wg]);
      const uint32_t astep = (uint32_t)(GW * d.La);
      const uint32_t Lx = (uint32_t)d.Lx;
      const bool rl = d.x_relu != 0;
      float va[RA], vb[RB];
      auto load = [&](int i) {
        const int kk = (kb_begin + i) * BKF + lane;
        const bool k_ok = kk < prm.p.Kred;
        uint32_t b = 0, l = 0;
        if (k_ok) prm.p.divLa.divmod((uint32_t)kk, b, l);
        const float* ab = d.Aact + ((size_t)b * d.M + m0 + wg) * d.La + l;
        const float* xb = d.X + ((size_t)b * d.Cred) * d.Lx;
        const int lp = k_ok ? (int)l * d.l_mul : -(1 << 29);   // (an invalid k never passes the bounds test below)
#pragma unroll
        for (int r = 0; r < RA; ++r) va[r] = (k_ok && r < na) ? __ldg(ab + r * astep) : 0.f;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          int xoff, pj;
          asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(xoff), "=r"(pj) : "r"(btab + (uint32_t)(r * GW * 8)));
          const int pn = lp + pj;
          vb[r] = ((uint32_t)pn < Lx) ? __ldg(xb + (uint32_t)(xoff + pn)) : 0.f;
        }
      };
      if (g < nkb) load(g);
      for (int i = g; i < nkb; i += S) {
        wait_empty(i);
        TC_ACC_BEGIN();
#pragma unroll
        for (int r = 0; r < RA; ++r) {
          const uint32_t o = ((r & 1) ? off_o : off_e) + (uint32_t)((r >> 1) * 1024);
          st_elem<PASSES>(sA_hi + o, sA_lo + o, va[r]);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const uint32_t o = ((r & 1) ? off_o : off_e) + (uint32_t)((r >> 1) * 1024);
          st_elem<PASSES>(sB_hi + o, sB_lo + o, rl ? fmaxf(vb[r], 0.f) : vb[r]);
        }
        TC_ACC_END(acc_store);
        publish();
        if (i + S < nkb) load(i + S);
      }
    }
    if (tid == 0) {
      TC_ACC_PUT(16, acc_empty);
      TC_ACC_PUT(17, acc_store);
      TC_ACC_PUT(18, acc_fence);
      TC_STAMP(19);
    }
  }
  } else {
  if (warp < N_PROD_WARPS) {
    // ================= producers =================
    const int ptid = tid;
    // per-thread constants of the conv B loader
    bool n_ok = false;
    const float* xb = nullptr;
    int lbase = 0, brow = 0, bk0 = 0;
    if constexpr (MODE == 0) {
      constexpr int TPR = N_PROD / BNL;
      brow = ptid / TPR;
      bk0 = (ptid % TPR) * (8 / TPR);
      const int n = n0 + (int)rank * BNL + brow;
      n_ok = n < prm.p.Ntot;
      uint32_t bb = 0, ll = 0;
      if (n_ok) prm.p.divL.divmod((uint32_t)n, bb, ll);
      xb = prm.p.d.X + (long long)bb * prm.p.d.x_sb;
      lbase = (int)ll * prm.p.d.l_mul + prm.p.d.off;
    }
    using Regs = typename std::conditional<MODE == 0, ConvRegs<BNL>, WgradRegs<BNL>>::type;
    WgradRows<BNL> rows;
    if constexpr (MODE == 1) {
      const vqs_wgrad_desc& d = prm.p.d;
#pragma unroll
      for (int i = 0; i < WgradRegs<BNL>::RB; ++i) {
        const int n = n0 + (int)rank * BNL + warp + i * N_PROD_WARPS;
        const int c = n / KSZ, j = n - c * KSZ;
        rows.xoff[i] = c * d.Lx;
        rows.pj[i] = (n < prm.p.Nw) ? j * d.j_mul + d.off : -(1 << 29);
      }
      int na = (d.M - m0 - warp + N_PROD_WARPS - 1) / N_PROD_WARPS;
      rows.na = na < 0 ? 0 : na;
    }
    auto load = [&](int kb, Regs& rg) {
      if constexpr (MODE == 0) load_conv<BNL>(prm.p, kb, rg, m0, ptid, n_ok, xb, lbase, bk0);
      else load_wgrad<BNL>(prm.p, kb, rg, rows, m0, warp, lane);
    };
    // Register ring of depth 3: while k-block i is stored, the loads of i+1 and i+2 are in flight.  With ~1.2k cycles of
    // loaded-L2 latency and 32 KB of operands per k-block, one block in flight caps the SM at ~27 B/clk (measured: 2.8k
    // cycles per k-block); two blocks in flight cover the 768-cycle MMA time of a 3xTF32 k-block.
    const uint32_t smem_base = smem_u32(smem);
    uint32_t offA = 0, off0 = 0;
    uint32_t offB[ConvRegs<BNL>::CH];
    if constexpr (MODE == 0) {
      const int r0a = ptid >> 3, ca = ptid & 7;
      offA = (uint32_t)(r0a * 128 + ((ca ^ (r0a & 7)) << 4));
#pragma unroll
      for (int ci = 0; ci < ConvRegs<BNL>::CH; ++ci) offB[ci] = (uint32_t)(brow * 128 + (((bk0 + ci) ^ (brow & 7)) << 4));
    } else {
      off0 = sw128_off(warp, lane);
#pragma unroll
      for (int ci = 0; ci < ConvRegs<BNL>::CH; ++ci) offB[ci] = 0;
    }
    [[maybe_unused]] long long acc_empty = 0, acc_store = 0, acc_fence = 0;
    auto stage = [&](int i, const Regs& rg) {
      const int s = i % Cfg::STAGES;
      const uint32_t ph = (uint32_t)(i / Cfg::STAGES) & 1u;
      {
        TC_ACC_BEGIN();
        mbar_wait(&sh->empty[s], ph ^ 1u);
        TC_ACC_END(acc_empty);
      }
      TC_ACC_BEGIN();
      const uint32_t sA_hi = smem_base + (uint32_t)(s * Cfg::STAGE_BYTES);
      const uint32_t sB_hi = sA_hi + Cfg::A_BYTES;
      const uint32_t sA_lo = sA_hi + Cfg::A_BYTES + Cfg::B_BYTES;
      const uint32_t sB_lo = sA_lo + Cfg::A_BYTES;
      if constexpr (MODE == 0) {
        if (a_image && tid == 0) {
          // A operand: one bulk copy per copy (hi, lo) of the pre-built 128 x 32 image block of this k-block; the copy
          // engine signals full[s] with complete_tx (async proxy: no fence needed), this thread adds the extra arrival
          const int nkb_all = prm.p.Ktot / BKF;
          const float* blk = prm.p.d.A + ((size_t)tile_m * nkb_all + (kb_begin + i)) * 8192;
          const uint32_t bar = smem_u32(&sh->full[s]);
          constexpr uint32_t bytes = (PASSES == 3) ? 2u * Cfg::A_BYTES : (uint32_t)Cfg::A_BYTES;
          asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
                       "r"(bytes)
                       : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           sA_hi),
                       "l"(blk), "r"((uint32_t)Cfg::A_BYTES), "r"(bar)
                       : "memory");
          if (PASSES == 3)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             sA_lo),
                         "l"(blk + 4096), "r"((uint32_t)Cfg::A_BYTES), "r"(bar)
                         : "memory");
        }
        store_conv<BNL, PASSES>(prm.p, rg, sA_hi, sA_lo, sB_hi, sB_lo, offA, offB);
      } else {
        store_wgrad<BNL, PASSES>(prm.p, rg, sA_hi, sA_lo, sB_hi, sB_lo, off0);
      }
      TC_ACC_END(acc_store);
      {
        TC_ACC_BEGIN();
        fence_proxy_async();  // this thread's generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->full[s]);          // one arrival per producer warp
        TC_ACC_END(acc_fence);
      }
    };
    Regs r0, r1, r2;
    if (nkb > 0) load(kb_begin, r0);
    if (nkb > 1) load(kb_begin + 1, r1);
    for (int i = 0; i < nkb; i += 3) {
      if (i + 2 < nkb) load(kb_begin + i + 2, r2);
      stage(i, r0);
      if (i + 1 >= nkb) break;
      if (i + 3 < nkb) load(kb_begin + i + 3, r0);
      stage(i + 1, r1);
      if (i + 2 >= nkb) break;
      if (i + 4 < nkb) load(kb_begin + i + 4, r1);
      stage(i + 2, r2);
    }
    if (tid == 0) {
      TC_ACC_PUT(16, acc_empty);
      TC_ACC_PUT(17, acc_store);
      TC_ACC_PUT(18, acc_fence);
      TC_STAMP(19);
    }
  }
  }
  if (warp >= N_PROD_WARPS) {
    // ================= MMA issuers =================
    // Measured with profiles/mma_rate.cu: the tensor core retires a 128 x 128 x 8 TF32 MMA every 64 cycles, but the round-1
    // issue loop (under `if (lane == 0)`: an ELECT / BRA.U.ANY loop per MMA, descriptors and the accumulator rotation
    // computed per k-block at run time, ~190 serially dependent instructions for 12 MMAs) needed well over the 768 cycles
    // of MMA time per k-block.  Now: whole converged warps run the loop with the MMAs guarded by elect_one() (bare
    // UTCHMMA), the ring is unrolled over all stages so that every descriptor is `base + constant`, and the 3xTF32 split
    // is issued by two warps with
    // DISJOINT accumulators (warp 16: main term Ah*Bh into accumulators 1..3 round-robin; warp 17: correction terms
    // Al*Bh + Ah*Bl into accumulator 0), so the order of additions into every accumulator stays fixed (deterministic) while
    // the issue rate doubles.  Both commit to empty[s] / tmem_full (barrier count = 2).
    if (PAIR && rank != 0) {
      // peer CTA of a pair: no MMAs; warp 16 relays "stage s of this CTA is full" to the leader's full[s]
      if (warp == N_PROD_WARPS) {
        for (int i = 0; i < nkb; ++i) {
          const int s = i % Cfg::STAGES;
          mbar_wait(&sh->full[s], (uint32_t)(i / Cfg::STAGES) & 1u);
          if (lane == 0 && i == 40) TC_STAMP_PEER(31);
          if (lane == 0) mbar_arrive_cluster(&sh->full[s], 0);
          __syncwarp();
        }
      }
    } else if (PASSES == 3) {
      if (warp == N_PROD_WARPS) issue_mmas<BN, PASSES, 1, PAIR>(sh, smem_u32(smem), tmem_base, nkb);
      else issue_mmas<BN, PASSES, 2, PAIR>(sh, smem_u32(smem), tmem_base, nkb);
    } else if (warp == N_PROD_WARPS) {
      issue_mmas<BN, PASSES, 0, PAIR>(sh, smem_u32(smem), tmem_base, nkb);
    }
  }

  // ================= epilogue: every producer warp takes part =================
  // TMEM lane quarter = warp % 4 (hardware rule); the BN/32 column blocks are spread over warp / 4, so the 16 producer
  // warps drain a 128 x 128 tile four times faster than 4 warps could (the epilogue with residual / mask operands was a
  // third of the dgrad kernels' time: 0.158 vs 0.103 ms for the same GEMM, profiles/r01j).
  if (warp < N_PROD_WARPS && (warp >> 2) < BN / 32) {
    const int q = warp & 3;                 // TMEM lane quarter -> rows m0 + 32 q .. + 31
    float* stg = reinterpret_cast<float*>(smem) + warp * (32 * 33);  // stage buffers are idle by now
    // lane r keeps the bias of row m0 + 32 q + r (fetched before the accumulators are waited for; the rows take it by
    // shuffle: a load per row inside the store loop cost one exposed L2 latency per ROW -- 19 k of the kernel's 115 k cycles)
    float bias_r = 0.f;
    if constexpr (MODE == 0) {
      const int mr = m0 + q * 32 + lane;
      if (prm.p.d.bias != nullptr && prm.p.partial == nullptr && mr < prm.p.d.M) bias_r = __ldg(prm.p.d.bias + mr);
    }
    if (nkb > 0) {
      mbar_wait(&sh->tmem_full, 0);
      tc_fence_after();
    }
    if (tid == 0) TC_STAMP(3);
    int Ncols;
    if constexpr (MODE == 0) Ncols = prm.p.Ntot;
    else Ncols = prm.p.Nw;
    constexpr int BLK_STEP = N_PROD_WARPS / 4;
#pragma unroll 1
    for (int blk = warp >> 2; blk < BN / 32; blk += BLK_STEP) {
      float v[32];
      if (nkb > 0) {
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(blk * 32);
        if (Cfg::NACC == 4) {
          float t1[32];
          tmem_ld32(ta + 1 * BN, v);
          tmem_ld32(ta + 2 * BN, t1);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += t1[j];
          tmem_ld32(ta + 3 * BN, t1);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += t1[j];
          tmem_ld32(ta, t1);
          // main term: undo the expected truncation loss of its (4 nkb / 3) MMAs per accumulator, then the correction terms
          // (a stride-2 dgrad, l_div = 2, reads a zero operand for every other tap: half of its MMAs add nothing)
          int n_mma = (4 * nkb + 2) / 3;
          if constexpr (MODE == 0) n_mma /= prm.p.d.l_div;
          const float comp = tf32x3_comp(n_mma);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], comp, t1[j]);
        } else {
          tmem_ld32(ta, v);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (tid == 0) TC_STAMP(8);
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
      __syncwarp();
      if (tid == 0) TC_STAMP(9);
      const int n = n0 + blk * 32 + lane;
      const bool n_ok = n < Ncols;
      if constexpr (MODE == 0) {
        const vqs_conv_gemm_desc& d = prm.p.d;
        if (prm.p.partial != nullptr) {      // split-K: raw accumulators, coalesced along n
          float* po = prm.p.partial + (size_t)blockIdx.z * d.M * prm.p.Ntot;
          for (int r = 0; r < 32; ++r) {
            const int m = m0 + q * 32 + r;
            if (m >= d.M) break;
            if (n_ok) po[(size_t)m * prm.p.Ntot + n] = stg[r * 33 + lane];
          }
          __syncwarp();
          continue;
        }
        uint32_t b = 0, l = 0;
        if (n_ok) prm.p.divL.divmod((uint32_t)n, b, l);
        const size_t col = (size_t)b * d.M * d.Lout + l;
        // Every tensor of the epilogue is addressed as (per-lane base) + (row offset, 32 bits); the optional operands are
        // warp-uniform branches around whole batches of 8 rows, so that a layer pays only for what it uses (the first version
        // evaluated all of them as predicated-off code with 64-bit index arithmetic per row and tensor: ~800 instructions
        // per batch, 16 k cycles per tile = a fifth of the kernel; measured with -DVQS_GEMM_TIMING).
        const int mrow0 = m0 + q * 32;
        const int rows_here = d.M - mrow0 < 32 ? d.M - mrow0 : 32;      // rows of this warp inside the matrix (<= 0: none)
        float* const outp = d.out + col;
        const bool has_pre = d.add_pre != nullptr, has_post = d.add_post != nullptr, has_mask = d.mask_kind != 0;
        const bool has_out2 = d.out2 != nullptr, has_mask2 = has_out2 && d.mask2_kind != 0, pre_relu = d.add_pre_relu != 0;
        const bool relu = d.relu != 0, has_mask_out = d.mask_out != nullptr;
        const uint32_t Lo = (uint32_t)d.Lout;
        if (!(has_pre || has_post || has_mask || has_out2)) {
          // bias (+ ReLU, + mask_out): one shared-memory read, one shuffle and one store per row
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const float bias_m = __shfl_sync(0xffffffffu, bias_r, r);
            float v1 = stg[r * 33 + lane] + bias_m;
            if (relu) v1 = fmaxf(v1, 0.f);
            if (r < rows_here && n_ok) {
              const uint32_t ro = (uint32_t)(mrow0 + r) * Lo;
              if (has_mask_out) d.mask_out[col + ro] = v1 > 0.f ? 1 : 0;
              outp[ro] = v1;
            }
          }
        } else {
          // 8 rows at a time: all global reads of the batch are issued before the first use (the serial version paid one
          // load latency per row: +50 us per 768x768x3 layer, profiles/r01c)
#pragma unroll 1
          for (int rb = 0; rb < 32; rb += 8) {
            if (rb >= rows_here) break;
            float x[8], pre[8], post[8];
            bool k1[8], k2[8];
            bool live[8];
            uint32_t ro[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              live[q] = n_ok && (rb + q < rows_here);
              ro[q] = (uint32_t)(mrow0 + rb + q) * Lo;
              x[q] = stg[(rb + q) * 33 + lane];
              pre[q] = 0.f;
              post[q] = 0.f;
              k1[q] = true;
              k2[q] = true;
            }
            if (has_pre) {
              const float* pp = d.add_pre + col;
#pragma unroll
              for (int q = 0; q < 8; ++q) if (live[q]) pre[q] = pp[ro[q]];
            }
            if (has_post) {
              const float* pp = d.add_post + col;
#pragma unroll
              for (int q = 0; q < 8; ++q) if (live[q]) post[q] = pp[ro[q]];
            }
            if (has_mask) {
              if (d.mask_kind == 1) {
                const float* pp = reinterpret_cast<const float*>(d.mask) + col;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (live[q]) k1[q] = pp[ro[q]] > 0.f;
              } else {
                const unsigned char* pp = reinterpret_cast<const unsigned char*>(d.mask) + col;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (live[q]) k1[q] = pp[ro[q]] != 0;
              }
            }
            if (has_mask2) {
              if (d.mask2_kind == 1) {
                const float* pp = reinterpret_cast<const float*>(d.mask2) + col;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (live[q]) k2[q] = pp[ro[q]] > 0.f;
              } else {
                const unsigned char* pp = reinterpret_cast<const unsigned char*>(d.mask2) + col;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (live[q]) k2[q] = pp[ro[q]] != 0;
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float bias_m = __shfl_sync(0xffffffffu, bias_r, rb + q);   // (all lanes: before the per-lane skip)
              float v1 = x[q] + bias_m;
              v1 += pre_relu ? fmaxf(pre[q], 0.f) : pre[q];
              if (relu) v1 = fmaxf(v1, 0.f);
              const bool pos = v1 > 0.f;
              if (!k1[q]) v1 = 0.f;
              v1 += post[q];
              if (live[q]) {
                if (has_mask_out) d.mask_out[col + ro[q]] = pos ? 1 : 0;
                outp[ro[q]] = v1;
                if (has_out2) d.out2[col + ro[q]] = k2[q] ? v1 : 0.f;
              }
            }
            if (tid == 0) TC_STAMP(10 + (rb >> 3));
          }
        }
      } else {
        const vqs_wgrad_desc& d = prm.p.d;
        float* out = prm.p.partial ? prm.p.partial + (size_t)blockIdx.z * d.M * prm.p.Nw : d.dW;
        const bool accum = (prm.p.partial == nullptr) && d.accumulate;
        const int rows_here = d.M - (m0 + q * 32) < 32 ? d.M - (m0 + q * 32) : 32;
        float* const op = out + (size_t)(m0 + q * 32) * prm.p.Nw + n;
        if (n_ok) {
          if (accum) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
              if (r < rows_here) op[(size_t)r * prm.p.Nw] += stg[r * 33 + lane];
          } else {
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
              if (r < rows_here) op[(size_t)r * prm.p.Nw] = stg[r * 33 + lane];
          }
        }
      }
      __syncwarp();
    }
  }
  if (tid == 0) {
    TC_STAMP(4);
    TC_STAMP_NS(7);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();     // neither CTA leaves (or frees its half of the pair's TMEM) while the other still works
  if (warp == N_PROD_WARPS) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// CTA pairs (VQS_GEMM_PAIR=0 disables; read once): 128-column tiles whose grid has an even number of 128-row tiles
bool pair_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("VQS_GEMM_PAIR");
    cached = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return cached == 1;
}

template <int MODE, int BN, int PASSES, int KSZ, int PAIR>
int launch_tc_t(const TcParams<MODE>& prm, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN, PASSES, PAIR>;
  auto kern = gemm_tc_kernel<MODE, BN, PASSES, KSZ, PAIR>;
  static DevCache configured;  // per instantiation and device
  if (dev_needs(configured, Cfg::SMEM_BYTES))
    VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  if (PAIR) {
    cudaError_t e = launch_pdl_cluster(kern, dim3(grid.y, grid.x, grid.z), dim3(TC_THREADS), Cfg::SMEM_BYTES, st, 2, 1, prm);
    if (e != cudaSuccess) {
      cudaFuncAttributes fa = {};
      cudaFuncGetAttributes(&fa, kern);
      set_error("pair GEMM launch failed: %s (grid %u x %u x %u, %d threads, %d B dynamic + %zu B static smem, %d registers, "
                "max dynamic smem %d)", cudaGetErrorString(e), grid.x, grid.y, grid.z, TC_THREADS, Cfg::SMEM_BYTES,
                fa.sharedSizeBytes, fa.numRegs, fa.maxDynamicSharedSizeBytes);
      return (int)e;
    }
  } else {
    VQS_CUDA(launch_pdl(kern, grid, dim3(TC_THREADS), Cfg::SMEM_BYTES, st, prm));
  }
  VQS_LAUNCH_CHECK();
  return 0;
}

template <int MODE, int BN, int PASSES, int PAIR>
int launch_tc_k(const TcParams<MODE>& prm, int ksz, dim3 grid, cudaStream_t st) {
  // (wgrad rows are (c, j) pairs: the kernel divides by KSZ; the conv-like GEMM does not use it)
  if constexpr (MODE == 0) {
    return launch_tc_t<MODE, BN, PASSES, 1, PAIR>(prm, grid, st);
  } else {
    switch (ksz) {
      case 1: return launch_tc_t<MODE, BN, PASSES, 1, PAIR>(prm, grid, st);
      case 2: return launch_tc_t<MODE, BN, PASSES, 2, PAIR>(prm, grid, st);
      case 3: return launch_tc_t<MODE, BN, PASSES, 3, PAIR>(prm, grid, st);
      case 4: return launch_tc_t<MODE, BN, PASSES, 4, PAIR>(prm, grid, st);
    }
    set_error("tcgen05 GEMM: kernel size %d not supported (1..4)", ksz);
    return VQS_ERR_ARG;
  }
}

// Half-wave conv GEMMs with few k-blocks (the 1 x 1 layers at T_q = 24: 72 tiles of 128 columns, 24 k-blocks) as ONE wave of CTA
// pairs with 64-column tiles (VQS_CONV_BN64_PAIR=0: single 128 x 64 CTAs as before; read once)
bool conv_bn64_pair_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("VQS_CONV_BN64_PAIR");
    cached = (e != nullptr && atoi(e) == 0) ? 0 : 1;
  }
  return cached == 1;
}

template <int MODE>
int launch_tc(const TcParams<MODE>& prm, int ksz, int bn, int precision, dim3 grid, cudaStream_t st) {
  bool pair = (bn == 128 || bn == 96) && grid.y % 2 == 0 && pair_enabled();
  if constexpr (MODE == 0) pair = pair && prm.p.d.a_tap_major != 0;     // an image or the tap-major matrix (conv_tc_supported)
  if constexpr (MODE == 0) {
    if (bn == 64 && grid.z == 1 && grid.y % 2 == 0 && pair_enabled() && conv_bn64_pair_enabled() && prm.p.d.a_tap_major != 0) {
      if (precision == 2) return launch_tc_k<MODE, 64, 3, 1>(prm, ksz, grid, st);
      return launch_tc_k<MODE, 64, 1, 1>(prm, ksz, grid, st);
    }
  }
  if constexpr (MODE == 1) {
    if (bn == 96) {        // 96-column tiles (launch_wgrad_tc): pairs, 3xTF32 only
      if (!pair || precision != 2) {
        set_error("tcgen05 wgrad: 96-column tiles need the pair kernel and the 3xTF32 engine");
        return VQS_ERR_ARG;
      }
      return launch_tc_k<MODE, 96, 3, 1>(prm, ksz, grid, st);
    }
  }
  if (precision == 2) {
    if (pair) return launch_tc_k<MODE, 128, 3, 1>(prm, ksz, grid, st);
    if (bn == 128) return launch_tc_k<MODE, 128, 3, 0>(prm, ksz, grid, st);
    return launch_tc_k<MODE, 64, 3, 0>(prm, ksz, grid, st);
  }
  if (pair) return launch_tc_k<MODE, 128, 1, 1>(prm, ksz, grid, st);
  if (bn == 128) return launch_tc_k<MODE, 128, 1, 0>(prm, ksz, grid, st);
  return launch_tc_k<MODE, 64, 1, 0>(prm, ksz, grid, st);
}

}  // namespace

bool conv_tc_supported(const ConvParams& p) {
  if ((long long)p.d.M * p.d.Lout >= (1ll << 31)) return false;              // (32-bit row offsets in the epilogue)
  if (p.d.a_tap_major == 2) return p.d.Cred % BKF == 0 && p.d.ksz >= 1;       // pre-built operand image
  return p.a_vec && p.d.a_tap_major == 1 && p.d.Cred % BKF == 0 && p.d.ksz >= 1;
}
bool wgrad_tc_supported(const WgradParams& p) { return p.d.ksz >= 1 && p.d.ksz <= 4 && p.Kred >= 32; }
bool wgrad_tc_pairs(const WgradParams& p) { return pair_enabled() && ((p.d.M + BM - 1) / BM) % 2 == 0; }

// Folds the split-K partials partial[z][m][n = b * Lout + l] in a fixed order and applies the descriptor's fused epilogue
// (vqs_b200.h: bias, add_pre (+relu), relu, mask_out, mask, add_post, out, out2 / mask2), coalesced along l.
__global__ void __launch_bounds__(256) conv_splitk_epilogue_kernel(const float* __restrict__ partial, int splits,
                                                                   int Ntot, const vqs_conv_gemm_desc d) {
  pdl_prologue_done();
  const long long total = (long long)d.M * Ntot;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int m = (int)(i / Ntot), n = (int)(i - (long long)m * Ntot);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(size_t)z * total + i];
    const int b = n / d.Lout, l = n - b * d.Lout;
    const size_t o = ((size_t)b * d.M + m) * d.Lout + l;
    if (d.bias) v += __ldg(d.bias + m);
    if (d.add_pre) {
      const float pre = d.add_pre[o];
      v += d.add_pre_relu ? fmaxf(pre, 0.f) : pre;
    }
    if (d.relu) v = fmaxf(v, 0.f);
    if (d.mask_out) d.mask_out[o] = v > 0.f ? 1 : 0;
    if (d.mask_kind && !tc_mask_on(d.mask, d.mask_kind, o)) v = 0.f;
    if (d.add_post) v += d.add_post[o];
    d.out[o] = v;
    if (d.out2) d.out2[o] = (!d.mask2_kind || tc_mask_on(d.mask2, d.mask2_kind, o)) ? v : 0.f;
  }
}

// The same fold for Ntot, Lout multiples of 4 and 16-byte aligned tensors (every T_q = 24 / 48 layer): four consecutive l per
// thread, 32-bit index arithmetic with precomputed reciprocals (the scalar kernel's 64-bit division per element made it
// 15.8 us for a 768 x 1536 layer under ncu).  Same operation order per element as the scalar kernel.
__device__ __forceinline__ void tc_mask4(const void* m, int kind, size_t o, bool (&on)[4]) {
  if (kind == 1) {
    const float4 f = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(m) + o);
    on[0] = f.x > 0.f; on[1] = f.y > 0.f; on[2] = f.z > 0.f; on[3] = f.w > 0.f;
  } else {
    const uchar4 u = *reinterpret_cast<const uchar4*>(reinterpret_cast<const unsigned char*>(m) + o);
    on[0] = u.x != 0; on[1] = u.y != 0; on[2] = u.z != 0; on[3] = u.w != 0;
  }
}
__global__ void __launch_bounds__(256) conv_splitk_epilogue4_kernel(const float* __restrict__ partial, int splits,
                                                                    int Ntot, const vqs_conv_gemm_desc d,
                                                                    const FastDiv divN4, const FastDiv divL4) {
  pdl_prologue_done();
  const uint32_t N4 = (uint32_t)Ntot >> 2;
  const uint32_t total4 = (uint32_t)d.M * N4;
  const size_t total = (size_t)d.M * Ntot;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gridDim.x * 256u) {
    uint32_t m, n4, b, l4;
    divN4.divmod(i, m, n4);
    divL4.divmod(n4, b, l4);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < splits; ++z) {
      const float4 pz = *reinterpret_cast<const float4*>(partial + (size_t)z * total + (size_t)i * 4);
      v[0] += pz.x; v[1] += pz.y; v[2] += pz.z; v[3] += pz.w;
    }
    const size_t o = ((size_t)b * d.M + m) * d.Lout + (size_t)l4 * 4;
    if (d.bias) {
      const float bs = __ldg(d.bias + m);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += bs;
    }
    if (d.add_pre) {
      const float4 pr = *reinterpret_cast<const float4*>(d.add_pre + o);
      const float pre[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += d.add_pre_relu ? fmaxf(pre[e], 0.f) : pre[e];
    }
    if (d.relu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    if (d.mask_out)
      *reinterpret_cast<uchar4*>(d.mask_out + o) = make_uchar4(v[0] > 0.f, v[1] > 0.f, v[2] > 0.f, v[3] > 0.f);
    if (d.mask_kind) {
      bool on[4];
      tc_mask4(d.mask, d.mask_kind, o, on);
#pragma unroll
      for (int e = 0; e < 4; ++e) if (!on[e]) v[e] = 0.f;
    }
    if (d.add_post) {
      const float4 po = *reinterpret_cast<const float4*>(d.add_post + o);
      v[0] += po.x; v[1] += po.y; v[2] += po.z; v[3] += po.w;
    }
    *reinterpret_cast<float4*>(d.out + o) = make_float4(v[0], v[1], v[2], v[3]);
    if (d.out2) {
      bool on[4] = {true, true, true, true};
      if (d.mask2_kind) tc_mask4(d.mask2, d.mask2_kind, o, on);
      *reinterpret_cast<float4*>(d.out2 + o) =
          make_float4(on[0] ? v[0] : 0.f, on[1] ? v[1] : 0.f, on[2] ? v[2] : 0.f, on[3] ? v[3] : 0.f);
    }
  }
}

int launch_conv_tc(const ConvParams& p, int precision, cudaStream_t st) {
  TcParams<0> prm;
  prm.p = p;
  const int mt = (p.d.M + BM - 1) / BM;
  // BN in {128, 64}: one CTA per SM, so the launch costs ceil(tiles / SMs) waves; a 128 x 64 tile costs ~0.6 of a
  // 128 x 128 one (same A operand, half the B operand and MMA work).  E.g. N = 3200: 150 tiles = 2 waves at BN = 128
  // but 3 cheaper waves at BN = 64.
  const long long sms = num_sms();
  const long long t128 = (long long)mt * ((p.Ntot + 127) / 128), t64 = (long long)mt * ((p.Ntot + 63) / 64);
  const double c128 = (double)((t128 + sms - 1) / sms), c64 = 0.6 * (double)((t64 + sms - 1) / sms);
  int bn = (c128 <= c64) ? 128 : 64;
  // Split-K (needs the caller's scratch, splitk_ws; partial sums are folded in a fixed order -> deterministic):
  //  * a handful of tiles (M <= 128 layers: pre_vq conv, the decoder's first dgrad, the last transposed conv: 12 - 24 CTAs
  //    walking 48 - 72 k-blocks each = 0.05 - 0.066 ms): 64-column tiles, as many splits as fill the machine;
  //  * half a wave of 128-column tiles (the T_q = 24 layers: N = 1536 -> 72 tiles): two splits of 128 x 128 tiles instead
  //    of one wave of 128 x 64 tiles that re-read the weight operand for half the work.
  const vqs_conv_gemm_desc& d = p.d;
  const int nkb = (p.Ktot + BKF - 1) / BKF;
  prm.p.splits = 1;
  prm.p.partial = nullptr;
  // half a wave of 128-column tiles, an even number of 128-row tiles and too few k-blocks for the split below (the 1 x 1 layers at
  // T_q = 24): one wave of 64-column CTA PAIRS instead of single 128 x 64 CTAs (0.029 -> 0.024 ms).  With 72 k-blocks the two
  // split-K halves of 128-column tiles stay faster (0.034 against 0.039 ms measured for 64-column pairs without a split).
  const bool half_wave_pairs = conv_bn64_pair_enabled() && pair_enabled() && mt % 2 == 0 && d.a_tap_major != 0 &&
                               t128 * 2 <= sms && t128 * 3 > sms && t64 <= sms && nkb < 48;
  if (half_wave_pairs) bn = 64;
  if (!half_wave_pairs && d.splitk_ws != nullptr && nkb >= 8) {
    int s = 0, sbn = 0;
    if (t64 * 3 <= sms) {
      sbn = 64;
      s = (int)(sms / t64);
    } else if (t128 * 2 <= sms && t128 * 3 > sms && nkb >= 48) {   // (k = 1 layers, 24 k-blocks: the split loses)
      sbn = 128;
      s = (int)(sms / t128);
    }
    if (s > nkb / 4) s = nkb / 4;
    const size_t need = (size_t)(s > 0 ? s : 0) * d.M * p.Ntot * sizeof(float);
    if (s >= 2 && need <= d.splitk_ws_bytes) {
      bn = sbn;
      prm.p.kt_per_split = (nkb + s - 1) / s;
      prm.p.splits = (nkb + prm.p.kt_per_split - 1) / prm.p.kt_per_split;
      prm.p.partial = (float*)d.splitk_ws;
    }
  }
  // (Tried: running the heavy dgrad epilogues -- masks, skip additions, second output -- as a separate pass over raw
  // accumulators instead of serially after each CTA's main loop: 3.42 vs 3.36 ms per step, so they stay in the kernel.)
  dim3 grid((p.Ntot + bn - 1) / bn, mt, prm.p.splits);
  if (int e = launch_tc<0>(prm, p.d.ksz, bn, precision, grid, st)) return e;
  if (prm.p.partial != nullptr) {
    const long long total = (long long)d.M * p.Ntot;
    const long long blocks = (total + 255) / 256;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    auto al4 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 3) == 0; };
    const bool vec4 = p.Ntot % 4 == 0 && d.Lout % 4 == 0 && total < (1ll << 31) && al16(prm.p.partial) && al16(d.out) &&
                      al16(d.add_pre) && al16(d.add_post) && al16(d.out2) && al4(d.mask_out) &&
                      (d.mask_kind == 1 ? al16(d.mask) : al4(d.mask)) && (d.mask2_kind == 1 ? al16(d.mask2) : al4(d.mask2));
    if (vec4) {
      const long long b4 = (total / 4 + 255) / 256;
      VQS_CUDA(launch_pdl(conv_splitk_epilogue4_kernel, dim3((unsigned)(b4 < 8 * sms ? b4 : 8 * sms)), dim3(256), 0, st,
                          prm.p.partial, prm.p.splits, p.Ntot, d, FastDiv((uint32_t)(p.Ntot / 4)),
                          FastDiv((uint32_t)(d.Lout / 4))));
    } else {
      VQS_CUDA(launch_pdl(conv_splitk_epilogue_kernel, dim3((unsigned)(blocks < 8 * sms ? blocks : 8 * sms)), dim3(256), 0, st,
                          prm.p.partial, prm.p.splits, p.Ntot, d));
    }
    VQS_LAUNCH_CHECK();
  }
  return 0;
}

int launch_wgrad_tc(const WgradParams& p, int precision, cudaStream_t st) {
  TcParams<1> prm;
  prm.p = p;
  int bn = 128;
  // 96-column tiles when 128-column tiles leave SMs without a tile although the kernel is paced by its producers, whose work
  // per k-block is proportional to the rows a CTA stages (128 of A + BN / 2 of B): 768 x 2304 = 108 tiles of 128 columns on 148
  // SMs, or 144 of 96 (each CTA stages 176 instead of 192 rows per k-block)
  {
    const int mt = (p.d.M + BM - 1) / BM;
    const int t128 = ((p.Nw + 127) / 128) * mt, t96 = ((p.Nw + 95) / 96) * mt;
    static const bool off = getenv("VQS_WGRAD_BN96") && atoi(getenv("VQS_WGRAD_BN96")) == 0;
    if (!off && precision == 2 && p.splits == 1 && mt % 2 == 0 && pair_enabled() && p.Nw % 96 == 0 && t96 <= num_sms() &&
        t128 < t96)
      bn = 96;
  }
  dim3 grid((p.Nw + bn - 1) / bn, (p.d.M + BM - 1) / BM, p.splits);
  return launch_tc<1>(prm, p.d.ksz, bn, precision, grid, st);
}

}  // namespace vqs

#ifdef VQS_GEMM_TIMING
extern "C" int vqs_debug_gemm_timing(long long* out8) {
  VQS_CUDA(cudaDeviceSynchronize());
  VQS_CUDA(cudaMemcpyFromSymbol(out8, vqs::g_tc_timing, sizeof(long long) * 48));
  return 0;
}
#endif
