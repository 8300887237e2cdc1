// Weight-gradient GEMM with TMA-fed operands (tcgen05, 3xTF32) -- the fast path of vqs_wgrad_gemm.
//
//   dW[m][c][j] = sum_{b, l} g[b][m][l] * x'[b][c][l + j*j_mul + off]          (autograd of nn.Conv1d / nn.ConvTranspose1d,
//                                                                               /root/reference/src/modules/conv1d_builder.py:33-44)
// Both operands are K-major as they lie in HBM: the reduction index l is the contiguous one of the (B, C, L) tensors.  So
// a TMA tensor map (l, b * C + channel) drops a [128 channels x 32 positions] box straight into the UMMA SWIZZLE_128B layout:
// the tap shift j*j_mul + off is a box coordinate, positions outside [0, L) arrive as zeros (= the conv padding), no
// thread computes an address.  LIMIT (found on the B200: the copy engine raises an illegal-instruction fault otherwise):
// a box origin must be 16-byte aligned, so every tap shift has to be a multiple of 4 positions -- true for the 1 x 1
// convolutions of the residual blocks only; the k = 3 layers (shifts -1 / 0 / +1) stay on the gather kernel until the
// shifted operand goes through a staged, re-aligned copy (DESIGN.md "next").  The raw fp32 words ARE the hi operands (kind::tf32 reads the top 19 bits); eight warps
// only derive the lo tiles  lo = x - tf32(x)  (a linear pass over 32 KB per k-block: the swizzle is position-preserving)
// and apply the fused input ReLU where the descriptor asks for it.  The gather-by-threads kernel (gemm_tc.cu, MODE 1)
// spends ~2000 warp-instructions per k-block on index arithmetic, scalar loads and swizzled stores; this one ~700.
// Numerics are those of gemm_tc.cu: D += Al*Bh + Ah*Bl (accumulator 0), Ah*Bh round-robin over accumulators 1-3, summed in
// fp32 by the epilogue.
//
// Tile: 128 (m) x 128 (channels c of ONE tap j); k-block = 32 positions of one utterance (rows past L are zero-filled, so
// L = 48 costs 2 blocks of 32).  Split-K over (b, l-block) as planned by plan_wgrad; every CTA writes its 128 x 128 tile
// to the workspace as one contiguous block (coalesced float4), wgrad_tma_reduce_kernel folds the splits into dW's
// (m, c, j) order.
// Warp roles (320 threads): warps 0-7 lo pass + epilogue (TMEM lane quarter = warp % 4), warp 8 TMA producer, warp 9 TMEM
// allocator + MMA issuer.
#include <cuda.h>
#include <stdlib.h>

#include "gemm_params.cuh"
#include "tc_common.cuh"

namespace vqs {
namespace {

constexpr int WT_T16 = 128 * 128;            // bytes of one [128 rows x 32 floats] operand tile
constexpr int WT_STAGE = 4 * WT_T16;         // A raw | B raw | A lo | B lo
constexpr int WT_STAGES = 3;
constexpr int WT_LO_WARPS = 8;
constexpr int WT_TMA_WARP = WT_LO_WARPS, WT_MMA_WARP = WT_LO_WARPS + 1;
constexpr int WT_THREADS = (WT_LO_WARPS + 3) * 32;   // lo warps | TMA warp | MMA issuer (main term) | MMA issuer (corrections)
constexpr int WT_SMEM = WT_STAGES * WT_STAGE + 1024 + 256;

struct WtShared {
  uint64_t full_raw[WT_STAGES], full_lo[WT_STAGES], empty[WT_STAGES], tmem_full;
  uint32_t tmem_base;
};

struct WgradTmaParams {
  float* partial;      // [splits][tiles_m][tiles_n][128][128]
  int nlb;             // 32-position blocks per utterance
  int kb_total;        // B * nlb
  int kt_per_split;
  int cpt;             // 128-channel tiles per tap: Cred / 128
  int j_mul, off, x_relu;
  int M, Cred;
};

__device__ __forceinline__ void wt_tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ float4 wt_lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void wt_tmem_ld16(uint32_t taddr, float* v) {
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

__global__ void __launch_bounds__(WT_THREADS, 1) wgrad_tma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapB,
                                                                  const WgradTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  WtShared* sh = reinterpret_cast<WtShared*>(smem + WT_STAGES * WT_STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x / p.cpt, c0 = (blockIdx.x - j * p.cpt) * 128, m0 = blockIdx.y * 128;
  const int kb_begin = blockIdx.z * p.kt_per_split;
  int kb_end = kb_begin + p.kt_per_split;
  if (kb_end > p.kb_total) kb_end = p.kb_total;
  const int nkb = kb_end > kb_begin ? kb_end - kb_begin : 0;

  if (tid == 0) {
    for (int s = 0; s < WT_STAGES; ++s) {
      mbar_init(&sh->full_raw[s], 1);
      mbar_init(&sh->full_lo[s], WT_LO_WARPS);
      mbar_init(&sh->empty[s], 2);       // one tcgen05.commit per MMA issuer
    }
    mbar_init(&sh->tmem_full, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WT_MMA_WARP) tmem_alloc(&sh->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  const uint32_t smem_a = smem_u32(smem);
  pdl_prologue_done();

  if (warp == WT_TMA_WARP) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int shift = j * p.j_mul + p.off;
      int b = kb_begin / p.nlb, lb = kb_begin - b * p.nlb;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % WT_STAGES;
        mbar_wait_sleep(&sh->empty[s], ((uint32_t)(i / WT_STAGES) & 1u) ^ 1u);
        const uint32_t bar = smem_u32(&sh->full_raw[s]);
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar),
                     "r"((uint32_t)(2 * WT_T16))
                     : "memory");
        const uint32_t dst = smem_a + (uint32_t)(s * WT_STAGE);
        wt_tma_load_2d(dst, &mapA, lb * 32, b * p.M + m0, &sh->full_raw[s]);                      // g[b][m0 ..][l0 ..]
        wt_tma_load_2d(dst + WT_T16, &mapB, lb * 32 + shift, b * p.Cred + c0, &sh->full_raw[s]);  // x[b][c0 ..][l0 + shift ..]
        if (++lb == p.nlb) {
          lb = 0;
          ++b;
        }
      }
    }
    __syncwarp();
  } else if (warp == WT_MMA_WARP || warp == WT_MMA_WARP + 1) {
    // ================= MMA issuers (gemm_tc.cu::issue_mmas: the issuing thread, not the operand feed, limited the round-1
    // kernel): ring unrolled over the stages -> every descriptor is base + constant; warp WT_MMA_WARP issues the main
    // term into accumulators 1..3 (round-robin), warp WT_MMA_WARP + 1 the two correction terms into accumulator 0 =================
    {
      constexpr uint32_t idesc = make_idesc_tf32(128);
      const bool main_role = warp == WT_MMA_WARP;
      const bool elected = elect_one();      // whole converged warp runs the loop, MMAs guarded (bare UTCHMMA)
      const uint64_t d0 = make_desc_sw128(smem_a);
      uint32_t par = 0;
#pragma unroll 1
      for (int i0 = 0; i0 < nkb; i0 += WT_STAGES) {
        const uint32_t nz = i0 > 0 ? 1u : 0u;
#pragma unroll
        for (int s = 0; s < WT_STAGES; ++s) {
          if (i0 + s < nkb) {
            mbar_wait(&sh->full_raw[s], par);
            mbar_wait(&sh->full_lo[s], par);
            tc_fence_after();
            const uint64_t a_hi = d0 + (uint64_t)((s * WT_STAGE) >> 4), b_hi = a_hi + (uint64_t)(WT_T16 >> 4);
            const uint64_t a_lo = a_hi + (uint64_t)((2 * WT_T16) >> 4), b_lo = a_hi + (uint64_t)((3 * WT_T16) >> 4);
            if (main_role) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                const int g = s * 4 + k;
                if (elected) umma_tf32(tmem_base + (uint32_t)((1 + g % 3) * 128), a_hi + adv, b_hi + adv, idesc, g < 3 ? nz : 1u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                if (elected) {
                  umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (s == 0 && k == 0) ? nz : 1u);
                  umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                }
              }
            }
            if (elected) umma_commit(&sh->empty[s]);
            __syncwarp();
          }
        }
        par ^= 1u;
      }
      if (elected) umma_commit(&sh->tmem_full);
    }
    __syncwarp();
  } else {
    // ================= lo pass: lo = x - tf32(x) over the 32 KB [A raw | B raw] of every stage =================
    const bool relu = p.x_relu != 0;
    for (int i = 0; i < nkb; ++i) {
      const int s = i % WT_STAGES;
      mbar_wait_sleep(&sh->full_raw[s], (uint32_t)(i / WT_STAGES) & 1u);
      const uint32_t src = smem_a + (uint32_t)(s * WT_STAGE) + (uint32_t)(tid * 16);
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = wt_lds_v4(src + (uint32_t)(u * WT_LO_WARPS * 32 * 16));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t a = src + (uint32_t)(u * WT_LO_WARPS * 32 * 16);
        float4 x = v[u];
        if (u >= 4 && relu) {      // the B tile (second 16 KB) carries the fused input ReLU: rewrite the hi operand too
          x = make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
          sts_v4(a, x);
        }
        sts_v4(a + 2 * WT_T16, make_float4(x.x - tf32_hi(x.x), x.y - tf32_hi(x.y), x.z - tf32_hi(x.z), x.w - tf32_hi(x.w)));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->full_lo[s]);
    }
    // ================= epilogue: (acc1 + acc2 + acc3) + acc0 -> this CTA's contiguous 128 x 128 block =================
    const int q = warp & 3, half = warp >> 2;
    float* out = p.partial + ((size_t)((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16384 +
                 (size_t)(q * 32 + lane) * 128 + half * 64;
    if (nkb > 0) {
      mbar_wait_sleep(&sh->tmem_full, 0);
      tc_fence_after();
    }
    const float comp = tf32x3_comp((4 * nkb + 2) / 3);
#pragma unroll 1
    for (int cb = 0; cb < 64; cb += 16) {
      float r[16];
      if (nkb > 0) {
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64 + cb);
        float t1[16], t2[16], t3[16];
        wt_tmem_ld16(ta + 128, r);
        wt_tmem_ld16(ta + 256, t1);
        wt_tmem_ld16(ta + 384, t2);
        wt_tmem_ld16(ta, t3);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = fmaf((r[e] + t1[e]) + t2[e], comp, t3[e]);   // tc_common.cuh::tf32x3_comp
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        *reinterpret_cast<float4*>(out + cb + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[m][c][j] (+)= sum over splits of the tile blocks written by wgrad_tma_kernel
__global__ void __launch_bounds__(256) wgrad_tma_reduce_kernel(const float* __restrict__ partial, int splits, int M,
                                                               int Cred, int ksz, int tiles_m, int tiles_n,
                                                               float* __restrict__ dW, int accumulate) {
  pdl_prologue_done();
  const long long total = (long long)M * Cred * ksz;
  const int cpt = Cred / 128;
  const size_t split_stride = (size_t)tiles_m * tiles_n * 16384;
  if (ksz == 1 && (Cred & 127) == 0 && ((reinterpret_cast<uintptr_t>(partial) | reinterpret_cast<uintptr_t>(dW)) & 15) == 0) {
    // 1 x 1 layers (the only users today): four consecutive channels per thread, shifts instead of divisions
    const uint32_t c4n = (uint32_t)Cred >> 2, total4 = (uint32_t)M * c4n;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total4; i += gridDim.x * 256u) {
      const uint32_t m = i / c4n, c = (i - m * c4n) << 2;
      const size_t o = ((size_t)(m >> 7) * tiles_n + (size_t)(c >> 7)) * 16384 + (size_t)(m & 127) * 128 + (c & 127);
      float4 a = accumulate ? *reinterpret_cast<const float4*>(dW + (size_t)i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < splits; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(partial + (size_t)s * split_stride + o);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      *reinterpret_cast<float4*>(dW + (size_t)i * 4) = a;
    }
    return;
  }
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int n = (int)(i % ((long long)Cred * ksz)), m = (int)(i / ((long long)Cred * ksz));
    const int c = n / ksz, j = n - c * ksz;
    const size_t o = ((size_t)(m >> 7) * tiles_n + (size_t)(j * cpt + (c >> 7))) * 16384 + (size_t)(m & 127) * 128 + (c & 127);
    float a = accumulate ? dW[i] : 0.f;
    for (int s = 0; s < splits; ++s) a += partial[(size_t)s * split_stride + o];
    dW[i] = a;
  }
}

typedef CUresult (*WtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
WtEncodeFn wt_encode_fn() {
  static WtEncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WtEncodeFn>(ptr);
  }
  return fn;
}

// (l, b * C + channel) view of a contiguous (B, C, L) fp32 tensor, box = 32 positions x 128 channels (a box never leaves
// its utterance because C % 128 == 0)
bool wt_make_map(CUtensorMap* map, const float* base, int B, int C, int L) {
  WtEncodeFn enc = wt_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)L, (cuuint64_t)B * C};
  const cuuint64_t strides[1] = {(cuuint64_t)L * 4};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// 3xTF32 only; unit position stride; rows of both tensors 16-byte aligned; whole 128 x 128 tiles
bool wgrad_tma_supported(const WgradParams& p) {
  const vqs_wgrad_desc& d = p.d;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // TMA needs 16-byte aligned box origins: every tap shift j*j_mul + off must be a multiple of 4 positions.  That holds
  // for the 1 x 1 convolutions only (k = 3, pad = 1 shifts by -1 / 0 / +1); shifted taps need a staged, re-aligned copy.
  // On by default (VQS_WGRAD_TMA=0 switches it off): on the step's 1 x 1 layers it beats the gather kernel by ~10 %
  // (0.0416 vs 0.0453 ms at Kred = 3072, 0.031 vs 0.032 ms at Kred = 1536) although a quarter of its k-blocks is zero
  // padding (L = 48 / 24 positions in blocks of 32).
  const char* off = getenv("VQS_WGRAD_TMA");
  if (off && off[0] == '0') return false;
  for (int j = 0; j < d.ksz; ++j)
    if ((j * d.j_mul + d.off) % 4 != 0) return false;
  return d.precision == VQS_PREC_TF32X3 && d.l_mul == 1 && d.La % 4 == 0 && d.Lx % 4 == 0 && d.M % 128 == 0 &&
         d.Cred % 128 == 0 && d.ksz >= 1 && d.ksz <= 4 && al16(d.Aact) && al16(d.X) && wt_encode_fn() != nullptr;
}

int wgrad_tma_kblocks(const WgradParams& p) { return p.d.B * ((p.d.La + 31) / 32); }

// p.splits / p.kt_per_split planned over wgrad_tma_kblocks(p) k-blocks; workspace holds splits * M * Nw floats
int launch_wgrad_tma(const WgradParams& p, float* workspace, cudaStream_t st) {
  const vqs_wgrad_desc& d = p.d;
  CUtensorMap mapA, mapB;
  if (!wt_make_map(&mapA, d.Aact, d.B, d.M, d.La) || !wt_make_map(&mapB, d.X, d.B, d.Cred, d.Lx)) {
    set_error("vqs_wgrad_gemm: cuTensorMapEncodeTiled failed (M=%d Cred=%d La=%d Lx=%d)", d.M, d.Cred, d.La, d.Lx);
    return VQS_ERR_ARG;
  }
  WgradTmaParams q;
  q.partial = workspace;
  q.nlb = (d.La + 31) / 32;
  q.kb_total = d.B * q.nlb;
  q.kt_per_split = p.kt_per_split;
  q.cpt = d.Cred / 128;
  q.j_mul = d.j_mul;
  q.off = d.off;
  q.x_relu = d.x_relu;
  q.M = d.M;
  q.Cred = d.Cred;
  static DevCache configured;
  if (dev_needs(configured, WT_SMEM))
    VQS_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
  const int tiles_n = d.ksz * q.cpt, tiles_m = d.M / 128;
  dim3 grid(tiles_n, tiles_m, p.splits);
  VQS_CUDA(launch_pdl(wgrad_tma_kernel, grid, dim3(WT_THREADS), WT_SMEM, st, mapA, mapB, q));
  VQS_LAUNCH_CHECK();
  const long long n = (long long)d.M * p.Nw;
  const long long blocks = (n + 255) / 256;
  VQS_CUDA(launch_pdl(wgrad_tma_reduce_kernel, dim3((unsigned)(blocks < 8 * num_sms() ? blocks : 8 * num_sms())), dim3(256), 0,
                      st, workspace, p.splits, d.M, d.Cred, d.ksz, tiles_m, tiles_n, d.dW, d.accumulate));
  VQS_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqs
