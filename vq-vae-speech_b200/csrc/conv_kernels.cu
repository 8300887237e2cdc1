// Conv1d / ConvTranspose1d forward, dgrad and wgrad as implicit GEMMs with fused epilogues -- exact-fp32 CUDA-core path.
// Replaces the ATen/cuDNN calls behind nn.Conv1d / nn.ConvTranspose1d / F.relu / residual adds of
//   /root/reference/src/models/convolutional_encoder.py:118-146, deconvolutional_decoder.py:100-137,
//   /root/reference/src/modules/residual.py:69-70, residual_stack.py:43-46  and their autograd backward.
//
// GEMM view (see include/vqs_b200.h): rows m = output channels, columns n = (batch, position), reduction over
// kk = (input channel, tap).  Operand A is a dense row-major weight matrix; operand B is gathered on the fly from the NCL
// activation tensor (padding, stride, transposed-conv tap flip, nearest-neighbour index and input ReLU are all index
// arithmetic in the loader), so no im2col buffer ever exists in HBM.
#include "vqs_common.cuh"
#include <math.h>

#include <stdlib.h>

#include "gemm_params.cuh"

namespace vqs {
namespace {

constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <int BM, int BN>
struct TileCfg {
  static constexpr int GM = BM / 64, GN = BN / 64;  // groups of 4 rows / 4 columns per thread
  static constexpr int TM = 4 * GM, TN = 4 * GN;
  static constexpr int LDA = BM + 4, LDB = BN + 4;
  static_assert(BM % 64 == 0 && BN % 64 == 0, "tile must be a multiple of 64");
};

template <int BM, int BN>
__device__ __forceinline__ void compute_tile(const float* __restrict__ As, const float* __restrict__ Bs,
                                             float (&acc)[TileCfg<BM, BN>::TM][TileCfg<BM, BN>::TN], int ty, int tx) {
  using C = TileCfg<BM, BN>;
#pragma unroll
  for (int k = 0; k < BK; ++k) {
    float a[C::TM], b[C::TN];
#pragma unroll
    for (int g = 0; g < C::GM; ++g) {
      float4 v = *reinterpret_cast<const float4*>(&As[k * C::LDA + g * 64 + ty * 4]);
      a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
    }
#pragma unroll
    for (int g = 0; g < C::GN; ++g) {
      float4 v = *reinterpret_cast<const float4*>(&Bs[k * C::LDB + g * 64 + tx * 4]);
      b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
    }
#pragma unroll
    for (int i = 0; i < C::TM; ++i)
#pragma unroll
      for (int j = 0; j < C::TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------------------------------------
// conv-like GEMM (forward conv, dgrad, transposed conv forward, transposed conv dgrad)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mask_on(const void* m, int kind, size_t i) {
  if (kind == 1) return reinterpret_cast<const float*>(m)[i] > 0.f;
  return reinterpret_cast<const unsigned char*>(m)[i] != 0;
}

template <int BM, int BN, int KSZ>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_kernel(const ConvParams p) {
  using C = TileCfg<BM, BN>;
  __shared__ __align__(16) float As[2][BK * C::LDA];
  __shared__ __align__(16) float Bs[2][BK * C::LDB];
  const vqs_conv_gemm_desc& d = p.d;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int Ktot = p.Ktot;
  const int ktiles = (Ktot + BK - 1) / BK;

  // ---- A loader: float4 f = tid + 256*i -> row f/4, k (f%4)*4 ----
  constexpr int A_LD = BM / 64;
  // ---- B loader: one column per thread, B_LD k-rows spaced B_STEP apart ----
  constexpr int B_LD = BN / 16;
  constexpr int B_STEP = NTHREADS / BN;
  const int bcol = tid % BN;
  const int brow0 = tid / BN;
  const int n_ld = n0 + bcol;
  const bool n_ok = n_ld < p.Ntot;
  uint32_t bb = 0, ll = 0;
  if (n_ok) p.divL.divmod((uint32_t)n_ld, bb, ll);
  const float* xb = d.X + (long long)bb * d.x_sb;
  const int lbase = (int)ll * d.l_mul + d.off;

  float4 areg[A_LD];
  float breg[B_LD];

  auto load_global = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int f = tid + NTHREADS * i;
      int row = f >> 2, kq = (f & 3) << 2;
      int m = m0 + row, k = k0 + kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < d.M) {
        const float* src = d.A + (size_t)m * Ktot + k;
        if (p.a_vec) {
          if (k < Ktot) v = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (k + 0 < Ktot) v.x = __ldg(src + 0);
          if (k + 1 < Ktot) v.y = __ldg(src + 1);
          if (k + 2 < Ktot) v.z = __ldg(src + 2);
          if (k + 3 < Ktot) v.w = __ldg(src + 3);
        }
      }
      areg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int kk = k0 + brow0 + B_STEP * i;
      int c, j;
      if (d.a_tap_major) {   // kk = j*Cred + c
        uint32_t jj = p.divCred.div((uint32_t)kk);
        j = (int)jj;
        c = kk - j * d.Cred;
      } else {               // kk = c*ksz + j
        c = kk / KSZ;
        j = kk - c * KSZ;
      }
      int pn = lbase + j * d.j_mul;
      bool ok = n_ok && kk < Ktot && pn >= 0;
      if (d.l_div == 2) {
        ok = ok && ((pn & 1) == 0);
        pn >>= 1;
      }
      ok = ok && pn < d.Lin;
      float v = 0.f;
      if (ok) {
        v = __ldg(xb + (long long)c * d.x_sc + (long long)pn * d.x_sl);
        if (d.x_relu) v = fmaxf(v, 0.f);
      }
      breg[i] = v;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int f = tid + NTHREADS * i;
      int row = f >> 2, kq = (f & 3) << 2;
      float* a = &As[buf][kq * C::LDA + row];
      a[0 * C::LDA] = areg[i].x;
      a[1 * C::LDA] = areg[i].y;
      a[2 * C::LDA] = areg[i].z;
      a[3 * C::LDA] = areg[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) Bs[buf][(brow0 + B_STEP * i) * C::LDB + bcol] = breg[i];
  };

  float acc[C::TM][C::TN];
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j] = 0.f;

  load_global(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) load_global(kt + 1);
    compute_tile<BM, BN>(As[buf], Bs[buf], acc, ty, tx);
    if (kt + 1 < ktiles) store_smem(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int gj = 0; gj < C::GN; ++gj) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int n = n0 + gj * 64 + tx * 4 + jj;
      if (n >= p.Ntot) continue;
      uint32_t b, l;
      p.divL.divmod((uint32_t)n, b, l);
      const size_t col_off = (size_t)b * d.M * d.Lout + l;
#pragma unroll
      for (int gi = 0; gi < C::GM; ++gi) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int m = m0 + gi * 64 + ty * 4 + ii;
          if (m >= d.M) continue;
          const size_t o = col_off + (size_t)m * d.Lout;
          float v = acc[gi * 4 + ii][gj * 4 + jj];
          if (d.bias) v += __ldg(d.bias + m);
          if (d.add_pre) {
            float r = d.add_pre[o];
            v += d.add_pre_relu ? fmaxf(r, 0.f) : r;
          }
          if (d.relu) v = fmaxf(v, 0.f);
          if (d.mask_out) d.mask_out[o] = v > 0.f ? 1 : 0;
          if (d.mask_kind && !mask_on(d.mask, d.mask_kind, o)) v = 0.f;
          if (d.add_post) v += d.add_post[o];
          d.out[o] = v;
          if (d.out2) d.out2[o] = (d.mask2_kind == 0 || mask_on(d.mask2, d.mask2_kind, o)) ? v : 0.f;
        }
      }
    }
  }
}

template <int BM, int BN>
int launch_conv(const ConvParams& p, cudaStream_t st) {
  dim3 grid((p.Ntot + BN - 1) / BN, (p.d.M + BM - 1) / BM, 1);
  switch (p.d.ksz) {
    case 1: conv_gemm_kernel<BM, BN, 1><<<grid, NTHREADS, 0, st>>>(p); break;
    case 2: conv_gemm_kernel<BM, BN, 2><<<grid, NTHREADS, 0, st>>>(p); break;
    case 3: conv_gemm_kernel<BM, BN, 3><<<grid, NTHREADS, 0, st>>>(p); break;
    case 4: conv_gemm_kernel<BM, BN, 4><<<grid, NTHREADS, 0, st>>>(p); break;
    default: set_error("vqs_conv_gemm: kernel size %d not supported (1..4)", p.d.ksz); return VQS_ERR_ARG;
  }
  VQS_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// wgrad GEMM: dW[m, (c, j)] = sum_{(b, l)} Aact[b, m, l] * X'[b, c, l*l_mul + j*j_mul + off]
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int KSZ>
__global__ void __launch_bounds__(NTHREADS) wgrad_gemm_kernel(const WgradParams p) {
  using C = TileCfg<BM, BN>;
  __shared__ __align__(16) float As[2][BK * C::LDA];
  __shared__ __align__(16) float Bs[2][BK * C::LDB];
  const vqs_wgrad_desc& d = p.d;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ktiles = (p.Kred + BK - 1) / BK;
  const int kt_begin = blockIdx.z * p.kt_per_split;
  int kt_end = kt_begin + p.kt_per_split;
  if (kt_end > ktiles) kt_end = ktiles;

  constexpr int A_LD = BM / 16, B_LD = BN / 16;
  const int kl = tid & 15, r = tid >> 4;
  int boff[B_LD], bjo[B_LD];
#pragma unroll
  for (int i = 0; i < B_LD; ++i) {
    int n = n0 + r + 16 * i;
    int c = n / KSZ, j = n - c * KSZ;
    bjo[i] = (n < p.Nw) ? j * d.j_mul + d.off : (-(1 << 29));  // forces pn < 0 -> zero
    boff[i] = c * d.Lx;
  }
  float areg[A_LD], breg[B_LD];

  auto load_global = [&](int kt) {
    const int kk = kt * BK + kl;
    const bool k_ok = kk < p.Kred;
    uint32_t b = 0, l = 0;
    if (k_ok) p.divLa.divmod((uint32_t)kk, b, l);
    const float* ab = d.Aact + ((size_t)b * d.M) * d.La + l;
    const float* xb = d.X + ((size_t)b * d.Cred) * d.Lx;
    const int lp = (int)l * d.l_mul;
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      int m = m0 + r + 16 * i;
      areg[i] = (k_ok && m < d.M) ? __ldg(ab + (size_t)m * d.La) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      int pn = lp + bjo[i];
      float v = 0.f;
      if (k_ok && pn >= 0 && pn < d.Lx) {
        v = __ldg(xb + boff[i] + pn);
        if (d.x_relu) v = fmaxf(v, 0.f);
      }
      breg[i] = v;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) As[buf][kl * C::LDA + r + 16 * i] = areg[i];
#pragma unroll
    for (int i = 0; i < B_LD; ++i) Bs[buf][kl * C::LDB + r + 16 * i] = breg[i];
  };

  float acc[C::TM][C::TN];
#pragma unroll
  for (int i = 0; i < C::TM; ++i)
#pragma unroll
    for (int j = 0; j < C::TN; ++j) acc[i][j] = 0.f;

  if (kt_begin < kt_end) {
    load_global(kt_begin);
    store_smem(0);
    __syncthreads();
    for (int kt = kt_begin; kt < kt_end; ++kt) {
      const int buf = (kt - kt_begin) & 1;
      if (kt + 1 < kt_end) load_global(kt + 1);
      compute_tile<BM, BN>(As[buf], Bs[buf], acc, ty, tx);
      if (kt + 1 < kt_end) store_smem(buf ^ 1);
      __syncthreads();
    }
  }
  float* out = p.partial ? p.partial + (size_t)blockIdx.z * d.M * p.Nw : d.dW;
  const bool accum = (p.partial == nullptr) && d.accumulate;
#pragma unroll
  for (int gi = 0; gi < C::GM; ++gi)
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int m = m0 + gi * 64 + ty * 4 + ii;
      if (m >= d.M) continue;
#pragma unroll
      for (int gj = 0; gj < C::GN; ++gj)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int n = n0 + gj * 64 + tx * 4 + jj;
          if (n >= p.Nw) continue;
          const size_t o = (size_t)m * p.Nw + n;
          float v = acc[gi * 4 + ii][gj * 4 + jj];
          out[o] = accum ? out[o] + v : v;
        }
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int S, long long n, float* __restrict__ out,
                                     int accumulate) {
  pdl_prologue_done();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = accumulate ? out[i] : 0.f;
    for (int s = 0; s < S; ++s) a += partial[(size_t)s * n + i];
    out[i] = a;
  }
}

// four elements per thread (n % 4 == 0, 16-byte aligned buffers): same summation order per element
__global__ void splitk_reduce4_kernel(const float4* __restrict__ partial, int S, long long n4, float4* __restrict__ out,
                                      int accumulate) {
  pdl_prologue_done();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = accumulate ? out[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float4 v = partial[(size_t)s * n4 + i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    out[i] = a;
  }
}

template <int BM, int BN>
int launch_wgrad(const WgradParams& p, cudaStream_t st) {
  dim3 grid((p.Nw + BN - 1) / BN, (p.d.M + BM - 1) / BM, p.splits);
  switch (p.d.ksz) {
    case 1: wgrad_gemm_kernel<BM, BN, 1><<<grid, NTHREADS, 0, st>>>(p); break;
    case 2: wgrad_gemm_kernel<BM, BN, 2><<<grid, NTHREADS, 0, st>>>(p); break;
    case 3: wgrad_gemm_kernel<BM, BN, 3><<<grid, NTHREADS, 0, st>>>(p); break;
    case 4: wgrad_gemm_kernel<BM, BN, 4><<<grid, NTHREADS, 0, st>>>(p); break;
    default: set_error("vqs_wgrad_gemm: kernel size %d not supported (1..4)", p.d.ksz); return VQS_ERR_ARG;
  }
  VQS_LAUNCH_CHECK();
  return 0;
}

struct WgradPlan {
  int bm, bn, splits, kt_per_split;
};
// bk = reduction elements per k-tile of the engine (16: CUDA cores, 32: tcgen05); tc tiles are always 128 x 128
WgradPlan plan_wgrad(int M, int Nw, int Kred, int bk = BK, bool tc = false, bool pair = false) {
  WgradPlan pl;
  pl.bm = (tc || M > 64) ? 128 : 64;
  pl.bn = (tc || Nw > 64) ? 128 : 64;
  long long tiles = (long long)((M + pl.bm - 1) / pl.bm) * ((Nw + pl.bn - 1) / pl.bn);
  int ktiles = (Kred + bk - 1) / bk;
  int max_s = ktiles / 4 > 0 ? ktiles / 4 : 1;
  if (max_s > 64) max_s = 64;
  if (tc) {
    // One CTA per SM: a launch costs ceil(tiles * splits / SMs) rounds of (k-blocks per CTA * 1.6 us + 4 us of
    // prologue / epilogue) -- fitted on the B200 (768 x 2304 x 3072: 4 splits = 3 rounds of 24 k-blocks = 0.132 ms,
    // 1 split of 48 k-blocks = 0.081 ms; 768 x 768 x 3072: 12 splits = 3 rounds of 8 = 0.060 ms) -- plus the reduction pass
    // over the partials.  "3 CTAs per SM" (the first rule) paid a whole extra round on most layers of the step (1.52 ms of
    // wgrad per step; 1.34 ms with the cheapest split).
    // CTA-pair kernel (gemm_tc.cu, even number of 128-row tiles): 0.55 us per k-block and ~7 us of prologue / first loads /
    // epilogue per round (-DVQS_GEMM_TIMING stamps: 24.5 k cycles for 24 k-blocks, 12.5 k around them), so fewer, longer
    // rounds and no reduction pass win more often than with the 1.6 us k-blocks of the single-CTA kernels
    const double kb_us = pair ? 0.55 : 1.6, round_us = pair ? 7.0 : 4.0;
    const double sms = (double)num_sms();
    double best = 1e30;
    int best_kps = ktiles;
    for (int s = 1; s <= max_s; ++s) {
      const int kps = (ktiles + s - 1) / s;
      const int se = (ktiles + kps - 1) / kps;
      const double rounds = ceil((double)tiles * se / sms);
      double cost = rounds * (kps * kb_us + round_us);
      if (se > 1) cost += 2.0 + se * ((double)M * Nw * 4.0) / 5.0e6;
      if (cost < best - 1e-9) {
        best = cost;
        best_kps = kps;
      }
    }
#ifdef VQS_DEBUG
    if (const char* e = getenv("VQS_WGRAD_SPLITS")) {   // profiling builds: force the split (profiles/probe_wgrad_splits.py)
      const int s = atoi(e);
      if (s >= 1 && s <= max_s) best_kps = (ktiles + s - 1) / s;
    }
#endif
    pl.kt_per_split = best_kps;
    pl.splits = (ktiles + best_kps - 1) / best_kps;
    return pl;
  }
  long long want = (2ll * num_sms() + tiles - 1) / tiles;
  int s = (int)(want < 1 ? 1 : want);
  if (s > max_s) s = max_s;
  pl.kt_per_split = (ktiles + s - 1) / s;
  pl.splits = (ktiles + pl.kt_per_split - 1) / pl.kt_per_split;
  return pl;
}

__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ g, int B, int M, int L,
                                                        float* __restrict__ db, int accumulate, const FastDiv divL,
                                                        const int vec4) {   // divL: L / 4 (vec4) or L
  // one block per channel m; fixed-order tree reduction -> deterministic
  __shared__ float red[256];
  pdl_prologue_done();
  const int m = blockIdx.x;
  float s = 0.f;
  if (vec4) {                                  // L % 4 == 0, 16-byte aligned g: whole 16-byte pieces of the (b, m) runs
    const uint32_t per4 = (uint32_t)B * divL.d;
    for (uint32_t i = threadIdx.x; i < per4; i += 256) {
      uint32_t b, l4;
      divL.divmod(i, b, l4);
      const float4 v = *reinterpret_cast<const float4*>(g + ((size_t)b * M + m) * L + (size_t)l4 * 4);
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    const uint32_t per = (uint32_t)B * L;
    for (uint32_t i = threadIdx.x; i < per; i += 256) {
      uint32_t b, l;
      divL.divmod(i, b, l);
      s += g[((size_t)b * M + m) * L + l];
    }
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) db[m] = accumulate ? db[m] + red[0] : red[0];
}

// Weight re-arrangement through a 32 x 32 x k shared-memory tile: global reads and writes are both contiguous runs
// (the first version read with stride d1*k in modes 0/2: 13.7 us per 7 MB weight, 8 % of the training step).
//   w[a][b][j]  ->  mode 0: out[b][a][j]   mode 1: out[a][j][b]   mode 2: out[b][j][a]
//   modes 3 / 4: tensor-core operand IMAGE of the tap-major matrix A[m][j*CredP + c] (mode 3: m = a, c = b; mode 4: m = b,
//   c = a; CredP = Cred rounded up to a multiple of 32, the extra channels are zeros): for every (128-row tile mt, 32-wide k-block kb) one contiguous 32 KB block [hi 16 KB | lo 16 KB], each
//   copy laid out exactly as the MMA reads it (128 rows x 128 B, SWIZZLE_128B), values pre-split x = hi + lo.  The GEMM
//   kernel then fetches its A operand with ONE cp.async.bulk per k-block instead of 512 thread loads + splits.
template <int KS>
__device__ __forceinline__ void permute_weight_tile(const float* __restrict__ w, int d0, int d1, int mode,
                                                    float* __restrict__ out, int a0, int b0, float* tile) {
  constexpr int ROW = 32 * KS + 1;
  const int tid = threadIdx.x;
  // load: for each a, the 32*KS contiguous floats w[a][b0 .. b0+31][*] (16 bytes per load on full, aligned tiles)
  if ((a0 + 32 <= d0) && (b0 + 32 <= d1) && ((d1 * KS) & 3) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    for (int i = tid; i < 32 * 8 * KS; i += 256) {
      const int ar = i / (8 * KS), q = i - ar * (8 * KS);
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + ((size_t)(a0 + ar) * d1 + b0) * KS) + q);
      float* t = tile + ar * ROW + q * 4;
      t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
  } else {
    for (int i = tid; i < 32 * 32 * KS; i += 256) {
      const int ar = i / (32 * KS), rem = i - ar * (32 * KS);
      const int a = a0 + ar, b = b0 + rem / KS;
      tile[ar * ROW + rem] = (a < d0 && b < d1) ? w[((size_t)a * d1 + b0) * KS + rem] : 0.f;
    }
  }
  __syncthreads();
  if (mode == 0) {          // out[b][a][j]: for each b, 32*KS contiguous floats over (a, j)
    for (int i = tid; i < 32 * 32 * KS; i += 256) {
      const int br = i / (32 * KS), rem = i - br * (32 * KS);
      const int ar = rem / KS, j = rem - ar * KS;
      const int a = a0 + ar, b = b0 + br;
      if (a < d0 && b < d1) out[((size_t)b * d0 + a0) * KS + rem] = tile[ar * ROW + br * KS + j];
    }
  } else if (mode == 1) {   // out[a][j][b]: for each (a, j), 32 contiguous b
    for (int i = tid; i < 32 * 32 * KS; i += 256) {
      const int br = i & 31, t = i >> 5;
      const int j = t % KS, ar = t / KS;
      const int a = a0 + ar, b = b0 + br;
      if (a < d0 && b < d1) out[((size_t)a * KS + j) * d1 + b] = tile[ar * ROW + br * KS + j];
    }
  } else if (mode == 2) {   // out[b][j][a]: for each (b, j), 32 contiguous a
    for (int i = tid; i < 32 * 32 * KS; i += 256) {
      const int ar = i & 31, t = i >> 5;
      const int j = t % KS, br = t / KS;
      const int a = a0 + ar, b = b0 + br;
      if (a < d0 && b < d1) out[((size_t)b * KS + j) * d0 + a] = tile[ar * ROW + br * KS + j];
    }
  } else {                  // operand image: 32 consecutive c of one (m, j) = one swizzled 128-byte row of a k-block
    // the image's reduction width is Cred rounded UP to whole 32-wide k-blocks (Cred = 39 -> 64): channels beyond Cred are
    // never written and stay zero (the host zero-fills padded images), the GEMM's activation loader reads them as zeros
    const int Cred = ((((mode == 3) ? d1 : d0) + 31) >> 5) << 5;
    const int nkb = KS * Cred / 32;
    // 16 bytes per store: the swizzle permutes whole 16-byte pieces of a row, so 4 consecutive c stay together (the first
    // version stored 4 bytes per thread: 24 store and 12 load instructions per thread and tile, 0.137 ms per training step
    // for the 25 images = 2.9 TB/s of the 6.5 the copy could have).  One task = (m, j, 4 consecutive c); full tiles only.
    const bool full = (a0 + 32 <= d0) && (b0 + 32 <= d1) && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (full) {
      for (int i = tid; i < 32 * KS * 8; i += 256) {
        const int q = i & 7, t = i >> 3;          // q: which 16-byte piece of the 128-byte row
        const int j = t % KS, mr = t / KS;
        float v[4];
        int m, c;
        if (mode == 3) {
          m = a0 + mr; c = b0 + q * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = tile[mr * ROW + (q * 4 + e) * KS + j];
        } else {
          m = b0 + mr; c = a0 + q * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = tile[(q * 4 + e) * ROW + mr * KS + j];
        }
        const int kk = j * Cred + c;
        const int kb = kk >> 5, kcol = kk & 31, r = m & 127, mt = m >> 7;
        const size_t base = ((size_t)mt * nkb + kb) * 8192;   // floats: 2 copies x 4096
        const int off = r * 32 + (((kcol >> 2) ^ (r & 7)) << 2);
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u);
        h.y = __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u);
        h.z = __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u);
        h.w = __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u);
        l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
        *reinterpret_cast<float4*>(out + base + off) = h;
        *reinterpret_cast<float4*>(out + base + 4096 + off) = l;
      }
    } else {
      // edge tiles, and tiles that lie wholly in the padding (the grid covers the PADDED image: rows up to the next multiple of
      // 128, channels up to the next multiple of 32): the tile load put zeros wherever w has no element, so every word of the
      // image is written by exactly one block and no memset is needed
      const int Mpad = ((((mode == 3) ? d0 : d1) + 127) >> 7) << 7;
      for (int i = tid; i < 32 * 32 * KS; i += 256) {
        const int cr = i & 31, t = i >> 5;        // cr: position along c (the fast index of the image row)
        const int j = t % KS, mr = t / KS;
        int m, c;
        float v;
        if (mode == 3) {
          m = a0 + mr; c = b0 + cr;
          v = tile[mr * ROW + cr * KS + j];
        } else {
          m = b0 + mr; c = a0 + cr;
          v = tile[cr * ROW + mr * KS + j];
        }
        if (m >= Mpad || c >= Cred) continue;
        const int kk = j * Cred + c;
        const int kb = kk >> 5, kcol = kk & 31, r = m & 127, mt = m >> 7;
        const size_t base = ((size_t)mt * nkb + kb) * 8192;   // floats: 2 copies x 4096
        const int off = r * 32 + ((((kcol >> 2) ^ (r & 7)) << 2) | (kcol & 3));
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        out[base + off] = h;
        out[base + 4096 + off] = v - h;
      }
    }
  }
}

// tile-grid extents of an item: the source (d0, d1) for the plain re-arrangements, the PADDED image for modes 3 / 4
__host__ __device__ __forceinline__ int permute_ext0(int d0, int mode) {
  return mode < 3 ? d0 : (mode == 3 ? (d0 + 127) / 128 * 128 : (d0 + 31) / 32 * 32);
}
__host__ __device__ __forceinline__ int permute_ext1(int d1, int mode) {
  return mode < 3 ? d1 : (mode == 3 ? (d1 + 31) / 32 * 32 : (d1 + 127) / 128 * 128);
}

template <int KS>
__global__ void __launch_bounds__(256) permute_weight_kernel(const float* __restrict__ w, int d0, int d1, int mode,
                                                             float* __restrict__ out) {
  __shared__ float tile[32 * (32 * KS + 1)];
  permute_weight_tile<KS>(w, d0, d1, mode, out, blockIdx.y * 32, blockIdx.x * 32, tile);
}

// All weight re-arrangements of a training step in ONE launch (25 separate launches cost 0.34 ms per step, mostly launch
// gaps and tails): block -> (item, 32 x 32 tile) through a prefix table in the kernel parameters.
struct PermuteBatch {
  const float* w[VQS_PERMUTE_MAX_ITEMS];
  float* out[VQS_PERMUTE_MAX_ITEMS];
  int d0[VQS_PERMUTE_MAX_ITEMS], d1[VQS_PERMUTE_MAX_ITEMS];
  unsigned char ks[VQS_PERMUTE_MAX_ITEMS], mode[VQS_PERMUTE_MAX_ITEMS];
  int first[VQS_PERMUTE_MAX_ITEMS + 1];   // first block of every item
  int n;
};

__global__ void __launch_bounds__(256) permute_weights_kernel(const __grid_constant__ PermuteBatch pb) {
  __shared__ float tile[32 * (32 * 4 + 1)];
  int it = 0;
  while (it + 1 < pb.n && (int)blockIdx.x >= pb.first[it + 1]) ++it;
  const int local = blockIdx.x - pb.first[it];
  const int d0 = pb.d0[it], d1 = pb.d1[it], mode = pb.mode[it];
  const int nbx = (permute_ext1(d1, mode) + 31) / 32;
  const int a0 = (local / nbx) * 32, b0 = (local % nbx) * 32;
  switch (pb.ks[it]) {
    case 1: permute_weight_tile<1>(pb.w[it], d0, d1, mode, pb.out[it], a0, b0, tile); break;
    case 2: permute_weight_tile<2>(pb.w[it], d0, d1, mode, pb.out[it], a0, b0, tile); break;
    case 3: permute_weight_tile<3>(pb.w[it], d0, d1, mode, pb.out[it], a0, b0, tile); break;
    default: permute_weight_tile<4>(pb.w[it], d0, d1, mode, pb.out[it], a0, b0, tile); break;
  }
}

}  // namespace
}  // namespace vqs

using namespace vqs;

extern "C" int vqs_conv_gemm(const vqs_conv_gemm_desc* d, vqs_stream_t stream) {
  VQS_CHECK_ARG(d != nullptr, "vqs_conv_gemm: NULL descriptor");
  VQS_CHECK_ARG(d->A && d->X && d->out, "vqs_conv_gemm: NULL tensor");
  VQS_CHECK_ARG(d->M > 0 && d->Cred > 0 && d->B > 0 && d->Lin > 0 && d->Lout > 0, "vqs_conv_gemm: bad shape");
  VQS_CHECK_ARG(d->l_div == 1 || d->l_div == 2, "vqs_conv_gemm: l_div must be 1 or 2");
  VQS_CHECK_ARG((long long)d->B * d->Lout < (1ll << 31) && (long long)d->B * d->M * d->Lout < (1ll << 40),
                "vqs_conv_gemm: problem too large");
  VQS_CHECK_ARG(!(d->mask_kind && !d->mask) && !(d->mask2_kind && !d->mask2), "vqs_conv_gemm: mask kind without mask");
  ConvParams p;
  p.d = *d;
  p.Ktot = d->Cred * d->ksz;
  p.Ntot = d->B * d->Lout;
  p.a_vec = (p.Ktot % 4 == 0) && ((reinterpret_cast<uintptr_t>(d->A) & 15) == 0);
  p.divL = FastDiv((uint32_t)d->Lout);
  p.cred_real = d->Cred;
  if (d->a_tap_major == 2 && d->Cred % 32 != 0) {
    // operand image of a layer whose channel count is not a multiple of the 32-wide k-block (the 39 MFCC channels): the image was
    // built for Cred rounded up (vqs_permute_weight), the GEMM runs over the padded width and its activation loader reads the
    // channels that do not exist as zeros (cred_real)
    p.d.Cred = (d->Cred + 31) / 32 * 32;
    p.Ktot = p.d.Cred * d->ksz;
    p.a_vec = (p.Ktot % 4 == 0) && ((reinterpret_cast<uintptr_t>(d->A) & 15) == 0);
  }
  p.cpb = p.d.Cred / 32 > 0 ? p.d.Cred / 32 : 1;
  p.divCpb = FastDiv((uint32_t)p.cpb);
  p.divCred = FastDiv((uint32_t)p.d.Cred);
  p.splits = 1;
  p.kt_per_split = 0;
  p.partial = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->a_tap_major == 2) {
    VQS_CHECK_ARG(d->precision != VQS_PREC_FP32 && conv_tc_supported(p),
                  "vqs_conv_gemm: an operand image (a_tap_major = 2) needs a tensor-core precision");
    count_engine(VQS_ENGINE_CONV_TC);
    return launch_conv_tc(p, d->precision, st);
  }
  if (d->precision != VQS_PREC_FP32 && conv_tc_supported(p)) {
    count_engine(VQS_ENGINE_CONV_TC);
    return launch_conv_tc(p, d->precision, st);
  }
  count_engine(VQS_ENGINE_CONV_CUDACORE);
  // tile choice: big tiles once they fill the machine, small tiles otherwise
  long long big = (long long)((d->M + 127) / 128) * ((p.Ntot + 127) / 128);
  if (d->M > 64 && p.Ntot > 64 && big >= num_sms()) return launch_conv<128, 128>(p, st);
  if (d->M > 64 && p.Ntot > 64 && big * 2 >= num_sms()) return launch_conv<128, 64>(p, st);
  return launch_conv<64, 64>(p, st);
}

extern "C" size_t vqs_wgrad_workspace_bytes(int M, int Cred, int ksz, int B, int La) {
  if (M <= 0 || Cred <= 0 || ksz <= 0 || B <= 0 || La <= 0) return 0;
  WgradPlan pl = plan_wgrad(M, Cred * ksz, B * La);
  WgradPlan pt = plan_wgrad(M, Cred * ksz, B * La, 32, true);
  WgradPlan pp = plan_wgrad(M, Cred * ksz, B * La, 32, true, true);
  if (pp.splits > pt.splits) pt = pp;
  int s = pl.splits > pt.splits ? pl.splits : pt.splits;
  size_t need = s > 1 ? (size_t)s * M * Cred * ksz * sizeof(float) : 0;
  if (M % 128 == 0 && Cred % 128 == 0) {   // the TMA-fed kernel always goes through the workspace (tile blocks + reduce)
    WgradPlan pm = plan_wgrad(M, Cred * ksz, B * ((La + 31) / 32) * 32, 32, true);
    size_t nt = (size_t)pm.splits * M * Cred * ksz * sizeof(float);
    if (nt > need) need = nt;
  }
  return need;
}

extern "C" int vqs_wgrad_gemm(const vqs_wgrad_desc* d, void* workspace, size_t workspace_bytes, vqs_stream_t stream) {
  VQS_CHECK_ARG(d != nullptr, "vqs_wgrad_gemm: NULL descriptor");
  VQS_CHECK_ARG(d->Aact && d->X && d->dW, "vqs_wgrad_gemm: NULL tensor");
  VQS_CHECK_ARG(d->M > 0 && d->Cred > 0 && d->B > 0 && d->La > 0 && d->Lx > 0, "vqs_wgrad_gemm: bad shape");
  VQS_CHECK_ARG((long long)d->B * d->La < (1ll << 31), "vqs_wgrad_gemm: problem too large");
  WgradParams p;
  p.d = *d;
  p.Nw = d->Cred * d->ksz;
  p.Kred = d->B * d->La;
  cudaStream_t st = (cudaStream_t)stream;
  if (wgrad_tma_supported(p)) {
    // both operands K-major in HBM: TMA boxes straight into the UMMA layout, threads only derive the lo tiles
    WgradPlan pm = plan_wgrad(d->M, p.Nw, wgrad_tma_kblocks(p) * 32, 32, true);
    p.splits = pm.splits;
    p.kt_per_split = pm.kt_per_split;
    const size_t need_t = (size_t)pm.splits * d->M * p.Nw * sizeof(float);
    if (need_t > workspace_bytes || !workspace) {
      set_error("vqs_wgrad_gemm: workspace %zu < %zu", workspace_bytes, need_t);
      return VQS_ERR_WORKSPACE;
    }
    count_engine(VQS_ENGINE_WGRAD_TMA);
    return launch_wgrad_tma(p, (float*)workspace, st);
  }
  const bool tc = d->precision != VQS_PREC_FP32 && wgrad_tc_supported(p);
  WgradPlan pl = tc ? plan_wgrad(d->M, p.Nw, p.Kred, 32, true, wgrad_tc_pairs(p)) : plan_wgrad(d->M, p.Nw, p.Kred);
  p.splits = pl.splits;
  p.kt_per_split = pl.kt_per_split;
  p.divLa = FastDiv((uint32_t)d->La);
  size_t need = pl.splits > 1 ? (size_t)pl.splits * d->M * p.Nw * sizeof(float) : 0;
  if (need > workspace_bytes || (need && !workspace)) {
    set_error("vqs_wgrad_gemm: workspace %zu < %zu", workspace_bytes, need);
    return VQS_ERR_WORKSPACE;
  }
  p.partial = pl.splits > 1 ? (float*)workspace : nullptr;
  int e;
  count_engine(tc ? VQS_ENGINE_WGRAD_TC : VQS_ENGINE_WGRAD_CUDACORE);
  if (tc) e = launch_wgrad_tc(p, d->precision, st);
  else if (pl.bm == 128 && pl.bn == 128) e = launch_wgrad<128, 128>(p, st);
  else if (pl.bm == 128) e = launch_wgrad<128, 64>(p, st);
  else if (pl.bn == 128) e = launch_wgrad<64, 128>(p, st);
  else e = launch_wgrad<64, 64>(p, st);
  if (e) return e;
  if (pl.splits > 1) {
    long long n = (long long)d->M * p.Nw;
    long long blocks = (n + 255) / 256;
    if (n % 4 == 0 && ((reinterpret_cast<uintptr_t>(p.partial) | reinterpret_cast<uintptr_t>(d->dW)) & 15) == 0) {
      const long long b4 = (n / 4 + 255) / 256;
      VQS_CUDA(launch_pdl(splitk_reduce4_kernel, dim3((unsigned)(b4 < 8 * num_sms() ? b4 : 8 * num_sms())), dim3(256), 0, st,
                          reinterpret_cast<const float4*>(p.partial), pl.splits, n / 4, reinterpret_cast<float4*>(d->dW),
                          d->accumulate));
    } else {
      VQS_CUDA(launch_pdl(splitk_reduce_kernel, dim3((unsigned)(blocks < 8 * num_sms() ? blocks : 8 * num_sms())), dim3(256), 0,
                          st, p.partial, pl.splits, n, d->dW, d->accumulate));
    }
    VQS_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int vqs_bias_grad(const float* g, int B, int M, int L, float* db, int accumulate, vqs_stream_t stream) {
  VQS_CHECK_ARG(g && db && B > 0 && M > 0 && L > 0, "vqs_bias_grad: bad arguments");
  const int vec4 = (L % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) ? 1 : 0;
  VQS_CHECK_ARG((long long)B * L < (1ll << 31), "vqs_bias_grad: B * L too large");
  VQS_CUDA(launch_pdl(bias_grad_kernel, dim3(M), dim3(256), 0, (cudaStream_t)stream, g, B, M, L, db, accumulate,
                      FastDiv((uint32_t)(vec4 ? L / 4 : L)), vec4));
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_permute_weight(const float* w, int d0, int d1, int k, int mode, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(w && out && d0 > 0 && d1 > 0 && k >= 1 && k <= 4 && mode >= 0 && mode <= 4,
                "vqs_permute_weight: bad arguments (kernel size 1..4, mode 0..4)");
  if (mode >= 3) {
    const int M = mode == 3 ? d0 : d1, Cred = mode == 3 ? d1 : d0;
    (void)M;
    (void)Cred;   // (rows beyond M and channels beyond Cred are written as zeros by the kernel: its grid covers the padded image)
  }
  dim3 grid((permute_ext1(d1, mode) + 31) / 32, (permute_ext0(d0, mode) + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  switch (k) {
    case 1: permute_weight_kernel<1><<<grid, 256, 0, st>>>(w, d0, d1, mode, out); break;
    case 2: permute_weight_kernel<2><<<grid, 256, 0, st>>>(w, d0, d1, mode, out); break;
    case 3: permute_weight_kernel<3><<<grid, 256, 0, st>>>(w, d0, d1, mode, out); break;
    default: permute_weight_kernel<4><<<grid, 256, 0, st>>>(w, d0, d1, mode, out); break;
  }
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_permute_weights(const vqs_permute_item* items, int n, vqs_stream_t stream) {
  VQS_CHECK_ARG(items && n >= 1 && n <= VQS_PERMUTE_MAX_ITEMS, "vqs_permute_weights: 1..%d items", VQS_PERMUTE_MAX_ITEMS);
  cudaStream_t st = (cudaStream_t)stream;
  PermuteBatch pb;
  int blocks = 0;
  for (int i = 0; i < n; ++i) {
    const vqs_permute_item& q = items[i];
    VQS_CHECK_ARG(q.w && q.out && q.d0 > 0 && q.d1 > 0 && q.k >= 1 && q.k <= 4 && q.mode >= 0 && q.mode <= 4,
                  "vqs_permute_weights: bad item %d (kernel size 1..4, mode 0..4)", i);
    if (q.mode >= 3) {
      const int M = q.mode == 3 ? q.d0 : q.d1, Cred = q.mode == 3 ? q.d1 : q.d0;
      (void)M;
      (void)Cred;   // (padding is written by the kernel)
    }
    pb.w[i] = q.w; pb.out[i] = q.out; pb.d0[i] = q.d0; pb.d1[i] = q.d1;
    pb.ks[i] = (unsigned char)q.k; pb.mode[i] = (unsigned char)q.mode;
    pb.first[i] = blocks;
    blocks += ((permute_ext0(q.d0, q.mode) + 31) / 32) * ((permute_ext1(q.d1, q.mode) + 31) / 32);
  }
  pb.first[n] = blocks;
  pb.n = n;
  permute_weights_kernel<<<blocks, 256, 0, st>>>(pb);
  VQS_LAUNCH_CHECK();
  return 0;
}
