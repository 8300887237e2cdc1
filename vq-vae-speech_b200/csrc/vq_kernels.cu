// VQ bottleneck kernels (CUDA-core fp32 path): nearest-code search + per-code statistics, EMA update,
// quantise / straight-through / loss, backward.  Replaces the ATen op chain of
//   /root/reference/src/models/vector_quantizer_ema.py:101-179 and vector_quantizer.py:88-150.
//
// Data layout in HBM: z / out / grads keep the caller's (B, D, T) layout; VQ rows are formed on the fly in shared memory
// in the reference's (D, T, B) flattening (flat f = d*T*B + t*B + b), so no permuted copy ever touches HBM.
#include <math.h>
#include <stdlib.h>

#include "vqs_common.cuh"

namespace vqs {
namespace {

constexpr int TILE = 128;  // rows per tile
constexpr int NT = 128;    // threads per CTA (4 warps)
constexpr int NWARP = NT / 32;
constexpr int RL = 4;      // row lanes per warp
constexpr int CL = 8;      // code lanes per warp
constexpr int RT = 8;      // rows per thread: a warp covers RL*RT = 32 rows

struct AssignParams {
  const float* z;
  const float* cb;
  int64_t* idx;
  float* partials;  // [grid][K*(D+1)] (smem_stats) or the final stats vector (global atomics)
  float* dmin2;
  float* dist;
  long long N;
  int layout, B, D, T, K;
  int Dp;        // smem row stride (floats)
  int ntiles;
  int nchunks;   // code chunks of CL*CT codes
  int resident;  // whole codebook resident in smem
  int smem_stats;
  FastDiv divD;
};

__device__ __forceinline__ void zero_smem(float* p, int n, int tid) {
  for (int i = tid; i < n; i += NT) p[i] = 0.f;
}

// Loads rows [r0, r0+TILE) into xs[row*Dp + j]; rows >= N are zero.
template <int VEC>
__device__ __forceinline__ void load_tile(const AssignParams& p, float* xs, long long r0, int tid) {
  const int D = p.D, Dp = p.Dp;
  long long rows_left = p.N - r0;
  int rows = rows_left < TILE ? (int)rows_left : TILE;
  if (rows < TILE) {
    for (int i = tid + rows * Dp; i < TILE * Dp; i += NT) xs[i] = 0.f;
  }
  if (p.layout == VQS_LAYOUT_FLAT_ND) {
    const float* src = p.z + r0 * D;
    if (VEC == 4 && ((reinterpret_cast<uintptr_t>(p.z) & 15) == 0)) {
      const int cpr = D >> 2;  // 16-byte chunks per row
      const int total = rows * cpr;
      for (int c = tid; c < total; c += NT) {
        uint32_t row = p.divD.div((uint32_t)c << 2);
        int c4 = c - row * cpr;
        cp_async16(&xs[row * Dp + (c4 << 2)], src + (size_t)c * 4);
      }
      cp_async_commit();
      cp_async_wait<0>();
    } else {
      const int total = rows * D;
      for (int e = tid; e < total; e += NT) {
        uint32_t row, j;
        p.divD.divmod((uint32_t)e, row, j);
        xs[row * Dp + j] = __ldg(src + e);
      }
    }
    return;
  }
  // BDT_AS_DTB: flat f = d*P + t*B + b  <-  z[(b*D + d)*T + t].  Walk the tile plane by plane (fixed d) with t fastest so
  // that global reads are coalesced along t.
  const int B = p.B, T = p.T;
  const long long P = (long long)T * B;
  const long long f0 = r0 * D;
  const long long f1 = f0 + (long long)rows * D;
  long long d_first = f0 / P;  // one 64-bit division per tile per thread
  long long d_last = (f1 - 1) / P;
  for (long long d = d_first; d <= d_last; ++d) {
    long long base = d * P;
    int p0 = (int)((f0 > base ? f0 : base) - base);
    int p1 = (int)((f1 < base + P ? f1 : base + P) - base);
    int t0 = p0 / B;
    int t1 = (p1 + B - 1) / B;
    int span = t1 - t0;
    int count = span * B;
    int qs = NT / span, rs = NT - qs * span;
    int b = tid / span, tt = tid - b * span;
    int foff = (int)(base - f0);  // may be negative
    // eight elements per pass: every load of a pass is issued before the first shared-memory store (one element at a time
    // left a single load in flight per thread: 38 us of pure latency for the 12 tiles of the N = 1536 training step)
    constexpr int U = 8;
    for (int e = tid; e < count; e += NT * U) {
      float val[U];
      int dst[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        dst[u] = -1;
        val[u] = 0.f;
        if (e + u * NT < count) {
          int t = t0 + tt;
          int pp = t * B + b;
          if (pp >= p0 && pp < p1) {
            uint32_t fl = (uint32_t)(foff + pp);
            uint32_t row, j;
            p.divD.divmod(fl, row, j);
            dst[u] = (int)(row * Dp + j);
            val[u] = __ldg(p.z + ((size_t)b * D + d) * T + t);
          }
          tt += rs;
          b += qs;
          if (tt >= span) {
            tt -= span;
            ++b;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (dst[u] >= 0) xs[dst[u]] = val[u];
    }
  }
}

template <int VEC, int CT>
__device__ __forceinline__ void load_codes(const AssignParams& p, float* es, float* se, int k0, int nk, int tid) {
  // es[kk*Dp + j] <- codebook[(k0+kk)*D + j] for kk < nk (zero rows beyond K), se[kk] = sum_j e^2 (+inf beyond K)
  const int D = p.D, Dp = p.Dp;
  constexpr int U = 8;                                  // loads of a pass in flight together (see load_tile)
  for (int e0 = tid; e0 < nk * D; e0 += NT * U) {
    float val[U];
    int dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * NT;
      dst[u] = -1;
      val[u] = 0.f;
      if (e < nk * D) {
        uint32_t kk, j;
        p.divD.divmod((uint32_t)e, kk, j);
        const int k = k0 + kk;
        dst[u] = (int)(kk * Dp + j);
        if (k < p.K) val[u] = __ldg(p.cb + (size_t)k * D + j);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] >= 0) es[dst[u]] = val[u];
  }
  __syncthreads();
  for (int kk = tid; kk < nk; kk += NT) {
    float s = 0.f;
    for (int j = 0; j < D; ++j) {
      float v = es[kk * Dp + j];
      s = fmaf(v, v, s);
    }
    se[kk] = (k0 + kk < p.K) ? s : INFINITY;
  }
}

template <int VEC, int CT, bool SECOND>
__global__ void __launch_bounds__(NT) vq_assign_kernel(const AssignParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kl = lane & (CL - 1), rl = lane >> 3;
  const int D = p.D, Dp = p.Dp, K = p.K;
  constexpr int CHUNK = CL * CT;
  const int ncodes_smem = p.resident ? p.nchunks * CHUNK : CHUNK;
  float* xs = smem;
  float* es = xs + TILE * Dp;
  float* se = es + ncodes_smem * Dp;
  int* sidx = reinterpret_cast<int*>(se + ncodes_smem);
  float* dws = reinterpret_cast<float*>(sidx + TILE);  // [K*D] then cnt[K]   (smem_stats only)
  float* cnts = dws + K * D;

  if (p.resident) load_codes<VEC, CT>(p, es, se, 0, ncodes_smem, tid);
  if (p.smem_stats) zero_smem(dws, K * (D + 1), tid);
  __syncthreads();

  // run-length accumulator of the stats phase (D <= 64): lane holds columns lane and lane+32 of code cur_k
  int cur_k = -1;
  float run0 = 0.f, run1 = 0.f, runc = 0.f;

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long r0 = (long long)tile * TILE;
    const long long rows_left = p.N - r0;
    const int rows = rows_left < TILE ? (int)rows_left : TILE;
    load_tile<VEC>(p, xs, r0, tid);
    __syncthreads();

    const int rbase = warp * 32 + rl;  // thread rows: rbase + RL*i
    // ---- sum_j x^2, split over the CL code lanes, butterfly-reduced (bitwise identical on all lanes) ----
    float sx[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) sx[i] = 0.f;
    for (int j = kl * VEC; j < D; j += CL * VEC) {
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float* xr = &xs[(rbase + RL * i) * Dp + j];
        if (VEC == 4) {
          float4 v = *reinterpret_cast<const float4*>(xr);
          sx[i] = fmaf(v.x, v.x, sx[i]);
          sx[i] = fmaf(v.y, v.y, sx[i]);
          sx[i] = fmaf(v.z, v.z, sx[i]);
          sx[i] = fmaf(v.w, v.w, sx[i]);
        } else {
          float v = xr[0];
          sx[i] = fmaf(v, v, sx[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      sx[i] += __shfl_xor_sync(0xffffffffu, sx[i], 1);
      sx[i] += __shfl_xor_sync(0xffffffffu, sx[i], 2);
      sx[i] += __shfl_xor_sync(0xffffffffu, sx[i], 4);
    }

    float bd[RT], b2[RT];
    int bk[RT];
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      bd[i] = INFINITY;
      b2[i] = INFINITY;
      bk[i] = 0x7fffffff;
    }

    for (int c = 0; c < p.nchunks; ++c) {
      const float* ec = es;
      const float* sec = se;
      if (p.resident) {
        ec = es + (size_t)c * CHUNK * Dp;
        sec = se + c * CHUNK;
      } else {
        __syncthreads();
        load_codes<VEC, CT>(p, es, se, c * CHUNK, CHUNK, tid);
        __syncthreads();
      }
      float acc[RT][CT];
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int q = 0; q < CT; ++q) acc[i][q] = 0.f;

      if (VEC == 4) {
#pragma unroll 2
        for (int j = 0; j < D; j += 4) {
          float4 xv[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) xv[i] = *reinterpret_cast<const float4*>(&xs[(rbase + RL * i) * Dp + j]);
#pragma unroll
          for (int q = 0; q < CT; ++q) {
            float4 ev = *reinterpret_cast<const float4*>(&ec[(kl + CL * q) * Dp + j]);
#pragma unroll
            for (int i = 0; i < RT; ++i) {
              float a = acc[i][q];
              a = fmaf(xv[i].x, ev.x, a);
              a = fmaf(xv[i].y, ev.y, a);
              a = fmaf(xv[i].z, ev.z, a);
              a = fmaf(xv[i].w, ev.w, a);
              acc[i][q] = a;
            }
          }
        }
      } else {
        for (int j = 0; j < D; ++j) {
          float xv[RT];
#pragma unroll
          for (int i = 0; i < RT; ++i) xv[i] = xs[(rbase + RL * i) * Dp + j];
#pragma unroll
          for (int q = 0; q < CT; ++q) {
            float ev = ec[(kl + CL * q) * Dp + j];
#pragma unroll
            for (int i = 0; i < RT; ++i) acc[i][q] = fmaf(xv[i], ev, acc[i][q]);
          }
        }
      }
      // d = (sx + se) - 2*dot, same expression shape as the reference (ema.py:109-111); codes ascend with q then c
#pragma unroll
      for (int q = 0; q < CT; ++q) {
        const int kk = kl + CL * q;
        const int k = c * CHUNK + kk;
        const float sek = sec[kk];
#pragma unroll
        for (int i = 0; i < RT; ++i) {
          float dd = __fsub_rn(__fadd_rn(sx[i], sek), __fmul_rn(2.0f, acc[i][q]));
          if (p.dist != nullptr && k < K) {
            long long r = r0 + rbase + RL * i;
            if (r < p.N) p.dist[r * K + k] = dd;
          }
          if (dd < bd[i]) {
            if (SECOND) b2[i] = bd[i];
            bd[i] = dd;
            bk[i] = k;
          } else if (SECOND && dd < b2[i]) {
            b2[i] = dd;
          }
        }
      }
    }
    // ---- argmin across the CL code lanes: lowest distance, lowest index on ties (torch.argmin) ----
#pragma unroll
    for (int i = 0; i < RT; ++i) {
#pragma unroll
      for (int o = 1; o < CL; o <<= 1) {
        float od = __shfl_xor_sync(0xffffffffu, bd[i], o);
        int ok = __shfl_xor_sync(0xffffffffu, bk[i], o);
        float o2 = SECOND ? __shfl_xor_sync(0xffffffffu, b2[i], o) : 0.f;
        bool take = (od < bd[i]) || (od == bd[i] && ok < bk[i]);
        if (SECOND) {
          float hi = take ? bd[i] : od;  // the loser of the two bests
          float s = fminf(b2[i], o2);
          b2[i] = fminf(s, hi);
        }
        if (take) {
          bd[i] = od;
          bk[i] = ok;
        }
      }
      if (kl == 0) {
        int r = rbase + RL * i;
        sidx[r] = bk[i] < K ? bk[i] : K - 1;  // only reachable when every distance is NaN/inf: stay in bounds
        if (SECOND && p.dmin2 != nullptr && r0 + r < p.N) {
          p.dmin2[(r0 + r) * 2 + 0] = bd[i];
          p.dmin2[(r0 + r) * 2 + 1] = b2[i];
        }
      }
    }
    __syncthreads();
    if (tid < rows) p.idx[r0 + tid] = (int64_t)sidx[tid];

    // ---- per-code statistics: warp w owns codes k % NWARP == w, rows visited in ascending order (deterministic) ----
    if (p.smem_stats) {
      for (int base = 0; base < rows; base += 32) {
        int r = base + lane;
        int k = (r < rows) ? sidx[r] : -1;
        unsigned m = __ballot_sync(0xffffffffu, k >= 0 && (k & (NWARP - 1)) == warp);
        while (m) {
          int bpos = __ffs(m) - 1;
          m &= m - 1;
          int kk = __shfl_sync(0xffffffffu, k, bpos);
          const float* xr = &xs[(base + bpos) * Dp];
          if (D <= 64) {
            if (kk != cur_k) {
              if (cur_k >= 0) {
                if (lane < D) dws[cur_k * D + lane] += run0;
                if (lane + 32 < D) dws[cur_k * D + lane + 32] += run1;
                if (lane == 0) cnts[cur_k] += runc;
              }
              cur_k = kk;
              run0 = run1 = runc = 0.f;
            }
            if (lane < D) run0 += xr[lane];
            if (lane + 32 < D) run1 += xr[lane + 32];
            runc += 1.f;
          } else {
            for (int j = lane; j < D; j += 32) dws[kk * D + j] += xr[j];
            if (lane == 0) cnts[kk] += 1.f;
          }
        }
      }
    } else {
      // large codebooks: low contention, straight global atomics into the (pre-zeroed) stats vector
      for (int r = warp; r < rows; r += NWARP) {
        int kk = sidx[r];
        const float* xr = &xs[r * Dp];
        for (int j = lane; j < D; j += 32) atomicAdd(&p.partials[K + (size_t)kk * D + j], xr[j]);
        if (lane == 0) atomicAdd(&p.partials[kk], 1.f);
      }
    }
    __syncthreads();
  }

  if (p.smem_stats) {
    if (D <= 64 && cur_k >= 0) {
      if (lane < D) dws[cur_k * D + lane] += run0;
      if (lane + 32 < D) dws[cur_k * D + lane + 32] += run1;
      if (lane == 0) cnts[cur_k] += runc;
    }
    __syncthreads();
    float* out = p.partials + (size_t)blockIdx.x * K * (D + 1);
    for (int i = tid; i < K; i += NT) out[i] = cnts[i];
    for (int i = tid; i < K * D; i += NT) out[K + i] = dws[i];
  }
}

__global__ void stats_reduce_kernel(const float* __restrict__ partials, int G, int S, float* __restrict__ stats) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float a = 0.f;
  for (int g = 0; g < G; ++g) a += partials[(size_t)g * S + s];
  stats[s] = a;
}

__global__ void one_hot_kernel(const int64_t* __restrict__ idx, long long N, int K, float* __restrict__ enc) {
  long long total = N * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long n = i / K;
    int k = (int)(i - n * K);
    enc[i] = (idx[n] == k) ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// EMA update (vector_quantizer_ema.py:143-156).
//   kernel 1 (one block): cs <- cs*decay + (1-decay)*counts ; n = sum(cs) ; cs <- (cs + eps)/(n + K*eps)*n   (in place)
//   kernel 2: ema_w <- ema_w*decay + (1-decay)*dw ; embedding <- ema_w / cs[:, None]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ema_cluster_size_kernel(float* __restrict__ cs, const float* __restrict__ stats,
                                                               float decay, float omd, float eps, float k_eps, int K) {
  __shared__ float red[256];
  const int tid = threadIdx.x;
  float part = 0.f;
  for (int k = tid; k < K; k += 256) part += __fadd_rn(__fmul_rn(cs[k], decay), __fmul_rn(omd, stats[k]));
  red[tid] = part;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  const float n = red[0];
  const float denom = __fadd_rn(n, k_eps);
  for (int k = tid; k < K; k += 256) {  // element k is read and written by this thread only
    float c1 = __fadd_rn(__fmul_rn(cs[k], decay), __fmul_rn(omd, stats[k]));
    cs[k] = __fmul_rn(__fdiv_rn(__fadd_rn(c1, eps), denom), n);
  }
}
__global__ void __launch_bounds__(256) ema_embedding_kernel(const float* __restrict__ cs, float* __restrict__ ema_w,
                                                            float* __restrict__ emb, const float* __restrict__ dw,
                                                            float decay, float omd, int K, int D) {
  const long long total = (long long)K * D;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    int k = (int)(i / D);
    float w = __fadd_rn(__fmul_rn(ema_w[i], decay), __fmul_rn(omd, dw[i]));
    ema_w[i] = w;
    emb[i] = __fdiv_rn(w, cs[k]);
  }
}

// ------------------------------------------------------------------------------------------------
// quantise / straight-through / SSE  and  backward: element-wise in the caller's memory order
// ------------------------------------------------------------------------------------------------
struct EwParams {
  const float* z;
  const float* g;  // backward: upstream gradient of the quantised output
  const float* gl; // backward: device scalar, upstream gradient of vq_loss
  const int64_t* idx;
  const float* cb;
  float* out;
  double* sse_partials;
  float coef;
  long long total;  // N*D
  int layout, B, D, T, K;
  int gather;       // forward only: out = codebook[idx] exactly, z is not read (no straight-through arithmetic, no SSE)
  FastDiv divD, divT;
  // blocked (B, D, T) order (BLK): a warp = 16 batch items x 2 groups of 4 frames of one channel
  FastDiv divNTP;   // NTP = ceil(T / 8) frame pairs
  long long nvec_blk;
};

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
#ifdef VQS_EW_PLAIN_LD
  return __ldg(reinterpret_cast<const float4*>(p));
#endif
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
#ifdef VQS_EW_PLAIN_ST
  *reinterpret_cast<float4*>(p) = v;
  return;
#endif
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Element-wise pass in the caller's memory order.  VEC = 4: each thread owns 4 consecutive floats (flat: 4 columns of one
// row, one index load; (B, D, T): 4 consecutive t of one (b, d) run = 4 different rows).  The codebook sits in shared
// memory with a row stride chosen per layout so that the gather is bank-conflict-free: D + 4 (16-byte aligned float4
// reads) for flat rows, D + 1 for (B, D, T), where the lanes of a warp hit the same column of different codes (the
// first build used stride D there: 32-way conflicts, 37 M conflict cycles per launch, profiles/r01c).
//
// BLK ((B, D, T) with B >= 4 D, VEC = 4): in memory order the 32 float4 of a warp are 128 frames of one (b, d), i.e. 128
// different VQ rows whose indices lie B / D * 8 bytes apart -- one 32-byte sector of idx per 4 bytes of data, eight times
// the traffic of the tensor itself (measured: 1.19 ms against 0.41 ms for flat rows at 2^22 rows).  The blocked order gives a
// warp 16 consecutive batch items x 8 frames of one channel instead: 8 rows (mostly one idx sector each, broadcast to the
// 16 lanes that share it) against 16 fully used 32-byte sectors of data; consecutive warps walk the frames, then the
// channels, of the same 16 batch items, so that HBM still sees long runs.
template <bool BWD, bool SMEM_CB, bool FLAT, int VEC, bool BLK>
__global__ void __launch_bounds__(256) vq_elementwise_kernel(const EwParams p) {
  extern __shared__ __align__(16) float cbs[];
  __shared__ double wred[8];
  const int D = p.D;
  const int Dp = SMEM_CB ? (FLAT ? D + 4 : D + 1) : D;
  if (SMEM_CB) {
    for (int i = threadIdx.x; i < p.K * D; i += 256) {
      uint32_t k, j;
      p.divD.divmod((uint32_t)i, k, j);
      cbs[k * Dp + j] = __ldg(p.cb + i);
    }
    __syncthreads();
  }
  const float* cb = SMEM_CB ? cbs : p.cb;
  float c = 0.f;
  if (BWD) c = p.gl[0] * p.coef;
  float sse = 0.f;
  const long long nvec = BLK ? p.nvec_blk : p.total / VEC;
  const long long stride = (long long)gridDim.x * 256;
  const int TB = p.T * p.B;
  // element offset of vector iv (-1: nothing there)
  auto offset_of = [&](const long long iv) -> long long {
    if (!BLK) return iv * VEC;
    const uint32_t task = (uint32_t)(iv >> 5), ln = (uint32_t)iv & 31u;
    uint32_t rest, tp, b16, d;
    p.divNTP.divmod(task, rest, tp);
    p.divD.divmod(rest, b16, d);
    const uint32_t t = (tp * 2 + (ln & 1u)) * 4, b = b16 * 16 + (ln >> 1);
    if (t >= (uint32_t)p.T || b >= (uint32_t)p.B) return -1;
    return (long long)((b * (uint32_t)D + d) * (uint32_t)p.T + t);
  };
  // one vector of VEC consecutive floats at offset o, inputs already in registers
  auto process = [&](const long long o, const float (&x)[VEC], const float (&g)[VEC]) {
    float q[VEC], r[VEC];
    if (FLAT) {
      uint32_t row, j;
      p.divD.divmod((uint32_t)o, row, j);
      const int k = (int)__ldg(p.idx + row);
      if (VEC == 4 && SMEM_CB) {
        float4 qv = *reinterpret_cast<const float4*>(&cb[k * Dp + j]);
        q[0] = qv.x; q[1] = qv.y; q[2] = qv.z; q[3] = qv.w;
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) q[e] = SMEM_CB ? cb[k * Dp + j + e] : __ldg(cb + (size_t)k * D + j + e);
      }
    } else {
      uint32_t bd, t, b, d;
      p.divT.divmod((uint32_t)o, bd, t);
      p.divD.divmod(bd, b, d);
      const uint32_t f0 = d * TB + t * p.B + b;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        uint32_t row, j;
        p.divD.divmod(f0 + e * p.B, row, j);
        const int k = (int)__ldg(p.idx + row);
        q[e] = SMEM_CB ? cb[k * Dp + j] : __ldg(cb + (size_t)k * D + j);
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (BWD) {
        const float df = __fsub_rn(x[e], q[e]);
        r[e] = fmaf(c, df, g[e]);
        sse = fmaf(df, df, sse);     // (x - q)^2: the commitment / codebook loss, when the caller asked for it here
      } else if (p.gather) {
        r[e] = q[e];
      } else {
        float df = __fsub_rn(q[e], x[e]);
        r[e] = __fadd_rn(x[e], df);  // inputs + (quantized - inputs).detach()  (ema.py:169)
        sse = fmaf(df, df, sse);
      }
    }
    if (VEC == 4) stg_stream4(p.out + o, make_float4(r[0], r[1], r[2], r[3]));
    else p.out[o] = r[0];
  };
  auto load = [&](const long long o, float (&x)[VEC], float (&g)[VEC]) {
    if (!BWD && p.gather) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) x[e] = 0.f;
      return;
    }
    if (VEC == 4) {
      float4 xv = ldg_stream4(p.z + o);
      x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
      if (BWD) {
        float4 gv = ldg_stream4(p.g + o);
        g[0] = gv.x; g[1] = gv.y; g[2] = gv.z; g[3] = gv.w;
      }
    } else {
      x[0] = __ldg(p.z + o);
      if (BWD) g[0] = __ldg(p.g + o);
    }
  };
  // The forward pass moves fewer bytes per vector than the backward pass (no upstream gradient): it keeps four vectors per
  // thread in flight to cover the HBM latency (one in flight: 80 % of the HBM peak at 64 resident warps).
  constexpr int UNR = BWD ? 1 : 4;
  for (long long iv = blockIdx.x * 256ll + threadIdx.x; iv < nvec; iv += stride * UNR) {
    float x[UNR][VEC], g[UNR][VEC];
    long long off[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      off[u] = (iv + u * stride < nvec) ? offset_of(iv + u * stride) : -1;
      if (off[u] >= 0) load(off[u], x[u], g[u]);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (off[u] >= 0) process(off[u], x[u], g[u]);
  }
  if (p.sse_partials != nullptr) {
    double s = warp_sum((double)sse);
    if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (int w = 0; w < 8; ++w) a += wred[w];
      p.sse_partials[blockIdx.x] = a;
    }
  }
}

// (B, 64, T) tensors with B % 64 == 0 and T % 4 == 0: a VQ row is the 64 batch items of one (d, t), so the index of a row
// serves 64 elements that lie D * T floats apart.  CTA tile = 64 batch items x 32 frames x DD channels: the DD * 32 indices
// are staged in shared memory (double-buffered: one barrier per tile), the data moves as full 128-byte lines (a warp = 4
// batch items x 32 frames), every load of a tile is issued before the barrier.  Tile order: frame blocks, then channel
// groups, of the same 64 batch items, so that HBM sees long runs.  Same arithmetic as vq_elementwise_kernel.
template <bool BWD, int DD>
__global__ void __launch_bounds__(256) vq_elementwise_bdt_tile_kernel(const EwParams p, const int ntiles,
                                                                      const FastDiv divNTB, const FastDiv divDG) {
  extern __shared__ __align__(16) float cbs[];
  __shared__ double wred[8];
  __shared__ __align__(16) int sidx[2][DD * 32];
  constexpr int D = 64, Dp = D + 1;
  for (int i = threadIdx.x; i < p.K * D; i += 256) cbs[(i >> 6) * Dp + (i & 63)] = __ldg(p.cb + i);
  const int T = p.T, Q = p.B >> 6;
  const int tg = threadIdx.x & 7, bq = threadIdx.x >> 3;       // 4-frame group, batch item (and + 32)
  float c = 0.f;
  if (BWD) c = p.gl[0] * p.coef;
  float sse = 0.f;
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    uint32_t rest, tb, b64, dg;
    divNTB.divmod((uint32_t)tile, rest, tb);
    divDG.divmod(rest, b64, dg);
    const int t0 = (int)tb * 32, d0 = (int)dg * DD;
    int kreg = 0;
    if (threadIdx.x < DD * 32) {
      const int dd = threadIdx.x >> 5, t = t0 + (threadIdx.x & 31);
      if (t < T) kreg = (int)__ldg(p.idx + ((long long)(d0 + dd) * T + t) * Q + b64);
    }
    const int t = t0 + tg * 4;
    const bool live = t < T;
    float4 xv[DD][2], gv[BWD ? DD : 1][2];
    const size_t o0 = ((size_t)(b64 * 64 + bq) * D + d0) * T + t;    // + (32 h * D + dd) * T
    if (live) {
#pragma unroll
      for (int dd = 0; dd < DD; ++dd)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const size_t o = o0 + (size_t)(32 * h * D + dd) * T;
          xv[dd][h] = (BWD || !p.gather) ? ldg_stream4(p.z + o) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (BWD) gv[dd][h] = ldg_stream4(p.g + o);
        }
    }
    if (threadIdx.x < DD * 32) sidx[buf][threadIdx.x] = kreg;
    __syncthreads();
    if (live) {
#pragma unroll
      for (int dd = 0; dd < DD; ++dd) {
        const int4 k4 = *reinterpret_cast<const int4*>(&sidx[buf][dd * 32 + tg * 4]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int j = bq + 32 * h;
          const float q[4] = {cbs[k4.x * Dp + j], cbs[k4.y * Dp + j], cbs[k4.z * Dp + j], cbs[k4.w * Dp + j]};
          const float x[4] = {xv[dd][h].x, xv[dd][h].y, xv[dd][h].z, xv[dd][h].w};
          float r[4];
          if (BWD) {
            const float g[4] = {gv[dd][h].x, gv[dd][h].y, gv[dd][h].z, gv[dd][h].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float df = __fsub_rn(x[e], q[e]);
              r[e] = fmaf(c, df, g[e]);
              sse = fmaf(df, df, sse);
            }
          } else if (p.gather) {
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = q[e];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float df = __fsub_rn(q[e], x[e]);
              r[e] = __fadd_rn(x[e], df);                      // inputs + (quantized - inputs).detach()  (ema.py:169)
              sse = fmaf(df, df, sse);
            }
          }
          stg_stream4(p.out + o0 + (size_t)(32 * h * D + dd) * T, make_float4(r[0], r[1], r[2], r[3]));
        }
      }
    }
  }
  if (p.sse_partials != nullptr) {
    double sd = warp_sum((double)sse);
    if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (int w = 0; w < 8; ++w) a += wred[w];
      p.sse_partials[blockIdx.x] = a;
    }
  }
}

// Flat (N, 64) rows, the same idea: CTA tile = 16 * RPT consecutive rows (RPT float4 per thread, a warp = two whole rows),
// the tile's indices arrive as one coalesced load and are staged in shared memory, all loads of a tile are in flight before
// the barrier.  Same arithmetic as vq_elementwise_kernel.
template <bool BWD, int RPT>
__global__ void __launch_bounds__(256) vq_elementwise_flat_tile_kernel(const EwParams p, const long long N, const int ntiles) {
  extern __shared__ __align__(16) float cbs[];
  __shared__ double wred[8];
  __shared__ int sidx[2][16 * RPT];
  constexpr int D = 64, Dp = D + 4;
  for (int i = threadIdx.x; i < p.K * D; i += 256) cbs[(i >> 6) * Dp + (i & 63)] = __ldg(p.cb + i);
  const int c4 = threadIdx.x & 15, rq = threadIdx.x >> 4;
  float c = 0.f;
  if (BWD) c = p.gl[0] * p.coef;
  float sse = 0.f;
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const long long r0 = (long long)tile * (16 * RPT);
    int kreg = 0;
    if (threadIdx.x < 16 * RPT && r0 + threadIdx.x < N) kreg = (int)__ldg(p.idx + r0 + threadIdx.x);
    float4 xv[RPT], gv[BWD ? RPT : 1];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const long long r = r0 + rq + 16 * i;
      if (r < N) {
        xv[i] = (BWD || !p.gather) ? ldg_stream4(p.z + r * D + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (BWD) gv[i] = ldg_stream4(p.g + r * D + c4 * 4);
      }
    }
    if (threadIdx.x < 16 * RPT) sidx[buf][threadIdx.x] = kreg;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const long long r = r0 + rq + 16 * i;
      if (r < N) {
        const float4 qv = *reinterpret_cast<const float4*>(&cbs[sidx[buf][rq + 16 * i] * Dp + c4 * 4]);
        const float q[4] = {qv.x, qv.y, qv.z, qv.w};
        const float x[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        float o[4];
        if (BWD) {
          const float g[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float df = __fsub_rn(x[e], q[e]);
            o[e] = fmaf(c, df, g[e]);
            sse = fmaf(df, df, sse);
          }
        } else if (p.gather) {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = q[e];
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float df = __fsub_rn(q[e], x[e]);
            o[e] = __fadd_rn(x[e], df);                      // inputs + (quantized - inputs).detach()  (ema.py:169)
            sse = fmaf(df, df, sse);
          }
        }
        stg_stream4(p.out + r * D + c4 * 4, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
  }
  if (p.sse_partials != nullptr) {
    double sd = warp_sum((double)sse);
    if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = sd;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (int w = 0; w < 8; ++w) a += wred[w];
      p.sse_partials[blockIdx.x] = a;
    }
  }
}

__global__ void gather_rows_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb, long long N, int D,
                                   float* __restrict__ out) {
  long long total = N * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long n = i / D;
    int j = (int)(i - n * D);
    out[i] = __ldg(cb + (size_t)idx[n] * D + j);
  }
}

__global__ void __launch_bounds__(256) vq_finalize_kernel(const double* __restrict__ sse_partials, int G,
                                                          const float* __restrict__ counts, int K, double n_rows_total,
                                                          double numel, float beta, float* __restrict__ scalars) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  // perplexity = exp(-sum p log(p + 1e-10)) in fp32 like the reference (ema.py:170-176)
  float ent = 0.f;
  if (counts != nullptr) {
    for (int k = tid; k < K; k += 256) {
      float pk = counts[k] / (float)n_rows_total;
      ent += pk * logf(pk + 1e-10f);
    }
  }
  red[tid] = (double)ent;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  double ent_all = red[0];
  __syncthreads();
  double a = 0.0;
  for (int g = tid; g < G; g += 256) a += sse_partials[g];
  red[tid] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    scalars[0] = (float)red[0];
    scalars[1] = (float)(red[0] / numel);
    scalars[2] = expf(-(float)ent_all);
    const float e_latent = scalars[1];
    const float commit = __fmul_rn(beta, e_latent);
    scalars[3] = commit;                         // VectorQuantizerEMA vq_loss (ema.py:166-167)
    scalars[4] = __fadd_rn(e_latent, commit);    // VectorQuantizer vq_loss = q_latent + beta * e_latent (vq.py:136-139)
  }
}

__global__ void grad_codebook_kernel(const float* __restrict__ stats, const float* __restrict__ cb,
                                     const float* __restrict__ gl, float coef, int K, int D, float* __restrict__ gE,
                                     int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * D) return;
  int k = i / D;
  float v = gl[0] * coef * (stats[k] * cb[i] - stats[K + i]);
  gE[i] = accumulate ? gE[i] + v : v;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr int EW_MAX_BLOCKS = 148 * 8;

struct AssignPlan {
  int VEC, CT, Dp, nchunks, resident, smem_stats;
  size_t smem;
};

bool plan_assign(int K, int D, AssignPlan& pl) {
  pl.VEC = (D % 4 == 0) ? 4 : 1;
  pl.Dp = pl.VEC == 4 ? D + 4 : (D | 1);
  const size_t SMEM_MAX = 200 * 1024;
  int ct = (K + CL - 1) / CL;
  if (ct <= 2) pl.CT = 2;
  else if (ct <= 4) pl.CT = 4;
  else if (ct <= 6) pl.CT = 6;
  else pl.CT = 8;
  int chunk = CL * pl.CT;
  pl.nchunks = (K + chunk - 1) / chunk;
  size_t xs = (size_t)TILE * pl.Dp * 4;
  size_t idxb = TILE * 4;
  size_t stats = (size_t)K * (D + 1) * 4;
  size_t es_res = (size_t)pl.nchunks * chunk * (pl.Dp + 1) * 4;
  size_t es_one = (size_t)chunk * (pl.Dp + 1) * 4;
  if (xs + es_one + idxb > SMEM_MAX) return false;  // D too large for this path
  const size_t TARGET = 100 * 1024;                 // leaves room for >= 2 CTAs per SM
  pl.resident = (xs + es_res + idxb <= TARGET) ? 1 : 0;
  size_t es = pl.resident ? es_res : es_one;
  pl.smem_stats = (xs + es + idxb + stats <= TARGET + 12 * 1024) ? 1 : 0;
  pl.smem = xs + es + idxb + (pl.smem_stats ? stats : 0);
  return true;
}

template <int VEC, int CT, bool SECOND>
int launch_assign_t(const AssignParams& p, const AssignPlan& pl, int& grid_out, cudaStream_t st) {
  auto kern = vq_assign_kernel<VEC, CT, SECOND>;
  VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int occ = 0;
  VQS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, pl.smem));
  if (occ < 1) occ = 1;
  int grid = occ * num_sms();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid > 4 * num_sms()) grid = 4 * num_sms();
  grid_out = grid;
  kern<<<grid, NT, pl.smem, st>>>(p);
  VQS_LAUNCH_CHECK();
  return 0;
}

template <bool SECOND>
int launch_assign(const AssignParams& p, const AssignPlan& pl, int& grid, cudaStream_t st) {
#define VQS_CASE(V, C) \
  if (pl.VEC == V && pl.CT == C) return launch_assign_t<V, C, SECOND>(p, pl, grid, st);
  VQS_CASE(4, 2) VQS_CASE(4, 4) VQS_CASE(4, 6) VQS_CASE(4, 8)
  VQS_CASE(1, 2) VQS_CASE(1, 4) VQS_CASE(1, 6) VQS_CASE(1, 8)
#undef VQS_CASE
  set_error("no assign kernel for VEC=%d CT=%d", pl.VEC, pl.CT);
  return VQS_ERR_ARG;
}

}  // namespace
}  // namespace vqs

namespace vqs {
// vq_assign_tc.cu
bool assign_tc_supported(int K, int D);
int launch_assign_tc(const float* z, int layout, int B, int D, int T, const float* cb, int K, int64_t* idx,
                     float* partials, int max_grid, int* grid_out, cudaStream_t st);
// vq_assign_tma.cu
bool assign_tma_supported(int layout, int B, int T, int K, int D);
int launch_assign_tma(const float* z, int layout, int B, int T, const float* cb, int K, int64_t* idx, float* partials,
                      int max_grid, int* grid_out, cudaStream_t st);
bool search_large_supported(int K, int D);
size_t search_large_workspace_bytes(int K, int D);
int launch_search_large(const float* z, int layout, int B, int D, int T, const float* cb, int K, int64_t* idx,
                        float* stats, void* workspace, cudaStream_t st);
static int g_vq_engine = 1;   // 0: tensor cores wherever supported, 1: auto (default), 2: CUDA cores only
}  // namespace vqs

using namespace vqs;

extern "C" int vqs_vq_set_engine(int engine) {
  VQS_CHECK_ARG(engine >= 0 && engine <= 2, "vqs_vq_set_engine: engine must be 0 (tensor cores), 1 (auto) or 2 (CUDA cores)");
  g_vq_engine = engine;
  return 0;
}

static size_t partials_bytes(int K, int D) {
  AssignPlan pl;
  const bool v1 = plan_assign(K, D, pl) && pl.smem_stats;
  if (!v1 && !assign_tc_supported(K, D)) return 0;
  return align_up((size_t)4 * num_sms() * K * (D + 1) * sizeof(float), 256);  // <= 4 CTAs per SM
}

extern "C" size_t vqs_vq_workspace_bytes(int K, int D) {
  if (K <= 0 || D <= 0) return 0;
  size_t front = partials_bytes(K, D);
  if (search_large_supported(K, D)) {
    size_t l = align_up(search_large_workspace_bytes(K, D), 256);
    if (l > front) front = l;
  }
  return front + align_up((size_t)EW_MAX_BLOCKS * sizeof(double), 256) + 256;
}

static int check_vq_shape(int layout, int B, int D, int T, int K) {
  VQS_CHECK_ARG(layout == VQS_LAYOUT_FLAT_ND || layout == VQS_LAYOUT_BDT_AS_DTB, "unknown layout %d", layout);
  VQS_CHECK_ARG(B > 0 && D > 0 && T > 0 && K > 0, "bad VQ shape B=%d D=%d T=%d K=%d", B, D, T, K);
  VQS_CHECK_ARG((long long)B * D * T < (1ll << 31), "VQ tensor too large for 32-bit indexing (B*D*T = %lld)",
                (long long)B * D * T);
  return 0;
}

extern "C" int vqs_vq_assign(const float* z, int layout, int B, int D, int T, const float* codebook, int K,
                             int64_t* idx, float* stats, float* dmin2, float* distances, void* workspace,
                             size_t workspace_bytes, vqs_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = check_vq_shape(layout, B, D, T, K)) return e;
  VQS_CHECK_ARG(z && codebook && idx && stats && workspace, "vqs_vq_assign: NULL pointer");
  if (workspace_bytes < vqs_vq_workspace_bytes(K, D)) {
    set_error("vqs_vq_assign: workspace %zu < %zu", workspace_bytes, vqs_vq_workspace_bytes(K, D));
    return VQS_ERR_WORKSPACE;
  }
  const long long Nrows = (long long)B * T;
  if (g_vq_engine <= 1 && dmin2 == nullptr && distances == nullptr && assign_tma_supported(layout, B, T, K, D) &&
      (reinterpret_cast<uintptr_t>(z) & 15) == 0 && (g_vq_engine == 0 || Nrows >= 4096)) {
    // streaming engine (flat rows by TMA, (B, D, T) rows with B % 64 == 0 by a cp.async gather): raw tf32 filter on
    // tcgen05 + exact fp32 settlement (same indices)
    int grid = 0;
    if (int e = launch_assign_tma(z, layout, B, T, codebook, K, idx, (float*)workspace, 4 * num_sms(), &grid, st))
      return e;
    const int S = K * (D + 1);
    stats_reduce_kernel<<<(S + 255) / 256, 256, 0, st>>>((const float*)workspace, grid, S, stats);
    VQS_LAUNCH_CHECK();
    return 0;
  }
  if (g_vq_engine == 0 && dmin2 == nullptr && distances == nullptr && assign_tc_supported(K, D) &&
      (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    // tensor-core search (tcgen05) + exact fp32 re-check: same indices and statistics as the CUDA-core search
    int grid = 0;
    if (int e = launch_assign_tc(z, layout, B, D, T, codebook, K, idx, (float*)workspace, 4 * num_sms(), &grid, st))
      return e;
    const int S = K * (D + 1);
    stats_reduce_kernel<<<(S + 255) / 256, 256, 0, st>>>((const float*)workspace, grid, S, stats);
    VQS_LAUNCH_CHECK();
    return 0;
  }
  if (g_vq_engine <= 1 && dmin2 == nullptr && distances == nullptr && !assign_tc_supported(K, D) &&
      search_large_supported(K, D) && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    // large codebooks: streamed tensor-core distance GEMM + exact fp32 settlement of near-ties
    return launch_search_large(z, layout, B, D, T, codebook, K, idx, stats, workspace, st);
  }
  AssignPlan pl;
  VQS_CHECK_ARG(plan_assign(K, D, pl), "vqs_vq_assign: embedding_dim %d too large for the shared-memory path", D);
  AssignParams p;
  p.z = z; p.cb = codebook; p.idx = idx; p.dmin2 = dmin2; p.dist = distances;
  p.N = (long long)B * T;
  p.layout = layout; p.B = B; p.D = D; p.T = T; p.K = K;
  p.Dp = pl.Dp;
  p.ntiles = (int)((p.N + TILE - 1) / TILE);
  p.nchunks = pl.nchunks;
  p.resident = pl.resident;
  p.smem_stats = pl.smem_stats;
  p.divD = FastDiv((uint32_t)D);
  const int S = K * (D + 1);
  if (pl.smem_stats) {
    p.partials = (float*)workspace;
  } else {
    p.partials = stats;
    VQS_CUDA(cudaMemsetAsync(stats, 0, (size_t)S * sizeof(float), st));
  }
  int grid = 0;
  int e = (dmin2 != nullptr) ? launch_assign<true>(p, pl, grid, st) : launch_assign<false>(p, pl, grid, st);
  if (e) return e;
  if (pl.smem_stats) {
    stats_reduce_kernel<<<(S + 255) / 256, 256, 0, st>>>((const float*)workspace, grid, S, stats);
    VQS_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int vqs_vq_one_hot(const int64_t* idx, long long N, int K, float* encodings, vqs_stream_t stream) {
  VQS_CHECK_ARG(idx && encodings && N > 0 && K > 0, "vqs_vq_one_hot: bad arguments");
  long long total = N * K;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  one_hot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, N, K, encodings);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_vq_ema_update(float* cluster_size, float* ema_w, float* embedding, const float* stats, float decay,
                                 float one_minus_decay, float eps, float k_eps, int K, int D, vqs_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  VQS_CHECK_ARG(cluster_size && ema_w && embedding && stats && K > 0 && D > 0, "vqs_vq_ema_update: bad arguments");
  ema_cluster_size_kernel<<<1, 256, 0, st>>>(cluster_size, stats, decay, one_minus_decay, eps, k_eps, K);
  VQS_LAUNCH_CHECK();
  long long total = (long long)K * D;
  int grid = (int)((total + 255) / 256);
  if (grid > 2 * num_sms()) grid = 2 * num_sms();
  ema_embedding_kernel<<<grid, 256, 0, st>>>(cluster_size, ema_w, embedding, stats + K, decay, one_minus_decay, K, D);
  VQS_LAUNCH_CHECK();
  return 0;
}

template <bool BWD, bool SM, bool FLAT, int VEC, bool BLK = false>
static int launch_ew_t(const EwParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = vq_elementwise_kernel<BWD, SM, FLAT, VEC, BLK>;
  if (smem > 48 * 1024) VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, 256, smem, st>>>(p);
  VQS_LAUNCH_CHECK();
  return 0;
}

template <bool BWD, bool SM>
static int launch_ew_2(const EwParams& p, bool flat, bool vec, bool blk, int grid, size_t smem, cudaStream_t st) {
  if (blk) return launch_ew_t<BWD, SM, false, 4, true>(p, grid, smem, st);
  if (flat) return vec ? launch_ew_t<BWD, SM, true, 4>(p, grid, smem, st) : launch_ew_t<BWD, SM, true, 1>(p, grid, smem, st);
  return vec ? launch_ew_t<BWD, SM, false, 4>(p, grid, smem, st) : launch_ew_t<BWD, SM, false, 1>(p, grid, smem, st);
}

static int launch_elementwise(bool bwd, EwParams& p, size_t cb_bytes, int& grid_out, cudaStream_t st) {
  const bool flat = p.layout == VQS_LAYOUT_FLAT_ND;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // float4 path: 4 consecutive floats never leave a row (flat) or a (b, d) run (B, D, T)
  const bool vec = al16(p.z) && al16(p.out) && (!bwd || al16(p.g)) && (flat ? (p.D % 4 == 0) : (p.T % 4 == 0));
  // flat (N, 64) rows: tiled kernels with staged indices (measured at 2^22 rows: forward 0.420 -> 0.350 ms = the HBM
  // roofline, backward 0.526 -> 0.503 ms).  VQS_EW_FLAT_TILE = 0 / 1 selects the grid-stride kernels for both / backward.
  {
    const char* ft = getenv("VQS_EW_FLAT_TILE");
    const int mode = ft ? atoi(ft) : 2;
    const long long N = p.total / 64;
    if (flat && vec && p.D == 64 && N >= 4096 && (size_t)p.K * 68 * sizeof(float) <= 96 * 1024 &&
        (bwd ? mode >= 2 : mode >= 1)) {
      const int RPT = bwd ? 4 : 8;
      const long long nt = (N + 16 * RPT - 1) / (16 * RPT);
      int grid = (int)(nt < EW_MAX_BLOCKS ? nt : EW_MAX_BLOCKS);
      grid_out = grid;
      const size_t smem = (size_t)p.K * 68 * sizeof(float);
      if (bwd) {
        auto kern = vq_elementwise_flat_tile_kernel<true, 4>;
        if (smem > 48 * 1024) VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, st>>>(p, N, (int)nt);
      } else {
        auto kern = vq_elementwise_flat_tile_kernel<false, 8>;
        if (smem > 48 * 1024) VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, st>>>(p, N, (int)nt);
      }
      VQS_LAUNCH_CHECK();
      return 0;
    }
  }
  // (B, 64, T) with whole 64-item batch groups: the tiled kernel (indices staged in shared memory, full-line accesses)
  if (!flat && vec && p.D == 64 && p.B % 64 == 0 && p.B >= 256 && (size_t)p.K * 65 * sizeof(float) <= 96 * 1024 &&
      getenv("VQS_EW_NO_TILE") == nullptr) {
    const int DD = bwd ? 2 : 4;
    const int NTB = (p.T + 31) / 32, DG = 64 / DD;
    const long long nt = (long long)(p.B / 64) * DG * NTB;
    int grid = (int)(nt < EW_MAX_BLOCKS ? nt : EW_MAX_BLOCKS);
    grid_out = grid;
    const size_t smem = (size_t)p.K * 65 * sizeof(float);
    const FastDiv dNTB((uint32_t)NTB), dDG((uint32_t)DG);
    if (bwd) {
      auto kern = vq_elementwise_bdt_tile_kernel<true, 2>;
      if (smem > 48 * 1024) VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 256, smem, st>>>(p, (int)nt, dNTB, dDG);
    } else {
      auto kern = vq_elementwise_bdt_tile_kernel<false, 4>;
      if (smem > 48 * 1024) VQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 256, smem, st>>>(p, (int)nt, dNTB, dDG);
    }
    VQS_LAUNCH_CHECK();
    return 0;
  }
  // blocked order for (B, D, T) tensors whose rows lie at least one idx sector apart along t (see the kernel)
  const bool blk = !flat && vec && p.B >= 4 * p.D && getenv("VQS_EW_NO_BLK") == nullptr;
  const int NTP = (p.T + 7) / 8;
  p.divNTP = FastDiv((uint32_t)NTP);
  p.nvec_blk = (long long)((p.B + 15) / 16) * p.D * NTP * 32;
  const long long nvec = blk ? p.nvec_blk : p.total / (vec ? 4 : 1);
  long long blocks = (nvec + 256 * 4 - 1) / (256 * 4);
  int grid = (int)(blocks < EW_MAX_BLOCKS ? blocks : EW_MAX_BLOCKS);
  if (grid < 1) grid = 1;
  grid_out = grid;
  const size_t padded = (size_t)p.K * (p.D + (flat ? 4 : 1)) * sizeof(float);
  const bool sm = padded <= 96 * 1024;
  const size_t smem = sm ? padded : 0;
  (void)cb_bytes;
  if (bwd) return sm ? launch_ew_2<true, true>(p, flat, vec, blk, grid, smem, st) : launch_ew_2<true, false>(p, flat, vec, blk, grid, smem, st);
  return sm ? launch_ew_2<false, true>(p, flat, vec, blk, grid, smem, st) : launch_ew_2<false, false>(p, flat, vec, blk, grid, smem, st);
}

extern "C" int vqs_vq_quantize(const float* z, int layout, int B, int D, int T, const int64_t* idx,
                               const float* codebook, int K, float* out, float* q_rows, const float* counts,
                               double n_rows_total, float beta, float* scalars, void* workspace, size_t workspace_bytes,
                               vqs_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = check_vq_shape(layout, B, D, T, K)) return e;
  if (z == nullptr) {
    // gather only: out = codebook[idx] in z's layout, exactly; z is not read, no loss is formed here (the training step
    // takes the loss from vqs_vq_backward_loss, which reads z anyway: 264 instead of 520 bytes per row in this pass)
    VQS_CHECK_ARG(idx && codebook && out, "vqs_vq_quantize: NULL pointer");
    EwParams g;
    g.z = nullptr; g.g = nullptr; g.gl = nullptr; g.idx = idx; g.cb = codebook; g.out = out; g.sse_partials = nullptr;
    g.coef = 0.f; g.gather = 1;
    g.total = (long long)B * D * T;
    g.layout = layout; g.B = B; g.D = D; g.T = T; g.K = K;
    g.divD = FastDiv((uint32_t)D);
    g.divT = FastDiv((uint32_t)T);
    int grid = 0;
    if (int e = launch_elementwise(false, g, (size_t)K * D * 4, grid, st)) return e;
    if (q_rows != nullptr) {
      long long N = (long long)B * T;
      long long blocks = (N * D + 255) / 256;
      gather_rows_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(idx, codebook, N, D, q_rows);
      VQS_LAUNCH_CHECK();
    }
    return 0;
  }
  VQS_CHECK_ARG(z && idx && codebook && out && scalars && workspace, "vqs_vq_quantize: NULL pointer");
  if (workspace_bytes < vqs_vq_workspace_bytes(K, D)) {
    set_error("vqs_vq_quantize: workspace too small");
    return VQS_ERR_WORKSPACE;
  }
  size_t front = partials_bytes(K, D);
  if (search_large_supported(K, D)) {
    size_t l = align_up(search_large_workspace_bytes(K, D), 256);
    if (l > front) front = l;
  }
  double* sse_part = (double*)((char*)workspace + front);
  EwParams p;
  p.z = z; p.g = nullptr; p.gl = nullptr; p.idx = idx; p.cb = codebook; p.out = out; p.sse_partials = sse_part;
  p.coef = 0.f; p.gather = 0;
  p.total = (long long)B * D * T;
  p.layout = layout; p.B = B; p.D = D; p.T = T; p.K = K;
  p.divD = FastDiv((uint32_t)D);
  p.divT = FastDiv((uint32_t)T);
  int grid = 0;
  if (int e = launch_elementwise(false, p, (size_t)K * D * 4, grid, st)) return e;
  vq_finalize_kernel<<<1, 256, 0, st>>>(sse_part, grid, counts, K, n_rows_total, (double)p.total, beta, scalars);
  VQS_LAUNCH_CHECK();
  if (q_rows != nullptr) {
    long long N = (long long)B * T;
    long long blocks = (N * D + 255) / 256;
    gather_rows_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(idx, codebook, N, D, q_rows);
    VQS_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int vqs_vq_backward(const float* g_out, const float* g_loss, float coef, const float* z, int layout, int B,
                               int D, int T, const int64_t* idx, const float* codebook, int K, float* grad_z,
                               vqs_stream_t stream) {
  if (int e = check_vq_shape(layout, B, D, T, K)) return e;
  VQS_CHECK_ARG(g_out && g_loss && z && idx && codebook && grad_z, "vqs_vq_backward: NULL pointer");
  EwParams p;
  p.z = z; p.g = g_out; p.gl = g_loss; p.idx = idx; p.cb = codebook; p.out = grad_z; p.sse_partials = nullptr;
  p.coef = coef; p.gather = 0;
  p.total = (long long)B * D * T;
  p.layout = layout; p.B = B; p.D = D; p.T = T; p.K = K;
  p.divD = FastDiv((uint32_t)D);
  p.divT = FastDiv((uint32_t)T);
  int grid = 0;
  return launch_elementwise(true, p, (size_t)K * D * 4, grid, (cudaStream_t)stream);
}

extern "C" int vqs_vq_backward_loss(const float* g_out, const float* g_loss, float coef, const float* z, int layout, int B,
                                    int D, int T, const int64_t* idx, const float* codebook, int K, float* grad_z,
                                    const float* counts, double n_rows_total, float beta, float* scalars, void* workspace,
                                    size_t workspace_bytes, vqs_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (int e = check_vq_shape(layout, B, D, T, K)) return e;
  VQS_CHECK_ARG(g_out && g_loss && z && idx && codebook && grad_z && scalars && workspace,
                "vqs_vq_backward_loss: NULL pointer");
  if (workspace_bytes < vqs_vq_workspace_bytes(K, D)) {
    set_error("vqs_vq_backward_loss: workspace too small");
    return VQS_ERR_WORKSPACE;
  }
  size_t front = partials_bytes(K, D);
  if (search_large_supported(K, D)) {
    size_t l = align_up(search_large_workspace_bytes(K, D), 256);
    if (l > front) front = l;
  }
  double* sse_part = (double*)((char*)workspace + front);
  EwParams p;
  p.z = z; p.g = g_out; p.gl = g_loss; p.idx = idx; p.cb = codebook; p.out = grad_z; p.sse_partials = sse_part;
  p.coef = coef; p.gather = 0;
  p.total = (long long)B * D * T;
  p.layout = layout; p.B = B; p.D = D; p.T = T; p.K = K;
  p.divD = FastDiv((uint32_t)D);
  p.divT = FastDiv((uint32_t)T);
  int grid = 0;
  if (int e = launch_elementwise(true, p, (size_t)K * D * 4, grid, st)) return e;
  vq_finalize_kernel<<<1, 256, 0, st>>>(sse_part, grid, counts, K, n_rows_total, (double)p.total, beta, scalars);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_vq_grad_codebook(const float* stats, const float* codebook, const float* g_loss, float coef, int K,
                                    int D, float* grad_E, int accumulate, vqs_stream_t stream) {
  VQS_CHECK_ARG(stats && codebook && g_loss && grad_E && K > 0 && D > 0, "vqs_vq_grad_codebook: bad arguments");
  grad_codebook_kernel<<<(K * D + 255) / 256, 256, 0, (cudaStream_t)stream>>>(stats, codebook, g_loss, coef, K, D,
                                                                             grad_E, accumulate);
  VQS_LAUNCH_CHECK();
  return 0;
}
