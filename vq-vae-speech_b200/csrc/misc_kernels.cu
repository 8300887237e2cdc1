// Library core (error string, launch counter) and the element-wise pieces of the path:
// nearest upsample, jitter gather, MSE forward+backward, ReLU / add helpers, (B,L,C)->(B,C,L) copy, fused AMSGrad.
#include <math.h>
#include <stdarg.h>

#include <atomic>

#include <stdlib.h>

#include "vqs_common.cuh"

namespace vqs {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<long long> g_engine[5];
void count_engine(int engine) {
  if (engine >= 0 && engine < 5) g_engine[engine].fetch_add(1, std::memory_order_relaxed);
}
long long engine_count(int engine) { return (engine >= 0 && engine < 5) ? g_engine[engine].load() : -1; }

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("VQS_PDL");
    cached = (e != nullptr && atoi(e) == 0) ? 0 : 1;   // on by default; VQS_PDL=0 restores plain stream order
  }
  return cached == 1;
}

int num_sms() {
  static int cached[16] = {0};                 // per device
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    if (dev >= 0 && dev < 16 && cached[dev] > 0) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
      if (dev >= 0 && dev < 16) cached[dev] = n;
      return n;
    }
  }
  (void)cudaGetLastError();
  return 148;  // B200; used only for workspace sizing when no device is visible
}

namespace {

constexpr int EW_T = 256;
inline int ew_grid(long long n, int per_thread = 4) {
  long long b = (n + (long long)EW_T * per_thread - 1) / ((long long)EW_T * per_thread);
  long long cap = (long long)num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define GRID_STRIDE(i, n) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void upsample2_fwd_kernel(const float* __restrict__ in, long long rows, int L, float* __restrict__ out) {
  const long long n = rows * L;  // one thread per input element, writes a float2
  GRID_STRIDE(i, n) {
    float v = in[i];
    reinterpret_cast<float2*>(out)[i] = make_float2(v, v);
  }
}
__global__ void upsample2_bwd_kernel(const float* __restrict__ g, long long rows, int L, float* __restrict__ gin) {
  const long long n = rows * L;
  GRID_STRIDE(i, n) {
    float2 v = reinterpret_cast<const float2*>(g)[i];
    gin[i] = v.x + v.y;
  }
}
__global__ void jitter_fwd_kernel(const float* __restrict__ in, long long rows, int L, const int* __restrict__ src,
                                  float* __restrict__ out) {
  const long long n = rows * L;
  GRID_STRIDE(i, n) {
    long long r = i / L;
    int t = (int)(i - r * L);
    out[i] = in[r * L + src[t]];
  }
}
__global__ void jitter_bwd_kernel(const float* __restrict__ g, long long rows, int L, const int* __restrict__ src,
                                  float* __restrict__ gin) {
  const long long n = rows * L;
  GRID_STRIDE(i, n) {
    long long r = i / L;
    int t = (int)(i - r * L);
    gin[i] = (src[t] == t) ? g[i] : 0.f;
  }
}
__global__ void relu_fwd_kernel(const float* __restrict__ in, long long n, float* __restrict__ out) {
  GRID_STRIDE(i, n) out[i] = fmaxf(in[i], 0.f);
}
__global__ void relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ act, long long n,
                                float* __restrict__ gin) {
  GRID_STRIDE(i, n) gin[i] = act[i] > 0.f ? g[i] : 0.f;
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                           float* __restrict__ out) {
  GRID_STRIDE(i, n) out[i] = a[i] + b[i];
}
__global__ void scale_kernel(const float* __restrict__ x, const float* __restrict__ s, long long n, float* __restrict__ out) {
  const float f = s[0];
  GRID_STRIDE(i, n) out[i] = x[i] * f;
}
// (B, L, C) -> (B, C, L) through a 32x32 smem tile (both sides coalesced)
__global__ void blc_to_ncl_kernel(const float* __restrict__ in, int B, int L, int C, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + (size_t)b * L * C;
  float* dst = out + (size_t)b * L * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int l = l0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (l < L && c < C) ? src[(size_t)l * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, l = l0 + threadIdx.x;
    if (c < C && l < L) dst[(size_t)c * L + l] = tile[threadIdx.x][i];
  }
}

// out[b, c, l] = c < Ca ? a[b, c, l] : v[b, c - Ca]: a per-utterance feature vector repeated over time and appended to the
// channels (speaker conditioning, deconvolutional_decoder.py:108-111 / global_conditioning.py:52-57)
__global__ void __launch_bounds__(EW_T) concat_channels_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                                              int Ca, int Cb, int L, long long n, float* __restrict__ out) {
  const int C = Ca + Cb;
  for (long long i = blockIdx.x * (long long)EW_T + threadIdx.x; i < n; i += (long long)gridDim.x * EW_T) {
    const long long bc = i / L;
    const int l = (int)(i - bc * L);
    const long long b = bc / C;
    const int c = (int)(bc - b * C);
    out[i] = c < Ca ? a[(b * Ca + c) * L + l] : __ldg(v + b * Cb + (c - Ca));
  }
}
// out[b, c, l] = g[b, c, l] for c < Ca: the first Ca channels of a (B, C, L) tensor (gradient of the concatenation above)
__global__ void __launch_bounds__(EW_T) slice_channels_kernel(const float* __restrict__ g, int C, int Ca, int L, long long n,
                                                             float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)EW_T + threadIdx.x; i < n; i += (long long)gridDim.x * EW_T) {
    const long long bc = i / L;
    const int l = (int)(i - bc * L);
    const long long b = bc / Ca;
    const int c = (int)(bc - b * Ca);
    out[i] = g[(b * C + c) * L + l];
  }
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ recon, const float* __restrict__ target,
                                                  int C, int L, long long t_sb, long long t_sc, long long t_sl,
                                                  float gmul, long long n, double* __restrict__ partials,
                                                  float* __restrict__ grad) {
  __shared__ double wred[8];
  float s = 0.f;
  const long long CL_ = (long long)C * L;
  GRID_STRIDE(i, n) {
    long long b = i / CL_;
    long long r = i - b * CL_;
    int c = (int)(r / L);
    int l = (int)(r - (long long)c * L);
    float d = recon[i] - target[b * t_sb + c * t_sc + l * t_sl];
    s = fmaf(d, d, s);
    if (grad != nullptr) grad[i] = gmul * d;
  }
  double w = warp_sum((double)s);
  if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int k = 0; k < 8; ++k) a += wred[k];
    partials[blockIdx.x] = a;
  }
}
__global__ void mse_finalize_kernel(const double* __restrict__ partials, int G, double numel, float* __restrict__ loss) {
  __shared__ double red[256];
  double a = 0.0;
  for (int g = threadIdx.x; g < G; g += 256) a += partials[g];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(red[0] / numel);
}

__global__ void normalize_features_kernel(const double* __restrict__ in, const double* __restrict__ mean,
                                          const double* __restrict__ stdev, long long n, int F,
                                          float* __restrict__ out) {
  GRID_STRIDE(i, n) {
    const int f = (int)(i % F);
    out[i] = (float)((in[i] - mean[f]) / stdev[f]);
  }
}

__global__ void step_inc_kernel(long long* step) { step[0] += 1; }

// torch.optim.Adam(amsgrad=True) single-tensor formulation (torch/optim/adam.py::_single_tensor_adam):
//   exp_avg.lerp_(grad, 1-b1); exp_avg_sq = b2*exp_avg_sq + (1-b2) grad^2; max_exp_avg_sq = max(..)
//   denom = sqrt(max_exp_avg_sq)/sqrt(bc2) + eps ; p -= (lr/bc1) * exp_avg / denom
__global__ void __launch_bounds__(256) amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v,
                                                      float* __restrict__ vmax, long long n4, long long n,
                                                      const long long* __restrict__ step, double lr_d, double b1_d,
                                                      double b2_d, float eps, float gscale) {
  // scalar algebra in double, then rounded to fp32 once -- what torch does with its Python-float hyper-parameters
  // (1 - beta2 = 0.001 exactly-rounded, not 1.f - 0.999f which is off by 1.3e-5 relative)
  const double t = (double)step[0];
  const float bc2s = (float)sqrt(1.0 - pow(b2_d, t));
  const float step_size = (float)(lr_d / (1.0 - pow(b1_d, t)));
  const float b2 = (float)b2_d;
  const float omb1 = (float)(1.0 - b1_d), omb2 = (float)(1.0 - b2_d);
  GRID_STRIDE(i, n4) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 X = reinterpret_cast<float4*>(vmax)[i];
#define VQS_ADAM(c)                                  \
  {                                                  \
    float gg = G.c * gscale;                         \
    M.c = M.c + omb1 * (gg - M.c);                   \
    V.c = b2 * V.c + omb2 * gg * gg;                 \
    X.c = fmaxf(X.c, V.c);                           \
    float den = sqrtf(X.c) / bc2s + eps;             \
    P.c = P.c - step_size * (M.c / den);             \
  }
    VQS_ADAM(x) VQS_ADAM(y) VQS_ADAM(z) VQS_ADAM(w)
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(vmax)[i] = X;
  }
  // scalar tail
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gg = g[i] * gscale;
    float mm = m[i] + omb1 * (gg - m[i]);
    float vv = b2 * v[i] + omb2 * gg * gg;
    float xx = fmaxf(vmax[i], vv);
    float den = sqrtf(xx) / bc2s + eps;
    p[i] = p[i] - step_size * (mm / den);
    m[i] = mm;
    v[i] = vv;
    vmax[i] = xx;
  }
}

// ------------------------------------------------------------------------------------------------
// weight normalisation (use_kaiming_normal: nn.utils.weight_norm, dim 0): one block per slice along dim 0
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];   // fixed order: deterministic
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              float* __restrict__ w, float* __restrict__ norm, int cols) {
  __shared__ float red[8];
  const size_t base = (size_t)blockIdx.x * cols;
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += 256) {
    const float x = v[base + i];
    s = fmaf(x, x, s);
  }
  const float n = sqrtf(block_sum_256(s, red));
  const float scale = g[blockIdx.x] / n;
  for (int i = threadIdx.x; i < cols; i += 256) w[base + i] = v[base + i] * scale;
  if (threadIdx.x == 0) norm[blockIdx.x] = n;
}

__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                              const float* __restrict__ g, const float* __restrict__ norm,
                                                              float* __restrict__ gv, float* __restrict__ gg, int cols) {
  __shared__ float red[8];
  const size_t base = (size_t)blockIdx.x * cols;
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += 256) s = fmaf(dw[base + i], v[base + i], s);
  const float dot = block_sum_256(s, red);
  const float n = norm[blockIdx.x];
  const float a = g[blockIdx.x] / n, c = dot / (n * n);
  for (int i = threadIdx.x; i < cols; i += 256) gv[base + i] = a * (dw[base + i] - v[base + i] * c);
  if (threadIdx.x == 0) gg[blockIdx.x] = dot / n;
}

// ------------------------------------------------------------------------------------------------
// eval-mode distance tables (vector_quantizer.py:108-127: torch.dist(x, y, 2) over itertools pairs)
// ------------------------------------------------------------------------------------------------
// element (row r, column j) of the VQ rows of `z` in either layout (vqs_b200.h: VQS_LAYOUT_*)
__device__ __forceinline__ float vq_row_elem(const float* __restrict__ z, int layout, int B, int D, int T, long long r,
                                             int j) {
  if (layout == VQS_LAYOUT_FLAT_ND) return __ldg(z + r * D + j);
  const long long f = r * D + j, P = (long long)T * B;
  const long long d = f / P, rem = f - d * P;
  const long long t = rem / B, b = rem - t * B;
  return __ldg(z + (b * D + d) * T + t);
}

__global__ void __launch_bounds__(256) pairwise_l2_kernel(const float* __restrict__ a, int layout, int B, int D, int T,
                                                          long long n, const float* __restrict__ bm, int m, int mode,
                                                          long long total, float* __restrict__ out) {
  for (long long p = blockIdx.x * 256ll + threadIdx.x; p < total; p += (long long)gridDim.x * 256) {
    long long i, j;
    if (mode == 0) {                 // itertools.product(rows of a, rows of b): p = i * m + j
      i = p / m;
      j = p - i * m;
    } else {                         // itertools.combinations(rows of a, 2): row i starts at i n - i (i + 1) / 2
      const double nn = 2.0 * (double)n - 1.0;
      i = (long long)floor((nn - sqrt(nn * nn - 8.0 * (double)p)) * 0.5);
      if (i < 0) i = 0;
      while (i > 0 && i * n - i * (i + 1) / 2 > p) --i;
      while ((i + 1) * n - (i + 1) * (i + 2) / 2 <= p) ++i;
      j = p - (i * n - i * (i + 1) / 2) + i + 1;
    }
    float s = 0.f;
    for (int c = 0; c < D; ++c) {
      const float x = vq_row_elem(a, layout, B, D, T, i, c);
      const float y = (mode == 0) ? __ldg(bm + j * D + c) : vq_row_elem(a, layout, B, D, T, j, c);
      const float df = __fsub_rn(x, y);
      s = fmaf(df, df, s);
    }
    out[p] = sqrtf(s);
  }
}

}  // namespace
}  // namespace vqs

using namespace vqs;

extern "C" int vqs_version(void) { return 100; }
extern "C" const char* vqs_last_error(void) { return g_err; }
extern "C" long long vqs_launch_count(void) { return g_launches.load(); }
extern "C" long long vqs_engine_count(int engine) { return engine_count(engine); }

extern "C" int vqs_upsample2_fwd(const float* in, long long rows, int L, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(in && out && rows > 0 && L > 0, "vqs_upsample2_fwd: bad arguments");
  upsample2_fwd_kernel<<<ew_grid(rows * L), EW_T, 0, (cudaStream_t)stream>>>(in, rows, L, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_upsample2_bwd(const float* g, long long rows, int L, float* gin, vqs_stream_t stream) {
  VQS_CHECK_ARG(g && gin && rows > 0 && L > 0, "vqs_upsample2_bwd: bad arguments");
  upsample2_bwd_kernel<<<ew_grid(rows * L), EW_T, 0, (cudaStream_t)stream>>>(g, rows, L, gin);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_jitter_fwd(const float* in, long long rows, int L, const int* src, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(in && out && src && rows > 0 && L > 0, "vqs_jitter_fwd: bad arguments");
  jitter_fwd_kernel<<<ew_grid(rows * L), EW_T, 0, (cudaStream_t)stream>>>(in, rows, L, src, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_jitter_bwd(const float* g, long long rows, int L, const int* src, float* gin, vqs_stream_t stream) {
  VQS_CHECK_ARG(g && gin && src && rows > 0 && L > 0, "vqs_jitter_bwd: bad arguments");
  jitter_bwd_kernel<<<ew_grid(rows * L), EW_T, 0, (cudaStream_t)stream>>>(g, rows, L, src, gin);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_relu_fwd(const float* in, long long n, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(in && out && n > 0, "vqs_relu_fwd: bad arguments");
  relu_fwd_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(in, n, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_relu_bwd(const float* g, const float* act, long long n, float* gin, vqs_stream_t stream) {
  VQS_CHECK_ARG(g && act && gin && n > 0, "vqs_relu_bwd: bad arguments");
  relu_bwd_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(g, act, n, gin);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_add(const float* a, const float* b, long long n, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(a && b && out && n > 0, "vqs_add: bad arguments");
  add_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(a, b, n, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_scale(const float* x, const float* s, long long n, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(x && s && out && n > 0, "vqs_scale: bad arguments");
  scale_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(x, s, n, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_blc_to_ncl(const float* in, int B, int L, int C, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(in && out && B > 0 && L > 0 && C > 0 && B <= 65535, "vqs_blc_to_ncl: bad arguments");
  dim3 grid((L + 31) / 32, (C + 31) / 32, B), block(32, 8);
  blc_to_ncl_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, B, L, C, out);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_concat_channels(const float* a, const float* v, int B, int Ca, int Cb, int L, float* out,
                                   vqs_stream_t stream) {
  VQS_CHECK_ARG(a && v && out && B > 0 && Ca > 0 && Cb > 0 && L > 0, "vqs_concat_channels: bad arguments");
  const long long n = (long long)B * (Ca + Cb) * L;
  concat_channels_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(a, v, Ca, Cb, L, n, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
extern "C" int vqs_slice_channels(const float* g, int B, int C, int Ca, int L, float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(g && out && B > 0 && C > 0 && Ca > 0 && Ca <= C && L > 0, "vqs_slice_channels: bad arguments");
  const long long n = (long long)B * Ca * L;
  slice_channels_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(g, C, Ca, L, n, out);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_normalize_features(const double* in, const double* mean, const double* stdev, long long n, int F,
                                      float* out, vqs_stream_t stream) {
  VQS_CHECK_ARG(in && mean && stdev && out && n > 0 && F > 0, "vqs_normalize_features: bad arguments");
  normalize_features_kernel<<<ew_grid(n), EW_T, 0, (cudaStream_t)stream>>>(in, mean, stdev, n, F, out);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_mse_fwd_bwd(const float* recon, const float* target, int B, int C, int L, long long t_sb,
                               long long t_sc, long long t_sl, float g_scale, float* loss, float* grad, void* workspace,
                               size_t workspace_bytes, vqs_stream_t stream) {
  VQS_CHECK_ARG(recon && target && loss && workspace && B > 0 && C > 0 && L > 0, "vqs_mse_fwd_bwd: bad arguments");
  const long long n = (long long)B * C * L;
  int grid = ew_grid(n);
  if (workspace_bytes < (size_t)grid * sizeof(double)) {
    set_error("vqs_mse_fwd_bwd: workspace %zu < %zu", workspace_bytes, (size_t)grid * sizeof(double));
    return VQS_ERR_WORKSPACE;
  }
  const float gmul = (float)((double)g_scale * 2.0 / (double)n);
  mse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(recon, target, C, L, t_sb, t_sc, t_sl, gmul, n,
                                                     (double*)workspace, grad);
  VQS_LAUNCH_CHECK();
  mse_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)workspace, grid, (double)n, loss);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_amsgrad_step(float* p, const float* g, float* m, float* v, float* vmax, long long n, long long* step,
                                int inc_step, double lr, double beta1, double beta2, double eps, double g_scale,
                                vqs_stream_t stream) {
  VQS_CHECK_ARG(p && g && m && v && vmax && step && n > 0, "vqs_amsgrad_step: bad arguments");
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  long long n4 = (al(p) && al(g) && al(m) && al(v) && al(vmax)) ? n / 4 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (inc_step) {
    step_inc_kernel<<<1, 1, 0, st>>>(step);
    VQS_LAUNCH_CHECK();
  }
  amsgrad_kernel<<<ew_grid(n, 8), 256, 0, st>>>(p, g, m, v, vmax, n4, n, step, lr, beta1, beta2, (float)eps,
                                                (float)g_scale);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_weight_norm_fwd(const float* v, const float* g, float* w, float* norm, int rows, int cols,
                                   vqs_stream_t stream) {
  VQS_CHECK_ARG(v && g && w && norm && rows > 0 && cols > 0, "vqs_weight_norm_fwd: bad arguments");
  weight_norm_fwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(v, g, w, norm, cols);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* norm, float* grad_v,
                                   float* grad_g, int rows, int cols, vqs_stream_t stream) {
  VQS_CHECK_ARG(dw && v && g && norm && grad_v && grad_g && rows > 0 && cols > 0, "vqs_weight_norm_bwd: bad arguments");
  weight_norm_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(dw, v, g, norm, grad_v, grad_g, cols);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_pairwise_l2(const float* a, int layout, int B, int D, int T, const float* b, int m, int mode, float* out,
                               vqs_stream_t stream) {
  VQS_CHECK_ARG(a && out && B > 0 && D > 0 && T > 0 && (mode == 0 || mode == 1), "vqs_pairwise_l2: bad arguments");
  VQS_CHECK_ARG(layout == VQS_LAYOUT_FLAT_ND || layout == VQS_LAYOUT_BDT_AS_DTB, "vqs_pairwise_l2: unknown layout %d", layout);
  VQS_CHECK_ARG(mode == 1 || (b && m > 0), "vqs_pairwise_l2: mode 0 needs the second row set");
  const long long n = (long long)B * T;
  const long long total = mode == 0 ? n * m : n * (n - 1) / 2;
  if (total <= 0) return 0;
  long long blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  pairwise_l2_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, layout, B, D, T, n, b, m, mode, total, out);
  VQS_LAUNCH_CHECK();
  return 0;
}
