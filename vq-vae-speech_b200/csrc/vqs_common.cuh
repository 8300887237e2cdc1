// Shared helpers for libvqs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqs_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvqs_b200 is written for sm_100a (B200) only"
#endif

namespace vqs {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void count_engine(int engine);
int num_sms();

// Per-DEVICE memo of the dynamic-shared-memory attribute of a kernel (cudaFuncSetAttribute is per device; a process that
// drives several GPUs must configure each one).  dev_needs(c, bytes): true when the current device has not been configured
// for at least `bytes` yet (and records it); devices beyond the table are configured on every call.
struct DevCache {
  size_t v[16] = {0};
};
inline bool dev_needs(DevCache& c, size_t bytes) {
  int d = -1;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 16) return true;
  if (bytes > c.v[d]) {
    c.v[d] = bytes;
    return true;
  }
  return false;
}

#define VQS_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      vqs::set_error(__VA_ARGS__);          \
      return VQS_ERR_ARG;                   \
    }                                       \
  } while (0)

#define VQS_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      vqs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define VQS_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      vqs::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
    vqs::count_launch();                                                                 \
  } while (0)

// Programmatic dependent launch (on by default, VQS_PDL=0 disables; read once): the kernels of the GEMM family call pdl_prologue_done() after a
// prologue that touches no global memory (barrier init, TMEM allocation), and are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that prologue and launch latency overlap the tail of the kernel
// before.  Without the attribute griddepcontrol.wait returns at once.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// The same for a kernel that runs as clusters of (cx, cy, 1) CTAs (grid dimensions must be multiples of the cluster's).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cx,
                                      unsigned cy, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cx;
  attr[0].val.clusterDim.y = cy;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// every global-memory access of the kernel comes after this; the next kernel of the stream may be scheduled from here on
__device__ __forceinline__ void pdl_prologue_done() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

// unsigned division by a runtime constant, exact for n < 2^31:
//   l = ceil(log2 d), s = l - 1, m = ceil(2^(32+s) / d) (< 2^32), n / d == umulhi(n, m) >> s
struct FastDiv {
  uint32_t mul, shift, d;
  FastDiv() : mul(0), shift(0), d(1) {}
  explicit FastDiv(uint32_t div) : mul(0), shift(0), d(div) {
    if (div <= 1) return;
    uint32_t l = 0;
    while ((1ull << l) < div) ++l;
    shift = l - 1;
    mul = (uint32_t)(((1ull << (32 + shift)) + div - 1) / div);
  }
  __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
    if (d == 1) return n;
#ifdef __CUDA_ARCH__
    return __umulhi(n, mul) >> shift;
#else
    return (uint32_t)(((uint64_t)n * mul) >> 32) >> shift;
#endif
  }
  __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace vqs
