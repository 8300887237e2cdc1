// Microbenchmark 2: the GEMM kernels' main-loop PROTOCOL without any data movement -- producer warps wait empty[s], (fence),
// arrive full[s]; one converged warp waits full[s], issues MMAS_PER_STAGE tcgen05.mma (guarded by elect.sync), commits to
// empty[s] -- to find what keeps the real kernel at ~950 cycles per k-block (12 MMAs = 768 cycles of tensor work) even
// with the operand feed switched off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/mma_pipe profiles/mma_pipe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
  for (uint32_t n = 0; n < (1u << 24); ++n) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* b) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\tselp.u32 %0, 1, 0, pe;\n\t}" : "=r"(e));
  return e != 0;
}

struct Args { int nkb, prod_warps, fence, stores, tcfence, two_issuers, data; };

template <int BN, int STAGES>
__global__ void __launch_bounds__(576, 1) pipe_kernel(Args a, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  __shared__ uint64_t full[STAGES], empty[STAGES], done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid * 16; i < STAGES * STAGE_BYTES; i += blockDim.x * 16) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (a.data == 1) {        // finite random floats in [1, 2)
      v = make_uint4(0x3F800000u | (h & 0x7FFFFF), 0x3F800000u | ((h * 3) & 0x7FFFFF), 0x3F800000u | ((h * 5) & 0x7FFFFF), 0x3F800000u | ((h * 7) & 0x7FFFFF));
    } else if (a.data == 2) { // random bit patterns (NaN, Inf, denormals included)
      v = make_uint4(h, h * 747796405u + 1u, h * 2891336453u + 7u, h ^ 0x7FC12345u);
    }
    *reinterpret_cast<uint4*>(smem + i) = v;
  }
  const int n_issuers = a.two_issuers ? 2 : 1;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], a.prod_warps); mbar_init(&empty[s], n_issuers); }
    mbar_init(&done, n_issuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t sbase = smem_u32(smem);
  if (warp < a.prod_warps) {
    for (int i = 0; i < a.nkb; ++i) {
      const int s = i % STAGES;
      if (!mbar_wait(&empty[s], (((uint32_t)(i / STAGES)) & 1u) ^ 1u)) break;
      if (a.stores) {   // 16-byte stores like the B producers: 4 per thread into this stage
        const uint32_t dst = sbase + (uint32_t)(s * STAGE_BYTES + A_BYTES) + (uint32_t)(tid * 16);
        for (int u = 0; u < a.stores; ++u)
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst + (uint32_t)(u * 8192) % (uint32_t)(B_BYTES)), "f"(0.f) : "memory");
      }
      if (a.fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else if (warp == 16 || (warp == 17 && a.two_issuers)) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t d0 = make_desc_sw128(sbase);
    const bool elected = elect_one();
    const bool main_role = warp == 16;
    const bool all = !a.two_issuers;
    uint32_t par = 0;
    bool ok = true;
    long long t0 = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < a.nkb && ok; i0 += STAGES) {
      const uint32_t nz = i0 > 0 ? 1u : 0u;
#pragma unroll
      for (int j = 0; j < STAGES; ++j) {
        if (i0 + j < a.nkb) {
          ok = mbar_wait(&full[j], par) && ok;
          if (i0 + j == 0) t0 = clock64();
          if (a.tcfence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a_hi = d0 + (uint64_t)((j * STAGE_BYTES) >> 4);
          const uint64_t b_hi = a_hi + (uint64_t)(A_BYTES >> 4);
          const uint64_t a_lo = a_hi + (uint64_t)((A_BYTES + B_BYTES) >> 4);
          const uint64_t b_lo = a_lo + (uint64_t)(A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            const int g = j * 4 + k;
            if (!elected) continue;
            if (all || !main_role) {
              umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, g == 0 ? nz : 1u);
              umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
            }
            if (all || main_role) umma_tf32(tmem_base + (uint32_t)((1 + g % 3) * BN), a_hi + adv, b_hi + adv, idesc, g < 3 ? nz : 1u);
          }
          if (elected) umma_commit(&empty[j]);
          __syncwarp();
        }
      }
      par ^= 1u;
    }
    if (elected) umma_commit(&done);
    __syncwarp();
    ok = mbar_wait(&done, 0) && ok;
    const long long t1 = clock64();
    if (elected && warp == 16) out[blockIdx.x] = ok ? (float)(t1 - t0) / (float)a.nkb : -1.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <int BN, int STAGES>
static void run(Args a, float* d_out) {
  constexpr int STAGE_BYTES = 2 * (128 * 128 + BN * 128);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  auto kern = pipe_kernel<BN, STAGES>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  float h[148];
  kern<<<148, 576, smem>>>(a, d_out);
  kern<<<148, 576, smem>>>(a, d_out);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  double c = 0; int bad = 0;
  for (int i = 0; i < 148; ++i) { if (h[i] < 0) ++bad; else c += h[i]; }
  c /= (148 - bad) > 0 ? (148 - bad) : 1;
  printf("BN=%3d stages=%d nkb=%d producers=%2d fence=%d stores=%d tcfence=%d issuers=%d data=%d | %7.1f cycles per k-block (12 MMAs: %d of tensor work)%s\n",
         BN, STAGES, a.nkb, a.prod_warps, a.fence, a.stores, a.tcfence, a.two_issuers ? 2 : 1, a.data, c, BN == 128 ? 768 : 384, bad ? " [TIMEOUT]" : "");
  fflush(stdout);
}

int main() {
  float* d_out;
  CK(cudaMalloc(&d_out, 148 * sizeof(float)));
  for (int data = 0; data < 3; ++data)
    for (int two = 0; two < 2; ++two) {
      Args a = {72, 16, 1, 0, 1, two, data};
      run<128, 3>(a, d_out);
      Args b = {72, 16, 1, 4, 1, two, data};
      run<128, 3>(b, d_out);
      run<64, 4>(a, d_out);
      run<64, 3>(a, d_out);
    }
  return 0;
}
