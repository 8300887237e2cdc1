// Microbenchmark: steady-state issue period of tcgen05.mma on one B200 SM (and on a CTA pair) for the operand
// configurations the conv / wgrad GEMMs could use.  One thread per CTA issues ITERS MMAs back to back on resident
// (zeroed) shared-memory operands -- no operand feed at all -- and the CTA reports cycles per MMA; optionally other warps
// hammer shared memory with 16-byte stores ("noise") to show how much of the shared-memory pipe the tensor core needs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o profiles/mma_rate profiles/mma_rate.cu
//   ./profiles/mma_rate
//
// Columns: kind (tf32 K=8 / f16 K=16 per instruction), M (128; 256 = cta_group::2 pair), N, A source (smem / tmem),
// number of accumulators visited round-robin, noise warps -> cycles per MMA, MACs per cycle per SM, % of the nominal
// pipe rate (tf32 2048, f16 4096 MAC/clk/SM).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Args {
  int kind;    // 0 tf32, 1 f16
  int m;       // 128 or 256 (pair)
  int n;       // 64..256
  int a_tmem;  // 0 smem, 1 tmem
  int nacc;    // accumulators visited round-robin
  int noise;   // warps storing to shared memory meanwhile
  int iters;
  int stage_bytes;  // distance between operand stages (ring of 4)
  int commit_every; // tcgen05.commit (to a barrier nobody waits on) after every n-th MMA; 0 = only at the end
};

template <int CG>
__device__ __forceinline__ void mma_issue(int kind, int a_tmem, uint32_t d, uint64_t adesc, uint32_t a_taddr,
                                          uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    if (kind == 0) {
      if (!a_tmem)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    } else {
      if (!a_tmem)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
  } else {
    if (kind == 0) {
      if (!a_tmem)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    } else {
      if (!a_tmem)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
  }
}


template <int CG, int KIND, int ATMEM, int N, int NACC, int CE = 0>
__global__ void __launch_bounds__(576, 1) mma_rate_kernel(Args a, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_done, bar_poll;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t cta_rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int total = 4 * a.stage_bytes + 16384;
  for (int i = tid * 16; i < total; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1048575;" ::"r"(smem_u32(&bar_poll)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop_flag = 0;
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0 && cta_rank == 0) {
    constexpr uint32_t fmt = KIND == 0 ? 2u : 0u;
    constexpr uint32_t M = CG * 128;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    constexpr uint32_t a_bytes = 128 * 128;
    const uint32_t base = smem_u32(smem);
    const uint32_t a_tcol = tmem_base + (uint32_t)(NACC * N);
    const uint64_t ad0 = make_desc_sw128(base), bd0 = make_desc_sw128(base + a_bytes);
    const uint32_t stage16 = (uint32_t)a.stage_bytes >> 4;
    // zero the accumulators once
#pragma unroll
    for (int c = 0; c < NACC; ++c)
      mma_issue<CG>(KIND, ATMEM, tmem_base + (uint32_t)(c * N), ad0, a_tcol, bd0, idesc, 0u);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < a.iters; it += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int s = u >> 2, k = u & 3;
        const uint64_t off = (uint64_t)(s * stage16 + k * 2);
        mma_issue<CG>(KIND, ATMEM, tmem_base + (uint32_t)((u % NACC) * N), ad0 + off, a_tcol + (uint32_t)(k * 8),
                      bd0 + off, idesc, 1u);
        if (CG == 1 && CE > 0 && (u + 1) % CE == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_poll)) : "memory");
      }
    }
    if (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_done)) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)), "h"((uint16_t)1) : "memory");
    uint32_t ok = 0, spins = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar_done)), "r"(0u) : "memory");
      if (++spins > (1u << 24)) break;
    }
    long long t1 = clock64();
    out[blockIdx.x] = ok ? (float)(t1 - t0) / (float)a.iters : -1.f;
    stop_flag = 1;
    if (CG == 2) {
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32((const void*)&stop_flag)), "r"(1u));
      asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
    }
  } else if (a.noise < 0 && warp >= 2 && warp < 2 - a.noise) {
    // pollers: spin on an mbarrier phase that never completes (what idle producer warps of a GEMM kernel do)
    unsigned n = 0;
    while (!stop_flag && n < (1u << 22)) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar_poll)), "r"(0u) : "memory");
      ++n;
    }
    if (tid == 64 && cta_rank == 0) out[gridDim.x + blockIdx.x] = (float)n;
  } else if (warp >= 2 && warp < 2 + a.noise) {
    const uint32_t nb = smem_u32(smem) + (uint32_t)(4 * a.stage_bytes) + (uint32_t)((tid & 31) * 16 + ((warp & 7) * 512));
    float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    unsigned n = 0;
    while (!stop_flag && n < (1u << 22)) {
#pragma unroll
      for (int u = 0; u < 16; ++u)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(nb + (uint32_t)((u & 3) * 4096)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      n += 16;
    }
    if (tid == 64 && cta_rank == 0) out[gridDim.x + blockIdx.x] = (float)n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

typedef void (*kern_t)(Args, float*);
template <int CG, int KIND, int ATMEM, int N>
static kern_t pick_nacc(int nacc) {
  switch (nacc) {
    case 1: return mma_rate_kernel<CG, KIND, ATMEM, N, 1>;
    case 3: if (N <= 128) return mma_rate_kernel<CG, KIND, ATMEM, N <= 128 ? N : 128, 3>; break;
    case 4: if (N <= 128 && !ATMEM) return mma_rate_kernel<CG, KIND, 0, N <= 128 ? N : 128, 4>; break;
  }
  return nullptr;
}
template <int CG, int KIND, int ATMEM>
static kern_t pick_n(int n, int nacc) {
  switch (n) {
    case 64: return pick_nacc<CG, KIND, ATMEM, 64>(nacc);
    case 128: return pick_nacc<CG, KIND, ATMEM, 128>(nacc);
    case 256: return pick_nacc<CG, KIND, ATMEM, 256>(nacc);
  }
  return nullptr;
}
static kern_t pick(const Args& a) {
  const int cg = a.m / 128;
  if (a.commit_every) {   // compile-time commit period: tf32, one CTA, operands in smem, one accumulator
    if (a.n == 64) return a.commit_every == 4 ? mma_rate_kernel<1, 0, 0, 64, 1, 4> : a.commit_every == 8 ? mma_rate_kernel<1, 0, 0, 64, 1, 8> : mma_rate_kernel<1, 0, 0, 64, 1, 16>;
    if (a.n == 128) return a.commit_every == 4 ? mma_rate_kernel<1, 0, 0, 128, 1, 4> : a.commit_every == 8 ? mma_rate_kernel<1, 0, 0, 128, 1, 8> : mma_rate_kernel<1, 0, 0, 128, 1, 16>;
    return a.commit_every == 4 ? mma_rate_kernel<1, 0, 0, 256, 1, 4> : a.commit_every == 8 ? mma_rate_kernel<1, 0, 0, 256, 1, 8> : mma_rate_kernel<1, 0, 0, 256, 1, 16>;
  }
  if (cg == 1) {
    if (a.kind == 0) return a.a_tmem ? pick_n<1, 0, 1>(a.n, a.nacc) : pick_n<1, 0, 0>(a.n, a.nacc);
    return a.a_tmem ? pick_n<1, 1, 1>(a.n, a.nacc) : pick_n<1, 1, 0>(a.n, a.nacc);
  }
  if (a.kind == 0) return a.a_tmem ? pick_n<2, 0, 1>(a.n, a.nacc) : pick_n<2, 0, 0>(a.n, a.nacc);
  return a.a_tmem ? pick_n<2, 1, 1>(a.n, a.nacc) : pick_n<2, 1, 0>(a.n, a.nacc);
}

static void run(Args a, int grid, float* d_out, float* h_out) {
  const int cg = a.m == 256 ? 2 : 1;
  const int nloc = a.n / cg;                     // B rows held by one CTA
  a.stage_bytes = 128 * 128 + nloc * 128;
  const size_t smem = 4 * (size_t)a.stage_bytes + 16384 + 1024;
  CK(cudaMemset(d_out, 0, 2 * 148 * sizeof(float)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kern_t kern = pick(a);
  if (kern == nullptr) { printf("no instantiation\n"); return; }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cg == 1) {
    kern<<<grid, 576, smem>>>(a, d_out);   // warm-up
    CK(cudaEventRecord(e0));
    kern<<<grid, 576, smem>>>(a, d_out);
    CK(cudaEventRecord(e1));
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(576);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, a, d_out));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, a, d_out));
    CK(cudaEventRecord(e1));
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaMemcpy(h_out, d_out, 2 * 148 * sizeof(float), cudaMemcpyDeviceToHost));
  double cyc = 0, noise = 0;
  int cnt = 0, bad = 0;
  for (int i = 0; i < grid; i += cg) {
    if (h_out[i] < 0) { ++bad; continue; }
    cyc += h_out[i];
    noise += h_out[grid + i];
    ++cnt;
  }
  cyc /= cnt > 0 ? cnt : 1;
  noise /= cnt > 0 ? cnt : 1;
  const int kper = a.kind == 0 ? 8 : 16;
  const double macs_per_sm = (double)a.m * a.n * kper / cg;     // per instruction per SM
  const double nominal = a.kind == 0 ? 2048.0 : 4096.0;
  const double tflops = 2.0 * a.m * a.n * kper * (double)a.iters * (grid / cg) / (ms * 1e-3) / 1e12;
  printf("%-4s M=%3d N=%3d A=%-4s nacc=%d noise=%2d | %7.1f cyc/MMA | %7.1f MAC/clk/SM = %5.1f %% of nominal | %7.1f TFLOP/s chip (%.3f ms)%s",
         a.kind == 0 ? "tf32" : "f16", a.m, a.n, a.a_tmem ? "tmem" : "smem", a.nacc, a.noise, cyc, macs_per_sm / cyc,
         100.0 * macs_per_sm / cyc / nominal, tflops, ms, bad ? "  [TIMEOUTS]" : "");
  if (a.noise > 0) printf(" | noise %.1f B/clk/SM", noise * 16.0 * 32.0 * a.noise / (cyc * a.iters));
  if (a.commit_every) printf(" | commit every %d MMAs", a.commit_every);
  if (a.noise < 0) printf(" | %d polling warps, %.0f polls per thread", -a.noise, noise);
  printf("\n");
  fflush(stdout);
}

int main(int argc, char** argv) {
  float *d_out, h_out[2 * 148];
  CK(cudaMalloc(&d_out, 2 * 148 * sizeof(float)));
  const int iters = 4096;
  const bool full = argc > 1;
  const int grids[2] = {148, 2};     // whole chip (power-capped clocks) and one SM / pair (boost clocks)
  for (int gi = 0; gi < (full ? 2 : 1); ++gi) {
    const int grid = grids[gi];
    printf("== grid %d CTAs ==\n", grid);
    if (full) {
      for (int kind = 0; kind < 2; ++kind)
        for (int m = 128; m <= 256; m += 128)
          for (int n = 64; n <= 256; n *= 2)
            for (int at = 0; at < 2; ++at) {
              const int maxacc = (512 - (at ? 32 : 0)) / n;
              Args a = {kind, m, n, at, 1, 0, iters, 0, 0};
              run(a, grid, d_out, h_out);
              if (n == 128 && maxacc >= 3) {
                a.nacc = maxacc >= 4 ? 4 : 3;
                run(a, grid, d_out, h_out);
              }
            }
      // shared-memory interference: 16-byte stores alongside, and warps polling an mbarrier alongside
      for (int kind = 0; kind < 2; ++kind)
        for (int n = 64; n <= 256; n *= 2)
          for (int noise = -16; noise <= 16; noise += 8) {
            if (noise == 0) continue;
            Args a = {kind, 128, n, 0, 1, noise, iters, 0, 0};
            run(a, grid, d_out, h_out);
          }
    }
    // does a tcgen05.commit drain the MMA pipe?  (the GEMM kernels commit once per k-block = 4 / 12 MMAs)
    for (int n = 64; n <= 256; n *= 2)
      for (int ce = 0; ce <= 16; ce = ce ? ce * 2 : 4) {
        Args a = {0, 128, n, 0, 1, 0, iters, 0, ce};
        run(a, grid, d_out, h_out);
      }
  }
  return 0;
}
