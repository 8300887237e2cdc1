// Data-parallel exchange of the training step over NVLink 5 / NVSwitch WITHOUT a collective library: this library's own
// kernels read and write peer memory -- plain peer pointers for the 11 KB of EMA statistics, the NVLS multicast mapping
// (multimem.ld_reduce / multimem.st: the switch reduces and replicates) for the 65 MB of gradients and parameters.
//
//   reference: none -- the reference's multi-GPU switch is dead code (pipeline_factory.py:56-61); the contract is SURVEY 8e:
//   EMA statistics summed over shards before the EMA update, gradients averaged, replicas bit-identical.
//
// Why not NCCL allreduces (round 1): an NCCL kernel's CTAs cannot share an SM with the one-wave tcgen05 GEMMs of the backward
// pass (200 KB of shared memory each), so every gradient bucket launched next to a GEMM pushed it into a second wave
// (conv_3 dgrad 0.111 -> 0.205 ms at 8 GPUs), the 11 KB statistics allreduce sat on the critical path of the forward pass at
// ~25 us, and all of it cost 8 % (2 GPUs) to 11 % (8 GPUs) of the step.  Here nothing runs beside the backward pass at all:
//
//   vqs_dp_allreduce_small   after a cross-GPU barrier every rank sums the peers' statistics vectors itself, in RANK ORDER
//                            (bit-identical result on every rank), through peer pointers: one tiny kernel
//   vqs_dp_amsgrad_step      the gradient allreduce is folded into the optimizer: after a barrier ("all backward passes are
//                            done") rank r owns slice r of the flat buffers, reads the SUM of the gradient over all GPUs with
//                            multimem.ld_reduce (the switch adds; 1/W of the bytes cross this GPU's links), applies
//                            Adam(amsgrad=True) to its slice (optimizer state is SHARDED: 1/W of the optimizer's HBM traffic
//                            per GPU) and broadcasts the new parameters with multimem.st; a second barrier ends the step
//   vqs_dp_barrier           signal-pad barrier: a release-store of this call's epoch into every peer's pad, an
//                            acquire-spin on the own pad (bounded: traps instead of hanging)
//
// All three are ordinary stream-ordered launches and are captured into the step's CUDA graph like every other kernel.
#include "vqs_common.cuh"

namespace vqs {
namespace {

constexpr unsigned SPIN_LIMIT_DP = 1u << 24;   // x ~100 ns: a few seconds, then trap (a dead peer must not hang the GPU)

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Thread r < world of the calling block: tells rank r "this rank reached epoch e of `channel`" and waits until rank r said
// the same (the body of dp_barrier_kernel, also run INSIDE the fused kernels below so that a barrier costs no launch).
__device__ __forceinline__ void barrier_signal_wait(const vqs_dp_ctx& ctx, int channel, unsigned e, int r) {
  __threadfence_system();
  unsigned* theirs = reinterpret_cast<unsigned*>(ctx.peer_pads[r]) + VQS_DP_PAD_WORD0 + channel * VQS_DP_MAX_WORLD + ctx.rank;
  st_release_sys(theirs, e);
  const unsigned* mine = reinterpret_cast<const unsigned*>(ctx.peer_pads[ctx.rank]) + VQS_DP_PAD_WORD0 +
                         channel * VQS_DP_MAX_WORLD + r;
  unsigned n = 0;
  while ((int)(ld_acquire_sys(mine) - e) < 0) {
    __nanosleep(32);
    if (++n > SPIN_LIMIT_DP) __trap();
  }
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One block.  Thread r < world: tells rank r "this rank reached epoch e of `channel`" and waits until rank r said the same.
// Everything this stream did before the launch is ordered before the signal (kernel boundary + system fence); everything a
// peer did before ITS signal is visible after the wait (acquire).
__global__ void __launch_bounds__(32) dp_barrier_kernel(vqs_dp_ctx ctx, int channel) {
  __shared__ unsigned epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = ctx.epochs[channel] + 1u;
    ctx.epochs[channel] = epoch_s;
  }
  __syncthreads();
  const unsigned e = epoch_s;
  if ((int)threadIdx.x < ctx.world) barrier_signal_wait(ctx, channel, e, threadIdx.x);
}

// Barrier + sum in ONE launch for vectors one block can cover (the 11 KB of EMA statistics sit on the critical path of the
// forward pass between the search and the EMA update: two launches cost 23 us at 8 GPUs).
__global__ void __launch_bounds__(1024) dp_barrier_sum_peers_kernel(vqs_dp_ctx ctx, int channel, vqs_dp_ptrs src,
                                                                    float* __restrict__ dst, int n) {
  __shared__ unsigned epoch_s;
  if (threadIdx.x == 0) {
    epoch_s = ctx.epochs[channel] + 1u;
    ctx.epochs[channel] = epoch_s;
  }
  __syncthreads();
  if ((int)threadIdx.x < ctx.world) barrier_signal_wait(ctx, channel, epoch_s, threadIdx.x);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 1024) {
    float v[VQS_DP_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < VQS_DP_MAX_WORLD; ++r)      // all loads in flight before the first add
      v[r] = r < ctx.world ? *reinterpret_cast<const volatile float*>(reinterpret_cast<const float*>(src.p[r]) + i) : 0.f;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < VQS_DP_MAX_WORLD; ++r)
      if (r < ctx.world) s += v[r];
    dst[i] = s;
  }
}

// dst[i] = sum over ranks (rank order) of the peers' src vectors; src_ptrs[r] = rank r's vector mapped into this process
__global__ void __launch_bounds__(256) dp_sum_peers_kernel(vqs_dp_ctx ctx, vqs_dp_ptrs src, float* __restrict__ dst,
                                                           long long n) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float v[VQS_DP_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < VQS_DP_MAX_WORLD; ++r)      // all loads in flight before the first add
      v[r] = r < ctx.world ? *reinterpret_cast<const volatile float*>(reinterpret_cast<const float*>(src.p[r]) + i) : 0.f;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < VQS_DP_MAX_WORLD; ++r)
      if (r < ctx.world) s += v[r];
    dst[i] = s;
  }
}

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Adam(amsgrad=True) on float4 elements [lo4, hi4) of the flat buffers: same arithmetic, in the same order, as
// amsgrad_kernel (misc_kernels.cu) with g = (sum over GPUs) * gscale.
__global__ void __launch_bounds__(256) dp_amsgrad_nvls_kernel(float* __restrict__ mc_p, const float* __restrict__ p_local,
                                                              const float* __restrict__ mc_g, float* __restrict__ m,
                                                              float* __restrict__ v, float* __restrict__ vmax,
                                                              long long lo4, long long hi4,
                                                              const long long* __restrict__ step, double lr_d, double b1_d,
                                                              double b2_d, float eps, float gscale) {
  const double t = (double)step[0];
  const float bc2s = (float)sqrt(1.0 - pow(b2_d, t));
  const float step_size = (float)(lr_d / (1.0 - pow(b1_d, t)));
  const float b2 = (float)b2_d;
  const float omb1 = (float)(1.0 - b1_d), omb2 = (float)(1.0 - b2_d);
  for (long long i = lo4 + blockIdx.x * 256ll + threadIdx.x; i < hi4; i += (long long)gridDim.x * 256) {
    const float4 G = multimem_ld_reduce_add(mc_g + i * 4);     // the switch adds the W copies
    float4 P = reinterpret_cast<const float4*>(p_local)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 X = reinterpret_cast<float4*>(vmax)[i];
#define VQS_ADAM_DP(c)                               \
  {                                                  \
    float gg = G.c * gscale;                         \
    M.c = M.c + omb1 * (gg - M.c);                   \
    V.c = b2 * V.c + omb2 * gg * gg;                 \
    X.c = fmaxf(X.c, V.c);                           \
    float den = sqrtf(X.c) / bc2s + eps;             \
    P.c = P.c - step_size * (M.c / den);             \
  }
    VQS_ADAM_DP(x) VQS_ADAM_DP(y) VQS_ADAM_DP(z) VQS_ADAM_DP(w)
#undef VQS_ADAM_DP
    multimem_st(mc_p + i * 4, P);                               // every GPU's copy of the parameters, this one included
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(vmax)[i] = X;
  }
  __threadfence_system();      // the multicast stores are performed before this stream's next kernel (the barrier) signals
}

__global__ void dp_step_inc_kernel(long long* step) { step[0] += 1; }

// The same update as a LIGHT kernel for the bucket-wise exchange that runs beside the backward pass: 128 threads, at most 40
// registers (__launch_bounds__(128, 12)), no shared memory -- a block fits into what a 576-thread x 96-register tcgen05 GEMM CTA
// leaves of an SM (10 240 registers), so the exchange takes issue slots and link bandwidth, not SMs, from the GEMMs.
__global__ void __launch_bounds__(128, 12) dp_amsgrad_nvls_light_kernel(float* __restrict__ mc_p, const float* __restrict__ p_local,
                                                                        const float* __restrict__ mc_g, float* __restrict__ m,
                                                                        float* __restrict__ v, float* __restrict__ vmax,
                                                                        long long lo4, long long hi4,
                                                                        const long long* __restrict__ step, double lr_d,
                                                                        double b1_d, double b2_d, float eps, float gscale) {
  const double t = (double)step[0];
  const float bc2s = (float)sqrt(1.0 - pow(b2_d, t));
  const float step_size = (float)(lr_d / (1.0 - pow(b1_d, t)));
  const float b2 = (float)b2_d;
  const float omb1 = (float)(1.0 - b1_d), omb2 = (float)(1.0 - b2_d);
  for (long long i = lo4 + blockIdx.x * 128ll + threadIdx.x; i < hi4; i += (long long)gridDim.x * 128) {
    const float4 G = multimem_ld_reduce_add(mc_g + i * 4);
    float4 P = reinterpret_cast<const float4*>(p_local)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float4 X = reinterpret_cast<float4*>(vmax)[i];
#define VQS_ADAM_DP(c)                               \
  {                                                  \
    float gg = G.c * gscale;                         \
    M.c = M.c + omb1 * (gg - M.c);                   \
    V.c = b2 * V.c + omb2 * gg * gg;                 \
    X.c = fmaxf(X.c, V.c);                           \
    float den = sqrtf(X.c) / bc2s + eps;             \
    P.c = P.c - step_size * (M.c / den);             \
  }
    VQS_ADAM_DP(x) VQS_ADAM_DP(y) VQS_ADAM_DP(z) VQS_ADAM_DP(w)
#undef VQS_ADAM_DP
    multimem_st(mc_p + i * 4, P);
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(vmax)[i] = X;
  }
  __threadfence_system();
}

// The same step as ONE launch (round 2; four launches -- step counter, barrier, update, barrier -- cost 197 us at 8 GPUs against
// 101 us for the single-GPU optimizer): block 0 runs the entry barrier and releases the others through a local flag, every
// thread keeps FOUR multimem.ld_reduce (and the 16 loads of its optimizer state) in flight before the first dependent
// instruction -- one 16-byte request per thread cannot fill a link whose round trip goes through the switch --, and the block
// that finishes last runs the exit barrier.  Local words (ctx.epochs): [GEN] launches completed, [GO] = GEN + 1 once the
// entry barrier has passed, [DONE] blocks finished.  Arithmetic and its order per element are those of the kernel above.
constexpr int EP_GO = VQS_DP_CHANNELS, EP_DONE = VQS_DP_CHANNELS + 1, EP_GEN = VQS_DP_CHANNELS + 2;
constexpr int DP_UNROLL = 4;
__global__ void __launch_bounds__(256) dp_amsgrad_nvls_fused_kernel(vqs_dp_ctx ctx, float* __restrict__ mc_p,
                                                                    const float* __restrict__ p_local,
                                                                    const float* __restrict__ mc_g, float* __restrict__ m,
                                                                    float* __restrict__ v, float* __restrict__ vmax,
                                                                    long long lo4, long long hi4, long long* step, int inc_step,
                                                                    double lr_d, double b1_d, double b2_d, float eps,
                                                                    float gscale, int ch_before, int ch_after) {
  __shared__ unsigned sh_e, sh_last;
  unsigned* ep = ctx.epochs;
  const unsigned gen = *reinterpret_cast<volatile unsigned*>(ep + EP_GEN);   // changes only after every block has finished
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      sh_e = ep[ch_before] + 1u;
      ep[ch_before] = sh_e;
      if (inc_step) step[0] += 1;
    }
    __syncthreads();
    if ((int)threadIdx.x < ctx.world) barrier_signal_wait(ctx, ch_before, sh_e, threadIdx.x);   // every rank's gradients are written
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(ep + EP_GO, gen + 1u);
  } else {
    if (threadIdx.x == 0) {
      unsigned n = 0;
      while (ld_acquire_gpu(ep + EP_GO) != gen + 1u) {
        __nanosleep(32);
        if (++n > SPIN_LIMIT_DP) __trap();
      }
    }
    __syncthreads();
  }
  const double t = (double)*reinterpret_cast<volatile long long*>(step);
  const float bc2s = (float)sqrt(1.0 - pow(b2_d, t));
  const float step_size = (float)(lr_d / (1.0 - pow(b1_d, t)));
  const float b2 = (float)b2_d;
  const float omb1 = (float)(1.0 - b1_d), omb2 = (float)(1.0 - b2_d);
  const long long stride = (long long)gridDim.x * 256;
  for (long long i0 = lo4 + blockIdx.x * 256ll + threadIdx.x; i0 < hi4; i0 += stride * DP_UNROLL) {
    float4 G[DP_UNROLL], P[DP_UNROLL], M[DP_UNROLL], V[DP_UNROLL], X[DP_UNROLL];
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi4) G[u] = multimem_ld_reduce_add(mc_g + i * 4);     // the switch adds the W copies
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi4) {
        P[u] = reinterpret_cast<const float4*>(p_local)[i];
        M[u] = reinterpret_cast<float4*>(m)[i];
        V[u] = reinterpret_cast<float4*>(v)[i];
        X[u] = reinterpret_cast<float4*>(vmax)[i];
      }
    }
#pragma unroll
    for (int u = 0; u < DP_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi4) {
#define VQS_ADAM_DP(c)                                     \
  {                                                        \
    float gg = G[u].c * gscale;                            \
    M[u].c = M[u].c + omb1 * (gg - M[u].c);                \
    V[u].c = b2 * V[u].c + omb2 * gg * gg;                 \
    X[u].c = fmaxf(X[u].c, V[u].c);                        \
    float den = sqrtf(X[u].c) / bc2s + eps;                \
    P[u].c = P[u].c - step_size * (M[u].c / den);          \
  }
        VQS_ADAM_DP(x) VQS_ADAM_DP(y) VQS_ADAM_DP(z) VQS_ADAM_DP(w)
#undef VQS_ADAM_DP
        multimem_st(mc_p + i * 4, P[u]);                            // every GPU's copy of the parameters, this one included
        reinterpret_cast<float4*>(m)[i] = M[u];
        reinterpret_cast<float4*>(v)[i] = V[u];
        reinterpret_cast<float4*>(vmax)[i] = X[u];
      }
    }
  }
  __threadfence_system();      // this thread's multicast stores are performed before its block counts itself done
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(ep + EP_DONE, 1u);
    sh_last = (done == gridDim.x - 1) ? 1u : 0u;
    if (sh_last) {
      __threadfence();          // (acquire side of the counter: every block's stores precede the signal below)
      ep[EP_DONE] = 0u;
      ep[EP_GEN] = gen + 1u;
      sh_e = ep[ch_after] + 1u;
      ep[ch_after] = sh_e;
    }
  }
  __syncthreads();
  if (sh_last && (int)threadIdx.x < ctx.world)
    barrier_signal_wait(ctx, ch_after, sh_e, threadIdx.x);          // all parameters written everywhere, all gradients consumed
}

bool ctx_ok(const vqs_dp_ctx* c) {
  if (!c || c->world < 1 || c->world > VQS_DP_MAX_WORLD || c->rank < 0 || c->rank >= c->world || !c->epochs) return false;
  for (int r = 0; r < c->world; ++r)
    if (!c->peer_pads[r]) return false;
  return true;
}

}  // namespace
}  // namespace vqs

using namespace vqs;

extern "C" int vqs_dp_barrier(const vqs_dp_ctx* ctx, int channel, vqs_stream_t stream) {
  VQS_CHECK_ARG(ctx_ok(ctx) && channel >= 0 && channel < VQS_DP_CHANNELS, "vqs_dp_barrier: bad context or channel");
  dp_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*ctx, channel);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_dp_allreduce_small(const vqs_dp_ctx* ctx, const vqs_dp_ptrs* src, float* dst, long long n, int channel,
                                      vqs_stream_t stream) {
  VQS_CHECK_ARG(ctx_ok(ctx) && src && dst && n > 0 && channel >= 0 && channel < VQS_DP_CHANNELS,
                "vqs_dp_allreduce_small: bad arguments");
  for (int r = 0; r < ctx->world; ++r) VQS_CHECK_ARG(src->p[r] != nullptr, "vqs_dp_allreduce_small: NULL peer pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 8 * 1024) {
    dp_barrier_sum_peers_kernel<<<1, 1024, 0, st>>>(*ctx, channel, *src, dst, (int)n);
    VQS_LAUNCH_CHECK();
    return 0;
  }
  dp_barrier_kernel<<<1, 32, 0, st>>>(*ctx, channel);          // every rank's vector is complete
  VQS_LAUNCH_CHECK();
  long long blocks = (n + 255) / 256;
  if (blocks > 64) blocks = 64;
  dp_sum_peers_kernel<<<(unsigned)blocks, 256, 0, st>>>(*ctx, *src, dst, n);
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_dp_amsgrad_step(const vqs_dp_ctx* ctx, float* mc_p, const float* p_local, const float* mc_g, float* m,
                                   float* v, float* vmax, long long n, long long* step, int inc_step, double lr,
                                   double beta1, double beta2, double eps, int ch_before, int ch_after,
                                   vqs_stream_t stream) {
  VQS_CHECK_ARG(ctx_ok(ctx) && mc_p && p_local && mc_g && m && v && vmax && step && n > 0 && n % 4 == 0,
                "vqs_dp_amsgrad_step: bad arguments (n must be a multiple of 4)");
  VQS_CHECK_ARG(ch_before >= 0 && ch_before < VQS_DP_CHANNELS && ch_after >= 0 && ch_after < VQS_DP_CHANNELS,
                "vqs_dp_amsgrad_step: bad barrier channel");
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  VQS_CHECK_ARG(al(mc_p) && al(p_local) && al(mc_g) && al(m) && al(v) && al(vmax), "vqs_dp_amsgrad_step: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n4 = n / 4;
  const long long per = (n4 + ctx->world - 1) / ctx->world;
  const long long lo4 = per * ctx->rank, hi4 = lo4 + per < n4 ? lo4 + per : n4;
  // VQS_DP_FUSED=1: the one-launch kernel.  Measured at 2 GPUs it is no faster than the four launches (0.187 - 0.199 ms against
  // 0.191 ms per call, profiles/r04o_dp_step_2gpu.txt: the step is bound by what crosses the links -- with NVLS every GPU's own
  // copy travels to the switch as well, 73 MB in and out per GPU at 8 GPUs at the ~400 GB/s per direction NCCL also reaches
  // here), so the simpler sequence of launches stays the default.
  static const bool fused = getenv("VQS_DP_FUSED") && atoi(getenv("VQS_DP_FUSED")) == 1;
  if (fused) {
    // one launch: entry barrier, sharded update with the reduce and the broadcast through the switch, exit barrier.  Every block
    // must be resident at once (the blocks wait for block 0): at most 2 per SM of 256 threads.
    long long blocks = hi4 > lo4 ? (hi4 - lo4 + 256ll * DP_UNROLL - 1) / (256ll * DP_UNROLL) : 1;
    static const int bps = getenv("VQS_DP_BLOCKS_PER_SM") ? atoi(getenv("VQS_DP_BLOCKS_PER_SM")) : 2;   // 1 or 2 (probe knob)
    const long long cap = (long long)(bps == 1 ? 1 : 2) * num_sms();
    if (blocks > cap) blocks = cap;
    dp_amsgrad_nvls_fused_kernel<<<(unsigned)blocks, 256, 0, st>>>(*ctx, mc_p, p_local, mc_g, m, v, vmax, lo4, hi4, step, inc_step,
                                                                   lr, beta1, beta2, (float)eps, 1.0f / (float)ctx->world,
                                                                   ch_before, ch_after);
    VQS_LAUNCH_CHECK();
    return 0;
  }
  if (inc_step) {
    dp_step_inc_kernel<<<1, 1, 0, st>>>(step);
    VQS_LAUNCH_CHECK();
  }
  dp_barrier_kernel<<<1, 32, 0, st>>>(*ctx, ch_before);        // every rank's backward pass has written its gradients
  VQS_LAUNCH_CHECK();
  if (hi4 > lo4) {
    long long blocks = (hi4 - lo4 + 255) / 256;
    const long long cap = 4ll * num_sms();
    if (blocks > cap) blocks = cap;
    dp_amsgrad_nvls_kernel<<<(unsigned)blocks, 256, 0, st>>>(mc_p, p_local, mc_g, m, v, vmax, lo4, hi4, step, lr, beta1,
                                                             beta2, (float)eps, 1.0f / (float)ctx->world);
    VQS_LAUNCH_CHECK();
  }
  dp_barrier_kernel<<<1, 32, 0, st>>>(*ctx, ch_after);         // all parameters written everywhere, all gradients consumed
  VQS_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqs_dp_amsgrad_range(const vqs_dp_ctx* ctx, float* mc_p, const float* p_local, const float* mc_g, float* m,
                                    float* v, float* vmax, long long lo, long long hi, long long* step, int inc_step, double lr,
                                    double beta1, double beta2, double eps, int ch_before, int ch_after, vqs_stream_t stream) {
  VQS_CHECK_ARG(ctx_ok(ctx) && mc_p && p_local && mc_g && m && v && vmax && step && lo >= 0 && hi >= lo && lo % 4 == 0 &&
                    hi % 4 == 0,
                "vqs_dp_amsgrad_range: bad arguments (lo, hi must be multiples of 4)");
  VQS_CHECK_ARG(ch_before >= 0 && ch_before < VQS_DP_CHANNELS && ch_after < VQS_DP_CHANNELS,
                "vqs_dp_amsgrad_range: bad barrier channel");
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  VQS_CHECK_ARG(al(mc_p) && al(p_local) && al(mc_g) && al(m) && al(v) && al(vmax), "vqs_dp_amsgrad_range: 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // These kernels run BESIDE GEMM CTAs that need the maximum shared-memory carve-out.  An SM cannot change its carve-out
    // while blocks are resident: a block of ours that lands on an idle SM with the default (L1-heavy) split would keep the next
    // ~200 KB GEMM CTA off that SM until it has finished, so they ask for the GEMMs' split.
    static DevCache carve;
    if (dev_needs(carve, 1)) {
      VQS_CUDA(cudaFuncSetAttribute(dp_amsgrad_nvls_light_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      VQS_CUDA(cudaFuncSetAttribute(dp_barrier_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      VQS_CUDA(cudaFuncSetAttribute(dp_step_inc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
  }
  if (inc_step) {
    dp_step_inc_kernel<<<1, 1, 0, st>>>(step);
    VQS_LAUNCH_CHECK();
  }
  dp_barrier_kernel<<<1, 32, 0, st>>>(*ctx, ch_before);        // every rank has finished the gradients of this bucket
  VQS_LAUNCH_CHECK();
  const long long b4 = lo / 4, n4 = (hi - lo) / 4;
  const long long per = (n4 + ctx->world - 1) / ctx->world;
  const long long lo4 = b4 + per * ctx->rank, hi4 = lo4 + per < b4 + n4 ? lo4 + per : b4 + n4;
  if (hi4 > lo4) {
    long long blocks = (hi4 - lo4 + 127) / 128;
    const long long cap = 2ll * num_sms();
    if (blocks > cap) blocks = cap;
    dp_amsgrad_nvls_light_kernel<<<(unsigned)blocks, 128, 0, st>>>(mc_p, p_local, mc_g, m, v, vmax, lo4, hi4, step, lr, beta1,
                                                                   beta2, (float)eps, 1.0f / (float)ctx->world);
    VQS_LAUNCH_CHECK();
  }
  if (ch_after >= 0) {
    dp_barrier_kernel<<<1, 32, 0, st>>>(*ctx, ch_after);       // all parameters written everywhere, all gradients consumed
    VQS_LAUNCH_CHECK();
  }
  return 0;
}
