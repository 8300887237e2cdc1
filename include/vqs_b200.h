/*
 * vqs_b200.h -- C ABI of libvqs_b200.so: hand-written sm_100a CUDA kernels for the VQ-VAE-Speech
 * training hot path (VQ bottleneck + Conv1d encoder / residual stack / jitter / deconvolutional decoder
 * + MSE + AMSGrad).  This is the drop-in boundary: everything above it is host code mirroring the
 * reference's torch.nn.Module interface; everything below it is CUDA.
 *
 * Conventions
 *   - every entry point returns 0 on success, non-zero on error (a cudaError_t or a VQS_ERR_* code);
 *     vqs_last_error() returns a thread-local message.  No exceptions cross the ABI.
 *   - all tensor pointers are DEVICE pointers owned by the caller (the PyTorch caching allocator in the
 *     Python host); the library allocates nothing persistent and never synchronises the device, so every
 *     call is stream-ordered on `stream` and CUDA-graph capturable.
 *   - all floating-point tensors are fp32, indices are int64 (the reference's torch.argmin dtype).
 *   - "NCL" = contiguous (batch, channels, length) as torch.nn.Conv1d uses.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   src/models/vector_quantizer_ema.py:83-183   VectorQuantizerEMA.forward   -> vqs_vq_assign, vqs_vq_ema_update,
 *                                                                              vqs_vq_quantize, vqs_vq_backward
 *   src/models/vector_quantizer.py:70-156       VectorQuantizer.forward      -> same + vqs_vq_grad_codebook
 *   src/modules/conv1d_builder.py:33-44 / conv_transpose1d_builder.py:33-44 (nn.Conv1d / nn.ConvTranspose1d
 *   forward and the autograd backward)                                       -> vqs_conv_gemm, vqs_wgrad_gemm,
 *                                                                              vqs_bias_grad, vqs_permute_weight
 *   src/modules/residual.py:31-70, residual_stack.py:34-46 (ReLU / skip adds) -> epilogue flags of vqs_conv_gemm
 *   src/modules/jitter.py:47-70                  Jitter.forward               -> vqs_jitter_fwd / vqs_jitter_bwd
 *   src/models/deconvolutional_decoder.py:66,117 nn.Upsample(scale_factor=2)  -> vqs_upsample2_fwd / _bwd
 *   src/experiments/convolutional_trainer.py:40,54 nn.MSELoss                 -> vqs_mse_fwd_bwd
 *   src/experiments/convolutional_trainer.py:41-42,68 optim.Adam(amsgrad=True) -> vqs_amsgrad_step
 */
#ifndef VQS_B200_H
#define VQS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vqs_stream_t; /* a cudaStream_t */

#define VQS_ERR_ARG 10001        /* invalid argument (shape, NULL pointer, unsupported size) */
#define VQS_ERR_WORKSPACE 10002  /* workspace too small */

/* ------------------------------------------------------------------------------------------------ */
/* library                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
int vqs_version(void);
const char* vqs_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py's gpu_launches). */
long long vqs_launch_count(void);
/* how many GEMM calls of the calling process were dispatched to each engine (parity tests assert that an eligible layer
 * really ran on tcgen05, and that nothing silently fell back to the CUDA-core kernel).  Unknown id -> -1. */
#define VQS_ENGINE_CONV_CUDACORE 0   /* vqs_conv_gemm on the exact-fp32 CUDA-core implicit GEMM */
#define VQS_ENGINE_CONV_TC 1         /* vqs_conv_gemm on tcgen05 (gemm_tc_kernel, conv mode) */
#define VQS_ENGINE_WGRAD_CUDACORE 2  /* vqs_wgrad_gemm on CUDA cores */
#define VQS_ENGINE_WGRAD_TC 3        /* vqs_wgrad_gemm on tcgen05, thread-gathered operands */
#define VQS_ENGINE_WGRAD_TMA 4       /* vqs_wgrad_gemm on tcgen05, both operands by TMA */
long long vqs_engine_count(int engine);

/* ------------------------------------------------------------------------------------------------ */
/* data-parallel exchange over NVLink / NVSwitch peer memory (csrc/dp_nvls.cu)                        */
/* ------------------------------------------------------------------------------------------------ */
/* The reference has no working multi-GPU path (its nn.DataParallel wrap is dead code, pipeline_factory.py:56-61); these
 * entry points implement the data-parallel contract of SURVEY.md 8e -- EMA statistics summed over shards, gradients
 * averaged, replicas bit-identical -- with this library's own kernels on symmetric buffers (every rank allocates the same
 * buffer; rank r's copy is mapped into every process, and the whole set additionally behind one NVLS multicast address).
 * The caller (parallel.py, through torch.distributed._symmetric_memory) owns the mappings.  One process per GPU, at most
 * VQS_DP_MAX_WORLD GPUs of one NVSwitch domain. */
#define VQS_DP_MAX_WORLD 8
#define VQS_DP_CHANNELS 4        /* independent barrier sequences */
#define VQS_DP_EPOCH_WORDS 8     /* local words per context: VQS_DP_CHANNELS epoch counters + flags of the fused step kernel */
#define VQS_DP_PAD_WORD0 256     /* first 32-bit word of the signal pad this library uses (VQS_DP_CHANNELS * VQS_DP_MAX_WORLD words) */
typedef struct {
  int rank, world;
  void* peer_pads[VQS_DP_MAX_WORLD];   /* peer_pads[r]: rank r's signal pad as mapped into THIS process (r == rank: own pad) */
  unsigned int* epochs;                /* VQS_DP_EPOCH_WORDS words in local device memory, zero-initialised once */
} vqs_dp_ctx;
typedef struct {
  const void* p[VQS_DP_MAX_WORLD];     /* p[r]: rank r's copy of a symmetric buffer as mapped into this process */
} vqs_dp_ptrs;
/* Cross-GPU barrier on `stream`: returns (on the device) once every rank's stream has reached its matching call. */
int vqs_dp_barrier(const vqs_dp_ctx* ctx, int channel, vqs_stream_t stream);
/* dst[i] = sum over ranks, in rank order, of rank r's src vector (n floats) -- bit-identical on every rank.  Starts with a
 * barrier on `channel` (all vectors complete).  dst is local memory.  Used for the packed [counts | dw] EMA statistics. */
int vqs_dp_allreduce_small(const vqs_dp_ctx* ctx, const vqs_dp_ptrs* src, float* dst, long long n, int channel,
                           vqs_stream_t stream);
/* Adam(amsgrad=True) step (convolutional_trainer.py:41-42,68) fused with the gradient allreduce: barrier(ch_before); rank r
 * updates elements [r n/W, (r+1) n/W) of the flat buffers with g = multimem.ld_reduce(mc_g) / W (the NVSwitch adds the W
 * gradient copies) and writes the new parameters to every GPU through mc_p (multimem.st); barrier(ch_after).  mc_p / mc_g:
 * NVLS multicast addresses of the symmetric parameter / gradient buffers; p_local: this rank's copy of the parameters;
 * m, v, vmax: LOCAL optimizer state (only this rank's slice is ever touched: the state is sharded over the ranks). */
int vqs_dp_amsgrad_step(const vqs_dp_ctx* ctx, float* mc_p, const float* p_local, const float* mc_g, float* m, float* v,
                        float* vmax, long long n, long long* step, int inc_step, double lr, double beta1, double beta2,
                        double eps, int ch_before, int ch_after, vqs_stream_t stream);
/* The same step for ONE BUCKET [lo, hi) of the flat buffers (element offsets, multiples of 4), meant for a side stream that
 * runs BESIDE the backward pass: barrier(ch_before) ("every rank has finished this bucket's gradients"), rank r updates the
 * r-th W-th of the bucket, barrier(ch_after) unless ch_after < 0.  The kernels are small (128 threads, <= 40 registers, no
 * shared memory), so their blocks fit next to the 576-thread / ~200 KB tcgen05 GEMM CTAs instead of taking SMs from them.
 * inc_step: add 1 to *step first (the first bucket of a step).  The optimizer state a rank owns is then the union of its
 * slices of the buckets (the caller keeps that list for checkpoints). */
int vqs_dp_amsgrad_range(const vqs_dp_ctx* ctx, float* mc_p, const float* p_local, const float* mc_g, float* m, float* v,
                         float* vmax, long long lo, long long hi, long long* step, int inc_step, double lr, double beta1,
                         double beta2, double eps, int ch_before, int ch_after, vqs_stream_t stream);

/* ------------------------------------------------------------------------------------------------ */
/* VQ bottleneck                                                                                    */
/* ------------------------------------------------------------------------------------------------ */
/* Row layouts.  FLAT_ND: z is a contiguous (N, D) matrix, N = B*T.
 * BDT_AS_DTB: z is a contiguous (B, D, T) tensor and rows are formed exactly as the reference does,
 * `inputs.permute(1, 2, 0).contiguous().view(-1, D)` (vector_quantizer_ema.py:101-106): row r holds flat
 * elements [r*D, (r+1)*D) of the (D, T, B) ordering, flat element f = d*T*B + t*B + b <- z[b, d, t]. */
#define VQS_LAYOUT_FLAT_ND 0
#define VQS_LAYOUT_BDT_AS_DTB 1

/* bytes of scratch the VQ entry points need for (K, D); pass at least this much as `workspace`. */
size_t vqs_vq_workspace_bytes(int K, int D);

/* Nearest-code search (vector_quantizer_ema.py:109-117): d[n,k] = (sum_j x_nj^2 + sum_j e_kj^2) - 2 sum_j x_nj e_kj
 * in fp32, idx[n] = first minimum over k.  Also accumulates the per-code statistics the EMA update and the
 * codebook gradient need: stats = [counts (K) | dw (K*D)], counts[k] = #rows assigned to k, dw[k] = sum of those rows
 * (`torch.sum(encodings, 0)`, `encodings.t() @ flat_input`, :145,:153) -- deterministic two-stage reduction.
 * Optional outputs (NULL to skip): dmin2 (N, 2) = best and second-best distance per row (near-tie report),
 * distances (N, K). */
int vqs_vq_assign(const float* z, int layout, int B, int D, int T, const float* codebook, int K,
                  int64_t* idx, float* stats, float* dmin2, float* distances,
                  void* workspace, size_t workspace_bytes, vqs_stream_t stream);

/* Search engine of vqs_vq_assign (process-wide).  All engines return identical indices and counts (tests/test_vq_gpu.py).
 *   1 = automatic (default): for D = 64, K <= 48 and at least 4096 rows the streaming engine (row tiles as raw tf32 tcgen05
 *       operands + exact fp32 settlement of every row the filter cannot decide): flat rows arrive by TMA, the reference's
 *       (B, 64, T) rows with B % 64 == 0 by a cp.async gather (2^22 rows: 0.42 / 0.58 ms against 1.23 / 3.7 ms on CUDA
 *       cores); the exact-fp32 CUDA-core search for every other shape with a shared-memory resident codebook; the streamed
 *       tcgen05 distance GEMM (persistent kernel, 3xTF32 scores + exact fp32 settlement of near-ties) for large codebooks
 *       (D = 32/64, K <= 8192; K = 512: 4x, K = 4096: 9x faster than the CUDA-core search, 254 TFLOP/s at 2^20 rows; the
 *       statistics of this engine are accumulated with fp32 atomics: counts exact, dw to rounding, order not fixed);
 *   0 = tensor cores wherever a tensor-core kernel exists (also the resident-codebook one);
 *   2 = CUDA cores only. */
int vqs_vq_set_engine(int engine);

/* encodings = zeros(N, K).scatter_(1, idx, 1)  (vector_quantizer_ema.py:118-119). */
int vqs_vq_one_hot(const int64_t* idx, long long N, int K, float* encodings, vqs_stream_t stream);

/* EMA codebook update with Laplace smoothing, in place (vector_quantizer_ema.py:143-156):
 *   cs <- cs*decay + (1-decay)*counts ; n = sum(cs) ; cs <- (cs + eps) / (n + K*eps) * n
 *   ema_w <- ema_w*decay + (1-decay)*dw ; embedding <- ema_w / cs[:, None]
 * `stats` is the [counts | dw] vector of vqs_vq_assign (allreduced over ranks under data parallelism). */
int vqs_vq_ema_update(float* cluster_size, float* ema_w, float* embedding, const float* stats,
                      float decay, float one_minus_decay, float eps, float k_eps, int K, int D, vqs_stream_t stream);

/* Quantise + straight-through + loss terms (vector_quantizer_ema.py:159-176):
 *   q = codebook[idx] ; out = x + (q - x) (same layout as z) ; sse = sum (q - x)^2
 *   scalars[0] = sse, [1] = e_latent = sse / (N*D), [2] = perplexity = exp(-sum p log(p + 1e-10)), p = counts / n_rows_total,
 *   [3] = beta * e_latent (VectorQuantizerEMA's vq_loss), [4] = e_latent + beta * e_latent (VectorQuantizer's vq_loss)
 * `counts` (K) are the (global) counts; n_rows_total is the number of rows they were taken over.
 * q_rows: optional (N, D) `concatenated_quantized` (:162), NULL to skip. */
int vqs_vq_quantize(const float* z, int layout, int B, int D, int T, const int64_t* idx, const float* codebook, int K,
                    float* out, float* q_rows, const float* counts, double n_rows_total, float beta, float* scalars,
                    void* workspace, size_t workspace_bytes, vqs_stream_t stream);

/* Backward of the bottleneck w.r.t. its input (autograd of :165-169):
 *   grad_z = g_out + g_loss[0] * coef * (x - q),  coef = 2*beta/(N*D) supplied by the caller; g_loss is a device scalar. */
int vqs_vq_backward(const float* g_out, const float* g_loss, float coef, const float* z, int layout, int B, int D, int T,
                    const int64_t* idx, const float* codebook, int K, float* grad_z, vqs_stream_t stream);

/* vqs_vq_backward that ALSO forms the losses of the forward pass in the same sweep (it reads z, idx and the codebook
 * anyway): scalars as vqs_vq_quantize writes them (SSE, e_latent = mean((q - x)^2), perplexity, vq_loss of either class).
 * Together with vqs_vq_quantize(z = NULL): the gather-only forward -- out = codebook[idx] exactly, z not read -- the fused
 * forward + backward moves 20 D + 16 bytes per row instead of 24 D + 16: z is read once.  The value of the
 * gather-only `quantized` is q itself where the reference's is fl(x + fl(q - x)) (vector_quantizer_ema.py:169): equal
 * within an ulp of max(|x|, |q|).  The drop-in modules keep the exact forward; the captured training step uses this pair. */
int vqs_vq_backward_loss(const float* g_out, const float* g_loss, float coef, const float* z, int layout, int B, int D, int T,
                         const int64_t* idx, const float* codebook, int K, float* grad_z, const float* counts,
                         double n_rows_total, float beta, float* scalars, void* workspace, size_t workspace_bytes,
                         vqs_stream_t stream);

/* Codebook gradient of the non-EMA VectorQuantizer (autograd of vector_quantizer.py:136-139):
 *   grad_E[k] (+)= g_loss[0] * coef * (counts[k] * E[k] - dw[k]),  coef = 2/(N*D). */
int vqs_vq_grad_codebook(const float* stats, const float* codebook, const float* g_loss, float coef, int K, int D,
                         float* grad_E, int accumulate, vqs_stream_t stream);

/* ------------------------------------------------------------------------------------------------ */
/* Conv1d / ConvTranspose1d as implicit GEMM                                                        */
/* ------------------------------------------------------------------------------------------------ */
/* GEMM engines.  FP32: exact-fp32 FMA on CUDA cores.  TF32 / TF32X3: tcgen05 tensor cores with TMEM accumulators --
 * single-pass TF32 (what cuDNN does by default for the reference's convs on a GPU) or the 3xTF32 split
 * (x = hi + lo, acc += Al*Bh + Ah*Bl + Ah*Bh, fp32-accurate: same 1e-5 parity bar as FP32).  Shapes the tensor-core
 * kernels do not cover (conv-like GEMM: A not tap-major, Cred % 32 != 0 or rows not 16-byte aligned; wgrad: reduction
 * shorter than 32) silently use FP32. */
#define VQS_PREC_FP32 0
#define VQS_PREC_TF32 1
#define VQS_PREC_TF32X3 2

/* One descriptor covers nn.Conv1d forward, its dgrad, nn.ConvTranspose1d (stride 1) forward and its dgrad:
 *
 *   acc[b, m, l] = sum_{c < Cred, j < ksz} A[m, c*ksz + j] * X'[b, c, p(l, j)]
 *   p(l, j): pn = l*l_mul + j*j_mul + off ; valid iff pn % l_div == 0 and 0 <= pn/l_div < Lin ; X' = 0 when invalid
 *   X'[b, c, p] = X[b*x_sb + c*x_sc + p*x_sl]  (relu'd first when x_relu)
 *
 * A is a row-major (M, Cred*ksz) matrix: the nn.Conv1d weight (Cout, Cin, k) as is for a forward conv and
 * the nn.ConvTranspose1d weight (Cin, Cout, k) as is for its dgrad; vqs_permute_weight produces the
 * (d1, d0, k) arrangement the other two cases need.  With a_tap_major != 0, A is instead (M, ksz, Cred) -- reduction
 * index j*Cred + c -- the arrangement the tensor-core engines require (vqs_permute_weight modes 1 and 2).  With
 * a_tap_major == 2, A is the pre-split, pre-swizzled tensor-core operand IMAGE of that matrix (vqs_permute_weight
 * modes 3 and 4): the GEMM fetches its A tiles with TMA bulk copies.  The image is Cred rounded up to a multiple of 32 wide;
 * the extra channels are zeros in the image and are read as zeros from X (Cred stays the real channel count here).
 *
 * Epilogue, per output element (b, m, l), out tensors are NCL (B, M, Lout):
 *   v = acc + bias[m]                       (bias may be NULL)
 *   v += add_pre  (relu(add_pre) when add_pre_relu)
 *   if relu: v = max(v, 0)
 *   if mask_out: mask_out = (v > 0)
 *   if mask: v = mask_on ? v : 0            (mask_kind 1: float tensor > 0; 2: uint8 tensor != 0)
 *   v += add_post
 *   out = v ;  if out2: out2 = (mask2 on) ? v : 0
 */
typedef struct vqs_conv_gemm_desc {
  const float* A;
  const float* X;
  int a_tap_major;
  int M, Cred, ksz;
  int B, Lin, Lout;
  long long x_sb, x_sc, x_sl;
  int l_mul, j_mul, off, l_div;
  int x_relu;
  const float* bias;
  const float* add_pre;
  int add_pre_relu;
  int relu;
  unsigned char* mask_out;
  const void* mask;
  int mask_kind;
  const float* add_post;
  float* out;
  float* out2;
  const void* mask2;
  int mask2_kind;
  int precision; /* VQS_PREC_* : which GEMM engine computes acc */
  /* Optional scratch (may be NULL / 0): lets the tensor-core engines split the reduction of few-tile GEMMs (M <= 128) with
   * a bias-only epilogue over several CTAs; needs splits * M * B * Lout floats, splits <= #SMs / tiles. */
  void* splitk_ws;
  size_t splitk_ws_bytes;
} vqs_conv_gemm_desc;

int vqs_conv_gemm(const vqs_conv_gemm_desc* d, vqs_stream_t stream);

/* Weight gradient of nn.Conv1d / nn.ConvTranspose1d (autograd of the above):
 *
 *   dW[m, c*ksz + j] (+)= sum_{b, l < La} Aact[b, m, l] * X'[b, c, l*l_mul + j*j_mul + off]   (0 outside [0, Lx))
 *
 * Aact is NCL (B, M, La), X is NCL (B, Cred, Lx) (relu'd on load when x_relu).  Conv1d: Aact = dy, X = x, l_mul = stride,
 * j_mul = 1, off = -pad, dW is (Cout, Cin, k).  ConvTranspose1d: Aact = x, X = dy, l_mul = 1, j_mul = 1, off = -pad,
 * dW is (Cin, Cout, k).  The (b, l) reduction is split over `workspace` partials and reduced in a fixed order. */
typedef struct vqs_wgrad_desc {
  const float* Aact;
  const float* X;
  int M, Cred, ksz;
  int B, La, Lx;
  int l_mul, j_mul, off;
  int x_relu;
  float* dW;
  int accumulate;
  int precision; /* VQS_PREC_* */
} vqs_wgrad_desc;

size_t vqs_wgrad_workspace_bytes(int M, int Cred, int ksz, int B, int La);
int vqs_wgrad_gemm(const vqs_wgrad_desc* d, void* workspace, size_t workspace_bytes, vqs_stream_t stream);

/* db[m] (+)= sum_{b, l} g[b, m, l]   (g NCL (B, M, L)). */
int vqs_bias_grad(const float* g, int B, int M, int L, float* db, int accumulate, vqs_stream_t stream);

/* Re-arrangements of a conv weight w[d0][d1][k]:
 *   mode 0: out[d1][d0][k]   (swap the channel dims)
 *   mode 1: out[d0][k][d1]   (tap-major, same orientation)
 *   mode 2: out[d1][k][d0]   (tap-major, channel dims swapped)
 *   mode 3 / 4: tensor-core operand image of the mode-1 / mode-2 matrix: ceil(M/128) * (k*ceil(Cred/32)) blocks of 8192
 *               floats ([hi | lo] x 128 rows x 32 floats, SWIZZLE_128B), M = d0 / d1, Cred = d1 / d0; rows beyond M and
 *               channels beyond Cred are zeros */
int vqs_permute_weight(const float* w, int d0, int d1, int k, int mode, float* out, vqs_stream_t stream);

/* The same re-arrangement for up to VQS_PERMUTE_MAX_ITEMS weights in ONE launch (the training step rebuilds ~25 GEMM
 * operands per step from the freshly updated parameters; replaces that many vqs_permute_weight calls). `items` is a
 * HOST array, read during the call. */
#define VQS_PERMUTE_MAX_ITEMS 32
typedef struct {
  const float* w; /* (d0, d1, k) parameter as stored by nn.Conv1d / nn.ConvTranspose1d */
  float* out;     /* as for vqs_permute_weight */
  int d0, d1, k, mode;
} vqs_permute_item;
int vqs_permute_weights(const vqs_permute_item* items, int n, vqs_stream_t stream);

/* ------------------------------------------------------------------------------------------------ */
/* element-wise pieces of the path                                                                  */
/* ------------------------------------------------------------------------------------------------ */
/* nn.Upsample(scale_factor=2), nearest: out[b,c,i] = in[b,c,i>>1]; backward: gin[b,c,l] = g[b,c,2l] + g[b,c,2l+1].
 * `rows` = B*C, L = input length. */
int vqs_upsample2_fwd(const float* in, long long rows, int L, float* out, vqs_stream_t stream);
int vqs_upsample2_bwd(const float* g, long long rows, int L, float* gin, vqs_stream_t stream);

/* Jitter (jitter.py:47-70) with a host-drawn plan src[t] (int32, device): out[b,c,t] = in[b,c,src[t]];
 * backward: gin[b,c,t] = (src[t] == t) ? g[b,c,t] : 0 (replaced columns were overwritten in place from a detached
 * clone, so they pass no gradient and the source columns receive none). */
int vqs_jitter_fwd(const float* in, long long rows, int L, const int* src, float* out, vqs_stream_t stream);
int vqs_jitter_bwd(const float* g, long long rows, int L, const int* src, float* gin, vqs_stream_t stream);

/* nn.MSELoss(recon, target) forward + backward in one pass.  recon is NCL (B, C, L); target element (b,c,l) is
 * target[b*t_sb + c*t_sc + l*t_sl] (lets the (B, T, F) feature batch be read without a permute copy).
 * loss[0] = mean((recon - target)^2) ; grad = g_scale * 2 (recon - target) / (B*C*L)  (grad may be NULL). */
int vqs_mse_fwd_bwd(const float* recon, const float* target, int B, int C, int L, long long t_sb, long long t_sc,
                    long long t_sl, float g_scale, float* loss, float* grad, void* workspace, size_t workspace_bytes,
                    vqs_stream_t stream);

/* out = relu(in) ; out = a + b ; gin = (act > 0) ? g : 0  -- used by the module-level (autograd) path only. */
int vqs_relu_fwd(const float* in, long long n, float* out, vqs_stream_t stream);
int vqs_relu_bwd(const float* g, const float* act, long long n, float* gin, vqs_stream_t stream);
int vqs_add(const float* a, const float* b, long long n, float* out, vqs_stream_t stream);
/* out = x * s[0], s a DEVICE scalar: the upstream gradient of a loss node (autograd of nn.MSELoss on the module path). */
int vqs_scale(const float* x, const float* s, long long n, float* out, vqs_stream_t stream);
/* strided (B, L, C) -> NCL (B, C, L) copy: the `.permute(0, 2, 1).contiguous().float()` of convolutional_vq_vae.py:118. */
int vqs_blc_to_ncl(const float* in, int B, int L, int C, float* out, vqs_stream_t stream);
/* Speaker conditioning of the decoder input (deconvolutional_decoder.py:108-111; global_conditioning.py:52-57 repeats the
 * per-utterance feature vector over time and `torch.cat([x, speaker_embedding], dim=1)` appends it to the channels):
 * out (B, Ca + Cb, L) = [ a (B, Ca, L) | v (B, Cb) repeated over L ];  vqs_slice_channels is its gradient with respect to
 * a: the first Ca channels of g (B, C, L). */
int vqs_concat_channels(const float* a, const float* v, int B, int Ca, int Cb, int L, float* out, vqs_stream_t stream);
int vqs_slice_channels(const float* g, int B, int C, int Ca, int L, float* out, vqs_stream_t stream);

/* Eval-mode distance tables of the bottleneck (src/models/vector_quantizer.py:108-127; the EMA class raises NameError on
 * the same lines): Euclidean distances torch.dist(x, y, 2) between VQ rows, in itertools order.
 *   mode 0: product(rows of a, rows of b)   -> out[n * m],          out[i * m + j] = |a_i - b_j|      (frames vs codes)
 *   mode 1: combinations(rows of a, 2)      -> out[n (n - 1) / 2],  pairs (0,1), (0,2), ..., (1,2), ...
 * The rows of `a` follow `layout` exactly as in vqs_vq_assign (n = B * T rows; FLAT_ND: pass B = n, T = 1); b is a plain
 * (m, D) matrix (the codebook). */
int vqs_pairwise_l2(const float* a, int layout, int B, int D, int T, const float* b, int m, int mode, float* out,
                    vqs_stream_t stream);

/* Weight normalisation of the use_kaiming_normal configurations (nn.utils.weight_norm, dim 0: src/modules/
 * conv1d_builder.py:41-43, conv_transpose1d_builder.py:41-43, residual.py:45-47,57-59).  The weight (rows, cols) is the
 * (d0, d1*k) view of the conv parameter; g has `rows` entries.
 *   fwd: w = v * g / ||v||_row, norm[row] = ||v||_row (kept for the backward)
 *   bwd: grad_g = <dw, v>_row / norm,  grad_v = g / norm * (dw - v <dw, v>_row / norm^2)      (autograd of the above) */
int vqs_weight_norm_fwd(const float* v, const float* g, float* w, float* norm, int rows, int cols, vqs_stream_t stream);
int vqs_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* norm, float* grad_v,
                        float* grad_g, int rows, int cols, vqs_stream_t stream);

/* Feature normalisation of the data feed, on the GPU (reference: `(dic['input_features'] - train_mean) / train_std` in
 * numpy float64, src/dataset/vctk_features_dataset.py:56-58, then `.float()` in convolutional_vq_vae.py:118):
 *   out[i] = (float)((in[i] - mean[i % F]) / std[i % F])      in float64 -> bit-identical to the reference's values. */
int vqs_normalize_features(const double* in, const double* mean, const double* stdev, long long n, int F, float* out,
                           vqs_stream_t stream);

/* torch.optim.Adam(amsgrad=True, weight_decay=0) over one flat fp32 buffer, one launch:
 *   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; vmax <- max(vmax, v)
 *   p <- p - (lr / bc1) * m / (sqrt(vmax) / sqrt(bc2) + eps),  bc1 = 1 - b1^step, bc2 = 1 - b2^step
 * `step` is read from a device int64 (incremented by the kernel launch BEFORE use when inc_step != 0) so that the call
 * is CUDA-graph replayable.  g_scale multiplies the gradient first (1/world_size after a sum-allreduce).  Hyper-parameters
 * are doubles: the scalar algebra (1 - beta, bias corrections, lr / bc1) is done in double and rounded to fp32 once, as
 * torch does with its Python-float hyper-parameters. */
int vqs_amsgrad_step(float* p, const float* g, float* m, float* v, float* vmax, long long n, long long* step,
                     int inc_step, double lr, double beta1, double beta2, double eps, double g_scale, vqs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VQS_B200_H */
