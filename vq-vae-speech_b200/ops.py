"""Tensor-level wrappers over the C ABI (include/vqs_b200.h).  torch is used for device memory and streams only:
every function here launches hand-written kernels from libvqs_b200.so on torch's current CUDA stream and never falls
back to a torch/ATen implementation (CPU tensors raise)."""
import ctypes

import torch

from . import _lib
from ._lib import LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND, ConvGemmDesc, WgradDesc

MASK_NONE, MASK_FLOAT, MASK_U8 = 0, 1, 2

# GEMM engine used by conv_gemm / wgrad_gemm when the caller does not pass one: 'fp32' (exact, CUDA cores),
# 'tf32' (tcgen05, single pass) or '3xtf32' (tcgen05, fp32-accurate split).  Captured into recorded schedules.
_PRECISION = _lib.PREC_TF32X3


def set_precision(name):
    """Selects the default GEMM engine; returns the previous one (name)."""
    global _PRECISION
    prev = [k for k, v in _lib.PRECISIONS.items() if v == _PRECISION][0]
    _PRECISION = _lib.PRECISIONS[name]
    return prev


def get_precision():
    return [k for k, v in _lib.PRECISIONS.items() if v == _PRECISION][0]


def _stream():
    return torch.cuda.current_stream().cuda_stream


# When a recorder list is installed (trainer.FusedTrainStep builds its schedule this way) launches are appended to it as
# (c_function, args, keepalive) instead of being issued; replaying the list re-issues them with the then-current stream.
_RECORDER = None


def set_recorder(rec):
    global _RECORDER
    prev = _RECORDER
    _RECORDER = rec
    return prev


def record_callable(fn):
    """Adds a host callable (e.g. an NCCL collective) to the schedule being recorded; runs it now when not recording."""
    if _RECORDER is not None:
        _RECORDER.append((None, fn, None))
    else:
        fn()


def _call(name, args, keep=None):
    fn = getattr(_lib.load(), name)
    if _RECORDER is not None:
        _RECORDER.append((fn, args, keep))
        return
    _lib.check(fn(*args, _stream()))


def replay(schedule, events=None):
    """Re-issues a recorded schedule on the current stream.  With `events` (a list) every launch is bracketed by a pair of
    CUDA events recorded on that same stream and (index, start, end) is appended -- bench.py's per-kernel timing."""
    st = _stream()
    if events is None:
        for fn, args, _ in schedule:
            if fn is None:
                args()
            else:
                rc = fn(*args, st)
                if rc != 0:
                    _lib.check(rc)
        return
    cur = torch.cuda.current_stream()
    for i, (fn, args, _) in enumerate(schedule):
        if fn is None:
            args()
            continue
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        rc = fn(*args, st)
        e1.record(cur)
        if rc != 0:
            _lib.check(rc)
        events.append((i, e0, e1))


def _p(t, dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor of the expected dtype (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('libvqs_b200 has no CPU path: got a tensor on %s' % t.device)
    if t.dtype != dtype:
        raise RuntimeError('expected %s tensor, got %s' % (dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError('expected a contiguous tensor, got strides %s for shape %s' % (t.stride(), tuple(t.shape)))
    return t.data_ptr()


def _pany(t):
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise RuntimeError('expected a contiguous CUDA tensor')
    return t.data_ptr()


# ------------------------------------------------------------------------------------------------
# VQ bottleneck
# ------------------------------------------------------------------------------------------------
def vq_set_engine(name):
    """'auto' (default: CUDA cores for resident codebooks, tcgen05 distance GEMM for large ones), 'tensor_core' (tcgen05
    wherever a kernel exists) or 'cuda_core'."""
    _lib.check(_lib.load().vqs_vq_set_engine({'tensor_core': 0, 'auto': 1, 'cuda_core': 2}[name]))


def vq_workspace_bytes(K, D):
    return int(_lib.load().vqs_vq_workspace_bytes(K, D))


def vq_workspace(K, D, device):
    return torch.empty(vq_workspace_bytes(K, D), dtype=torch.uint8, device=device)


def vq_shape(z, layout, D):
    """(B, D, T) triple the C ABI expects.  FLAT_ND: z is (N, D) -> (N, D, 1) i.e. B=N, T=1."""
    if layout == LAYOUT_FLAT_ND:
        if z.dim() != 2 or z.shape[1] != D:
            raise RuntimeError('flat layout expects (N, %d), got %s' % (D, tuple(z.shape)))
        return z.shape[0], D, 1
    if z.dim() != 3 or z.shape[1] != D:
        raise RuntimeError('VQ input must be (B, %d, T), got %s' % (D, tuple(z.shape)))
    return z.shape[0], D, z.shape[2]


def vq_assign(z, codebook, layout, ws, idx=None, stats=None, dmin2=None, distances=None):
    K, D = codebook.shape
    B, _, T = vq_shape(z, layout, D)
    N = B * T
    if idx is None:
        idx = torch.empty(N, dtype=torch.int64, device=z.device)
    if stats is None:
        stats = torch.empty(K * (D + 1), dtype=torch.float32, device=z.device)
    _call('vqs_vq_assign', (_p(z), layout, B, D, T, _p(codebook), K, _p(idx, torch.int64), _p(stats),
                                         _p(dmin2), _p(distances), _pany(ws), ws.numel()))
    return idx, stats


def vq_one_hot(idx, K, out=None):
    N = idx.numel()
    if out is None:
        out = torch.empty(N, K, dtype=torch.float32, device=idx.device)
    _call('vqs_vq_one_hot', (_p(idx, torch.int64), N, K, _p(out)))
    return out


def vq_ema_update(cluster_size, ema_w, embedding, stats, decay, eps):
    K, D = embedding.shape
    _call('vqs_vq_ema_update', (_p(cluster_size), _p(ema_w), _p(embedding), _p(stats), float(decay),
                                             float(1 - decay), float(eps), float(K * eps), K, D))


def vq_quantize(z, idx, codebook, layout, ws, counts, n_rows_total, beta, out=None, q_rows=None, scalars=None):
    K, D = codebook.shape
    B, _, T = vq_shape(z, layout, D)
    if out is None:
        out = torch.empty_like(z)
    if scalars is None:
        scalars = torch.empty(8, dtype=torch.float32, device=z.device)
    _call('vqs_vq_quantize', (_p(z), layout, B, D, T, _p(idx, torch.int64), _p(codebook), K, _p(out),
                                           _p(q_rows), _p(counts), float(n_rows_total), float(beta), _p(scalars),
                                           _pany(ws), ws.numel()))
    return out, scalars


def vq_backward(g_out, g_loss, coef, z, idx, codebook, layout, out=None):
    K, D = codebook.shape
    B, _, T = vq_shape(z, layout, D)
    if out is None:
        out = torch.empty_like(z)
    _call('vqs_vq_backward', (_p(g_out), _p(g_loss), float(coef), _p(z), layout, B, D, T,
                                           _p(idx, torch.int64), _p(codebook), K, _p(out)))
    return out


def vq_gather(idx, codebook, layout, shape, out=None):
    """Gather-only forward: out (shape = z's shape, z's layout) = codebook[idx] exactly; z is not read (vqs_vq_quantize with
    z = NULL).  The losses then come from vq_backward_loss."""
    K, D = codebook.shape
    if out is None:
        out = torch.empty(*shape, dtype=torch.float32, device=codebook.device)
    B, _, T = vq_shape(out, layout, D)
    _call('vqs_vq_quantize', (None, layout, B, D, T, _p(idx, torch.int64), _p(codebook), K, _p(out), None, None, 0.0, 0.0,
                              None, None, 0))
    return out


def vq_backward_loss(g_out, g_loss, coef, z, idx, codebook, layout, ws, counts, n_rows_total, beta, out=None, scalars=None):
    """grad_z and, in the same sweep over z, the forward's losses (scalars as vq_quantize fills them)."""
    K, D = codebook.shape
    B, _, T = vq_shape(z, layout, D)
    if out is None:
        out = torch.empty_like(z)
    if scalars is None:
        scalars = torch.empty(8, dtype=torch.float32, device=z.device)
    _call('vqs_vq_backward_loss', (_p(g_out), _p(g_loss), float(coef), _p(z), layout, B, D, T, _p(idx, torch.int64),
                                   _p(codebook), K, _p(out), _p(counts), float(n_rows_total), float(beta), _p(scalars),
                                   _pany(ws), ws.numel()))
    return out, scalars


def vq_grad_codebook(stats, codebook, g_loss, coef, out=None, accumulate=False):
    K, D = codebook.shape
    if out is None:
        out = torch.empty_like(codebook)
    _call('vqs_vq_grad_codebook', (_p(stats), _p(codebook), _p(g_loss), float(coef), K, D, _p(out),
                                                int(accumulate)))
    return out


# ------------------------------------------------------------------------------------------------
# conv-like implicit GEMM
# ------------------------------------------------------------------------------------------------
def conv_gemm(A, X, out, M, Cred, ksz, B, Lin, Lout, l_mul, j_mul, off, l_div=1, x_strides=None, x_relu=False,
              bias=None, add_pre=None, add_pre_relu=False, relu=False, mask_out=None, mask=None, mask_kind=MASK_NONE,
              add_post=None, out2=None, mask2=None, mask2_kind=MASK_NONE, precision=None, a_tap_major=False,
              splitk_ws=None):
    """acc[b,m,l] = sum_{c,j} A[m, c*ksz+j] * X'[b, c, (l*l_mul + j*j_mul + off)/l_div] followed by the fused epilogue
    documented in include/vqs_b200.h.  x_strides = (batch, channel, position) element strides of X (default NCL)."""
    d = ConvGemmDesc()
    d.A = _p(A)
    d.X = X.data_ptr() if x_strides is not None else _p(X)
    d.a_tap_major = int(a_tap_major)          # 0 canonical, 1 tap-major matrix, 2 tensor-core operand image
    d.M, d.Cred, d.ksz = M, Cred, ksz
    d.B, d.Lin, d.Lout = B, Lin, Lout
    if x_strides is None:
        x_strides = (Cred * Lin, Lin, 1)
    d.x_sb, d.x_sc, d.x_sl = x_strides
    d.l_mul, d.j_mul, d.off, d.l_div = l_mul, j_mul, off, l_div
    d.x_relu = int(x_relu)
    d.bias = _p(bias)
    d.add_pre = _p(add_pre)
    d.add_pre_relu = int(add_pre_relu)
    d.relu = int(relu)
    d.mask_out = _p(mask_out, torch.uint8)
    d.mask = _pany(mask)
    d.mask_kind = mask_kind if mask is not None else MASK_NONE
    d.add_post = _p(add_post)
    d.out = _p(out)
    d.out2 = _p(out2)
    d.mask2 = _pany(mask2)
    d.mask2_kind = mask2_kind if mask2 is not None else MASK_NONE
    d.precision = _PRECISION if precision is None else _lib.PRECISIONS[precision]
    d.splitk_ws = _pany(splitk_ws)
    d.splitk_ws_bytes = splitk_ws.numel() * splitk_ws.element_size() if splitk_ws is not None else 0
    d._splitk_ws_ref = splitk_ws                # keeps the scratch tensor alive with the recorded descriptor
    _call('vqs_conv_gemm', (ctypes.byref(d),), d)
    return out


def wgrad_workspace_bytes(M, Cred, ksz, B, La):
    return int(_lib.load().vqs_wgrad_workspace_bytes(M, Cred, ksz, B, La))


def wgrad_gemm(Aact, X, dW, M, Cred, ksz, B, La, Lx, l_mul, j_mul, off, ws, x_relu=False, accumulate=False,
               precision=None):
    d = WgradDesc()
    d.Aact = _p(Aact)
    d.X = _p(X)
    d.M, d.Cred, d.ksz = M, Cred, ksz
    d.B, d.La, d.Lx = B, La, Lx
    d.l_mul, d.j_mul, d.off = l_mul, j_mul, off
    d.x_relu = int(x_relu)
    d.dW = _p(dW)
    d.accumulate = int(accumulate)
    d.precision = _PRECISION if precision is None else _lib.PRECISIONS[precision]
    _call('vqs_wgrad_gemm', (ctypes.byref(d), _pany(ws), 0 if ws is None else ws.numel()), d)
    return dW


def bias_grad(g, db, accumulate=False):
    B, M, L = g.shape
    _call('vqs_bias_grad', (_p(g), B, M, L, _p(db), int(accumulate)))
    return db


def permute_weight(w, out=None, mode=0):
    """w[d0][d1][k] -> mode 0: [d1][d0][k]; mode 1: [d0][k][d1] (tap-major); mode 2: [d1][k][d0] (tap-major, swapped);
    modes 3 / 4: tensor-core operand image of the mode-1 / mode-2 matrix (see include/vqs_b200.h)."""
    d0, d1, k = w.shape
    if out is None:
        shape = {0: (d1, d0, k), 1: (d0, k, d1), 2: (d1, k, d0),
                 3: (((d0 + 127) // 128) * (k * ((d1 + 31) // 32)) * 8192,),
                 4: (((d1 + 127) // 128) * (k * ((d0 + 31) // 32)) * 8192,)}[mode]
        out = torch.empty(*shape, dtype=torch.float32, device=w.device)
    _call('vqs_permute_weight', (_p(w), d0, d1, k, mode, _p(out)))
    return out


def permute_weights(items):
    """[(w, out, mode), ...] -> the same re-arrangements as permute_weight, _lib.PERMUTE_MAX_ITEMS of them per launch."""
    items = list(items)
    for lo in range(0, len(items), _lib.PERMUTE_MAX_ITEMS):
        chunk = items[lo:lo + _lib.PERMUTE_MAX_ITEMS]
        arr = (_lib.PermuteItem * len(chunk))()
        for i, (w, out, mode) in enumerate(chunk):
            d0, d1, k = w.shape
            arr[i].w, arr[i].out = _p(w), _p(out)
            arr[i].d0, arr[i].d1, arr[i].k, arr[i].mode = d0, d1, k, int(mode)
        _call('vqs_permute_weights', (arr, len(chunk)), keep=(arr, chunk))


def pairwise_l2(a, layout, D, b=None):
    """Euclidean distances between VQ rows in itertools order: product(rows(a), rows(b)) when b is given (n * m values),
    combinations(rows(a), 2) otherwise (n (n - 1) / 2 values).  vector_quantizer.py:108-127."""
    B, D_, T = vq_shape(a, layout, D)
    n = B * T
    if b is not None:
        out = torch.empty(n * b.shape[0], dtype=torch.float32, device=a.device)
        _call('vqs_pairwise_l2', (_p(a), layout, B, D_, T, _p(b), b.shape[0], 0, _p(out)))
    else:
        out = torch.empty(n * (n - 1) // 2, dtype=torch.float32, device=a.device)
        if out.numel():
            _call('vqs_pairwise_l2', (_p(a), layout, B, D_, T, None, 0, 1, _p(out)))
    return out


def weight_norm_fwd(v, g, w, norm):
    """w = v * g / ||v|| per slice along dim 0 (nn.utils.weight_norm); norm receives ||v||."""
    rows = v.shape[0]
    _call('vqs_weight_norm_fwd', (_p(v), _p(g), _p(w), _p(norm), rows, v.numel() // rows))


def weight_norm_bwd(dw, v, g, norm, grad_v, grad_g):
    rows = v.shape[0]
    _call('vqs_weight_norm_bwd', (_p(dw), _p(v), _p(g), _p(norm), _p(grad_v), _p(grad_g), rows, v.numel() // rows))


def tensor_core_engine():
    return _PRECISION != _lib.PREC_FP32


# ------------------------------------------------------------------------------------------------
# element-wise pieces
# ------------------------------------------------------------------------------------------------
def upsample2_fwd(x, out=None):
    B, C, L = x.shape
    if out is None:
        out = torch.empty(B, C, 2 * L, dtype=torch.float32, device=x.device)
    _call('vqs_upsample2_fwd', (_p(x), B * C, L, _p(out)))
    return out


def upsample2_bwd(g, out=None):
    B, C, L2 = g.shape
    if out is None:
        out = torch.empty(B, C, L2 // 2, dtype=torch.float32, device=g.device)
    _call('vqs_upsample2_bwd', (_p(g), B * C, L2 // 2, _p(out)))
    return out


def jitter_fwd(x, src, out=None):
    B, C, L = x.shape
    if out is None:
        out = torch.empty_like(x)
    _call('vqs_jitter_fwd', (_p(x), B * C, L, _p(src, torch.int32), _p(out)))
    return out


def jitter_bwd(g, src, out=None):
    B, C, L = g.shape
    if out is None:
        out = torch.empty_like(g)
    _call('vqs_jitter_bwd', (_p(g), B * C, L, _p(src, torch.int32), _p(out)))
    return out


def relu_fwd(x, out=None):
    if out is None:
        out = torch.empty_like(x)
    _call('vqs_relu_fwd', (_p(x), x.numel(), _p(out)))
    return out


def relu_bwd(g, act, out=None):
    if out is None:
        out = torch.empty_like(g)
    _call('vqs_relu_bwd', (_p(g), _p(act), g.numel(), _p(out)))
    return out


def add(a, b, out=None):
    if out is None:
        out = torch.empty_like(a)
    _call('vqs_add', (_p(a), _p(b), a.numel(), _p(out)))
    return out


def scale(x, s, out=None):
    """x * s, s a one-element device tensor (the upstream gradient of a loss)."""
    if out is None:
        out = torch.empty_like(x)
    _call('vqs_scale', (_p(x), _p(s), x.numel(), _p(out)))
    return out


def blc_to_ncl(x, out=None):
    B, L, C = x.shape
    if out is None:
        out = torch.empty(B, C, L, dtype=torch.float32, device=x.device)
    _call('vqs_blc_to_ncl', (_p(x), B, L, C, _p(out)))
    return out


def concat_channels(a, v, out=None):
    """(B, Ca, L) ++ (B, Cb) repeated over L -> (B, Ca + Cb, L)   (speaker conditioning, deconvolutional_decoder.py:108-111)."""
    B, Ca, L = a.shape
    Cb = v.shape[1]
    if out is None:
        out = torch.empty(B, Ca + Cb, L, dtype=torch.float32, device=a.device)
    _call('vqs_concat_channels', (_p(a), _p(v), B, Ca, Cb, L, _p(out)))
    return out


def slice_channels(g, Ca, out=None):
    """First Ca channels of a (B, C, L) tensor as a contiguous (B, Ca, L) tensor."""
    B, C, L = g.shape
    if out is None:
        out = torch.empty(B, Ca, L, dtype=torch.float32, device=g.device)
    _call('vqs_slice_channels', (_p(g), B, C, Ca, L, _p(out)))
    return out


def mse_workspace(device):
    return torch.empty(148 * 8 * 8 * 2, dtype=torch.uint8, device=device)


def mse_fwd_bwd(recon, target, target_strides, g_scale, loss, grad, ws):
    """loss[0] = mean((recon - target)^2); grad = g_scale * 2 (recon - target) / numel (grad may be None)."""
    B, C, L = recon.shape
    sb, sc, sl = target_strides
    _call('vqs_mse_fwd_bwd', (_p(recon), target.data_ptr(), B, C, L, sb, sc, sl, float(g_scale), _p(loss),
                                           _p(grad), _pany(ws), ws.numel()))
    return loss, grad


def amsgrad_step(p, g, m, v, vmax, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, g_scale=1.0, inc_step=True):
    _call('vqs_amsgrad_step', (_p(p), _p(g), _p(m), _p(v), _p(vmax), p.numel(), _p(step, torch.int64),
                                            int(inc_step), float(lr), float(beta1), float(beta2), float(eps),
                                            float(g_scale)))


def dp_barrier(ctx, channel):
    """Cross-GPU barrier on the current stream (csrc/dp_nvls.cu).  ctx: parallel.NvlsExchange.ctx (kept alive by the caller)."""
    _call('vqs_dp_barrier', (ctypes.byref(ctx), int(channel)), ctx)


def dp_allreduce_small(ctx, src_ptrs, dst, channel):
    """dst (local) = sum over ranks, in rank order, of the symmetric vector behind src_ptrs (a _lib.DpPtrs)."""
    _call('vqs_dp_allreduce_small', (ctypes.byref(ctx), ctypes.byref(src_ptrs), _p(dst), dst.numel(), int(channel)),
          (ctx, src_ptrs))


def dp_amsgrad_step(ctx, mc_p, p_local, mc_g, m, v, vmax, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, inc_step=True,
                    ch_before=1, ch_after=2):
    """Adam(amsgrad=True) fused with the gradient allreduce over the NVLS multicast mappings mc_p / mc_g (addresses)."""
    _call('vqs_dp_amsgrad_step', (ctypes.byref(ctx), ctypes.c_void_p(int(mc_p)), _p(p_local), ctypes.c_void_p(int(mc_g)),
                                  _p(m), _p(v), _p(vmax), p_local.numel(), _p(step, torch.int64), int(inc_step), float(lr),
                                  float(beta1), float(beta2), float(eps), int(ch_before), int(ch_after)), ctx)


def dp_amsgrad_range_on(stream, ctx, mc_p, p_local, mc_g, m, v, vmax, lo, hi, step, lr, beta1=0.9, beta2=0.999, eps=1e-8,
                        inc_step=False, ch_before=1, ch_after=-1):
    """The optimizer + exchange of ONE BUCKET [lo, hi) of the flat buffers (vqs_dp_amsgrad_range), issued NOW on `stream` (a
    torch.cuda.Stream: the side stream that runs beside the backward pass) -- never recorded: callers wrap it in
    record_callable so that the fork / join with the main stream is part of the same host callable."""
    fn = getattr(_lib.load(), 'vqs_dp_amsgrad_range')
    _lib.check(fn(ctypes.byref(ctx), ctypes.c_void_p(int(mc_p)), _p(p_local), ctypes.c_void_p(int(mc_g)), _p(m), _p(v),
                  _p(vmax), int(lo), int(hi), _p(step, torch.int64), int(inc_step), float(lr), float(beta1), float(beta2),
                  float(eps), int(ch_before), int(ch_after), stream.cuda_stream))


def normalize_features(x64, mean64, std64, out=None):
    """(x - mean) / std in float64 on the device, stored as float32 (the reference normalises in numpy float64 and casts
    with .float()): x64 (..., F) float64, mean64 / std64 (F,) float64."""
    F = x64.shape[-1]
    if out is None:
        out = torch.empty(x64.shape, dtype=torch.float32, device=x64.device)
    _call('vqs_normalize_features', (_p(x64, torch.float64), _p(mean64, torch.float64), _p(std64, torch.float64),
                                     x64.numel(), F, _p(out)))
    return out
