"""Conv / VQ building blocks on top of ops.py: the six conv-like GEMM mappings and the torch.autograd.Functions that give
the drop-in nn.Modules (modules.py, vector_quantizer.py, convolutional_vq_vae.py) a backward.  Forward and backward both
run hand-written kernels only."""
import torch
from torch.autograd import Function

from . import ops
from ._lib import LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
from .ops import MASK_FLOAT, MASK_NONE, MASK_U8


# ------------------------------------------------------------------------------------------------
# conv-like GEMM mappings (see include/vqs_b200.h for the index convention)
# ------------------------------------------------------------------------------------------------
def conv_out_len(Lin, k, stride, pad):
    return (Lin + 2 * pad - k) // stride + 1


def convT_out_len(Lin, k, pad):
    return Lin - 1 + k - 2 * pad


# GEMM-ready weight arrangements.  role -> permute mode for (operand image, tap-major matrix, canonical [None = as is]).
# The tensor-core engines take the pre-split, pre-swizzled operand IMAGE (vqs_permute_weight modes 3 / 4; Cred is padded with
# zero channels to whole 32-wide k-blocks); everything else runs the canonical arrangement on the exact-fp32 CUDA-core engine.
import collections
import os

GemmW = collections.namedtuple('GemmW', 'A tap M Cred ksz')   # tap: 0 canonical, 1 tap-major matrix, 2 operand image
_NO_IMAGE = os.environ.get('VQS_NO_IMAGE', '0') == '1'    # A/B probe: tap-major matrices instead of operand images
_ROLES = {'conv_fwd': (3, None), 'conv_dgrad': (4, 0), 'convT_fwd': (4, 0), 'convT_dgrad': (3, None)}


def _role_dims(w_shape, role):
    d0, d1, k = w_shape
    return (d0, d1, k) if role in ('conv_fwd', 'convT_dgrad') else (d1, d0, k)     # (M, Cred, ksz)


def conv_tc_eligible(M, Cred, precision=None):
    """True when a conv-like GEMM with M output rows and reduction width Cred (channels) runs on tcgen05: a tensor-core
    precision and either whole 32-wide k-blocks, or more than one of them under at least one full 128-row tile (the 39 MFCC
    channels of the 768-wide model: an operand image padded to 64 with zero channels; small layers would be mostly padding in
    both directions and stay on the exact CUDA-core kernel)."""
    prec = ops.get_precision() if precision is None else precision
    return prec != 'fp32' and (Cred % 32 == 0 or (Cred > 32 and M >= 128 and not _NO_IMAGE))


def gemm_weight_layout(w_shape, role, precision=None):
    """(tap, permute mode or None, number of floats of the GEMM operand buffer) for a weight of shape w_shape."""
    M, Cred, k = _role_dims(w_shape, role)
    if conv_tc_eligible(M, Cred, precision):
        if _NO_IMAGE:     # plain tap-major fp32 matrix [M][ksz][Cred]: the GEMM's producer warps load and split it themselves
            return 1, {3: 1, 4: 2}[_ROLES[role][0]], M * Cred * k
        return 2, _ROLES[role][0], ((M + 127) // 128) * (k * ((Cred + 31) // 32)) * 8192
    return 0, _ROLES[role][1], M * Cred * k


def gemm_weight(w, role, out=None, precision=None):
    """The dense GEMM operand of `role` for weight w: a GemmW (re-arranged by vqs_permute_weight when needed)."""
    tap, mode, numel = gemm_weight_layout(tuple(w.shape), role, precision)
    M, Cred, k = _role_dims(tuple(w.shape), role)
    if mode is None:
        return GemmW(w, tap, M, Cred, k)
    if out is None:
        out = torch.empty(numel, dtype=torch.float32, device=w.device)
    ops.permute_weight(w, out=out, mode=mode)
    return GemmW(out, tap, M, Cred, k)


def conv1d_forward(x, A, b, stride, pad, out=None, x_strides=None, x_shape=None, **epi):
    """nn.Conv1d forward: y[b,o,l] = bias[o] + sum_{c,j} w[o,c,j] x[b,c,l*stride + j - pad].   A = gemm_weight(w, 'conv_fwd')."""
    Cout, Cin, k = A.M, A.Cred, A.ksz
    B, _, Lin = x_shape if x_shape is not None else x.shape
    Lout = conv_out_len(Lin, k, stride, pad)
    if out is None:
        out = torch.empty(B, Cout, Lout, dtype=torch.float32, device=A.A.device)
    return ops.conv_gemm(A.A, x, out, Cout, Cin, k, B, Lin, Lout, stride, 1, -pad, 1, x_strides=x_strides, bias=b,
                         a_tap_major=A.tap, **epi)


def conv1d_dgrad(gy, A, Lx, stride, pad, out=None, **epi):
    """dx[b,c,i] = sum_{o,j} w[o,c,j] gy[b,o,(i + pad - j)/stride].   A = gemm_weight(w, 'conv_dgrad')."""
    Cin, Cout, k = A.M, A.Cred, A.ksz
    B, _, Ly = gy.shape
    if out is None:
        out = torch.empty(B, Cin, Lx, dtype=torch.float32, device=gy.device)
    return ops.conv_gemm(A.A, gy, out, Cin, Cout, k, B, Ly, Lx, 1, -1, pad, stride, a_tap_major=A.tap, **epi)


def conv1d_wgrad(gy, x, dW, stride, pad, ws, x_relu=False, accumulate=False):
    """dW[o,c,j] (+)= sum_{b,l} gy[b,o,l] x[b,c,l*stride + j - pad]."""
    Cout, Cin, k = dW.shape
    B, _, Ly = gy.shape
    return ops.wgrad_gemm(gy, x, dW, Cout, Cin, k, B, Ly, x.shape[2], stride, 1, -pad, ws, x_relu=x_relu,
                          accumulate=accumulate)


def convT1d_forward(x, A, b, pad, out_len=None, out=None, **epi):
    """nn.ConvTranspose1d (stride 1) forward: y[b,o,i] = bias[o] + sum_{c,j} x[b,c,i - j + pad] w[c,o,j].
    A = gemm_weight(w, 'convT_fwd').  out_len < full length computes only the first out_len positions."""
    Cout, Cin, k = A.M, A.Cred, A.ksz
    B, _, Lin = x.shape
    Lout = convT_out_len(Lin, k, pad) if out_len is None else out_len
    if out is None:
        out = torch.empty(B, Cout, Lout, dtype=torch.float32, device=x.device)
    return ops.conv_gemm(A.A, x, out, Cout, Cin, k, B, Lin, Lout, 1, -1, pad, 1, bias=b, a_tap_major=A.tap, **epi)


def convT1d_dgrad(gy, A, Lx, pad, out=None, **epi):
    """dx[b,c,l] = sum_{o,j} gy[b,o,l + j - pad] w[c,o,j]  (gy may be the trimmed tensor: positions beyond it are 0).
    A = gemm_weight(w, 'convT_dgrad')."""
    Cin, Cout, k = A.M, A.Cred, A.ksz
    B, _, Ly = gy.shape
    if out is None:
        out = torch.empty(B, Cin, Lx, dtype=torch.float32, device=gy.device)
    return ops.conv_gemm(A.A, gy, out, Cin, Cout, k, B, Ly, Lx, 1, 1, -pad, 1, a_tap_major=A.tap, **epi)


def convT1d_wgrad(gy, x, dW, pad, ws, accumulate=False):
    """dW[c,o,j] (+)= sum_{b,l} x[b,c,l] gy[b,o,l + j - pad]."""
    Cin, Cout, k = dW.shape
    B, _, Lx = x.shape
    return ops.wgrad_gemm(x, gy, dW, Cin, Cout, k, B, Lx, gy.shape[2], 1, 1, -pad, ws, accumulate=accumulate)


def _wgrad_ws(M, Cred, k, B, La, device):
    n = ops.wgrad_workspace_bytes(M, Cred, k, B, La)
    return torch.empty(max(n, 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------
# autograd Functions (module-level path)
# ------------------------------------------------------------------------------------------------
class Conv1dFn(Function):
    """y = [relu](conv1d(x, w) + b) with hand-written forward / dgrad / wgrad kernels."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, relu):
        x = x.contiguous()
        out = conv1d_forward(x, gemm_weight(w.contiguous(), 'conv_fwd'), None if b is None else b.contiguous(), stride,
                             pad, relu=relu)
        ctx.save_for_backward(x, w, out if relu else None)
        ctx.cfg = (stride, pad, relu, b is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        stride, pad, relu, has_b = ctx.cfg
        g = g.contiguous()
        if relu:
            g = ops.relu_bwd(g, out)
        Cout, Cin, k = w.shape
        B = x.shape[0]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = conv1d_dgrad(g, gemm_weight(w.contiguous(), 'conv_dgrad'), x.shape[2], stride, pad)
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(w)
            conv1d_wgrad(g, x, dW, stride, pad, _wgrad_ws(Cout, Cin, k, B, g.shape[2], g.device))
        if has_b and ctx.needs_input_grad[2]:
            db = ops.bias_grad(g, torch.empty(Cout, dtype=torch.float32, device=g.device))
        return dx, dW, db, None, None, None


class ConvTranspose1dFn(Function):
    """y = [relu](conv_transpose1d(x, w, stride=1) + b); out_len trims the output to its first out_len positions."""

    @staticmethod
    def forward(ctx, x, w, b, pad, relu, out_len):
        x = x.contiguous()
        out = convT1d_forward(x, gemm_weight(w.contiguous(), 'convT_fwd'), None if b is None else b.contiguous(), pad,
                              out_len=out_len, relu=relu)
        ctx.save_for_backward(x, w, out if relu else None)
        ctx.cfg = (pad, relu, b is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out = ctx.saved_tensors
        pad, relu, has_b = ctx.cfg
        g = g.contiguous()
        if relu:
            g = ops.relu_bwd(g, out)
        Cin, Cout, k = w.shape
        B = x.shape[0]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = convT1d_dgrad(g, gemm_weight(w.contiguous(), 'convT_dgrad'), x.shape[2], pad)
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(w)
            convT1d_wgrad(g, x, dW, pad, _wgrad_ws(Cin, Cout, k, B, x.shape[2], g.device))
        if has_b and ctx.needs_input_grad[2]:
            db = ops.bias_grad(g, torch.empty(Cout, dtype=torch.float32, device=g.device))
        return dx, dW, db, None, None, None


class ReluFn(Function):
    @staticmethod
    def forward(ctx, x):
        out = ops.relu_fwd(x.contiguous())
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        return ops.relu_bwd(g.contiguous(), out)


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, g):
        return g, g


class Upsample2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample2_fwd(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.upsample2_bwd(g.contiguous())


class JitterFn(Function):
    @staticmethod
    def forward(ctx, x, src):
        ctx.save_for_backward(src)
        return ops.jitter_fwd(x.contiguous(), src)

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        return ops.jitter_bwd(g.contiguous(), src), None


class MseLossFn(Function):
    """nn.MSELoss()(recon, target) with target given as a strided (b, c, l) view; one pass computes loss and gradient."""

    @staticmethod
    def forward(ctx, recon, target):
        recon = recon.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=recon.device)
        grad = torch.empty_like(recon)
        ops.mse_fwd_bwd(recon, target, target.stride(), 1.0, loss, grad, ops.mse_workspace(recon.device))
        ctx.save_for_backward(grad)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        # g is the scalar upstream gradient (1.0 in the trainer)
        return ops.scale(grad, g.reshape(1).to(torch.float32).contiguous()), None


class WeightNormFn(Function):
    """w = g v / ||v|| over all dims but 0 (torch._weight_norm, dim = 0) and its gradient: vqs_weight_norm_fwd / _bwd."""

    @staticmethod
    def forward(ctx, v, g):
        v, g = v.contiguous(), g.contiguous()
        w = torch.empty_like(v)
        norm = torch.empty(v.shape[0], dtype=torch.float32, device=v.device)
        ops.weight_norm_fwd(v, g, w, norm)
        ctx.save_for_backward(v, g, norm)
        return w

    @staticmethod
    def backward(ctx, dw):
        v, g, norm = ctx.saved_tensors
        gv, gg = torch.empty_like(v), torch.empty_like(g)
        ops.weight_norm_bwd(dw.contiguous(), v, g, norm, gv, gg)
        return gv, gg


def weight_norm(v, g):
    return WeightNormFn.apply(v, g)


class ConcatChannelsFn(Function):
    """(B, Ca, L) ++ (B, Cb) repeated over L (speaker conditioning, deconvolutional_decoder.py:108-111); the features are
    constants (a fresh embedding per call), so only the first Ca channels pass a gradient."""

    @staticmethod
    def forward(ctx, a, v):
        ctx.Ca = a.shape[1]
        return ops.concat_channels(a.contiguous(), v.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.slice_channels(g.contiguous(), ctx.Ca), None


def concat_channels(a, v):
    return ConcatChannelsFn.apply(a, v)


def conv1d(x, w, b, stride=1, pad=0, relu=False):
    return Conv1dFn.apply(x, w, b, stride, pad, relu)


def conv_transpose1d(x, w, b, pad=0, relu=False, out_len=None):
    return ConvTranspose1dFn.apply(x, w, b, pad, relu, out_len)


def relu(x):
    return ReluFn.apply(x)


def add(a, b):
    return AddFn.apply(a, b)


def upsample2(x):
    return Upsample2Fn.apply(x)


def jitter(x, src):
    return JitterFn.apply(x, src)


def mse_loss(recon, target):
    return MseLossFn.apply(recon, target)


# ------------------------------------------------------------------------------------------------
# VQ bottleneck Function
# ------------------------------------------------------------------------------------------------
class VQFn(Function):
    """The bottleneck of VectorQuantizer / VectorQuantizerEMA (reference vector_quantizer_ema.py:101-179).

    forward(z, codebook, state) -> (quantized_ste, scalars) where scalars = [sse, e_latent, perplexity, beta*e_latent,
    e_latent + beta*e_latent].  `state` is a dict owned by the module: ws, layout, beta, ema (None or dict with
    cluster_size, ema_w, decay, eps), training, stats_allreduce (callable or None), want (dict of optional output
    buffers filled in place: idx, stats, dmin2, distances, q_rows).
    backward: grad_z = g_q + g_loss * 2 beta (x - q) / (N D); non-EMA also grad_E (vector_quantizer.py:136-139).
    """

    @staticmethod
    def forward(ctx, z, codebook, state):
        z = z.contiguous()
        layout, beta, ws = state['layout'], state['beta'], state['ws']
        K, D = codebook.shape
        ema = state['ema']
        cb = codebook.detach()
        want = state['want']
        idx, stats = ops.vq_assign(z, cb, layout, ws, idx=want.get('idx'), stats=want.get('stats'),
                                   dmin2=want.get('dmin2'), distances=want.get('distances'))
        N = idx.numel()
        n_total = N
        if state['training'] and ema is not None:
            if state.get('stats_allreduce') is not None:
                n_total = state['stats_allreduce'](stats, N)
            ops.vq_ema_update(ema['cluster_size'], ema['ema_w'], cb, stats, ema['decay'], ema['eps'])
        out, scalars = ops.vq_quantize(z, idx, cb, layout, ws, stats[:K], n_total, beta, q_rows=want.get('q_rows'))
        ctx.save_for_backward(z, idx, cb, stats)
        ctx.cfg = (layout, beta, ema is not None, N, D)
        want['idx'] = idx
        want['stats'] = stats
        return out, scalars

    @staticmethod
    def backward(ctx, g_out, g_scalars):
        z, idx, cb, stats = ctx.saved_tensors
        layout, beta, is_ema, N, D = ctx.cfg
        dev = z.device
        if g_scalars is None:
            g_scalars = torch.zeros(8, dtype=torch.float32, device=dev)
        g_scalars = g_scalars.contiguous()
        # the loss the modules expose is scalars[3] (EMA) or scalars[4] (non-EMA); both are affine in e_latent:
        #   EMA:      vq_loss = beta * e_latent                       -> d/dz = g * 2 beta (x - q) / (N D)
        #   non-EMA:  vq_loss = q_latent + beta * e_latent (same value; q_latent's input is detached) -> same d/dz
        gl = g_scalars[3:4] if is_ema else g_scalars[4:5]
        if g_out is None:
            g_out = torch.zeros_like(z)
        gz = ops.vq_backward(g_out.contiguous(), gl, 2.0 * beta / (N * D), z, idx, cb, layout)
        gE = None
        if not is_ema and ctx.needs_input_grad[1]:
            gE = ops.vq_grad_codebook(stats, cb, gl, 2.0 / (N * D))
        return gz, gE, None
