"""Accuracy of the conv-like GEMMs at the benchmarked shapes against fp64 (torch double on the GPU = test infrastructure):
relative L2 error, max-norm error, and the SIGNED relative bias mean((ours - ref) * sign(ref)) / mean|ref| -- a negative bias
is the tensor core's truncating accumulation shrinking every sum."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, functional as F
dev = torch.device('cuda:0')
torch.manual_seed(0)
def stats(a, r):
    a = a.double(); d = a - r
    return (d.norm() / r.norm()).item(), (d.abs().max() / r.abs().max()).item(), ((d * r.sign()).mean() / r.abs().mean()).item()
for (B, C, L, k, pad) in ((64, 768, 48, 3, 1), (64, 768, 24, 3, 1), (64, 768, 48, 1, 0), (8, 768, 48, 3, 1)):
    x = torch.randn(B, C, L, device=dev); w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    gy = torch.randn(B, C, L, device=dev)
    y64 = torch.nn.functional.conv1d(x.double(), w.double(), None, 1, pad)
    dx64 = torch.nn.functional.conv_transpose1d(gy.double(), w.double(), None, 1, pad)
    dw64 = torch.nn.grad.conv1d_weight(x.double(), w.shape, gy.double(), stride=1, padding=pad)
    y32 = torch.nn.functional.conv1d(x, w, None, 1, pad)
    for prec in ('fp32', '3xtf32'):
        ops.set_precision(prec)
        y = F.conv1d_forward(x, F.gemm_weight(w, 'conv_fwd'), None, 1, pad)
        dx = F.conv1d_dgrad(gy, F.gemm_weight(w, 'conv_dgrad'), L, 1, pad)
        dW = torch.empty_like(w)
        F.conv1d_wgrad(gy, x, dW, 1, pad, F._wgrad_ws(C, C, k, B, L, dev))
        for name, a, r in (('fwd', y, y64), ('dgrad', dx, dx64), ('wgrad', dW, dw64)):
            print('ACC B=%d C=%d L=%d k=%d %-6s %-5s L2 %.2e  max %.2e  signed bias %+.2e' % ((B, C, L, k, prec, name) + stats(a, r)), flush=True)
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    print('ACC B=%d C=%d L=%d k=%d %-6s %-5s L2 %.2e  max %.2e  signed bias %+.2e' % ((B, C, L, k, 'cudnn', 'fwd') + stats(y32, y64)), flush=True)
# transposed convolutions (canonical weight layout on the CUDA-core engine), incl. the trimmed conv_trans_2 of the step
for (B, C, L, k, pad, keep) in ((64, 768, 48, 3, 0, 47), (64, 768, 48, 3, 1, 48), (8, 768, 48, 3, 0, 47)):
    x = torch.randn(B, C, L, device=dev); w = torch.randn(C, C, k, device=dev) / (C * k) ** 0.5
    y64 = torch.nn.functional.conv_transpose1d(x.double(), w.double(), None, 1, pad)[:, :, :keep]
    gy = torch.randn(B, C, keep, device=dev)
    gfull = torch.zeros(B, C, L - 1 + k - 2 * pad, device=dev, dtype=torch.float64); gfull[:, :, :keep] = gy.double()
    dx64 = torch.nn.functional.conv1d(gfull, w.double().transpose(0, 1).contiguous().transpose(0, 1), None, 1, pad) if False else None
    xr = x.double().requires_grad_(True); wr = w.double().requires_grad_(True)
    yy = torch.nn.functional.conv_transpose1d(xr, wr, None, 1, pad)
    (yy[:, :, :keep] * gy.double()).sum().backward()
    for prec in ('fp32', '3xtf32'):
        ops.set_precision(prec)
        y = F.convT1d_forward(x, F.gemm_weight(w, 'convT_fwd'), None, pad, out_len=keep)
        dx = F.convT1d_dgrad(gy, F.gemm_weight(w, 'convT_dgrad'), L, pad)
        dW = torch.empty_like(w)
        F.convT1d_wgrad(gy, x, dW, pad, F._wgrad_ws(C, C, k, B, L, dev))
        for name, a, r in (('fwd', y, y64), ('dgrad', dx, xr.grad), ('wgrad', dW, wr.grad)):
            print('ACC convT B=%d C=%d L=%d k=%d pad=%d keep=%d %-6s %-5s L2 %.2e  max %.2e  signed bias %+.2e' % ((B, C, L, k, pad, keep, prec, name) + stats(a, r)), flush=True)
