"""Small driver for `ncu --set full`: a few launches of the hot kernels at bench shapes.
    python profiles/prof_kernels.py [conv|vq|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import functional as F, ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
dev = torch.device('cuda:0')
torch.manual_seed(0)
if what in ('conv', 'all'):
    B, C, L = 64, 768, 47
    x = torch.randn(B, C, L, device=dev)
    w = torch.randn(C, C, 3, device=dev) / 48
    b = torch.randn(C, device=dev)
    res = torch.randn(B, C, L, device=dev)
    mask_out = torch.empty(B, C, L, dtype=torch.uint8, device=dev)
    out = torch.empty(B, C, L, device=dev)
    dW = torch.empty_like(w)
    ws = F._wgrad_ws(C, C, 3, B, L, dev)
    for prec in ('3xtf32', 'tf32', 'fp32'):
        ops.set_precision(prec)
        A = F.gemm_weight(w, 'conv_fwd')
        for _ in range(2):
            F.conv1d_forward(x, A, b, 1, 1, out=out, relu=True, mask_out=mask_out, add_post=res)  # enc conv_2
            F.conv1d_forward(x, A, b, 1, 1, out=out)                                             # plain epilogue
            F.conv1d_wgrad(out, x, dW, 1, 1, ws)
    torch.cuda.synchronize()
if what in ('vq', 'all'):
    K, D = 44, 64
    Bv, T = 8192, 128
    N = Bv * T
    W = torch.randn(K, D, device=dev)
    ws = ops.vq_workspace(K, D, dev)
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    stats = torch.empty(K * (D + 1), device=dev)
    one = torch.ones(1, device=dev)
    sc = torch.zeros(8, device=dev)
    for layout, shape in ((LAYOUT_BDT_AS_DTB, (Bv, D, T)), (LAYOUT_FLAT_ND, (N, D))):
        z = torch.randn(*shape, device=dev)
        g = torch.randn(*shape, device=dev)
        q = torch.empty_like(z)
        for _ in range(2):
            ops.vq_assign(z, W, layout, ws, idx=idx, stats=stats)
            ops.vq_quantize(z, idx, W, layout, ws, stats[:K], N, 0.25, out=q, scalars=sc)
            ops.vq_backward(g, one, 1e-6, z, idx, W, layout, out=q)
        del z, g, q
    torch.cuda.synchronize()
print('done')
