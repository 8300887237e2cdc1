"""One launch each of the VQ kernels of the fused metric (flat rows and the reference's (B, D, T) rows, N = 2^22, K = 44) for
    ncu --set full --clock-control none --import-source on -k regex:"vq_assign_tma_kernel|vq_elementwise" -c 6 -o OUT python profiles/probe_ncu_vq.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
K, D, B, T = 44, 64, 32768, 128
N = B * T
gen = torch.Generator(device=dev).manual_seed(7)
W = torch.randn(K, D, device=dev, generator=gen)
ws = ops.vq_workspace(K, D, dev)
one = torch.ones(1, device=dev)
for layout, shape in ((LAYOUT_FLAT_ND, (N, D)), (LAYOUT_BDT_AS_DTB, (B, D, T))):
    z = torch.randn(*shape, device=dev, generator=gen)
    g = torch.randn(*shape, device=dev, generator=gen)
    idx, st = ops.vq_assign(z, W, layout, ws)
    out, sc = ops.vq_quantize(z, idx, W, layout, ws, st[:K], N, 0.25)
    gz = ops.vq_backward(g, one, 2 * 0.25 / (N * D), z, idx, W, layout)
    torch.cuda.synchronize()
    del z, g, out, gz
print('done')
