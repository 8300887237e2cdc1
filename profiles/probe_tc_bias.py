"""Signed bias of the 3xTF32 tcgen05 conv GEMM against fp64 as a function of the reduction length (calibration / check of the
truncation-loss compensation in the epilogue of gemm_tc.cu): mean((ours - ref) * sign(ref)) / mean|ref|."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, functional as F
dev = torch.device('cuda:0')
torch.manual_seed(1)
ops.set_precision('3xtf32')
B, L, Cout = 64, 48, 768
for dist in ('normal', 'positive'):
    for Cin in (32, 64, 128, 256, 512, 768, 1024, 1536, 3072):
        for k in (1, 3):
            x = torch.randn(B, Cin, L, device=dev); w = torch.randn(Cout, Cin, k, device=dev) / (Cin * k) ** 0.5
            if dist == 'positive':
                x, w = x.abs(), w.abs()
            r = torch.nn.functional.conv1d(x.double(), w.double(), None, 1, k // 2)
            y = F.conv1d_forward(x, F.gemm_weight(w, 'conv_fwd'), None, 1, k // 2)
            d = y.double() - r
            nkb = Cin * k // 32
            print('BIAS %-8s K=%5d k-blocks=%3d adds/acc~%5.1f  signed bias %+.3e  L2 %.2e' % (
                dist, Cin * k, nkb, 4 * nkb / 3.0, ((d * r.sign()).mean() / r.abs().mean()).item(), (d.norm() / r.norm()).item()), flush=True)
