"""Times the large-codebook tensor-core search (flat and (B, D, T) layouts) with CUDA events."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND, LAYOUT_BDT_AS_DTB
dev = torch.device('cuda:0')
ops.vq_set_engine('tensor_core')
for K in (512, 4096):
    D, B, T = 64, 512, 128
    N = B * T
    W = torch.randn(K, D, device=dev)
    ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
    for name, layout, z in (('flat', LAYOUT_FLAT_ND, torch.randn(N, D, device=dev)), ('bdt', LAYOUT_BDT_AS_DTB, torch.randn(B, D, T, device=dev))):
        for _ in range(2): ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print('PROBE K=%d N=%d %s: %.3f ms  %.1f TFLOP/s (algorithmic 2NDK)' % (K, N, name, ms, 2.0 * N * D * K / (ms * 1e-3) / 1e12))
