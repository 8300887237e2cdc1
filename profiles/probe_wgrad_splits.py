"""wgrad split-K sweep (profiling build: VQS_EXTRA_NVCC_FLAGS=-DVQS_DEBUG); VQS_WGRAD_SPLITS=n forces the split."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import functional as F, ops  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
ops.set_precision('3xtf32')
for (B, C, L, k) in ((64, 768, 48, 3), (64, 768, 24, 3), (64, 768, 47, 3)):
    x = torch.randn(B, C, L, device=dev)
    gy = torch.randn(B, C, L, device=dev)
    dW = torch.empty(C, C, k, device=dev)
    ws = torch.empty(64 * C * C * k * 4, dtype=torch.uint8, device=dev)
    fn = lambda: F.conv1d_wgrad(gy, x, dW, 1, k // 2, ws)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print('splits=%s B%d C%d L%d k%d wgrad %.4f ms' % (os.environ.get('VQS_WGRAD_SPLITS', 'plan'), B, C, L, k, e0.elapsed_time(e1) / 30), flush=True)
