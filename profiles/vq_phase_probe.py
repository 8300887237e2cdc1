"""Times vqs_vq_assign (flat rows, N = 2^22, K = 44, D = 64) with CUDA events; run under VQS_TC_DEBUG=0/1/2/3 to see what
each phase of the tensor-core search costs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
K, D, N = 44, 64, 1 << 22
W = torch.randn(K, D, device=dev); z = torch.randn(N, D, device=dev)
ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
for eng in ('tensor_core', 'cuda_core'):
    ops.vq_set_engine(eng)
    for _ in range(3): ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
    e1.record(); torch.cuda.synchronize()
    print('VQS_TC_DEBUG=%s engine=%s  %.3f ms' % (os.environ.get('VQS_TC_DEBUG', '0'), eng, e0.elapsed_time(e1) / 10))
