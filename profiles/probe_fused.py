"""CUDA-event timing of the four kernels of the fused VQ metric (flat rows, N = 2^22, K = 44), as bench.py's vq object."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
K, D, N = 44, 64, 1 << 22
gen = torch.Generator(device=dev).manual_seed(7)
W = torch.randn(K, D, device=dev, generator=gen); z = torch.randn(N, D, device=dev, generator=gen); g = torch.randn(N, D, device=dev, generator=gen)
cs = torch.zeros(K, device=dev); ew = torch.randn(K, D, device=dev, generator=gen)
ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
q = torch.empty_like(z); gz = torch.empty_like(z); sc = torch.zeros(8, device=dev); one = torch.ones(1, device=dev)
fs = [('assign', lambda: ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)),
      ('ema', lambda: ops.vq_ema_update(cs, ew, W, st, 0.99, 1e-5)),
      ('quantize', lambda: ops.vq_quantize(z, idx, W, LAYOUT_FLAT_ND, ws, st[:K], N, 0.25, out=q, scalars=sc)),
      ('backward', lambda: ops.vq_backward(g, one, 2 * 0.25 / (N * D), z, idx, W, LAYOUT_FLAT_ND, out=gz))]
for mode in ('0', '1', '2'):       # VQS_EW_FLAT_TILE: 0 grid-stride kernels, 1 tiled forward, 2 tiled forward + backward (default)
    os.environ['VQS_EW_FLAT_TILE'] = mode
    for _ in range(3):
        for n, f in fs: f()
    torch.cuda.synchronize()
    acc = dict((n, 0.0) for n, _ in fs); iters = 20; evs = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t0.record()
    for _ in range(iters):
        for n, f in fs:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); evs.append((n, a, b))
    t1.record(); torch.cuda.synchronize()
    for n, a, b in evs: acc[n] += a.elapsed_time(b) / iters
    tot = t0.elapsed_time(t1) / iters
    print('FUSED flat_tile=%s %s  total %.4f ms  %.3f G rows/s  %.1f %% of 6536 GB/s' % (mode, ' '.join('%s %.4f' % (n, acc[n]) for n, _ in fs), tot, N / tot / 1e6, 1296 * N / tot / 1e6 / 6536.4 * 100))
