"""Correctness + CUDA-event timing of the streaming (TMA-fed) VQ search against the CUDA-core search, flat rows."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
torch.manual_seed(0)
D = 64


def run(K, N, data):
    W = torch.randn(K, D, device=dev)
    if data == 'near_dup' and K > 4:
        W[K // 2] = W[1]; W[K - 1] = W[2] * (1 + 1e-7); W[3] = W[2] + 1e-6 * torch.randn(D, device=dev)
    if data == 'trained':
        z = W[torch.randint(0, K, (N,), device=dev)] + 0.1 * torch.randn(N, D, device=dev)
    else:
        z = torch.randn(N, D, device=dev)
    ws = ops.vq_workspace(K, D, dev)
    ops.vq_set_engine('cuda_core'); i1, s1 = ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws); i1, s1 = i1.clone(), s1.clone()
    ops.vq_set_engine('tensor_core'); i2, s2 = ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws); i2, s2 = i2.clone(), s2.clone()
    i3, s3 = ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws)
    torch.cuda.synchronize()
    d1, d2 = s1[K:], s2[K:]
    rel = float((d1 - d2).abs().max() / d1.abs().max().clamp_min(1e-30))
    print('K=%d N=%d %s: idx diff %d, counts equal %s, dw rel %.2e, deterministic %s' % (
        K, N, data, int((i1 != i2).sum()), torch.equal(s1[:K], s2[:K]), rel, torch.equal(s2, s3) and torch.equal(i2, i3)), flush=True)


if len(sys.argv) < 2 or sys.argv[1] == 'check':
    for K, N, data in [(44, 128, 'randn'), (44, 100, 'randn'), (44, 777, 'near_dup'), (29, 5000, 'trained'), (64, 100000, 'randn'),
                       (1, 300, 'randn'), (17, 4097, 'randn'), (44, 1 << 20, 'randn'), (44, 1 << 20, 'trained')]:
        run(K, N, data)
if len(sys.argv) < 2 or sys.argv[1] == 'time':
    K, N = 44, 1 << 22
    W = torch.randn(K, D, device=dev); z = torch.randn(N, D, device=dev)
    ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
    cs = torch.zeros(K, device=dev); ew = torch.randn(K, D, device=dev)
    for state in ('randn codebook', 'after 10 EMA updates'):
      if state != 'randn codebook':
        ops.vq_set_engine('auto')
        for _ in range(10):
            ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st); ops.vq_ema_update(cs, ew, W, st, 0.99, 1e-5)
      print(state, flush=True)
      for eng in ('tensor_core', 'cuda_core'):
        ops.vq_set_engine(eng)
        for _ in range(3): ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print('engine=%s N=%d K=%d: %.3f ms  %.2f G rows/s  %.0f GB/s read' % (eng, N, K, ms, N / ms / 1e6, N * 264 / ms / 1e6), flush=True)
