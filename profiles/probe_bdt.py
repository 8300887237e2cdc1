"""CUDA-event timing of the fused VQ metric in the reference's (B, D, T) row layout: streaming engine with the cp.async
gather vs the CUDA-core search, blocked vs memory-order element-wise kernels (VQS_EW_NO_BLK=1 in the environment selects the
latter), for a few (B, T) shapes.  Also checks that both engines return the same indices.

    python profiles/probe_bdt.py            # all shapes
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB
dev = torch.device('cuda:0')
K, D = 44, 64
for B, T in ((32768, 128), (1 << 16, 64), (43648, 96), (8192, 47), (256, 96)):
    N = B * T
    gen = torch.Generator(device=dev).manual_seed(7)
    W = torch.randn(K, D, device=dev, generator=gen); z = torch.randn(B, D, T, device=dev, generator=gen); g = torch.randn(B, D, T, device=dev, generator=gen)
    cs = torch.zeros(K, device=dev); ew = torch.randn(K, D, device=dev, generator=gen)
    ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
    q = torch.empty_like(z); gz = torch.empty_like(z); sc = torch.zeros(8, device=dev); one = torch.ones(1, device=dev)
    ops.vq_set_engine('cuda_core'); i_cc, s_cc = ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws); i_cc = i_cc.clone(); s_cc = s_cc.clone()
    ops.vq_set_engine('auto'); i_tc, s_tc = ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws)
    same = bool(torch.equal(i_cc, i_tc)) and bool(torch.equal(s_cc[:K], s_tc[:K]))
    for eng in ('cuda_core', 'auto'):
        ops.vq_set_engine(eng)
        fs = [('assign', lambda: ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws, idx=idx, stats=st)),
              ('ema', lambda: ops.vq_ema_update(cs, ew, W, st, 0.99, 1e-5)),
              ('quantize', lambda: ops.vq_quantize(z, idx, W, LAYOUT_BDT_AS_DTB, ws, st[:K], N, 0.25, out=q, scalars=sc)),
              ('backward', lambda: ops.vq_backward(g, one, 2 * 0.25 / (N * D), z, idx, W, LAYOUT_BDT_AS_DTB, out=gz))]
        W.copy_(torch.randn(K, D, device=dev, generator=gen))
        for _ in range(3):
            for n, f in fs: f()
        torch.cuda.synchronize()
        acc = dict((n, 0.0) for n, _ in fs); iters = 10; evs = []
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t0.record()
        for _ in range(iters):
            for n, f in fs:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); f(); b.record(); evs.append((n, a, b))
        t1.record(); torch.cuda.synchronize()
        for n, a, b in evs: acc[n] += a.elapsed_time(b) / iters
        tot = t0.elapsed_time(t1) / iters
        print('BDT B=%d T=%d N=%d engine=%s blk=%s same_idx=%s  %s  total %.4f ms  %.3f G rows/s  %.1f %% of 6536 GB/s' % (
            B, T, N, eng, os.environ.get('VQS_EW_NO_BLK') is None, same, ' '.join('%s %.4f' % (n, acc[n]) for n, _ in fs), tot,
            N / tot / 1e6, 1296 * N / tot / 1e6 / 6536.4 * 100), flush=True)
    ops.vq_set_engine('auto')
    if N >= 1 << 22:                                   # producer depth of the cp.async gather (tiles in flight per thread)
        for lag in ('1', '2', '3'):
            os.environ['VQS_TMA_LAG'] = lag
            for _ in range(3): ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws, idx=idx, stats=st)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws, idx=idx, stats=st)
            b.record(); torch.cuda.synchronize()
            print('BDT B=%d T=%d assign VQS_TMA_LAG=%s %.4f ms same_idx=%s' % (B, T, lag, a.elapsed_time(b) / 10, bool(torch.equal(idx, i_cc))), flush=True)
        os.environ.pop('VQS_TMA_LAG')
        for dbg in ('0', '8'):                         # 8: sector-wise gather (4 batch items x 8 frames per request)
            os.environ['VQS_TMA_DEBUG'] = dbg
            for _ in range(3): ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws, idx=idx, stats=st)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): ops.vq_assign(z, W, LAYOUT_BDT_AS_DTB, ws, idx=idx, stats=st)
            b.record(); torch.cuda.synchronize()
            print('BDT B=%d T=%d assign VQS_TMA_DEBUG=%s %.4f ms' % (B, T, dbg, a.elapsed_time(b) / 10), flush=True)
        os.environ.pop('VQS_TMA_DEBUG')
    del z, g, q, gz
