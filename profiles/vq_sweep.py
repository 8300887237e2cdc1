"""Codebook-size sweep of the VQ bottleneck (BASELINE.json configs[3]): fwd (assign + EMA + quantise) + bwd at VCTK shape
(B=256, T_q=96 -> N=24576 rows) and at N=2^20 rows, K in {44, 512, 4096}, D=64, in the reference's (B,D,T) row layout and
on ready-made flat rows, with both search engines where available.  Prints one JSON line per case."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
D = 64


def case(B, T, K, lname, layout):
    N = B * T
    g = torch.Generator(device=dev).manual_seed(K)
    shape = (B, D, T) if lname == 'bdt' else (N, D)
    W = torch.randn(K, D, device=dev, generator=g); z = torch.randn(*shape, device=dev, generator=g)
    gq = torch.randn(*shape, device=dev, generator=g)
    cs = torch.zeros(K, device=dev); ew = torch.randn(K, D, device=dev, generator=g)
    ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev)
    st = torch.empty(K * (D + 1), device=dev); q = torch.empty_like(z); gz = torch.empty_like(z)
    sc = torch.zeros(8, device=dev); one = torch.ones(1, device=dev)
    for eng in ('cuda_core', 'tensor_core'):
        ops.vq_set_engine(eng)

        def once():
            ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
            ops.vq_ema_update(cs, ew, W, st, 0.99, 1e-5)
            ops.vq_quantize(z, idx, W, layout, ws, st[:K], N, 0.25, out=q, scalars=sc)
            ops.vq_backward(gq, one, 2 * 0.25 / (N * D), z, idx, W, layout, out=gz)
        for _ in range(3): once()
        torch.cuda.synchronize()
        it = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): once()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(it): ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
        a1.record(); torch.cuda.synchronize()
        ams = a0.elapsed_time(a1) / it
        print(json.dumps({'N': N, 'K': K, 'layout': lname, 'engine': eng, 'ms_fwd_bwd_ema': round(ms, 4),
                          'rows_per_s': N / (ms * 1e-3), 'assign_ms': round(ams, 4),
                          'assign_tflops_alg': 2.0 * N * D * K / (ams * 1e-3) / 1e12}), flush=True)


for (B, T) in ((256, 96), (8192, 128)):
    for lname, layout in (('bdt', LAYOUT_BDT_AS_DTB), ('flat', LAYOUT_FLAT_ND)):
        for K in (44, 512, 4096):
            case(B, T, K, lname, layout)
ops.vq_set_engine('auto')
