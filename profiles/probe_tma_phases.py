"""Phase probe of the streaming VQ search (flat rows, N = 2^22, K = 44): VQS_TMA_DEBUG bit 0 skips the statistics pass,
bit 1 the settlement, bit 2 shortens the scan (results are wrong with any bit set: timing only)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND, LAYOUT_BDT_AS_DTB
dev = torch.device('cuda:0')
K, D, N = 44, 64, 1 << 22
gen = torch.Generator(device=dev).manual_seed(7)
W = torch.randn(K, D, device=dev, generator=gen)
ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
for lname, layout, shape in (('flat', LAYOUT_FLAT_ND, (N, D)), ('bdt', LAYOUT_BDT_AS_DTB, (32768, D, 128))):
    z = torch.randn(*shape, device=dev, generator=gen)
    for dbg in ('0', '1', '2', '4', '3', '5', '7'):
        os.environ['VQS_TMA_DEBUG'] = dbg
        for _ in range(3): ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
        b.record(); torch.cuda.synchronize()
        print('PHASES %s VQS_TMA_DEBUG=%s %.4f ms' % (lname, dbg, a.elapsed_time(b) / 20), flush=True)
    os.environ.pop('VQS_TMA_DEBUG')
    del z
