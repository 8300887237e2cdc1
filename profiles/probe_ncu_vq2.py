"""Driver for `ncu --set full` (round 2): the large-codebook tcgen05 search (K = 4096, N = 2^18 rows) and the four passes of
the K = 44 bottleneck as the training step runs them (search + statistics, gather-only forward, backward with losses), flat
rows and the reference's (B, 64, T) rows, N = 2^22."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND, LAYOUT_BDT_AS_DTB
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(3)
D = 64
# K = 4096
K, N = 4096, 1 << 18
W = torch.randn(K, D, device=dev, generator=g); z = torch.randn(N, D, device=dev, generator=g)
ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
for _ in range(2): ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
torch.cuda.synchronize()
del z, W
# K = 44
K, N = 44, 1 << 22
W = torch.randn(K, D, device=dev, generator=g)
ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
one = torch.ones(1, device=dev); sc = torch.zeros(8, device=dev)
for layout, shape in ((LAYOUT_FLAT_ND, (N, D)), (LAYOUT_BDT_AS_DTB, (N // 128, D, 128))):
    z = torch.randn(*shape, device=dev, generator=g); gq = torch.randn(*shape, device=dev, generator=g)
    q = torch.empty_like(z); gz = torch.empty_like(z)
    for _ in range(2):
        ops.vq_assign(z, W, layout, ws, idx=idx, stats=st)
        ops.vq_gather(idx, W, layout, shape, out=q)
        ops.vq_backward_loss(gq, one, 1e-6, z, idx, W, layout, ws, st[:K], N, 0.25, out=gz, scalars=sc)
    torch.cuda.synchronize()
    del z, gq, q, gz
print('done')
