"""Driver for `ncu --set full` of the large-codebook tcgen05 search: K = 4096, D = 64, N = 2^20 flat rows, 3 launches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(3)
K, D, N = 4096, 64, 1 << 20
W = torch.randn(K, D, device=dev, generator=g); z = torch.randn(N, D, device=dev, generator=g)
ws = ops.vq_workspace(K, D, dev); idx = torch.empty(N, dtype=torch.int64, device=dev); st = torch.empty(K * (D + 1), device=dev)
ops.vq_set_engine('tensor_core')
for _ in range(3):
    ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
torch.cuda.synchronize()
print('done', int(idx.sum()))
