"""K = 512 / 4096 search: indices against the CUDA-core engine (exact fp32) and CUDA-event timing of the tcgen05 search."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND  # noqa: E402

dev = torch.device('cuda:0')
D = 64
for K, N in ((512, 1 << 18), (4096, 1 << 18), (4096, 1 << 20), (1000, 5000)):
    g = torch.Generator(device=dev).manual_seed(K)
    W = torch.randn(K, D, device=dev, generator=g)
    z = torch.randn(N, D, device=dev, generator=g)
    ws = ops.vq_workspace(K, D, dev)
    out = {}
    for eng in ('cuda_core', 'tensor_core'):
        ops.vq_set_engine(eng)
        idx = torch.empty(N, dtype=torch.int64, device=dev)
        st = torch.empty(K * (D + 1), device=dev)
        ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
        e1.record()
        torch.cuda.synchronize()
        out[eng] = (idx.clone(), st[:K].clone(), e0.elapsed_time(e1) / 5)
    same = torch.equal(out['cuda_core'][0], out['tensor_core'][0])
    ndiff = int((out['cuda_core'][0] != out['tensor_core'][0]).sum())
    ms = out['tensor_core'][2]
    print('K=%d N=%d: indices identical %s (%d differ), counts identical %s | tensor-core search %.4f ms = %.1f TFLOP/s algorithmic | '
          'CUDA cores %.4f ms' % (K, N, same, ndiff, torch.equal(out['cuda_core'][1], out['tensor_core'][1]), ms,
                                  2.0 * N * D * K / ms / 1e9, out['cuda_core'][2]), flush=True)
ops.vq_set_engine('auto')
