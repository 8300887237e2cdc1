"""Data-parallel parity of the REAL engine: two ranks (two processes, gloo process group over CUDA tensors so that the
test also runs on a single-GPU box) each run FusedTrainStep on their shard; results must equal the oracle emulation
of the contract (tests/test_parallel_cpu.py::emulate_dp).  The multi-GPU NCCL path is the same code with backend
'nccl' (bench.py under torch.distributed.run)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, rel_err
from test_parallel_cpu import _free_port, emulate_dp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from test_model_gpu import _build
        from vq_vae_speech_b200.trainer import FusedTrainStep
        from conftest import load_golden as lg
        g = lg('model_ema_k44')
        dev = torch.device('cuda', rank % ndev)
        model, cfg = _build(g, dev)
        x = torch.from_numpy(np.concatenate([g['x0'], g['x1']], 0))
        eng = FusedTrainStep(model, x.shape[0] // world, x.shape[1], cfg['learning_rate'], precision='fp32')
        assert eng.world == world
        eng.use_graph = False          # gloo over CUDA tensors is not graph-capturable; NCCL (bench.py) is
        eng.step(eng.comm.shard(x))
        out = eng.losses()
        grads = {k: (v * eng.comm.grad_scale).cpu().numpy() for k, v in eng.gradients().items()}
        vq = model._vq
        ret[rank] = dict(grads=grads, cs=vq._ema_cluster_size.cpu().numpy(), ema_w=vq._ema_w.detach().cpu().numpy(),
                         W=vq._embedding.weight.detach().cpu().numpy(), vq_loss=out['vq_loss'],
                         idx=eng.encoding_indices().cpu().numpy().reshape(-1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('fine_buckets', [False, True])
def test_fused_step_data_parallel_two_ranks(fine_buckets):
    """fine_buckets: VQS_DP_FINE=1 sends the gradients of conv_3 / conv_2 / conv_1 as separate trailing buckets."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    world = 2
    g = load_golden('model_ema_k44')
    avg, states, _ = emulate_dp(g, world)
    ret = mp.Manager().dict()
    saved = os.environ.get('VQS_DP_FINE')
    os.environ['VQS_DP_FINE'] = '1' if fine_buckets else '0'      # the spawned workers inherit the environment
    try:
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    finally:
        os.environ.pop('VQS_DP_FINE', None)
        if saved is not None:
            os.environ['VQS_DP_FINE'] = saved
    for r in range(world):
        res = ret[r]
        cs, ew, W, vq_loss, idx = states[r]
        assert np.array_equal(res['idx'], idx)
        assert rel_err(res['cs'], cs) < 1e-5 and rel_err(res['ema_w'], ew) < 1e-5 and rel_err(res['W'], W) < 1e-5
        assert rel_err(res['vq_loss'], vq_loss) < 1e-5
        for n, ref in avg.items():
            if '_layers.1.' in n:
                continue
            assert rel_err(res['grads'][n], ref) < 2e-5, n
    assert np.array_equal(ret[0]['W'], ret[1]['W'])          # replicated codebook stays bit-identical
    assert np.array_equal(ret[0]['cs'], ret[1]['cs'])
    for n in ret[0]['grads']:
        assert np.array_equal(ret[0]['grads'][n], ret[1]['grads'][n]), n
