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


def _nccl_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['RANK'], os.environ['WORLD_SIZE'], os.environ['LOCAL_RANK'] = str(rank), str(world), str(rank)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    from vq_vae_speech_b200 import parallel
    try:
        comm = parallel.init_nccl(dev)
        from test_model_gpu import _build
        from vq_vae_speech_b200.trainer import FusedTrainStep
        from conftest import load_golden as lg
        g = lg('model_ema_k44_h64')                       # num_hiddens 64: every conv GEMM is tcgen05-eligible
        model, cfg = _build(g, dev)
        x = torch.from_numpy(g['x0'])                      # global batch of 4 utterances
        eng = FusedTrainStep(model, x.shape[0] // world, x.shape[1], cfg['learning_rate'], precision='3xtf32',
                             use_graph=True)
        assert eng.world == world and eng.use_graph
        res = []
        for s in range(3):                                 # step 0 eager, step 1 captures the graph (NCCL inside), step 2 replays
            eng.step(comm.shard(torch.from_numpy(g[f'x{s}'])))
            out = eng.losses()
            res.append(dict(vq_loss=out['vq_loss'], recon_loss=out['reconstruction_loss'],
                            idx=eng.encoding_indices().cpu().numpy().reshape(-1)))
            if s == 0:
                grads0 = {k: (v * eng.comm.grad_scale).cpu().numpy() for k, v in eng.gradients().items()}
        assert eng.graph is not None
        vq = model._vq
        checks = [(eng.flat_p, 'parameters'), (vq._embedding.weight.data, 'codebook'), (vq._ema_cluster_size, 'cluster sizes')]
        if eng.nvls is None:                               # (the NVLS exchange shards the AMSGrad moments over the ranks)
            checks += [(eng.flat_m, 'exp_avg'), (eng.flat_vmax, 'max_exp_avg_sq')]
        for t, what in checks:
            comm.assert_replicated(t, what)
        # the optimizer state a checkpoint would hold is complete and identical on both ranks either way (collective)
        from vq_vae_speech_b200.trainer import optimizer_state_dict
        sd = optimizer_state_dict(eng)
        flat = torch.cat([st['exp_avg_sq'].reshape(-1) for _, st in sorted(sd['state'].items())])
        comm.assert_replicated(flat, 'gathered exp_avg_sq')
        assert float(flat.abs().sum()) > 0
        ret[rank] = dict(steps=res, grads0=grads0, cs=vq._ema_cluster_size.cpu().numpy(), nvls=eng.nvls is not None,
                         W=vq._embedding.weight.detach().cpu().numpy())
        dist.barrier()
        torch.cuda.synchronize()
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    # destroy_process_group can block after collectives were captured into a CUDA graph (seen in bench.py): leave directly
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


EXCHANGES = {   # environment of the spawned ranks -> which exchange the step runs (parallel.py, trainer.py, csrc/dp_nvls.cu)
    'nvls': {},                                   # default: own kernels over the NVLS multicast mapping, four launches after backward
    'nvls_one_launch': {'VQS_DP_FUSED': '1'},     # the same exchange as ONE launch with in-kernel barriers
    'nvls_beside_backward': {'VQS_DP_OVERLAP': '1'},   # bucket-wise on a side stream (vqs_dp_amsgrad_range)
    'nccl': {'VQS_DP_NVLS': '0'},                 # NCCL allreduces captured into the graph
}


@pytest.mark.parametrize('exchange', sorted(EXCHANGES))
def test_fused_step_data_parallel_nccl_in_cuda_graph(exchange):
    """The path bench.py times under torch.distributed.run: the gradient / statistics exchange (every variant the library has)
    captured into the step's CUDA graph, tcgen05 3xTF32 GEMMs, two GPUs.  Step 0 (eager) is checked against the per-shard
    oracle emulation of the contract; steps 1-2 (capture + replay) must keep the replicated state bit-identical on
    both ranks and finite (bench.py's parity_check additionally asserts on every multi-GPU run that a graph replay equals the
    eager-NCCL replay bit for bit)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices (run with gpurun --gpus 2)')
    from oracle import model_oracle as mo
    from test_parallel_cpu import _oracle_cfg
    world = 2
    g = load_golden('model_ema_k44_h64')
    # per-shard emulation of step 0 on the 4-utterance batch x0: statistics summed, gradients averaged
    cfg = _oracle_cfg(g)
    p = {k[5:]: v.astype(np.float64) for k, v in g.items() if k.startswith('init.')}
    x = g['x0']
    per = x.shape[0] // world
    shards = [x[r * per:(r + 1) * per] for r in range(world)]
    local = []
    for xs in shards:
        out, c = mo.model_forward(dict(p), xs, cfg)
        local.append((c['vq']['counts'], c['vq']['dw'], c['vq']['N']))
    tot = (sum(l[0] for l in local), sum(l[1] for l in local), sum(l[2] for l in local))
    cfg2 = dict(cfg, stats_allreduce=lambda c_, d_, n_: tot)
    grads, states = [], []
    for xs in shards:
        out, c = mo.model_forward(dict(p), xs, cfg2)
        gr, _ = mo.model_backward(p, c, out, xs.transpose(0, 2, 1), cfg2)
        grads.append(gr)
        states.append((out['vq_loss'], c['vq']['idx']))
    avg = {k: sum(gr[k] for gr in grads) / world for k in grads[0]}
    ret = mp.Manager().dict()
    keys = ('VQS_DP_FUSED', 'VQS_DP_OVERLAP', 'VQS_DP_NVLS')
    saved = dict((k, os.environ.get(k)) for k in keys)
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(EXCHANGES[exchange])               # the spawned ranks inherit the environment
    try:
        mp.spawn(_nccl_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]
    assert ret[0]['nvls'] == (exchange != 'nccl')
    for r in range(world):
        vq_loss, idx = states[r]
        assert np.array_equal(ret[r]['steps'][0]['idx'], idx)
        assert rel_err(ret[r]['steps'][0]['vq_loss'], vq_loss) < 1e-5
        for n, ref in avg.items():
            if '_layers.1.' in n:
                continue
            assert rel_err(ret[r]['grads0'][n], ref) < 2e-5, n
        for s in range(3):
            assert np.isfinite(ret[r]['steps'][s]['vq_loss']) and np.isfinite(ret[r]['steps'][s]['recon_loss'])
    assert np.array_equal(ret[0]['W'], ret[1]['W']) and np.array_equal(ret[0]['cs'], ret[1]['cs'])
