"""world_size-2 gloo tests (CPU) of the data-parallel contract and its host-side plumbing
(vq-vae-speech_b200/parallel.py): per rank the step is the oracle's step on that rank's shard; EMA statistics are
summed over shards before the EMA update; gradients are averaged.  The per-rank compute stands in through the numpy
oracle (the CUDA engine runs the same plumbing on a B200: tests/test_parallel_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_cfg(g):
    return dict(commitment_cost=float(g['cfg_commitment_cost']), decay=float(g['cfg_decay']),
                num_residual_layers=int(g['cfg_num_residual_layers']), learning_rate=float(g['cfg_learning_rate']),
                epsilon=1e-5)


def emulate_dp(g, world):
    """Single-process emulation: oracle per shard, statistics summed, gradients averaged."""
    from oracle import model_oracle as mo
    cfg = _oracle_cfg(g)
    p = {k[5:]: v.astype(np.float64) for k, v in g.items() if k.startswith('init.')}
    x = np.concatenate([g['x0'], g['x1']], 0)          # global batch of 4 utterances
    per = x.shape[0] // world
    shards = [x[r * per:(r + 1) * per] for r in range(world)]
    local = []
    for xs in shards:                                   # pass 1: local statistics (old codebook)
        out, c = mo.model_forward(dict(p), xs, cfg)
        local.append((c['vq']['counts'], c['vq']['dw'], c['vq']['N']))
    tot = (sum(l[0] for l in local), sum(l[1] for l in local), sum(l[2] for l in local))
    cfg2 = dict(cfg, stats_allreduce=lambda c_, d_, n_: tot)
    grads, states = [], []
    for xs in shards:                                   # pass 2: the step with the global statistics
        out, c = mo.model_forward(dict(p), xs, cfg2)
        gr, _ = mo.model_backward(p, c, out, xs.transpose(0, 2, 1), cfg2)
        grads.append(gr)
        states.append((c['vq']['cluster_size'], c['vq']['ema_w'], c['vq']['W_used'], out['vq_loss'], c['vq']['idx']))
    avg = {k: sum(gr[k] for gr in grads) / world for k in grads[0]}
    return avg, states, tot


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import model_oracle as mo
        from vq_vae_speech_b200.parallel import DataParallelComm
        from conftest import load_golden as lg
        g = lg('model_ema_k44')
        comm = DataParallelComm()
        assert comm.world == world and comm.rank == rank
        cfg = _oracle_cfg(g)
        p = {k[5:]: v.astype(np.float64) for k, v in g.items() if k.startswith('init.')}
        x = torch.from_numpy(np.concatenate([g['x0'], g['x1']], 0))
        xs = comm.shard(x).numpy()

        def stats_allreduce(counts, dw, n):
            packed = torch.from_numpy(np.concatenate([counts, dw.reshape(-1)]))
            comm.allreduce_stats(packed)
            K = counts.shape[0]
            packed = packed.numpy()
            return packed[:K], packed[K:].reshape(dw.shape), comm.total_rows(n)

        cfg['stats_allreduce'] = stats_allreduce
        out, c = mo.model_forward(dict(p), xs, cfg)
        grads, _ = mo.model_backward(p, c, out, xs.transpose(0, 2, 1), cfg)
        names = sorted(grads)
        flat = torch.from_numpy(np.concatenate([grads[n].reshape(-1) for n in names]))
        split = flat.numel() // 3
        comm.start_bucket(flat, split, flat.numel())       # "decoder" bucket first
        comm.start_bucket(flat, 0, split)
        comm.wait_buckets()
        flat = flat * comm.grad_scale
        comm.assert_replicated(flat, 'averaged gradients')
        comm.assert_replicated(torch.from_numpy(c['vq']['W_used']), 'codebook')
        ret[rank] = dict(flat=flat.numpy(), names=names, shapes=[grads[n].shape for n in names],
                         cs=c['vq']['cluster_size'], ema_w=c['vq']['ema_w'], W=c['vq']['W_used'],
                         vq_loss=float(out['vq_loss']), idx=c['vq']['idx'])
    finally:
        dist.destroy_process_group()


def test_dp_contract_world2_gloo():
    world = 2
    g = load_golden('model_ema_k44')
    avg, states, tot = emulate_dp(g, world)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert sorted(ret.keys()) == [0, 1]
    for r in range(world):
        res = ret[r]
        off = 0
        for n, shp in zip(res['names'], res['shapes']):
            sz = int(np.prod(shp))
            got = res['flat'][off:off + sz].reshape(shp)
            off += sz
            assert np.allclose(got, avg[n], rtol=1e-12, atol=1e-15), n
        cs, ew, W, vq_loss, idx = states[r]
        assert np.allclose(res['cs'], cs, rtol=1e-12) and np.allclose(res['ema_w'], ew, rtol=1e-12)
        assert np.allclose(res['W'], W, rtol=1e-12)
        assert abs(res['vq_loss'] - vq_loss) <= 1e-12 * abs(vq_loss)
        assert np.array_equal(res['idx'], idx)
    # EMA state identical on both ranks
    assert np.array_equal(ret[0]['W'], ret[1]['W']) and np.array_equal(ret[0]['cs'], ret[1]['cs'])
    # and the shards really saw different data (different indices) while sharing the statistics
    assert not np.array_equal(ret[0]['idx'], ret[1]['idx'])
    assert float(tot[0].sum()) == 2 * len(ret[0]['idx'])


def test_shard_rejects_indivisible_batch():
    from vq_vae_speech_b200.parallel import DataParallelComm
    comm = DataParallelComm()
    assert comm.world == 1 and comm.grad_scale == 1.0
    x = torch.zeros(5, 3)
    assert comm.shard(x).shape[0] == 5
    comm.world = 2
    with pytest.raises(ValueError):
        comm.shard(x)


def test_optimizer_shards_cover_every_range_exactly_once():
    """parallel.shard_bounds (the host mirror of the split inside vqs_dp_amsgrad_step / vqs_dp_amsgrad_range): the ranks' slices
    of a range are disjoint, 16-byte aligned, in rank order and cover it completely -- also when trailing ranks get nothing."""
    from vq_vae_speech_b200.parallel import shard_bounds
    for world in (1, 2, 3, 4, 8):
        for lo, hi in ((0, 16_400_384), (64, 64), (128, 132), (256, 256 + 4 * 5), (4096, 4096 + 4 * 1001), (0, 4 * world)):
            cur = lo
            for rank in range(world):
                a, b = shard_bounds(lo, hi, world, rank)
                assert a % 4 == 0 and b % 4 == 0 and a <= b
                if b > a:                                   # (an empty slice may start beyond the range: never touched)
                    assert a == cur and b <= hi
                    cur = b
            assert cur == hi


def test_operand_image_layouts_pad_channels_and_rows():
    """functional.gemm_weight_layout: operand images are ceil(M / 128) x (k * ceil(Cred / 32)) blocks of 8192 floats; the 39-channel
    layers of the 768-wide model take padded images (tcgen05), narrow or small layers stay on the exact CUDA-core engine."""
    from vq_vae_speech_b200 import functional as F
    assert F.gemm_weight_layout((768, 768, 3), 'conv_fwd', '3xtf32') == (2, 3, 6 * 72 * 8192)
    assert F.gemm_weight_layout((768, 39, 3), 'conv_fwd', '3xtf32') == (2, 3, 6 * 6 * 8192)        # Cred 39 -> 64
    assert F.gemm_weight_layout((768, 39, 2), 'convT_dgrad', '3xtf32') == (2, 3, 6 * 4 * 8192)     # M = 768, Cred = 39
    assert F.gemm_weight_layout((768, 39, 2), 'convT_fwd', '3xtf32') == (2, 4, 1 * 48 * 8192)      # M = 39 (one tile), Cred = 768
    assert F.gemm_weight_layout((48, 39, 3), 'conv_fwd', '3xtf32')[0] == 0                          # M < 128: CUDA cores
    assert F.gemm_weight_layout((768, 7, 3), 'conv_fwd', '3xtf32')[0] == 0                          # one partial k-block: CUDA cores
    assert F.gemm_weight_layout((768, 768, 3), 'conv_fwd', 'fp32')[0] == 0                          # exact engine: canonical weights
    assert F.conv_tc_eligible(768, 39, '3xtf32') and not F.conv_tc_eligible(64, 39, '3xtf32')
    assert F.conv_tc_eligible(39, 768, '3xtf32') and not F.conv_tc_eligible(768, 768, 'fp32')
