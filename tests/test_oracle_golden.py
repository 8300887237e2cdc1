"""Pins the CPU oracle (oracle/*.py) against vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rel_err
from oracle import model_oracle as mo
from oracle import vq_oracle as vqo

TOL = 1e-5          # north_star: losses, quantized outputs and gradients within 1e-5 relative (fp32)

VQ_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'vq_*.npz')))
MODEL_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'model_*.npz')))


def test_golden_present():
    assert len(VQ_CASES) >= 10 and len(MODEL_CASES) >= 5


@pytest.mark.parametrize('case', VQ_CASES)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_vq_oracle_matches_reference(case, dtype):
    g = load_golden(case)
    K, D, B, T = int(g['K']), int(g['D']), int(g['B']), int(g['T'])
    is_ema = bool(g['ema'])
    beta = float(g['commitment_cost'])
    W = g['W0']
    state = dict(cluster_size=g['cs0'], ema_w=g['ema_w0'], decay=float(g['decay']), eps=float(g['epsilon'])) if is_ema else None
    n_near = 0
    for s in range(int(g['steps'])):
        z = g[f'z{s}']
        f = vqo.vq_forward(z, W, beta, ema=state, training=True, dtype=dtype)
        ref_idx = g[f'idx{s}'].reshape(-1)
        mism = f['idx'] != ref_idx
        assert not np.any(mism & ~f['near_tie']), 'index mismatch outside near-tie rows'
        n_near += int(f['near_tie'].sum())
        ste_scale = max(np.abs(z).max(), np.abs(g[f'quantized{s}']).max())
        assert rel_err(f['quantized'], g[f'quantized{s}'], ste_scale) < TOL
        if dtype == np.float32 and not is_ema and not mism.any():   # fixed codebook + same IEEE ops: bit-identical
            assert np.array_equal(f['quantized'], g[f'quantized{s}'])
        assert rel_err(f['vq_loss'], g[f'vq_loss{s}']) < TOL
        assert rel_err(f['perplexity'], g[f'perplexity{s}']) < TOL
        assert rel_err(f['q_rows'], g[f'concat{s}']) < TOL
        gz, gE = vqo.vq_backward(z, f, beta, g[f'g{s}'], float(g['g_loss']), ema=is_ema, dtype=dtype)
        assert rel_err(gz, g[f'grad_z{s}']) < TOL
        if s == 0:
            assert np.array_equal(f['encodings'], g['encodings0'])
            assert rel_err(f['distances'], g['distances0']) < TOL
        if is_ema:
            assert rel_err(f['cluster_size'], g[f'cs{s + 1}']) < TOL
            assert rel_err(f['ema_w'], g[f'ema_w{s + 1}']) < TOL
            assert rel_err(f['W_used'], g[f'W{s + 1}']) < TOL
            state = dict(state, cluster_size=g[f'cs{s + 1}'], ema_w=g[f'ema_w{s + 1}'])
        else:
            assert rel_err(gE, g[f'grad_E{s}']) < TOL
        W = g[f'W{s + 1}']
    # eval mode: no EMA update, codebook as is
    f = vqo.vq_forward(g['z_eval'], W, beta, ema=state, training=False, dtype=dtype)
    assert np.array_equal(f['idx'], g['eval_idx'].reshape(-1))
    assert rel_err(f['quantized'], g['eval_quantized'], max(np.abs(g['z_eval']).max(), np.abs(g['eval_quantized']).max())) < TOL
    assert rel_err(f['vq_loss'], g['eval_vq_loss']) < TOL
    assert rel_err(f['perplexity'], g['eval_perplexity']) < TOL


def test_vq_true_ties_resolve_to_lowest_index():
    g = load_golden('vq_ema_k44_d64_b2_t24_dup')
    idx = g['idx0'].reshape(-1)
    K = int(g['K'])
    assert not np.any(idx == K // 2) and not np.any(idx == K - 1)   # duplicates of code 1 never win
    f = vqo.vq_forward(g['z0'], g['W0'], 0.25, ema=None)
    assert np.array_equal(f['idx'], idx)


def test_row_layout_is_dtb_order():
    z = np.arange(2 * 4 * 3, dtype=np.float32).reshape(2, 4, 3)      # B=2, D=4, T=3
    rows = vqo.rows_from_bdt(z)
    assert rows.shape == (6, 4)
    # flat element f = d*T*B + t*B + b  <-  z[b, d, t]   
    flat = rows.reshape(-1)
    for b in range(2):
        for d in range(4):
            for t in range(3):
                assert flat[d * 6 + t * 2 + b] == z[b, d, t]
    assert np.array_equal(vqo.bdt_from_rows(rows, 2, 4, 3), z)


def _params(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize('case', MODEL_CASES)
def test_model_oracle_matches_reference(case):
    g = load_golden(case)
    cfg = dict(commitment_cost=float(g['cfg_commitment_cost']), decay=float(g['cfg_decay']),
               num_residual_layers=int(g['cfg_num_residual_layers']), learning_rate=float(g['cfg_learning_rate']),
               epsilon=1e-5)
    p = {k: v.astype(np.float64) for k, v in _params(g, 'init.').items()}
    opt = dict(step=0, m={}, v={}, vmax={})
    use_jitter = bool(g['cfg_use_jitter'])
    if use_jitter:   # the plan is reproducible from the seed with the reference's RNG call order
        np.random.seed(int(g['seed']))
    for s in range(int(g['steps'])):
        src = None
        if use_jitter:
            src = mo.jitter_plan(int(g['T']) // 2 + 1, float(g['cfg_jitter_probability']))
            assert np.array_equal(src, g[f'jitter_src{s}'])
        if 'speakers' in g and int(g['speakers']):          # the step's speaker features (drawn by the reference's forward)
            cfg['speaker_features'] = g[f'speaker_features{s}']
        r = mo.train_step(p, opt, g[f'x{s}'], cfg, jitter_src=src)
        ref_idx = g[f'idx{s}'].reshape(-1)
        assert not np.any((r['encoding_indices'].reshape(-1) != ref_idx) & ~r['near_tie'])
        assert rel_err(r['reconstructed_x'], g[f'recon{s}']) < TOL
        assert rel_err(r['vq_loss'], g[f'vq_loss{s}']) < TOL
        assert rel_err(r['reconstruction_loss'], g[f'recon_loss{s}']) < TOL
        assert rel_err(r['perplexity'], g[f'perplexity{s}']) < TOL
        if s == 0:
            for n, ref in _params(g, 'grad0.').items():
                if '_layers.1.' in n:
                    continue
                assert rel_err(r['grads'][n], ref) < 2e-5, n
    # Adam divides by sqrt(v): where |grad| is at fp32-noise level the update direction itself is noise, so
    # post-step parameters are compared against the size of the cumulative update (lr * steps), not 1e-5.
    budget = 0.05 * cfg['learning_rate'] * int(g['steps'])
    for n, ref in _params(g, 'final.').items():
        if n.startswith('_vq.') and cfg['decay'] > 0:
            assert rel_err(p[n], ref) < TOL, n          # EMA state is not touched by Adam
        else:
            assert float(np.max(np.abs(p[n] - ref))) < budget, n


@pytest.mark.parametrize('case', MODEL_CASES)
def test_torch_port_matches_reference(case):
    """oracle/torch_port.py (what bench.py times as the CPU baseline) reproduces the reference's training steps."""
    import torch
    from oracle.torch_port import PortTrainer
    g = load_golden(case)
    cfg = dict(output_features_filters=13, augment_output_features=True, input_features_filters=13,
               augment_input_features=True, use_jitter=bool(g['cfg_use_jitter']),
               use_kaiming_normal=bool(g['cfg_use_kaiming_normal']),
               use_speaker_conditioning=bool(g['cfg_use_speaker_conditioning']))
    for k in ('num_hiddens', 'num_residual_layers', 'embedding_dim', 'num_embeddings', 'residual_channels'):
        cfg[k] = int(g['cfg_' + k])
    for k in ('decay', 'commitment_cost', 'jitter_probability', 'learning_rate'):
        cfg[k] = float(g['cfg_' + k])
    torch.set_num_threads(1)
    tr = PortTrainer(cfg, seed=int(g['seed']))
    tr.model.load_reference_state({k[5:]: v for k, v in g.items() if k.startswith('init.')})
    np.random.seed(int(g['seed']))
    for s in range(int(g['steps'])):
        sf = torch.from_numpy(g[f'speaker_features{s}']) if cfg['use_speaker_conditioning'] else None
        r = tr.step(torch.from_numpy(g[f'x{s}']), speaker_features=sf)
        if cfg['use_jitter']:
            assert np.array_equal(r['jitter_src'], g[f'jitter_src{s}'])
        assert np.array_equal(r['encoding_indices'].numpy().reshape(-1), g[f'idx{s}'].reshape(-1))
        assert rel_err(r['reconstructed_x'].numpy(), g[f'recon{s}']) < TOL
        assert rel_err(r['vq_loss'], g[f'vq_loss{s}']) < TOL
        assert rel_err(r['reconstruction_loss'], g[f'recon_loss{s}']) < TOL
        assert rel_err(r['perplexity'], g[f'perplexity{s}']) < TOL


@pytest.mark.parametrize('case', ['evaltables_noema_k44_d64_b2_t24', 'evaltables_noema_k10_d2_b3_t17'])
def test_oracle_eval_distance_tables(case):
    """Eval-only pairwise distance tables (vector_quantizer.py:108-127) against the reference's own output."""
    g = load_golden(case)
    enc, emb, fve = vqo.eval_distance_tables(g['z'], g['W'])
    assert enc.shape == g['encoding_distances'].shape and emb.shape == g['embedding_distances'].shape
    assert fve.shape == g['frames_vs_embedding_distances'].shape
    assert rel_err(enc, g['encoding_distances']) < TOL
    assert rel_err(emb, g['embedding_distances']) < TOL
    assert rel_err(fve, g['frames_vs_embedding_distances']) < TOL
