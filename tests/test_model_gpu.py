"""GPU parity of the full encoder + VQ + decoder training step against golden vectors produced by the unmodified
reference (tests/golden/model_*.npz: `ConvolutionalVQVAE` + `Adam(amsgrad=True)` for several consecutive steps), for
both product paths: the drop-in nn.Modules under autograd, and the fused / CUDA-graph training step."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
MODEL_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, 'model_*.npz')))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def _cfg(g):
    cfg = dict(output_features_filters=13, augment_output_features=True, output_features_dim=47, verbose=False,
               input_features_dim=47, use_kaiming_normal=False, input_features_type='mfcc', input_features_filters=13,
               augment_input_features=True, sampling_rate=16000, use_speaker_conditioning=False,
               record_codebook_stats=False)
    for k in ('num_hiddens', 'num_residual_layers', 'embedding_dim', 'num_embeddings', 'residual_channels'):
        cfg[k] = int(g['cfg_' + k])
    for k in ('decay', 'commitment_cost', 'jitter_probability', 'learning_rate'):
        cfg[k] = float(g['cfg_' + k])
    cfg['use_jitter'] = bool(g['cfg_use_jitter'])
    cfg['use_kaiming_normal'] = bool(g['cfg_use_kaiming_normal'])
    cfg['use_speaker_conditioning'] = bool(g['cfg_use_speaker_conditioning'])
    return cfg


def _build(g, dev):
    from vq_vae_speech_b200.convolutional_vq_vae import ConvolutionalVQVAE
    cfg = _cfg(g)
    model = ConvolutionalVQVAE(cfg, dev)
    sd = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('init.')}
    model.load_state_dict(sd)
    return model.to(dev).train(), cfg


def _check_final(model, g, cfg, steps):
    # Adam divides by sqrt(v): where |grad| is at fp32-noise level the direction itself is noise, so post-step parameters
    # are compared against the size of the cumulative update (lr * steps), like tests/test_oracle_golden.py
    budget = 0.05 * cfg['learning_rate'] * steps
    sd = model.state_dict()
    for k, ref in g.items():
        if not k.startswith('final.'):
            continue
        n = k[6:]
        got = sd[n].detach().cpu().numpy()
        if n.startswith('_vq.') and cfg['decay'] > 0:
            assert rel_err(got, ref) < TOL, n
        else:
            assert float(np.max(np.abs(got - ref))) < budget, n


@pytest.mark.parametrize('case', MODEL_CASES)
def test_module_path_matches_reference(case):
    """Drop-in nn.Modules + autograd + our fused AMSGrad == reference modules + torch Adam(amsgrad=True)."""
    dev = _dev()
    from vq_vae_speech_b200 import functional as F, ops
    g = load_golden(case)
    model, cfg = _build(g, dev)
    params = [p for p in model.parameters()]
    state = {}
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    if cfg['use_jitter']:
        np.random.seed(int(g['seed']))
    steps = int(g['steps'])
    for s in range(steps):
        x = torch.from_numpy(g[f'x{s}']).to(dev)
        for p in params:
            p.grad = None
        speaker_dic = speaker_id = None
        if cfg['use_speaker_conditioning']:
            # the decoder draws a fresh random speaker embedding from the host RNG inside forward, like the reference
            # (global_conditioning.py:34); make_golden.py pinned the RNG to 1000 + s right before the call
            speaker_dic = {'p%03d' % i: i for i in range(int(g['speakers']))}
            speaker_id = torch.from_numpy(g[f'speaker_id{s}'])
            torch.manual_seed(1000 + s)
        recon, vq_loss, losses, perplexity, idx, _ = model(x, speaker_dic, speaker_id)
        recon_loss = F.mse_loss(recon, x.permute(0, 2, 1))
        loss = vq_loss + recon_loss
        loss.backward()
        if cfg['use_jitter']:
            assert np.array_equal(model._decoder._jitter.last_plan, g[f'jitter_src{s}'])
        assert np.array_equal(idx.cpu().numpy().reshape(-1), g[f'idx{s}'].reshape(-1))
        assert tuple(recon.shape) == g[f'recon{s}'].shape
        assert rel_err(recon.detach().cpu().numpy(), g[f'recon{s}']) < TOL
        assert rel_err(vq_loss.item(), g[f'vq_loss{s}']) < TOL
        assert rel_err(recon_loss.item(), g[f'recon_loss{s}']) < TOL
        assert rel_err(perplexity.item(), g[f'perplexity{s}']) < TOL
        if s == 0:
            for n, p in model.named_parameters():
                key = 'grad0.' + n
                if key in g:
                    assert p.grad is not None, n
                    assert rel_err(p.grad.cpu().numpy(), g[key]) < 2e-5, n
                else:
                    assert p.grad is None, n
        first = True
        for p in params:            # Adam(amsgrad=True): parameters without a gradient are skipped (EMA codebook)
            if p.grad is None:
                continue
            if p not in state:
                state[p] = [torch.zeros_like(p) for _ in range(3)]
            m, v, vm = state[p]
            ops.amsgrad_step(p.data, p.grad.contiguous(), m, v, vm, step, cfg['learning_rate'], inc_step=first)
            first = False
    _check_final(model, g, cfg, steps)


@pytest.mark.parametrize('case', MODEL_CASES)
@pytest.mark.parametrize('use_graph', [False, True])
@pytest.mark.parametrize('precision', ['fp32', '3xtf32'])
def test_fused_step_matches_reference(case, use_graph, precision):
    """The hand-scheduled (and CUDA-graph captured) training step == the reference trainer iteration, with the conv GEMMs
    on CUDA cores (exact fp32) and on tcgen05 tensor cores (3xTF32 split): same 1e-5 bar for both."""
    dev = _dev()
    from vq_vae_speech_b200.trainer import FusedTrainStep
    g = load_golden(case)
    model, cfg = _build(g, dev)
    B, T = int(g['B']), int(g['T'])
    eng = FusedTrainStep(model, B, T, cfg['learning_rate'], use_graph=use_graph, precision=precision)
    if cfg['use_jitter']:
        np.random.seed(int(g['seed']))
    steps = int(g['steps'])
    for s in range(steps):
        if cfg['use_speaker_conditioning']:
            # the reference draws a fresh random speaker embedding on the host RNG inside forward (global_conditioning.py:34);
            # make_golden.py pinned the RNG to 1000 + s right before the call, the fused step draws the same one
            speaker_dic = {'p%03d' % i: i for i in range(int(g['speakers']))}
            torch.manual_seed(1000 + s)
            eng.step(torch.from_numpy(g[f'x{s}']), speaker_dic=speaker_dic, speaker_id=torch.from_numpy(g[f'speaker_id{s}']))
            assert np.array_equal(eng.buf['spk'].cpu().numpy(), g[f'speaker_features{s}'])
        else:
            eng.step(torch.from_numpy(g[f'x{s}']))
        out = eng.losses()
        if cfg['use_jitter']:
            assert np.array_equal(eng.last_plan, g[f'jitter_src{s}'])
        assert np.array_equal(eng.encoding_indices().cpu().numpy().reshape(-1), g[f'idx{s}'].reshape(-1))
        assert rel_err(eng.buf['recon'].cpu().numpy(), g[f'recon{s}']) < TOL
        assert rel_err(out['vq_loss'], g[f'vq_loss{s}']) < TOL
        assert rel_err(out['reconstruction_loss'], g[f'recon_loss{s}']) < TOL
        assert rel_err(out['perplexity'], g[f'perplexity{s}']) < TOL
        if s == 0:
            grads = eng.gradients()
            for k, ref in g.items():
                if k.startswith('grad0.') and '_layers.1.' not in k:
                    assert rel_err(grads[k[6:]].cpu().numpy(), ref) < 2e-5, k
    assert eng.graph is not None or not use_graph
    _check_final(model, g, cfg, steps)


def test_state_dict_keys_and_init_match_reference():
    """Same constructor order -> same initial weights under the same torch seed; same state_dict keys."""
    _dev()
    from vq_vae_speech_b200.convolutional_vq_vae import ConvolutionalVQVAE
    for case in MODEL_CASES:
        g = load_golden(case)
        cfg = _cfg(g)
        torch.manual_seed(int(g['seed']))
        np.random.seed(int(g['seed']))
        model = ConvolutionalVQVAE(cfg, 'cpu')
        sd = model.state_dict()
        keys = sorted(k[5:] for k in g if k.startswith('init.'))
        assert sorted(sd.keys()) == keys
        for k in keys:
            assert np.array_equal(sd[k].numpy(), g['init.' + k]), k


@pytest.mark.parametrize('case', ['model_ema_k44', 'model_noema_k44'])
def test_fused_step_tf32_stated_tolerance(case):
    """Single-pass TF32 (cuDNN's default for the reference's convs on a GPU): 10-bit mantissa operands.  Stated tolerance:
    reconstruction and losses within 1e-2 relative of the fp32 reference over the golden steps; indices are compared only
    as an agreement rate (TF32 perturbs z by ~1e-3 relative, so rows with a small top-2 gap legitimately flip)."""
    dev = _dev()
    from vq_vae_speech_b200.trainer import FusedTrainStep
    g = load_golden(case)
    model, cfg = _build(g, dev)
    eng = FusedTrainStep(model, int(g['B']), int(g['T']), cfg['learning_rate'], use_graph=False, precision='tf32')
    eng.step(torch.from_numpy(g['x0']))
    out = eng.losses()
    assert rel_err(eng.buf['recon'].cpu().numpy(), g['recon0']) < 1e-2
    assert rel_err(out['reconstruction_loss'], g['recon_loss0']) < 1e-2
    assert rel_err(out['vq_loss'], g['vq_loss0']) < 5e-2
    agree = np.mean(eng.encoding_indices().cpu().numpy().reshape(-1) == g['idx0'].reshape(-1))
    assert agree >= 0.9, agree


def test_checkpoint_roundtrip_and_torch_adam_compat(tmp_path):
    """the checkpoint dict has the reference's keys (convolutional_trainer.py:76-86); its 'optimizer' entry
    loads into torch.optim.Adam(amsgrad=True) (how the reference resumes, pipeline_factory.py:118-120); and a step after
    save -> load reproduces the step of the uninterrupted run bit for bit."""
    dev = _dev()
    from vq_vae_speech_b200 import trainer as tr
    g = load_golden('model_ema_k44')
    model, cfg = _build(g, dev)
    eng = tr.FusedTrainStep(model, int(g['B']), int(g['T']), cfg['learning_rate'], use_graph=False)
    eng.step(torch.from_numpy(g['x0']))
    path = str(tmp_path / 'exp_1_checkpoint.pth')
    tr.save_checkpoint(eng, path, 'exp', 0)
    ck = torch.load(path, weights_only=False)
    assert sorted(ck.keys()) == sorted(['experiment_name', 'epoch', 'model', 'optimizer', 'train_res_recon_error',
                                        'train_res_perplexity'])
    assert ck['epoch'] == 1
    ref_opt = torch.optim.Adam(model.parameters(), lr=cfg['learning_rate'], amsgrad=True)
    ref_opt.load_state_dict(ck['optimizer'])                      # torch accepts the format
    assert len(ref_opt.state_dict()['state']) == len(eng.grads)
    eng.step(torch.from_numpy(g['x1']))                            # uninterrupted run: step 2
    want = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model2, _ = _build(g, dev)
    eng2 = tr.FusedTrainStep(model2, int(g['B']), int(g['T']), cfg['learning_rate'], use_graph=False)
    tr.load_checkpoint(eng2, path)
    eng2.step(torch.from_numpy(g['x1']))
    for k, v in model2.state_dict().items():
        assert torch.equal(v, want[k]), k


def test_checkpoint_roundtrip_weight_normalised_model(tmp_path):
    """use_kaiming_normal (weight_g / weight_v parameters, effective-weight scratch buffers in the step): the optimizer
    state dict covers exactly the named parameters, torch Adam accepts it, and save -> load -> step is bit-exact."""
    dev = _dev()
    from vq_vae_speech_b200 import trainer as tr
    g = load_golden('model_ema_k29_kaiming')
    model, cfg = _build(g, dev)
    eng = tr.FusedTrainStep(model, int(g['B']), int(g['T']), cfg['learning_rate'], use_graph=False)
    eng.step(torch.from_numpy(g['x0']))
    path = str(tmp_path / 'wn_1_checkpoint.pth')
    tr.save_checkpoint(eng, path, 'wn', 0)
    ck = torch.load(path, weights_only=False)
    names = [n for n, _ in model.named_parameters()]
    assert any(n.endswith('weight_g') for n in names)
    stepped = [names[i] for i in sorted(ck['optimizer']['state'])]
    assert stepped == [n for n in names if not n.startswith('_vq.')]          # EMA codebook: no optimizer state
    ref_opt = torch.optim.Adam(model.parameters(), lr=cfg['learning_rate'], amsgrad=True)
    ref_opt.load_state_dict(ck['optimizer'])
    eng.step(torch.from_numpy(g['x1']))
    want = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model2, _ = _build(g, dev)
    eng2 = tr.FusedTrainStep(model2, int(g['B']), int(g['T']), cfg['learning_rate'], use_graph=False)
    tr.load_checkpoint(eng2, path)
    eng2.step(torch.from_numpy(g['x1']))
    for k, v in model2.state_dict().items():
        assert torch.equal(v, want[k]), k


def test_load_optimizer_state_restores_hyper_parameters():
    """The reference resumes through optimizer.load_state_dict, which restores lr / betas / eps of the checkpoint
    (pipeline_factory.py:118-120): a step built with another learning rate must continue with the checkpoint's."""
    dev = _dev()
    from vq_vae_speech_b200 import trainer as tr
    g = load_golden('model_ema_k44')
    B, T = int(g['B']), int(g['T'])
    model, cfg = _build(g, dev)
    eng = tr.FusedTrainStep(model, B, T, cfg['learning_rate'], use_graph=True)
    eng.step(torch.from_numpy(g['x0']))
    sd = tr.optimizer_state_dict(eng)
    after0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    eng.step(torch.from_numpy(g['x1']))
    want = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model2, _ = _build(g, dev)
    eng2 = tr.FusedTrainStep(model2, B, T, 10 * cfg['learning_rate'], use_graph=True)
    with torch.no_grad():
        sd2 = model2.state_dict()
        for k, v in after0.items():
            sd2[k].copy_(v)
    tr.load_optimizer_state_dict(eng2, sd)
    assert abs(eng2.lr - cfg['learning_rate']) < 1e-12
    eng2.step(torch.from_numpy(g['x1']))
    for k, v in model2.state_dict().items():
        assert torch.equal(v, want[k]), k


def test_fused_step_separate_target_matches_reference_formula():
    """convolutional_trainer.py:47,54: the reconstruction loss is taken against data['output_features'].  With a target
    different from the input, losses and the gradient of the decoder's last layer must equal the numpy oracle's."""
    dev = _dev()
    from oracle import model_oracle as mo
    from vq_vae_speech_b200.trainer import FusedTrainStep
    g = load_golden('model_ema_k44')
    model, cfg = _build(g, dev)
    B, T = int(g['B']), int(g['T'])
    p = {k[5:]: v.astype(np.float64) for k, v in g.items() if k.startswith('init.')}
    eng = FusedTrainStep(model, B, T, cfg['learning_rate'], use_graph=False, separate_target=True)
    x, target = g['x0'], g['x1']
    with pytest.raises(ValueError):
        eng.step(torch.from_numpy(x))                                    # target missing
    eng.step(torch.from_numpy(x), torch.from_numpy(target))
    got = eng.losses()
    ocfg = dict(commitment_cost=cfg['commitment_cost'], decay=cfg['decay'], num_residual_layers=cfg['num_residual_layers'],
                learning_rate=cfg['learning_rate'], epsilon=1e-5)
    out, c = mo.model_forward(dict(p), x, ocfg)
    grads, _ = mo.model_backward(p, c, out, target.transpose(0, 2, 1), ocfg)
    recon_loss = float(np.mean((out['reconstructed_x'] - target.transpose(0, 2, 1)) ** 2))
    assert rel_err(got['reconstruction_loss'], recon_loss) < TOL
    assert rel_err(got['vq_loss'], float(out['vq_loss'])) < TOL
    mine = eng.gradients()
    for n in ('_decoder._conv_trans_3.weight', '_decoder._conv_trans_3.bias', '_encoder._conv_1.weight', '_pre_vq_conv.weight'):
        assert rel_err(mine[n].cpu().numpy(), grads[n]) < 2e-5, n
    eng_same = FusedTrainStep(_build(g, dev)[0], B, T, cfg['learning_rate'], use_graph=False)
    with pytest.raises(ValueError):
        eng_same.step(torch.from_numpy(x), torch.from_numpy(target))     # built without separate_target


def test_feature_batcher_matches_reference_normalisation():
    """on-GPU normalisation == the reference's numpy float64 `(x - mean) / std` followed by .float()."""
    dev = _dev()
    from vq_vae_speech_b200.data import FeatureBatcher
    rng = np.random.RandomState(3)
    mean, std = rng.randn(39) * 10, np.abs(rng.randn(39)) * 5 + 0.5
    items = [{'input_features': rng.randn(47, 39) * 12 + 3, 'speaker_id': i} for i in range(8)]
    fb = FeatureBatcher(4, 47, dev, {'train_mean': mean, 'train_std': std}, rank=1, world_size=2)
    mine = fb.shard(items)
    assert [m['speaker_id'] for m in mine] == [4, 5, 6, 7]
    got = fb.collate(mine).cpu().numpy()
    ref = np.stack([((m['input_features'] - mean) / std) for m in mine]).astype(np.float32)
    assert np.array_equal(got, ref)
    with pytest.raises(ValueError):
        fb.collate(mine[:2])


def test_feature_loader_reads_pickles_and_survives_a_pipelined_loop(tmp_path):
    """SURVEY 8f N3 end to end: per-utterance pickles -> epoch order -> rank shard -> pinned staging -> GPU normalisation.
    The consumer never synchronises between batches (as FusedTrainStep.step() does not): every batch must still equal the
    reference's numpy normalisation of exactly its own utterances (the pinned staging slots rotate behind CUDA events)."""
    dev = _dev()
    from test_data_cpu import _write_features
    from vq_vae_speech_b200.data import FeatureLoader, FeaturePickleDataset
    n, B, T = 40, 4, 47
    _write_features(str(tmp_path), 'train', n, T=T, seed=5)
    ds = FeaturePickleDataset(str(tmp_path), 'train')
    rng = np.random.RandomState(1)
    norm = {'train_mean': rng.randn(39) * 5, 'train_std': np.abs(rng.randn(39)) * 3 + 0.5}
    for rank in range(2):
        torch.manual_seed(21)
        loader = FeatureLoader(ds, B, T, dev, normalizer=norm, rank=rank, world_size=2)
        assert len(loader) == n // (2 * B)
        kept, want = [], []
        for x, target, items in loader:                    # no synchronisation inside the loop
            assert target is x                              # output_features is input_features in these pickles
            kept.append(x.clone())
            want.append(np.stack([(it['input_features'] - norm['train_mean']) / norm['train_std'] for it in items]
                                 ).astype(np.float32))
        assert len(kept) == len(loader)
        for a, b in zip(kept, want):
            assert np.array_equal(a.cpu().numpy(), b)
    # the two ranks of one epoch see disjoint utterances
    seen = []
    for rank in range(2):
        torch.manual_seed(21)
        for _, _, items in FeatureLoader(ds, B, T, dev, rank=rank, world_size=2):
            seen += [it['index'] for it in items]
    assert len(seen) == len(set(seen)) == n
