"""GPU parity of the BENCHMARKED configuration itself: num_hiddens = 768, per-GPU batch 64 x 47 frames (and 16 x 191), GEMM
engine '3xtf32' on tcgen05, the step captured into a CUDA graph -- what bench.py times -- against the reference run live on
the CPU: the unmodified reference modules + ConvolutionalTrainer.iterate from oracle/_ref when those files travelled with
the snapshot (oracle/build_ref.py), else the torch-CPU port oracle/torch_port.py (itself pinned to the reference by
tests/test_oracle_golden.py::test_torch_port_matches_reference).

Bars (north_star): indices exact outside near-ties -- a flipped row must have an fp64 top-2 distance gap below
NEAR_TIE (relative to |x|^2 + |e|^2) on the REFERENCE's z; z, reconstruction, losses, EMA state 1e-5; parameters after
the optimizer steps within the update budget (Adam divides by sqrt(v): where |grad| is at fp32-noise level the direction
is noise, tests/test_oracle_golden.py uses the same budget).

Gradients at this size cannot be held to 2e-5 tensor by tensor against ANY other implementation, the reference's own
arithmetic included: each layer has 2.4 M ReLU pre-activations, the forward passes of two fp32 implementations differ by
~1e-6 relative, so O(1) pre-activations per layer sit closer to zero than that and their masks flip -- the gradient's
analogue of a near-tie VQ frame.  ONE flipped element of typical size moves the relative L2 error of every gradient
downstream to ~1e-4 (measured: torch-CPU fp32 vs the same step in fp64 differs by 4e-4 on `_encoder._conv_2.weight`; our
exact-fp32 CUDA-core engine, whose GEMMs are each within 9e-7 of fp64, shows the same pattern: profiles/
r02k_fullsize_grads.txt; the tcgen05 engine with its truncation-loss compensation is within 8e-6 of the fp64 step on every
tensor there, closer than the reference's own fp32 arithmetic).  So the full-size step asserts flip-tolerant bounds
(relative L2 <= 1e-3 per tensor against the fp64 step, cosine of the whole flat gradient >= 1 - 1e-6), prints the
per-tensor table, and the 2e-5 gradient bar is enforced where it is well defined: on every backward GEMM of the step in
isolation at exactly these shapes against fp64 (test_backward_gemms_at_benchmarked_shapes_match_fp64), and on the whole
step at the reference-fixture sizes (tests/test_model_gpu.py).  Also asserts that every tcgen05-eligible GEMM of the step
really ran on tcgen05 (vqs_engine_count), so a silent CUDA-core fallback cannot pass.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
# z reaches the bottleneck through ~8 GEMM layers that agree with the CPU reference to ~1e-6 relative each (different
# summation order: oneDNN / MKL vs tcgen05 3xTF32); a row whose two best codes are closer than this can legitimately flip
NEAR_TIE = 2e-5


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


class _Reference(object):
    """The live CPU reference behind one interface: step(x) -> dict, grads(), state().  exact=True: the port in fp64."""

    def __init__(self, cfg, sd, seed, exact=False):
        from oracle import ref_harness
        self.kind = 'reference' if (ref_harness.available() and not exact) else 'port'
        self.exact = exact
        self.cap = {}
        if self.kind == 'reference':
            self.tr = ref_harness.RefTrainer(cfg, seed=seed)
            self.model = self.tr.model
            self.model.load_state_dict(sd)
            self.model._vq.register_forward_pre_hook(self._pre_vq)
            self.model.register_forward_hook(self._post_model)
        else:
            from oracle.torch_port import PortTrainer
            self.tr = PortTrainer(cfg, seed=seed)
            self.model = self.tr.model
            self.model.load_reference_state(sd)
            if exact:
                self.model.double()
                self.tr.opt = torch.optim.Adam(self.model.parameters(), lr=cfg['learning_rate'], amsgrad=True)
            self.model.pre.register_forward_hook(lambda m, i, o: self.cap.__setitem__('z', o.detach().clone()))

    def _pre_vq(self, module, inputs):
        self.cap['z'] = inputs[0].detach().clone()
        self.cap['W'] = module._embedding.weight.detach().clone()

    def _post_model(self, module, inputs, outputs):
        self.cap['recon'] = outputs[0].detach().clone()
        self.cap['idx'] = outputs[4].detach().clone()

    def step(self, x):
        if self.kind == 'port':
            self.cap['W'] = self.model.emb.weight.detach().clone()
        out = dict(self.tr.step(x.double() if self.exact else x))
        if self.kind == 'port':
            self.cap['recon'], self.cap['idx'] = out['reconstructed_x'], out['encoding_indices']
        out.update(z=self.cap['z'].numpy(), W=self.cap['W'].numpy(), recon=self.cap['recon'].numpy(),
                   idx=self.cap['idx'].numpy().reshape(-1))
        return out

    def _names(self):
        if self.kind == 'reference':
            seen = set()
            for n, p in self.model.named_parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    yield n, p
            return
        km = dict((v, k) for k, v in self.model.KEYMAP.items())
        for ref, mine in km.items():
            mod = getattr(self.model, mine)
            yield ref + '.weight', mod.weight
            yield ref + '.bias', mod.bias
        for mine, ref in (('eres', '_encoder'), ('dres', '_decoder')):
            yield ref + '._residual_stack._layers.0._block.1.weight', getattr(self.model, mine).c1.weight
            yield ref + '._residual_stack._layers.0._block.3.weight', getattr(self.model, mine).c2.weight

    def grads(self):
        return dict((n, p.grad.detach().double().numpy().copy()) for n, p in self._names() if p.grad is not None)

    def params(self):
        return dict((n, p.detach().numpy().copy()) for n, p in self._names())

    def vq_state(self):
        if self.kind == 'reference':
            vq = self.model._vq
            return dict(W=vq._embedding.weight.detach().numpy(), ema_w=vq._ema_w.detach().numpy(),
                        cs=vq._ema_cluster_size.detach().numpy())
        return dict(W=self.model.emb.weight.detach().numpy(), ema_w=self.model.ema_w.detach().numpy(),
                    cs=self.model.cs.detach().numpy())


def _near_tie_rows(z_bdt, W):
    """fp64 top-2 gap of every VQ row of the reference's z (rows formed as vector_quantizer_ema.py:101-106), relative to
    |x|^2 + |e|^2."""
    B, D, T = z_bdt.shape
    rows = np.ascontiguousarray(np.transpose(z_bdt.astype(np.float64), (1, 2, 0))).reshape(-1, D)
    W = W.astype(np.float64)
    d = (rows ** 2).sum(1, keepdims=True) + (W ** 2).sum(1)[None, :] - 2.0 * rows @ W.T
    part = np.partition(d, 1, axis=1)
    gap = part[:, 1] - part[:, 0]
    scale = (rows ** 2).sum(1) + (W[np.argmin(d, 1)] ** 2).sum(1)
    return gap / np.maximum(scale, 1e-30)


@pytest.mark.parametrize('B,T', [(64, 47), (16, 191)])
def test_benchmarked_config_matches_live_reference(B, T):
    dev = _dev()
    from vq_vae_speech_b200 import _lib
    from vq_vae_speech_b200.convolutional_vq_vae import ConvolutionalVQVAE
    from vq_vae_speech_b200.trainer import FusedTrainStep, reference_config
    cfg = reference_config(decay=0.99, batch_size=B)          # bench.py's model: num_hiddens 768, EMA codebook 44 x 64
    seed, steps = 1234, 3
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = ConvolutionalVQVAE(cfg, 'cpu')
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref = _Reference(cfg, sd, seed)
    ref64 = _Reference(cfg, sd, seed, exact=True)
    model = model.to(dev).train()
    eng = FusedTrainStep(model, B, T, cfg['learning_rate'], use_graph=True, precision='3xtf32')
    # every conv-like GEMM must run on tcgen05, the two 39-channel layers through operand images padded to 64 channels
    from vq_vae_speech_b200 import functional as F
    conv = [e[2] for e in eng.schedule if e[0] is not None and e[0].__name__ == 'vqs_conv_gemm']
    wgr = [e[2] for e in eng.schedule if e[0] is not None and e[0].__name__ == 'vqs_wgrad_gemm']
    assert sum(1 for d in conv if d.Cred % 32 != 0) == 2
    want_cc = sum(1 for d in conv if not F.conv_tc_eligible(d.M, d.Cred, '3xtf32'))
    assert want_cc == 0 and len(conv) >= 32
    gen = torch.Generator().manual_seed(seed)
    flipped_total = 0
    for s in range(steps):
        x = torch.randn(B, T, 39, generator=gen)
        c0 = _lib.engine_counts()
        eng.step(x)
        got = eng.losses()
        c1 = _lib.engine_counts()
        if s == 0:       # the first step runs launch by launch through the C ABI (later ones replay the captured graph)
            assert c1['conv_cudacore'] - c0['conv_cudacore'] == want_cc
            assert c1['conv_tc'] - c0['conv_tc'] == len(conv) - want_cc
            assert c1['wgrad_cudacore'] == c0['wgrad_cudacore']
            assert (c1['wgrad_tc'] - c0['wgrad_tc']) + (c1['wgrad_tma'] - c0['wgrad_tma']) == len(wgr)
        r = ref.step(x)
        idx = eng.encoding_indices().cpu().numpy().reshape(-1)
        if s > 0:
            # After the first optimizer step the two runs no longer hold the same parameters to 1e-5: Adam's first updates
            # are lr * g / (|g| + eps) ~ lr * sign(g), so an entry whose gradient is at rounding-noise level moves by
            # lr = 2e-4 in a direction that is noise in ANY implementation (the reference's included; this is why the
            # post-step parameters are judged against the update budget).  Later steps are therefore held to the size of
            # that perturbation: forward quantities 2e-3, at least 99 % of the indices equal.
            assert np.mean(idx == r['idx']) >= 0.99, (s, float(np.mean(idx == r['idx'])))
            assert rel_err(eng.buf['z'].cpu().numpy(), r['z']) < 2e-3
            assert rel_err(eng.buf['recon'].cpu().numpy(), r['recon']) < 2e-3
            for k in ('reconstruction_loss', 'vq_loss', 'perplexity', 'loss'):
                assert rel_err(got[k], r[k]) < 2e-3, (s, k, got[k], r[k])
            flipped_total += int((idx != r['idx']).sum())
            continue
        flipped = np.nonzero(idx != r['idx'])[0]
        if flipped.size:
            gaps = _near_tie_rows(r['z'], r['W'])
            assert (gaps[flipped] < NEAR_TIE).all(), ('index mismatch outside near-ties', flipped[:8], gaps[flipped][:8])
            flipped_total += flipped.size
            print('step %d: %d near-tie rows flipped (gap < %g relative): reported' % (s, flipped.size, NEAR_TIE))
        assert rel_err(eng.buf['z'].cpu().numpy(), r['z']) < TOL
        assert rel_err(eng.buf['recon'].cpu().numpy(), r['recon']) < TOL
        for k in ('reconstruction_loss', 'vq_loss', 'perplexity', 'loss'):
            assert rel_err(got[k], r[k]) < TOL, (s, k, got[k], r[k])
        if s == 0:
            ref64.step(x)
            grads, rg, rg64 = eng.gradients(), ref.grads(), ref64.grads()
            assert len(rg) >= 24 and sorted(rg) == sorted(rg64)
            worst_ref = max(rel_err(rg[n], rg64[n]) for n in rg)
            num = den_a = den_b = 0.0
            table = []
            for n, g_ref in rg.items():
                mine = grads[n].cpu().numpy().astype(np.float64)
                l2 = float(np.linalg.norm(mine - rg64[n]) / max(np.linalg.norm(rg64[n]), 1e-300))
                table.append((n, rel_err(mine, g_ref), rel_err(mine, rg64[n]), l2, rel_err(g_ref, rg64[n])))
                assert l2 <= 1e-3, (n, l2)
                num += float((mine * rg64[n]).sum())
                den_a += float((mine * mine).sum())
                den_b += float((rg64[n] * rg64[n]).sum())
            cosine = num / np.sqrt(den_a * den_b)
            print('gradients (name, max-norm vs fp32 reference, max-norm vs fp64 step, L2 vs fp64 step, reference vs fp64):')
            for row in table:
                print('  %-58s %.2e %.2e %.2e %.2e' % row)
            print('cosine of the flat gradient with the fp64 step: 1 - %.2e' % (1.0 - cosine))
            assert cosine >= 1.0 - 1e-6
            # (a single flipped mask contaminates every tensor downstream of it, so the number of tensors beyond 2e-5 says
            # nothing; 64 x 47 on this seed: none beyond 8e-6; 16 x 191: one flip early in the decoder's backward, and the
            # reference's own fp32 run has one worth 7e-3 on the same step)
            print('tensors beyond 2e-5 of the fp64 step: %d of %d; reference fp32 vs fp64 worst: %.2e' % (
                sum(1 for r_ in table if r_[2] > 2e-5), len(table), worst_ref))
    assert eng.graph is not None
    if flipped_total == 0:
        st = ref.vq_state()
        vq = model._vq
        assert rel_err(vq._embedding.weight.detach().cpu().numpy(), st['W']) < 2e-3
        assert rel_err(vq._ema_w.detach().cpu().numpy(), st['ema_w']) < 2e-3
        assert rel_err(vq._ema_cluster_size.detach().cpu().numpy(), st['cs']) < TOL       # functions of the counts only
    # parameters after three optimizer steps: within the update budget (each entry moved by at most ~ lr per step)
    budget = 2.5 * cfg['learning_rate'] * steps
    mine = dict(model.named_parameters())
    close, total = 0, 0
    for n, p_ref in ref.params().items():
        if n.startswith('_vq.'):
            continue
        d = np.abs(mine[n].detach().cpu().numpy() - p_ref)
        assert float(d.max()) < budget, n
        close += int((d < 0.05 * cfg['learning_rate'] * steps).sum())
        total += d.size
    assert close >= 0.97 * total, (close, total)      # all but the noise-gradient entries moved the same way


@pytest.mark.parametrize('case', ['model_ema_k44_h64', 'model_noema_jitter_k44_h96'])
def test_eligible_layers_of_the_reference_fixtures_run_on_tcgen05(case):
    """The num_hiddens % 32 == 0 reference fixtures (tests/golden/make_golden.py --only-h64; parity itself is checked by
    test_model_gpu.py::test_fused_step_matches_reference for every model_*.npz): with the '3xtf32' engine every conv-like
    GEMM whose reduction width is a multiple of 32 must be dispatched to tcgen05 -- none to the CUDA-core kernel."""
    dev = _dev()
    from conftest import load_golden
    from test_model_gpu import _build
    from vq_vae_speech_b200 import _lib
    from vq_vae_speech_b200.trainer import FusedTrainStep
    g = load_golden(case)
    model, cfg = _build(g, dev)
    eng = FusedTrainStep(model, int(g['B']), int(g['T']), cfg['learning_rate'], use_graph=False, precision='3xtf32')
    from vq_vae_speech_b200 import functional as F
    conv = [e[2] for e in eng.schedule if e[0] is not None and e[0].__name__ == 'vqs_conv_gemm']
    eligible = sum(1 for d in conv if F.conv_tc_eligible(d.M, d.Cred, '3xtf32'))   # (39-channel layers: padded images from M = 128 on)
    assert eligible >= len(conv) - 2 and eligible >= 30
    if cfg['use_jitter']:
        np.random.seed(int(g['seed']))
    c0 = _lib.engine_counts()
    eng.step(torch.from_numpy(g['x0']))
    eng.losses()
    c1 = _lib.engine_counts()
    assert c1['conv_tc'] - c0['conv_tc'] == eligible
    assert c1['conv_cudacore'] - c0['conv_cudacore'] == len(conv) - eligible
    assert np.array_equal(eng.encoding_indices().cpu().numpy().reshape(-1), g['idx0'].reshape(-1))
    assert rel_err(eng.buf['recon'].cpu().numpy(), g['recon0']) < TOL


BWD_SHAPES = [  # (B, Cin, Cout, L, k, stride, pad): the conv layers of the benchmarked step (num_hiddens 768, batch 64)
    (64, 768, 768, 47, 3, 1, 1), (64, 768, 768, 47, 4, 2, 2), (64, 768, 768, 24, 3, 1, 1), (64, 768, 768, 24, 1, 1, 0),
    (64, 768, 768, 48, 3, 1, 1), (64, 768, 768, 48, 1, 1, 0), (64, 768, 64, 24, 3, 1, 1), (64, 64, 768, 24, 3, 1, 1),
    (64, 39, 768, 47, 3, 1, 1),
]


@pytest.mark.parametrize('B,Cin,Cout,L,k,stride,pad', BWD_SHAPES)
@pytest.mark.parametrize('prec', ['3xtf32', 'fp32'])
def test_backward_gemms_at_benchmarked_shapes_match_fp64(B, Cin, Cout, L, k, stride, pad, prec):
    """Every conv layer of the benchmarked step, forward / dgrad / wgrad in isolation (masks play no role), against the
    same operation in fp64 (torch double on the GPU: test infrastructure): max-norm 1e-5 (the bar), relative L2 5e-6, and
    the SIGNED bias mean((ours - ref) sign(ref)) / mean|ref| -- the tensor core accumulates with truncation, which shrinks
    every sum systematically; the epilogue compensates the expected loss (gemm_tc.cu), so the residual bias must stay
    below 5e-7 (uncompensated it is -1.8e-6 at K = 2304 and compounds linearly over the ~17 GEMMs of the backward chain)."""
    dev = _dev()
    from vq_vae_speech_b200 import functional as F, ops
    prev = ops.set_precision(prec)
    try:
        g = torch.Generator(device=dev).manual_seed(B + Cin + L + k)
        x = torch.randn(B, Cin, L, device=dev, generator=g)
        w = torch.randn(Cout, Cin, k, device=dev, generator=g) / (Cin * k) ** 0.5
        y64 = torch.nn.functional.conv1d(x.double(), w.double(), None, stride, pad)
        gy = torch.randn(*y64.shape, device=dev, generator=g)
        xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
        (torch.nn.functional.conv1d(xr, wr, None, stride, pad) * gy.double()).sum().backward()
        y = F.conv1d_forward(x, F.gemm_weight(w, 'conv_fwd'), None, stride, pad)
        dx = F.conv1d_dgrad(gy, F.gemm_weight(w, 'conv_dgrad'), L, stride, pad)
        dW = torch.empty_like(w)
        F.conv1d_wgrad(gy, x, dW, stride, pad, F._wgrad_ws(Cout, Cin, k, B, y64.shape[2], dev))
        for name, a, r in (('fwd', y, y64), ('dgrad', dx, xr.grad), ('wgrad', dW, wr.grad)):
            d = a.double() - r
            l2 = (d.norm() / r.norm()).item()
            mx = (d.abs().max() / r.abs().max()).item()
            bias = ((d * r.sign()).mean() / r.abs().mean()).item()
            assert mx < 1e-5 and l2 < 5e-6 and abs(bias) < 5e-7, (name, prec, 'max', mx, 'L2', l2, 'signed bias', bias)
    finally:
        ops.set_precision(prev)
