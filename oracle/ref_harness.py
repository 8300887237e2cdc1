"""Runs the UNMODIFIED reference modules (oracle/_ref/src, or /root/reference/src in the build container) -- TEST /
BASELINE INFRASTRUCTURE ONLY: bench.py's `cpu_baseline` / `--impl reference` legs, its `gpu_eager_baseline` context
numbers and tests/ may import this; the product never does.

The timed call is the reference's own ConvolutionalTrainer.iterate (src/experiments/convolutional_trainer.py:44-74) on
the reference's own ConvolutionalVQVAE, fed through the two attributes of the data stream it reads (`speaker_dic`,
`training_batch_size`) and a progress-bar stand-in with set_description.  The only shim is an empty module for
`evaluation.gradient_stats` (imports matplotlib; reached only with record_codebook_stats on).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.path.join(HERE, '_ref', 'src'), os.environ.get('VQS_REFERENCE_SRC', '/root/reference/src')]


def ref_src():
    """Directory holding the reference's `models/`, `modules/`, ... packages, or None."""
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, 'models', 'convolutional_vq_vae.py')) and \
                os.path.isfile(os.path.join(c, 'experiments', 'convolutional_trainer.py')):
            return c
    return None


def available():
    return ref_src() is not None


_loaded = {}


def load():
    """Imports the reference's ConvolutionalVQVAE and ConvolutionalTrainer.  Returns (ConvolutionalVQVAE, ConvolutionalTrainer)."""
    if _loaded:
        return _loaded['model'], _loaded['trainer']
    src = ref_src()
    if src is None:
        raise RuntimeError('reference files not found (run python oracle/build_ref.py where /root/reference exists)')
    if src not in sys.path:
        sys.path.insert(0, src)
    try:
        import matplotlib  # noqa: F401  (present: let the reference import its real gradient_stats)
        have_mpl = os.path.isfile(os.path.join(src, 'evaluation', 'gradient_stats.py'))
    except Exception:
        have_mpl = False
    if not have_mpl and 'evaluation.gradient_stats' not in sys.modules:
        pkg = sys.modules.setdefault('evaluation', types.ModuleType('evaluation'))
        stub = types.ModuleType('evaluation.gradient_stats')

        class GradientStats(object):   # reached only when record_codebook_stats is on
            @staticmethod
            def build_gradient_entry(named_parameters):
                raise RuntimeError('evaluation.gradient_stats needs matplotlib (stand-in module)')
        stub.GradientStats = GradientStats
        pkg.gradient_stats = stub
        sys.modules['evaluation.gradient_stats'] = stub
    from models.convolutional_vq_vae import ConvolutionalVQVAE
    from experiments.convolutional_trainer import ConvolutionalTrainer
    _loaded['model'], _loaded['trainer'], _loaded['src'] = ConvolutionalVQVAE, ConvolutionalTrainer, src
    return ConvolutionalVQVAE, ConvolutionalTrainer


class _Stream(object):
    """The two data-stream attributes the trainer touches (convolutional_trainer.py:52, base_trainer.py:84)."""

    def __init__(self, batch_size):
        self.speaker_dic = {}
        self.training_batch_size = batch_size


class _Bar(object):
    def set_description(self, text):
        self.text = text


class RefTrainer(object):
    """The reference model + trainer on `device` ('cpu' or 'cuda'); step(x) = one ConvolutionalTrainer.iterate."""

    def __init__(self, cfg, seed=1234, device='cpu'):
        import numpy as np
        import torch
        Model, Trainer = load()
        cfg = dict(cfg)
        cfg.setdefault('record_codebook_stats', False)
        cfg.setdefault('start_epoch', 0)
        cfg.setdefault('num_epochs', 1)
        self.device = torch.device(device)
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.model = Model(cfg, self.device).to(self.device).train()
        self.trainer = Trainer(self.device, _Stream(cfg.get('batch_size', 2)), cfg, '/tmp', 'bench', model=self.model)
        self.bar = _Bar()
        self.source = ref_src()

    def step(self, x_btf, target_btf=None, speaker_id=None):
        import torch
        B = x_btf.shape[0]
        data = {'input_features': x_btf, 'output_features': x_btf if target_btf is None else target_btf,
                'speaker_id': speaker_id if speaker_id is not None else torch.zeros(B, dtype=torch.long)}
        losses, perplexity = self.trainer.iterate(data, 0, 0, [], self.bar)
        losses = dict(losses)
        losses['perplexity'] = perplexity
        return losses
