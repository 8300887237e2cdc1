"""CPU port of the reference training step on the reference's own numeric backend -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure PyTorch (requirements.txt:2): on CPU its hot path is ATen's oneDNN convolutions, MKL matmul and
element-wise kernels.  /root/reference cannot travel to the GPU box, so this file restates the same step with the same
torch CPU operators, independently written and compact:

  ConvolutionalEncoder.forward    /root/reference/src/models/convolutional_encoder.py:118-146
  Residual / ResidualStack        /root/reference/src/modules/residual.py:69-70, residual_stack.py:43-46
  VectorQuantizer{,EMA}.forward   /root/reference/src/models/vector_quantizer.py:88-150, vector_quantizer_ema.py:101-179
  Jitter.forward                  /root/reference/src/modules/jitter.py:47-70
  DeconvolutionalDecoder.forward  /root/reference/src/models/deconvolutional_decoder.py:100-137
  ConvolutionalVQVAE.forward      /root/reference/src/models/convolutional_vq_vae.py:117-139
  ConvolutionalTrainer.iterate    /root/reference/src/experiments/convolutional_trainer.py:44-74

It is what bench.py times as `cpu_baseline` (kind "port") and as `--impl reference`, with all host threads.  It is pinned
against the reference's golden vectors by tests/test_oracle_golden.py::test_torch_port_matches_reference.  Only tests/,
__graft_entry__.smoke() and bench.py may import it; the product never does.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .model_oracle import jitter_plan


class _Res(nn.Module):
    def __init__(self, c, r):
        super().__init__()
        self.c1 = nn.Conv1d(c, r, 3, padding=1, bias=False)
        self.c2 = nn.Conv1d(r, c, 1, bias=False)

    def forward(self, x):
        x = F.relu(x)                       # nn.ReLU(inplace=True) at the head of the block: the skip sees relu(x)
        return x + self.c2(F.relu(self.c1(x)))


class PortVQVAE(nn.Module):
    """Layer construction order follows the reference so that the same torch seed gives the same initial weights."""

    def __init__(self, cfg):
        super().__init__()
        C, D, K = cfg['num_hiddens'], cfg['embedding_dim'], cfg['num_embeddings']
        Fi = cfg['input_features_filters'] * (3 if cfg['augment_input_features'] else 1)
        Fo = cfg['output_features_filters'] * (3 if cfg['augment_output_features'] else 1)
        self.cfg = cfg
        self.e1 = nn.Conv1d(Fi, C, 3, padding=1)
        self.e2 = nn.Conv1d(C, C, 3, padding=1)
        self.e3 = nn.Conv1d(C, C, 4, stride=2, padding=2)
        self.e4 = nn.Conv1d(C, C, 3, padding=1)
        self.e5 = nn.Conv1d(C, C, 3, padding=1)
        self.eres = _Res(C, C)
        self.pre = nn.Conv1d(C, D, 3, padding=1)
        self.emb = nn.Embedding(K, D)
        self.ema = cfg['decay'] > 0.0
        if self.ema:
            self.emb.weight.data.normal_()
            self.register_buffer('cs', torch.zeros(K))
            self.ema_w = nn.Parameter(torch.Tensor(K, D))
            self.ema_w.data.normal_()
        else:
            self.emb.weight.data.uniform_(-1 / K, 1 / K)
        # deconvolutional_decoder.py:56: +40 channels of speaker features when use_speaker_conditioning
        self.d1 = nn.Conv1d(D + (40 if cfg.get('use_speaker_conditioning', False) else 0), C, 3, padding=1)
        self.dres = _Res(C, cfg['residual_channels'])
        self.t1 = nn.ConvTranspose1d(C, C, 3, padding=1)
        self.t2 = nn.ConvTranspose1d(C, C, 3, padding=0)
        self.t3 = nn.ConvTranspose1d(C, Fo, 2, padding=0)
        self.weight_norm = bool(cfg.get('use_kaiming_normal', False))
        if self.weight_norm:     # conv1d_builder.py:41-43, conv_transpose1d_builder.py:41-43, residual.py:45-47,57-59
            # (_pre_vq_conv is a plain nn.Conv1d: convolutional_vq_vae.py:61-66)
            for name in ('e1', 'e2', 'e3', 'e4', 'e5', 'd1', 't1', 't2', 't3'):
                setattr(self, name, nn.utils.weight_norm(getattr(self, name)))
            for res in (self.eres, self.dres):
                res.c1 = nn.utils.weight_norm(res.c1)
                res.c2 = nn.utils.weight_norm(res.c2)

    # name map to the reference's state_dict keys (for loading golden initial weights)
    KEYMAP = {'e1': '_encoder._conv_1', 'e2': '_encoder._conv_2', 'e3': '_encoder._conv_3', 'e4': '_encoder._conv_4',
              'e5': '_encoder._conv_5', 'pre': '_pre_vq_conv', 'd1': '_decoder._conv_1',
              't1': '_decoder._conv_trans_1', 't2': '_decoder._conv_trans_2', 't3': '_decoder._conv_trans_3'}

    def load_reference_state(self, sd):
        def put(mod, key):
            for suffix in (('weight_g', 'weight_v') if hasattr(mod, 'weight_g') else ('weight',)):
                getattr(mod, suffix).copy_(torch.as_tensor(sd[key + '.' + suffix]))

        with torch.no_grad():
            for mine, ref in self.KEYMAP.items():
                put(getattr(self, mine), ref)
                getattr(self, mine).bias.copy_(torch.as_tensor(sd[ref + '.bias']))
            for mine, ref in (('eres', '_encoder'), ('dres', '_decoder')):
                put(getattr(self, mine).c1, ref + '._residual_stack._layers.0._block.1')
                put(getattr(self, mine).c2, ref + '._residual_stack._layers.0._block.3')
            self.emb.weight.copy_(torch.as_tensor(sd['_vq._embedding.weight']))
            if self.ema:
                self.ema_w.copy_(torch.as_tensor(sd['_vq._ema_w']))
                self.cs.copy_(torch.as_tensor(sd['_vq._ema_cluster_size']))

    def quantize(self, z):
        """The bottleneck; returns (vq_loss, quantized (B, D, T), perplexity, idx (N, 1))."""
        cfg = self.cfg
        inp = z.permute(1, 2, 0).contiguous()                        # (D, T, B)  -- the reference's row definition
        flat = inp.view(-1, cfg['embedding_dim'])
        W = self.emb.weight
        dist = (flat.pow(2).sum(1, keepdim=True) + W.pow(2).sum(1)) - 2 * flat @ W.t()
        idx = dist.argmin(1, keepdim=True)
        enc = torch.zeros(idx.shape[0], cfg['num_embeddings'], dtype=flat.dtype).scatter_(1, idx, 1)
        if self.ema and self.training:
            g, eps, K = cfg['decay'], cfg.get('epsilon', 1e-5), cfg['num_embeddings']
            with torch.no_grad():
                cs = self.cs * g + (1 - g) * enc.sum(0)
                n = cs.sum()
                self.cs = (cs + eps) / (n + K * eps) * n
                self.ema_w.data = self.ema_w.data * g + (1 - g) * (enc.t() @ flat.detach())
                self.emb.weight.data = self.ema_w.data / self.cs.unsqueeze(1)
            W = self.emb.weight.detach()
        q = (enc @ W).view(inp.shape)
        e_latent = F.mse_loss(q.detach(), inp)
        if self.ema:
            loss = cfg['commitment_cost'] * e_latent
        else:
            loss = F.mse_loss(q, inp.detach()) + cfg['commitment_cost'] * e_latent
        q = inp + (q - inp).detach()
        p = enc.mean(0)
        ppl = torch.exp(-(p * torch.log(p + 1e-10)).sum())
        return loss, q.permute(2, 0, 1).contiguous(), ppl, idx

    def forward(self, x_btf, jitter_src=None, speaker_features=None):
        nl = self.cfg['num_residual_layers']
        x = x_btf.permute(0, 2, 1).contiguous().to(self.emb.weight.dtype)   # .float() in the reference (vq_vae.py:118)
        a1 = F.relu(self.e1(x))
        h = F.relu(self.e2(a1)) + a1
        a3 = F.relu(self.e3(h))
        h = F.relu(self.e4(a3)) + a3
        h5 = F.relu(self.e5(h)) + h
        r = h5
        for _ in range(nl):
            r = self.eres(r)                                         # the SAME block applied nl times
        z = self.pre(F.relu(r) + h5)
        vq_loss, q, ppl, idx = self.quantize(z)
        if jitter_src is not None:
            keep = torch.as_tensor(jitter_src == np.arange(len(jitter_src)))
            src = torch.as_tensor(jitter_src, dtype=torch.long)
            q = torch.where(keep[None, None, :], q, q.detach()[:, :, src])   # replaced columns carry no gradient
        if speaker_features is not None:     # (B, 40), repeated over time (global_conditioning.py:52-57)
            q = torch.cat([q, speaker_features[:, :, None].expand(-1, -1, q.shape[2])], dim=1)
        y = F.interpolate(self.d1(q), scale_factor=2)
        for _ in range(nl):
            y = self.dres(y)
        y = F.relu(self.t1(F.relu(y)))
        y = F.relu(self.t2(y))
        y = self.t3(y)
        return y[:, :, :x.shape[2]], vq_loss, ppl, idx


class PortTrainer(object):
    """zero_grad -> forward -> MSE + vq_loss -> backward -> Adam(amsgrad=True).step()  (trainer.py:44-74)."""

    def __init__(self, cfg, seed=1234):
        torch.manual_seed(seed)
        np.random.seed(seed)
        self.cfg = cfg
        self.model = PortVQVAE(cfg).train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=cfg['learning_rate'], amsgrad=True)

    def step(self, x_btf, speaker_features=None):
        cfg = self.cfg
        src = None
        if cfg['use_jitter']:
            src = jitter_plan(x_btf.shape[1] // 2 + 1, cfg['jitter_probability'])
        self.opt.zero_grad()
        recon, vq_loss, ppl, idx = self.model(x_btf, src, speaker_features)
        target = x_btf.permute(0, 2, 1).contiguous().to(recon.dtype)         # .float() in the reference (trainer.py:47)
        recon_loss = F.mse_loss(recon, target)
        loss = vq_loss + recon_loss
        loss.backward()
        self.opt.step()
        return dict(loss=loss.item(), reconstruction_loss=recon_loss.item(), vq_loss=vq_loss.item(),
                    perplexity=ppl.item(), encoding_indices=idx, reconstructed_x=recon.detach(), jitter_src=src)
