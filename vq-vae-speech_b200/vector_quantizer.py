"""Drop-in VectorQuantizer / VectorQuantizerEMA backed by the sm_100a kernels.

Same constructors, forward signature, 11-tuple return order, attribute and state_dict names as the reference
(/root/reference/src/models/vector_quantizer.py:58-160, vector_quantizer_ema.py:65-187).  Differences, all deliberate:
  * the EMA update writes `_ema_cluster_size`, `_ema_w` and `_embedding.weight` IN PLACE instead of re-creating
    nn.Parameter objects every step (ema.py:154,156) -- values, names and requires_grad are identical, storage is stable
    (CUDA graphs, data parallelism); only `id()` differs;
  * inputs must be CUDA fp32 tensors: there is no CPU path;
  * the eval-only O(N^2) Python distance loops (vector_quantizer.py:108-127 / ema.py:122-140) are ONE kernel launch each
    (vqs_pairwise_l2, same itertools ordering and shapes).  The reference's EMA class raises NameError on those lines
    (`product` is never imported there); here both classes return the tables.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
from .functional import VQFn


class _VQBase(nn.Module):
    #: materialise the (B, T, K) `encodings` / `distances` tuple slots (the reference always does; ConvolutionalVQVAE
    #: discards them, convolutional_vq_vae.py:128, so it switches this off)
    materialize_outputs = True
    #: fill the float dict with .item() like the reference (one device sync per forward)
    sync_losses = True
    #: optional callable(stats, n_rows) -> n_rows_total that sums the [counts | dw] statistics over data-parallel ranks
    stats_allreduce = None
    #: row layout; LAYOUT_FLAT_ND treats a 2-D (N, D) input as N ready-made rows (sharding-invariant)
    layout = LAYOUT_BDT_AS_DTB
    #: also record (best, second-best) distance per row in `last_dmin2` (near-tie report)
    record_near_ties = False

    def _workspace(self, device):
        ws = getattr(self, '_ws', None)
        if ws is None or ws.device != device:
            ws = ops.vq_workspace(self._num_embeddings, self._embedding_dim, device)
            object.__setattr__(self, '_ws', ws)
        return ws

    def _run(self, inputs, ema, record_codebook_stats):
        if not inputs.is_cuda:
            raise RuntimeError('%s: inputs must be on a CUDA device (no CPU fallback)' % type(self).__name__)
        K, D = self._num_embeddings, self._embedding_dim
        layout = LAYOUT_FLAT_ND if inputs.dim() == 2 else self.layout
        if layout == LAYOUT_FLAT_ND:
            B, T = 1, inputs.shape[0]
            N = inputs.shape[0]
        else:
            B, _, T = inputs.shape
            N = B * T
        dev = inputs.device
        want = {}
        if self.materialize_outputs:
            want['distances'] = torch.empty(N, K, dtype=torch.float32, device=dev)
        if self.record_near_ties:
            want['dmin2'] = torch.empty(N, 2, dtype=torch.float32, device=dev)
        need_concat = (not self.training) or record_codebook_stats
        if need_concat:
            want['q_rows'] = torch.empty(N, D, dtype=torch.float32, device=dev)
        state = dict(layout=layout, beta=float(self._commitment_cost), ws=self._workspace(dev), ema=ema,
                     training=self.training, stats_allreduce=self.stats_allreduce, want=want)
        quantized, scalars = VQFn.apply(inputs.float(), self._embedding.weight, state)
        idx = want['idx']
        encodings = distances = None
        if self.materialize_outputs:
            encodings = ops.vq_one_hot(idx, K).view(B, T, K)
            distances = want['distances'].view(B, T, K)
        if self.record_near_ties:
            self.last_dmin2 = want['dmin2']
        self.last_stats = want['stats']
        return quantized, scalars, encodings, distances, idx.view(N, 1), want.get('q_rows')

    def _eval_tables(self, inputs, compute_distances_if_possible):
        """(encoding_distances (B, -1), embedding_distances (K (K-1)/2,), frames_vs_embedding_distances (B, T, K)) in eval
        mode, (None, None, None) otherwise -- vector_quantizer.py:108-127."""
        if self.training or not compute_distances_if_possible:
            return None, None, None
        D = self._embedding_dim
        x = inputs.float().contiguous()
        W = self._embedding.weight.detach().contiguous()
        layout = LAYOUT_FLAT_ND if x.dim() == 2 else self.layout
        if layout == LAYOUT_FLAT_ND:
            batch, time = 1, x.shape[0]
        else:
            batch, time = x.shape[0], x.shape[2]
        enc = ops.pairwise_l2(x, layout, D).view(batch, -1)
        emb = ops.pairwise_l2(W, LAYOUT_FLAT_ND, D)
        fve = ops.pairwise_l2(x, layout, D, W).view(batch, time, -1)
        return enc, emb, fve

    @property
    def embedding(self):
        return self._embedding


class VectorQuantizerEMA(_VQBase):
    """reference: src/models/vector_quantizer_ema.py:39-187."""

    def __init__(self, num_embeddings, embedding_dim, commitment_cost, decay, device, epsilon=1e-5):
        super(VectorQuantizerEMA, self).__init__()
        self._num_embeddings = num_embeddings
        self._embedding_dim = embedding_dim
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.normal_()
        self._commitment_cost = commitment_cost
        self.register_buffer('_ema_cluster_size', torch.zeros(num_embeddings))
        self._ema_w = nn.Parameter(torch.Tensor(num_embeddings, self._embedding_dim))
        self._ema_w.data.normal_()
        self._decay = decay
        self._device = device
        self._epsilon = epsilon

    def forward(self, inputs, compute_distances_if_possible=True, record_codebook_stats=False):
        ema = dict(cluster_size=self._ema_cluster_size, ema_w=self._ema_w.data, decay=self._decay, eps=self._epsilon)
        quantized, scalars, encodings, distances, idx, concat = self._run(inputs, ema, record_codebook_stats)
        vq_loss = scalars[3]
        perplexity = scalars[2].detach()
        losses = {'vq_loss': vq_loss.item()} if self.sync_losses else {'vq_loss': vq_loss.detach()}
        enc_d, emb_d, fve_d = self._eval_tables(inputs, compute_distances_if_possible)
        return (vq_loss, quantized, perplexity, encodings, distances, idx, losses, enc_d, emb_d, fve_d, concat)


class VectorQuantizer(_VQBase):
    """reference: src/models/vector_quantizer.py:38-160."""

    def __init__(self, num_embeddings, embedding_dim, commitment_cost, device):
        super(VectorQuantizer, self).__init__()
        self._embedding_dim = embedding_dim
        self._num_embeddings = num_embeddings
        self._embedding = nn.Embedding(self._num_embeddings, self._embedding_dim)
        self._embedding.weight.data.uniform_(-1 / self._num_embeddings, 1 / self._num_embeddings)
        self._commitment_cost = commitment_cost
        self._device = device

    def forward(self, inputs, compute_distances_if_possible=True, record_codebook_stats=False):
        quantized, scalars, encodings, distances, idx, concat = self._run(inputs, None, record_codebook_stats)
        vq_loss = scalars[4]
        perplexity = scalars[2].detach()
        if self.sync_losses:
            s = scalars.detach().cpu()          # one D2H copy for the four floats of vector_quantizer.py:154-155
            losses = {'e_latent_loss': float(s[1]), 'q_latent_loss': float(s[1]),
                      'commitment_loss': float(s[3]), 'vq_loss': float(s[4])}
        else:
            d = scalars.detach()
            losses = {'e_latent_loss': d[1], 'q_latent_loss': d[1], 'commitment_loss': d[3], 'vq_loss': d[4]}
        enc_d, emb_d, fve_d = self._eval_tables(inputs, compute_distances_if_possible)
        return (vq_loss, quantized, perplexity, encodings, distances, idx, losses, enc_d, emb_d, fve_d, concat)
