"""Data feed for the fused training step: per-utterance feature dicts, as the reference's pickles hold
them (`input_features` (T, 39) float64; src/dataset/vctk_features_dataset.py:43-64), are collated into pinned host
batches, sharded over data-parallel ranks, copied to the device and normalised THERE with the reference's float64
arithmetic (`(x - train_mean) / train_std`, :56-58) -- bit-identical to normalising in numpy and casting with .float().
Reading the pickle files themselves is left to the caller (any iterable of dicts works; the reference's DataLoader does)."""
import numpy as np
import torch

from . import ops


class FeatureBatcher(object):
    def __init__(self, batch_size, num_frames, device, normalizer=None, rank=0, world_size=1, key='input_features',
                 num_filters=39):
        """normalizer: None or the dict the reference pickles as data/vctk/vctk-mfcc-stats.pickle
        ({'train_mean': (39,), 'train_std': (39,)}).  rank / world_size: this process's shard of every global batch."""
        self.B, self.T, self.F = int(batch_size), int(num_frames), int(num_filters)
        self.device = torch.device(device)
        self.rank, self.world, self.key = int(rank), int(world_size), key
        self.host = torch.empty(self.B, self.T, self.F, dtype=torch.float64).pin_memory()
        self.dev64 = torch.empty(self.B, self.T, self.F, dtype=torch.float64, device=self.device)
        self.out = torch.empty(self.B, self.T, self.F, dtype=torch.float32, device=self.device)
        if normalizer is not None:
            mean = np.asarray(normalizer['train_mean'], np.float64).reshape(-1)
            std = np.asarray(normalizer['train_std'], np.float64).reshape(-1)
        else:
            mean, std = np.zeros(self.F), np.ones(self.F)
        if mean.shape[0] != self.F or std.shape[0] != self.F:
            raise ValueError('normalizer statistics must have %d entries' % self.F)
        self.mean = torch.from_numpy(mean).to(self.device)
        self.std = torch.from_numpy(std).to(self.device)

    def shard(self, global_items):
        """This rank's utterances of a global batch of world_size * batch_size items (contiguous slices, like
        DataParallelComm.shard)."""
        if len(global_items) != self.B * self.world:
            raise ValueError('global batch must hold %d utterances, got %d' % (self.B * self.world, len(global_items)))
        return global_items[self.rank * self.B:(self.rank + 1) * self.B]

    def collate(self, items):
        """items: batch_size dicts (or arrays) of (T, F) features -> normalised float32 (B, T, F) device tensor."""
        if len(items) != self.B:
            raise ValueError('expected %d utterances, got %d' % (self.B, len(items)))
        for i, it in enumerate(items):
            a = np.asarray(it[self.key] if isinstance(it, dict) else it, np.float64)
            if a.shape != (self.T, self.F):
                raise ValueError('utterance %d has shape %s, expected %s' % (i, a.shape, (self.T, self.F)))
            self.host[i].copy_(torch.from_numpy(a))
        self.dev64.copy_(self.host, non_blocking=True)
        return ops.normalize_features(self.dev64, self.mean, self.std, out=self.out)
