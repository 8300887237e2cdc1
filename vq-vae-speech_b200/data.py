"""Data feed for the fused training step, mirroring the reference's feature pipeline:

  FeaturePickleDataset   src/dataset/vctk_features_dataset.py:31-64   <vctk_path>/<features_path>/<subset>/<index>.pickle,
                                                                       one dict per utterance
  EpochSampler           src/dataset/vctk_features_stream.py:56-62     DataLoader(shuffle=True): a fresh permutation per
                                                                       epoch drawn exactly like torch's RandomSampler, cut
                                                                       into global batches, each rank taking its slice
  FeatureBatcher         vctk_features_dataset.py:56-58 + collate      pinned host batch -> device -> `(x - train_mean) /
                                                                       train_std` in the reference's float64 arithmetic,
                                                                       stored as float32 (bit-identical to numpy + .float())
  FeatureLoader          the three together: iterate (input, target) device batches of one rank for an epoch

The reference normalises on the host inside __getitem__; here the raw float64 features travel and the GPU normalises
(vqs_normalize_features), which gives the same bits.  No arithmetic happens on the host.
"""
import os
import pickle
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import ops


class FeaturePickleDataset(object):
    """The per-utterance pickles the reference's feature export writes (vctk_features_dataset.py:31-64): item `index` is
    `<vctk_path>/<features_path>/<subset>/<index>.pickle`, a dict with at least 'input_features' and 'output_features'
    ((T, 39) float64), 'speaker_id', 'quantized', 'one_hot', ...  Same errors as the reference for a missing / empty file;
    same None -> empty-array substitution; 'index' added.  Normalisation is NOT applied here (FeatureBatcher does it on the
    GPU); pass normalizer=... to get the reference's host-side behaviour instead."""

    def __init__(self, vctk_path, subset, normalizer=None, features_path='features'):
        if subset not in ('train', 'val'):
            raise ValueError("subset must be 'train' or 'val'")
        self._sub_features_path = os.path.join(vctk_path, features_path, subset)
        if not os.path.isdir(self._sub_features_path):
            raise OSError("No such directory '{}'".format(self._sub_features_path))
        self._files_number = len(os.listdir(self._sub_features_path))
        self._normalizer = normalizer

    def __getitem__(self, index):
        path = self._sub_features_path + os.sep + str(index) + '.pickle'
        if not os.path.isfile(path):
            raise OSError("No such file '{}'".format(path))
        if os.path.getsize(path) == 0:
            raise OSError("Empty file '{}'".format(path))
        with open(path, 'rb') as file:
            dic = pickle.load(file)
        if self._normalizer:
            dic['input_features'] = (dic['input_features'] - self._normalizer['train_mean']) / self._normalizer['train_std']
            dic['output_features'] = (dic['output_features'] - self._normalizer['train_mean']) / self._normalizer['train_std']
        dic['quantized'] = np.array([]) if dic.get('quantized') is None else dic['quantized']
        dic['one_hot'] = np.array([]) if dic.get('one_hot') is None else dic['one_hot']
        dic['index'] = index
        return dic

    def __len__(self):
        return self._files_number


class EpochSampler(object):
    """Index order of one epoch, sharded over data-parallel ranks.

    shuffle=True reproduces `DataLoader(dataset, batch_size, shuffle=True)` of the reference (vctk_features_stream.py:56-62):
    the DataLoader iterator first draws its worker base seed (one int64) from the global torch RNG, then torch's
    RandomSampler seeds a private generator with a second int64 and takes `torch.randperm(n, generator=...)` -- the same
    draws are made here, so under the same `torch.manual_seed` a single rank sees the reference's batches in the
    reference's order (tests/test_data_cpu.py compares against a real DataLoader).  The permutation is cut into GLOBAL batches of
    world_size * batch_size utterances; rank r takes items [r * B, (r + 1) * B) of each (the contiguous split of
    parallel.DataParallelComm.shard).  Every rank must draw the same permutation: seed the global RNG identically on all
    ranks (bench.py and the trainer do) or pass `generator`.  drop_last=True (default) drops the final partial global batch:
    the captured step has a fixed shape; drop_last=False yields it (shorter) as the reference's DataLoader does."""

    def __init__(self, num_items, batch_size, rank=0, world_size=1, shuffle=True, drop_last=True, generator=None):
        self.n, self.B = int(num_items), int(batch_size)
        self.rank, self.world = int(rank), int(world_size)
        self.shuffle, self.drop_last, self.generator = bool(shuffle), bool(drop_last), generator
        if not (0 <= self.rank < self.world):
            raise ValueError('rank %d outside world of %d' % (self.rank, self.world))

    def __len__(self):
        g = self.B * self.world
        return self.n // g if self.drop_last else (self.n + g - 1) // g

    def permutation(self):
        if not self.shuffle:
            return list(range(self.n))
        if self.generator is None:
            torch.empty((), dtype=torch.int64).random_()                         # torch/utils/data/dataloader.py: _base_seed
            seed = int(torch.empty((), dtype=torch.int64).random_().item())      # torch/utils/data/sampler.py RandomSampler
            gen = torch.Generator()
            gen.manual_seed(seed)
        else:
            gen = self.generator
        return torch.randperm(self.n, generator=gen).tolist()

    def __iter__(self):
        """Yields this rank's index list for every global batch of the epoch (a new permutation per call = per epoch)."""
        perm = self.permutation()
        g = self.B * self.world
        for k in range(len(self)):
            chunk = perm[k * g:(k + 1) * g]
            per = len(chunk) // self.world if len(chunk) < g else self.B
            yield chunk[self.rank * per:(self.rank + 1) * per]


class FeatureBatcher(object):
    NBUF = 4      # staging slots in rotation (collate_pair uses two per batch when the target differs from the input)

    def __init__(self, batch_size, num_frames, device, normalizer=None, rank=0, world_size=1, key='input_features',
                 num_filters=39):
        """normalizer: None or the dict the reference pickles as data/vctk/vctk-mfcc-stats.pickle
        ({'train_mean': (39,), 'train_std': (39,)}).  rank / world_size: this process's shard of every global batch.

        Staging is a rotation of NBUF (pinned host, float64 device, float32 output) slots, each guarded by a CUDA event
        recorded after its H2D copy + normalisation: collate() never rewrites pinned memory a pending DMA still reads, and
        the tensor it returns stays valid until NBUF - 1 further collate() calls (FusedTrainStep.load_batch copies it into
        its own static buffer on the same stream, so a pipelined loop may run ahead of the GPU safely)."""
        self.B, self.T, self.F = int(batch_size), int(num_frames), int(num_filters)
        self.device = torch.device(device)
        self.rank, self.world, self.key = int(rank), int(world_size), key
        self.host = [torch.empty(self.B, self.T, self.F, dtype=torch.float64).pin_memory() for _ in range(self.NBUF)]
        self.dev64 = [torch.empty(self.B, self.T, self.F, dtype=torch.float64, device=self.device) for _ in range(self.NBUF)]
        self.out = [torch.empty(self.B, self.T, self.F, dtype=torch.float32, device=self.device) for _ in range(self.NBUF)]
        self.done = [None] * self.NBUF
        self.slot = 0
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

    def collate(self, items, key=None):
        """items: batch_size dicts (or arrays) of (T, F) features -> normalised float32 (B, T, F) device tensor (valid until
        NBUF - 1 further calls)."""
        if len(items) != self.B:
            raise ValueError('expected %d utterances, got %d' % (self.B, len(items)))
        key = self.key if key is None else key
        s = self.slot
        self.slot = (s + 1) % self.NBUF
        if self.done[s] is not None:
            self.done[s].synchronize()         # the previous use of this slot (DMA out of pinned memory, normalise) is over
        host = self.host[s]
        for i, it in enumerate(items):
            a = np.asarray(it[key] if isinstance(it, dict) else it, np.float64)
            if a.shape != (self.T, self.F):
                raise ValueError('utterance %d has shape %s, expected %s' % (i, a.shape, (self.T, self.F)))
            host[i].copy_(torch.from_numpy(a))
        self.dev64[s].copy_(host, non_blocking=True)
        out = ops.normalize_features(self.dev64[s], self.mean, self.std, out=self.out[s])
        ev = torch.cuda.Event()
        ev.record()
        self.done[s] = ev
        return out

    def collate_pair(self, items):
        """(input, target) device batches: data['input_features'] and data['output_features'] of the reference's loader
        (convolutional_trainer.py:45-47).  When both keys hold the same array the target is the input tensor itself."""
        x = self.collate(items, 'input_features')
        same = all(isinstance(it, dict) and it.get('output_features') is it.get('input_features') for it in items)
        return x, (x if same else self.collate(items, 'output_features'))


class FeatureLoader(object):
    """dataset + EpochSampler + FeatureBatcher: `for x, target, items in loader:` yields one rank's normalised device batches
    for one epoch (a new permutation every time it is iterated).  Pickles are read by `num_workers` threads one batch ahead
    (file reads release the GIL); `items` are the raw dicts (speaker_id etc.)."""

    def __init__(self, dataset, batch_size, num_frames, device, normalizer=None, rank=0, world_size=1, shuffle=True,
                 drop_last=True, num_workers=2, generator=None):
        self.dataset = dataset
        self.sampler = EpochSampler(len(dataset), batch_size, rank, world_size, shuffle, drop_last, generator)
        self.batcher = FeatureBatcher(batch_size, num_frames, device, normalizer, rank, world_size)
        self.num_workers = max(int(num_workers), 1)

    def __len__(self):
        return len(self.sampler)

    def _read(self, indices):
        return [self.dataset[i] for i in indices]

    def __iter__(self):
        with ThreadPoolExecutor(self.num_workers) as pool:
            pending = None
            for indices in self.sampler:
                if len(indices) != self.batcher.B:
                    break                                   # partial last batch: fixed-shape consumers stop here
                nxt = pool.submit(self._read, indices)
                if pending is not None:
                    items = pending.result()
                    x, t = self.batcher.collate_pair(items)
                    yield x, t, items
                pending = nxt
            if pending is not None:
                items = pending.result()
                x, t = self.batcher.collate_pair(items)
                yield x, t, items
