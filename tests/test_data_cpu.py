"""Host logic of the data feed (vq-vae-speech_b200/data.py) -- no CUDA: the per-utterance pickle reader
against files written the way the reference's export writes them (vctk_features_dataset.py:43-64), the epoch order against
a real torch DataLoader(shuffle=True) (vctk_features_stream.py:56-62), and the rank sharding."""
import os
import pickle

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from vq_vae_speech_b200.data import EpochSampler, FeaturePickleDataset


def _write_features(root, subset, n, T=47, F=39, seed=0):
    rng = np.random.RandomState(seed)
    d = os.path.join(root, 'features', subset)
    os.makedirs(d)
    for i in range(n):
        feats = rng.randn(T, F) * 10 + 3
        dic = {'preprocessed_audio': None, 'wav_filename': 'p%03d_%03d.wav' % (225 + i % 3, i), 'input_features': feats,
               'output_features': feats, 'speaker_id': np.array(i % 3), 'quantized': None, 'one_hot': None,
               'shifting_time': 0.0}
        with open(os.path.join(d, '%d.pickle' % i), 'wb') as f:
            pickle.dump(dic, f)
    return d


def test_pickle_dataset_reads_like_the_reference(tmp_path):
    d = _write_features(str(tmp_path), 'train', 5)
    ds = FeaturePickleDataset(str(tmp_path), 'train')
    assert len(ds) == 5
    it = ds[3]
    assert it['index'] == 3 and it['input_features'].shape == (47, 39) and it['input_features'].dtype == np.float64
    assert it['quantized'].size == 0 and it['one_hot'].size == 0           # None -> empty array (dataset.py:60-61)
    with pytest.raises(OSError):
        ds[7]                                                              # "No such file"
    open(os.path.join(d, '4.pickle'), 'wb').close()
    with pytest.raises(OSError):
        ds[4]                                                              # "Empty file"
    # host-side normalisation on request == the reference's __getitem__ arithmetic (dataset.py:56-58)
    norm = {'train_mean': np.linspace(-1, 1, 39), 'train_std': np.linspace(0.5, 2, 39)}
    raw = ds[2]['input_features']
    got = FeaturePickleDataset(str(tmp_path), 'train', normalizer=norm)[2]['input_features']
    assert np.array_equal(got, (raw - norm['train_mean']) / norm['train_std'])
    with pytest.raises(ValueError):
        FeaturePickleDataset(str(tmp_path), 'test')
    with pytest.raises(OSError):
        FeaturePickleDataset(str(tmp_path), 'val')                         # no such directory


@pytest.mark.parametrize('n,B', [(23, 4), (64, 16), (7, 2)])
def test_epoch_order_equals_torch_dataloader(n, B):
    """Same torch seed -> the same batches in the same order as DataLoader(shuffle=True), over two epochs."""
    torch.manual_seed(1234)
    dl = DataLoader(list(range(n)), batch_size=B, shuffle=True, num_workers=0)
    ref = [[b.tolist() for b in dl] for _ in range(2)]
    torch.manual_seed(1234)
    s = EpochSampler(n, B, drop_last=False)
    mine = [list(s) for _ in range(2)]
    assert mine == ref
    assert ref[0] != ref[1]                                                # a new permutation every epoch
    torch.manual_seed(1234)
    dropped = list(EpochSampler(n, B, drop_last=True))
    assert dropped == [b for b in ref[0] if len(b) == B] and len(EpochSampler(n, B)) == n // B


def test_rank_shards_partition_every_global_batch():
    n, B, W = 50, 4, 3
    per_rank = []
    for r in range(W):
        torch.manual_seed(9)                                               # every rank seeds the global RNG identically
        per_rank.append(list(EpochSampler(n, B, rank=r, world_size=W)))
    torch.manual_seed(9)
    whole = list(EpochSampler(n, B * W, drop_last=True))                   # the global batches
    assert len(per_rank[0]) == n // (B * W) == len(whole)
    for k, g in enumerate(whole):
        assert sum((per_rank[r][k] for r in range(W)), []) == g            # contiguous slices, in rank order
        assert all(len(per_rank[r][k]) == B for r in range(W))
    with pytest.raises(ValueError):
        EpochSampler(n, B, rank=3, world_size=3)
    assert list(EpochSampler(6, 2, shuffle=False)) == [[0, 1], [2, 3], [4, 5]]
