"""Generate golden input/output vectors by running the UNMODIFIED reference (/root/reference/src).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/vq_*.npz and tests/golden/model_*.npz.  The reference's own tests hold no
vectors for this path, so these files are what pins the oracle
(tests/test_oracle_golden.py) and, through it, the CUDA path (tests/test_*_gpu.py).
Everything is CPU fp32, torch 2.11.0+cu128.
"""
import os
import sys

import numpy as np
import torch

REF_SRC = '/root/reference/src'
sys.path.insert(0, REF_SRC)
HERE = os.path.dirname(os.path.abspath(__file__))

from models.vector_quantizer import VectorQuantizer  # noqa: E402
from models.vector_quantizer_ema import VectorQuantizerEMA  # noqa: E402
from models.convolutional_vq_vae import ConvolutionalVQVAE  # noqa: E402
from modules.jitter import Jitter  # noqa: E402


def np32(t):
    return t.detach().cpu().numpy().copy()


def vq_case(name, K, D, B, T, ema, steps=3, seed=0, trained_like=False, dup_codes=False):
    """`steps` consecutive training-mode forwards+backwards of the VQ module alone, then one eval forward."""
    torch.manual_seed(seed)
    if ema:
        vq = VectorQuantizerEMA(K, D, 0.25, 0.99, 'cpu')
    else:
        vq = VectorQuantizer(K, D, 0.25, 'cpu')
    if dup_codes:  # true ties: duplicated code vectors must resolve to the lowest index (torch.argmin)
        with torch.no_grad():
            vq._embedding.weight[K // 2] = vq._embedding.weight[1]
            vq._embedding.weight[K - 1] = vq._embedding.weight[1]
    vq.train()
    rec = dict(K=K, D=D, B=B, T=T, ema=int(ema), steps=steps, commitment_cost=0.25, decay=0.99, epsilon=1e-5, g_loss=1.5)
    rec['W0'] = np32(vq._embedding.weight)
    if ema:
        rec['ema_w0'] = np32(vq._ema_w)
        rec['cs0'] = np32(vq._ema_cluster_size)
    for s in range(steps):
        if trained_like:
            W = vq._embedding.weight.detach()
            rows = W[torch.randint(0, K, (B * T,))] + 0.1 * torch.randn(B * T, D)
            z = rows.view(D, T, B).permute(2, 0, 1).contiguous()
        else:
            z = torch.randn(B, D, T)
        if dup_codes and s == 0:
            z[:, :, 0] = 0.0   # includes all-zero rows
        z.requires_grad_(True)
        g = torch.randn(B, D, T)
        W_before = vq._embedding.weight
        outs = vq(z, record_codebook_stats=True)
        vq_loss, quantized, perplexity, encodings, distances, idx, losses = outs[:7]
        (vq_loss * 1.5 + (quantized * g).sum()).backward()
        rec[f'z{s}'] = np32(z)
        rec[f'g{s}'] = np32(g)
        rec[f'vq_loss{s}'] = np32(vq_loss)
        rec[f'quantized{s}'] = np32(quantized)
        rec[f'perplexity{s}'] = np32(perplexity)
        rec[f'idx{s}'] = np32(idx)
        rec[f'grad_z{s}'] = np32(z.grad)
        rec[f'concat{s}'] = np32(outs[10])
        if s == 0:
            rec['encodings0'] = np32(encodings)
            rec['distances0'] = np32(distances)
        if ema:
            rec[f'W{s + 1}'] = np32(vq._embedding.weight)
            rec[f'ema_w{s + 1}'] = np32(vq._ema_w)
            rec[f'cs{s + 1}'] = np32(vq._ema_cluster_size)
        else:
            rec[f'grad_E{s}'] = np32(W_before.grad)
            W_before.grad = None
            with torch.no_grad():   # plain SGD so that the codebook moves between steps
                W_before -= 0.1 * torch.from_numpy(rec[f'grad_E{s}'])
            rec[f'W{s + 1}'] = np32(vq._embedding.weight)
    vq.eval()
    z = torch.randn(B, D, T)
    outs = vq(z, compute_distances_if_possible=False)
    rec['z_eval'] = np32(z)
    rec['eval_vq_loss'] = np32(outs[0])
    rec['eval_quantized'] = np32(outs[1])
    rec['eval_perplexity'] = np32(outs[2])
    rec['eval_idx'] = np32(outs[5])
    rec['eval_concat'] = np32(outs[10])
    np.savez_compressed(os.path.join(HERE, f'vq_{name}.npz'), **rec)
    print('wrote', name)


def vq_eval_tables_case(name, K, D, B, T, seed):
    """Eval-mode forward of the reference VectorQuantizer with compute_distances_if_possible=True: the three O(N^2)
    distance tables (vector_quantizer.py:108-127; the EMA class raises NameError there)."""
    torch.manual_seed(seed)
    vq = VectorQuantizer(K, D, 0.25, 'cpu').eval()
    with torch.no_grad():
        vq._embedding.weight.normal_()
    z = torch.randn(B, D, T)
    outs = vq(z, compute_distances_if_possible=True)
    rec = dict(K=K, D=D, B=B, T=T, W=np32(vq._embedding.weight), z=np32(z), idx=np32(outs[5]),
               encoding_distances=np32(outs[7]), embedding_distances=np32(outs[8]),
               frames_vs_embedding_distances=np32(outs[9]), quantized=np32(outs[1]), concat=np32(outs[10]))
    np.savez_compressed(os.path.join(HERE, f'evaltables_{name}.npz'), **rec)
    print('wrote', name)


def model_cfg(**over):
    cfg = dict(output_features_filters=13, augment_output_features=True, output_features_dim=47, verbose=False,
               input_features_dim=47, num_hiddens=48, num_residual_layers=2, use_kaiming_normal=False,
               input_features_type='mfcc', input_features_filters=13, augment_input_features=True,
               sampling_rate=16000, embedding_dim=64, decay=0.99, num_embeddings=44, commitment_cost=0.25,
               residual_channels=48, use_jitter=False, jitter_probability=0.12, use_speaker_conditioning=False,
               record_codebook_stats=False, learning_rate=2e-4)
    cfg.update(over)
    return cfg


def model_case(name, B=2, T=47, steps=3, seed=1234, speakers=0, **over):
    """`steps` full trainer iterations (convolutional_trainer.py:44-74 restated: the trainer module itself
    needs matplotlib/tqdm-free imports that are absent here)."""
    cfg = model_cfg(**over)
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = ConvolutionalVQVAE(cfg, 'cpu').train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg['learning_rate'], amsgrad=True)
    crit = torch.nn.MSELoss()
    rec = {f'cfg_{k}': v for k, v in cfg.items() if isinstance(v, (int, float, bool))}
    rec.update(B=B, T=T, steps=steps, seed=seed, speakers=speakers)
    speaker_dic = {'p%03d' % i: i for i in range(speakers)} if speakers else None
    for k, v in model.state_dict().items():
        rec['init.' + k] = np32(v)
    for s in range(steps):
        x = torch.randn(B, T, 39)
        target = x.permute(0, 2, 1).contiguous().float()
        rng_state = np.random.get_state()
        opt.zero_grad()
        speaker_id = None
        if speakers:
            # the reference draws a fresh random speaker embedding inside forward (global_conditioning.py:34): pin the
            # host RNG right before the call so that a replay can draw the same one
            speaker_id = torch.randint(0, speakers, (B, 1))
            torch.manual_seed(1000 + s)
            probe = torch.nn.Embedding(speakers, 40)
            probe.weight.data.normal_(0, 0.1)
            rec[f'speaker_id{s}'] = np32(speaker_id)
            rec[f'speaker_features{s}'] = np32(probe.weight[speaker_id.view(-1)])
            torch.manual_seed(1000 + s)
        recon, vq_loss, losses, perplexity, idx, _ = model(x, speaker_dic, speaker_id)
        recon_loss = crit(recon, target)
        loss = vq_loss + recon_loss
        loss.backward()
        if cfg['use_jitter']:   # recover the plan the reference just drew, for the record
            np.random.set_state(rng_state)
            from oracle.model_oracle import jitter_plan
            rec[f'jitter_src{s}'] = jitter_plan(T // 2 + 1, cfg['jitter_probability'])
        if s == 0:
            for n, prm in model.named_parameters():
                if prm.grad is not None:
                    rec['grad0.' + n] = np32(prm.grad)
        opt.step()
        rec[f'x{s}'] = np32(x)
        rec[f'recon{s}'] = np32(recon)
        rec[f'vq_loss{s}'] = np32(vq_loss)
        rec[f'recon_loss{s}'] = np32(recon_loss)
        rec[f'perplexity{s}'] = np32(perplexity)
        rec[f'idx{s}'] = np32(idx)
    for k, v in model.state_dict().items():
        rec['final.' + k] = np32(v)
    np.savez_compressed(os.path.join(HERE, f'model_{name}.npz'), **rec)
    print('wrote', name)


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    torch.set_num_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == '--only-evaltables':   # added later as well
        vq_eval_tables_case('noema_k44_d64_b2_t24', 44, 64, 2, 24, seed=11)
        vq_eval_tables_case('noema_k10_d2_b3_t17', 10, 2, 3, 17, seed=12)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == '--only-speaker':
        model_case('ema_k29_speaker', decay=0.99, num_embeddings=29, use_speaker_conditioning=True, speakers=5, seed=99)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == '--only-h64':          # round 2: num_hiddens % 32 == 0 -> the conv-mode tcgen05
        # GEMMs of the fused step are eligible for every layer (the 48-wide fixtures run them on the CUDA-core kernel)
        model_case('ema_k44_h64', decay=0.99, num_hiddens=64, residual_channels=64, B=4, seed=2024)
        model_case('noema_jitter_k44_h96', decay=0.0, num_hiddens=96, residual_channels=64, B=3, use_jitter=True, seed=2025)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == '--only-kaiming':      # added after the other fixtures were committed
        model_case('ema_k29_kaiming', decay=0.99, num_embeddings=29, use_kaiming_normal=True, seed=777)
        sys.exit(0)
    # VQ bottleneck alone (BASELINE.json configs[2]: vq44, vq29, vq10x2) + layout / tie edge cases
    vq_case('ema_k44_d64_b2_t24', 44, 64, 2, 24, True)
    vq_case('ema_k29_d64_b2_t24', 29, 64, 2, 24, True, seed=1)
    vq_case('ema_k10_d2_b2_t24', 10, 2, 2, 24, True, seed=2)
    vq_case('ema_k44_d64_b3_t17', 44, 64, 3, 17, True, seed=3)          # T*B % D != 0: rows straddle channels
    vq_case('ema_k44_d64_b16_t96_trained', 44, 64, 16, 96, True, seed=4, trained_like=True, steps=2)
    vq_case('ema_k44_d64_b2_t24_dup', 44, 64, 2, 24, True, seed=5, dup_codes=True, steps=1)
    vq_case('noema_k44_d64_b2_t24', 44, 64, 2, 24, False, seed=6)
    vq_case('noema_k10_d2_b2_t24', 10, 2, 2, 24, False, seed=7)
    vq_case('noema_k100_d64_b5_t24', 100, 64, 5, 24, False, seed=8, steps=2)
    vq_case('ema_k512_d64_b8_t24', 512, 64, 8, 24, True, seed=9, steps=2)
    # full model, reduced width so the fixtures stay small (arithmetic path is identical)
    model_case('ema_k44', decay=0.99)
    model_case('noema_k44', decay=0.0)
    model_case('ema_jitter_k29', decay=0.99, num_embeddings=29, use_jitter=True, seed=5678)
    model_case('noema_k10_d2', decay=0.0, num_embeddings=10, embedding_dim=2, seed=4242)
    model_case('ema_k44_b5_t191', decay=0.99, B=5, T=191, steps=2, num_hiddens=32, residual_channels=24)
    vq_eval_tables_case('noema_k44_d64_b2_t24', 44, 64, 2, 24, seed=11)
    vq_eval_tables_case('noema_k10_d2_b3_t17', 10, 2, 3, 17, seed=12)
    # speaker conditioning (deconvolutional_decoder.py:108-111; experiments_vq29-mfcc39.json:21)
    model_case('ema_k29_speaker', decay=0.99, num_embeddings=29, use_speaker_conditioning=True, speakers=5, seed=99)
    # weight-normalised convs (use_kaiming_normal: conv1d_builder.py:41-43, residual.py:45-47,57-59)
    model_case('ema_k29_kaiming', decay=0.99, num_embeddings=29, use_kaiming_normal=True, seed=777)
    # widths that are multiples of 32: every conv-mode GEMM of the fused step is tcgen05-eligible
    model_case('ema_k44_h64', decay=0.99, num_hiddens=64, residual_channels=64, B=4, seed=2024)
    model_case('noema_jitter_k44_h96', decay=0.0, num_hiddens=96, residual_channels=64, B=3, use_jitter=True, seed=2025)
