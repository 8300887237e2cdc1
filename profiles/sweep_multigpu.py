"""BASELINE.json configs[4] on the whole machine: jittered batch-size / utterance-length sweep of the full training step,
weak scaling (fixed per-GPU batch) and strong scaling (global batch 256), data parallel over WORLD_SIZE GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/sweep_multigpu.py out.jsonl
    python profiles/sweep_multigpu.py out.jsonl          (one GPU: the baseline of every point)

One process group, one FusedTrainStep per point (use_jitter = true, 3xtf32 tcgen05 GEMMs, CUDA graph); 20 timed steps after
5 warm-up steps, CUDA events, max over ranks.  One JSON line per point (rank 0)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
import torch.distributed as dist
from vq_vae_speech_b200 import parallel
if world > 1:
    parallel.init_nccl(dev)
from vq_vae_speech_b200.convolutional_vq_vae import ConvolutionalVQVAE
from vq_vae_speech_b200.trainer import FusedTrainStep, reference_config
out_path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/sweep_multigpu.jsonl'
# (per-GPU batch, frames, kind)
POINTS = [(2, 47, 'weak'), (16, 47, 'weak'), (32, 47, 'weak'), (64, 47, 'weak'), (128, 47, 'weak'), (256, 47, 'weak'),
          (16, 191, 'weak'), (64, 191, 'weak')]
if 256 % world == 0:
    POINTS.append((256 // world, 47, 'strong(global 256)'))
STEPS, WARM = 20, 5
f = open(out_path, 'w') if rank == 0 else None
for B, T, kind in POINTS:
    cfg = reference_config(decay=0.99, batch_size=B, use_jitter=True)
    torch.manual_seed(1234); np.random.seed(1234 + rank)
    model = ConvolutionalVQVAE(cfg, dev).to(dev).train()
    eng = FusedTrainStep(model, B, T, cfg['learning_rate'], precision='3xtf32')
    gen = torch.Generator().manual_seed(1234 + rank)
    xs = [torch.randn(B, T, 39, generator=gen).to(dev) for _ in range(4)]
    for i in range(WARM):
        eng.step(xs[i % 4])
    eng.losses()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(STEPS):
        eng.step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    losses = eng.losses()
    if rank == 0:
        rec = {'n_gpus': world, 'per_gpu_batch': B, 'global_batch': B * world, 'frames': T, 'scaling': kind, 'use_jitter': True,
               'ms_per_step': round(ms, 4), 'utterances_per_s': round(B * world / (ms * 1e-3), 1),
               'frames_per_s': round(B * world * T / (ms * 1e-3), 1), 'exchange': 'nvls' if getattr(eng, 'nvls', None) is not None else ('nccl' if world > 1 else None),
               'loss': losses['loss']}
        f.write(json.dumps(rec) + '\n'); f.flush()
        print(json.dumps(rec), file=sys.stderr, flush=True)
    del eng, model, xs
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
if f: f.close()
os._exit(0)
