#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native VQ-VAE-Speech training hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): vq44-mfcc39 full training step -- ConvolutionalEncoder + VectorQuantizerEMA(44x64,
decay 0.99) + DeconvolutionalDecoder, MSE + vq_loss, Adam(lr 2e-4, amsgrad) -- on synthetic normalised-MFCC-39 batches of
VCTK shape (T = 47 frames), fp32, per-GPU batch fixed (weak scaling), gradients and EMA statistics allreduced over NCCL.
Prints ONE JSON line (rank 0).  A "step" is one full training iteration on one batch.

  value      utterances/s over all ranks, inputs resident in HBM, K steps timed with CUDA events, max over ranks
  e2e        the same through FusedTrainStep.step() with HOST (pinned) batches: H2D copy + step + D2H of the losses
  roofline   the dominant kernel family of the timed region (per-launch CUDA events recorded inside the timed region)
  vq         the fused VQ fwd+bwd+EMA metric of BASELINE.json on N = 2^22 rows against the HBM roofline
  cpu_baseline  the reference's own modules + ConvolutionalTrainer.iterate (oracle/_ref, copied unmodified from
                /root/reference/src by oracle/build_ref.py) on the box's host cores (kind "reference"); the torch-CPU port
                oracle/torch_port.py (kind "port") only when oracle/_ref is absent
  gpu_eager_baseline  context only: the same unmodified reference modules on cuda:0 (cuDNN / cuBLAS eager), allow_tf32 off / on
  parity_check  (N > 1) replicated parameters / codebook bit-identical across ranks after the timed region, loss of every
                rank finite and equal to what the rank's own shard gives
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='utterances per GPU per step')
    ap.add_argument('--frames', type=int, default=47, help='MFCC frames per utterance (length 7680 -> 47)')
    ap.add_argument('--decay', type=float, default=0.99)
    ap.add_argument('--codes', type=int, default=44)
    ap.add_argument('--jitter', action='store_true', help='use_jitter = true (jitter.py:47-70, p = 0.12); BASELINE configs[4]')
    ap.add_argument('--precision', default='3xtf32', choices=['fp32', '3xtf32', 'tf32'],
                    help='GEMM engine of the conv GEMMs: 3xtf32 = tcgen05 with the fp32-accurate split (1e-5 parity)')
    ap.add_argument('--vq-rows', type=int, default=1 << 22)
    ap.add_argument('--skip-vq', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-eager', action='store_true', help='skip the gpu_eager_baseline context numbers')
    ap.add_argument('--nccl-max-ctas', type=int, default=None,
                    help='CTA cap of the NCCL communicator (default: parallel.NCCL_MAX_CTAS = 4; 0 = NCCL default)')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=float(p['hbm_gbs']), tensor=float(p['bf16_tflops_sustained']),
                    tensor_burst=float(p['bf16_tflops']), src='measured')
    except Exception:
        return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, src='fallback')


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.3] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)[kernel]['dram_bytes_per_launch']
    except Exception:
        return None


def model_config(args):
    from vq_vae_speech_b200.trainer import reference_config
    return reference_config(decay=args.decay, num_embeddings=args.codes, batch_size=args.batch,
                            use_jitter=bool(getattr(args, 'jitter', False)))


def flops_of(entry):
    """Algorithmic FLOPs of one recorded launch (2 * MACs that touch real data), from its descriptor."""
    fn, _, d = entry
    name = fn.__name__
    if name == 'vqs_conv_gemm':
        return 2.0 * d.M * d.Cred * d.ksz * d.B * d.Lout / d.l_div
    if name == 'vqs_wgrad_gemm':
        return 2.0 * d.M * d.Cred * d.ksz * d.B * d.La
    return 0.0


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle/torch_port.py on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_trainer(cfg, seed=1234, device='cpu'):
    """(trainer, kind): the unmodified reference (oracle/_ref or /root/reference) when present, else the torch-CPU port."""
    from oracle import ref_harness
    if ref_harness.available():
        return ref_harness.RefTrainer(cfg, seed=seed, device=device), 'reference'
    if device != 'cpu':
        return None, None
    from oracle.torch_port import PortTrainer
    return PortTrainer(cfg, seed=seed), 'port'


def cpu_train_throughput(cfg, batch, frames, seconds, warmup=2, min_steps=3, max_steps=200, fixed_steps=None, threads=None):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(threads or cores)
    tr, kind = cpu_trainer(cfg)
    gen = torch.Generator().manual_seed(1234)
    xs = [torch.randn(batch, frames, 39, generator=gen) for _ in range(4)]
    for i in range(warmup):
        tr.step(xs[i % 4])
    n, t0 = 0, time.perf_counter()
    while True:
        tr.step(xs[n % 4])
        n += 1
        el = time.perf_counter() - t0
        if fixed_steps is not None:
            if n >= fixed_steps:
                break
        elif (el >= seconds and n >= min_steps) or n >= max_steps:
            break
    el = time.perf_counter() - t0
    r = dict(value=batch * n / el, ms_per_step=1e3 * el / n, steps=n, cores=cores, threads=torch.get_num_threads(), kind=kind)
    torch.set_num_threads(cores)
    return r


def cpu_sample_text(kind, r, batch, frames):
    what = {'reference': 'the UNMODIFIED reference (oracle/_ref: models/*.py + ConvolutionalTrainer.iterate, '
                         'convolutional_trainer.py:44-74) on torch CPU',
            'port': 'torch-CPU port of the reference step (oracle/torch_port.py; oracle/_ref absent)'}[kind]
    return '%s, %d threads, %d steps of batch %d x %d frames (%.1f ms/step)' % (what, r['threads'], r['steps'], batch,
                                                                               frames, r['ms_per_step'])


def gpu_eager_baseline(cfg, batch, frames, dev, steps=10, warmup=3):
    """Context numbers, not the baseline: the unmodified reference modules on the SAME B200 through stock PyTorch eager
    (cuDNN convolutions, cuBLAS matmuls), once with TF32 off (fp32 parity numerics) and once with torch's TF32 switches on."""
    import torch
    from oracle import ref_harness
    if not ref_harness.available():
        return None
    out = {}
    gen = torch.Generator().manual_seed(1234)
    xs = [torch.randn(batch, frames, 39, generator=gen).to(dev) for _ in range(4)]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (('allow_tf32_false', False), ('allow_tf32_true', True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            tr = ref_harness.RefTrainer(cfg, seed=1234, device=str(dev))
            sid = torch.zeros(batch, dtype=torch.long, device=dev)
            for i in range(warmup):
                tr.step(xs[i % 4], speaker_id=sid)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                tr.step(xs[i % 4], speaker_id=sid)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {'ms_per_step': ms, 'utterances_per_s': batch / (ms * 1e-3)}
            del tr
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    out['what'] = ('unmodified reference modules + ConvolutionalTrainer.iterate on cuda (torch %s eager: cuDNN / cuBLAS, '
                   '3 .item() syncs per step), batch %d x %d frames, %d steps' % (torch.__version__, batch, frames, steps))
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref: unmodified modules + trainer; the
    port only if those files are absent) on the host cores, all threads.  Rank 0 alone runs; each step is a bounded sample
    (a smaller batch when needed) of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = model_config(args)
    # bound the run to a few minutes: calibrate one step at the full per-GPU batch, shrink the sample batch if needed
    batch = args.batch
    probe = cpu_train_throughput(cfg, batch, args.frames, 0.0, warmup=1, fixed_steps=1)
    budget_s = 150.0
    need = probe['ms_per_step'] * 1e-3 * (args.steps + args.warmup)
    while need > budget_s and batch > 2:
        batch = max(2, batch // 2)
        need /= 2
    r = cpu_train_throughput(cfg, batch, args.frames, 0.0, warmup=args.warmup, fixed_steps=args.steps)
    sample = cpu_sample_text(r['kind'], r, batch, args.frames)
    line = {
        'impl': 'reference', 'metric': 'train_utterances_per_sec', 'value': r['value'], 'unit': 'utterances/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 1, batch_override=batch),
        'cpu_baseline': {'value': r['value'], 'unit': 'utterances/s', 'cores': r['threads'], 'kind': r['kind'],
                         'sample': sample},
        'e2e': {'value': r['value'], 'unit': 'utterances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(args, world, batch_override=None):
    b = batch_override if batch_override is not None else args.batch
    return {'workload': 'vq44-mfcc39 full training step (encoder + VectorQuantizerEMA %dx64 + decoder, MSE + vq_loss, '
                        'AMSGrad lr 2e-4), synthetic MFCC-39, T=%d' % (args.codes, args.frames),
            'per_gpu_batch': b, 'global_batch': b * world, 'frames': args.frames, 'decay': args.decay,
            'use_jitter': bool(getattr(args, 'jitter', False)),
            'parallelism': 'dp%d' % world, 'gemm_engine': getattr(args, 'precision', None),
            'dp_exchange': getattr(args, 'dp_exchange', None),
            'l2': 'no explicit flush: every step streams weights + optimizer state + activations >> 126 MB L2'}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def vq_bench(dev, pk, rows, K=44, D=64, iters=20):
    """BASELINE.json's first metric: VQ rows/s for fused fwd + bwd + EMA, N rows resident in HBM (1 GiB at N = 2^22, far
    beyond L2).  Kernels: vqs_vq_assign, vqs_vq_ema_update, vqs_vq_quantize, vqs_vq_backward, each bracketed by events."""
    import torch
    from vq_vae_speech_b200 import ops, LAYOUT_BDT_AS_DTB, LAYOUT_FLAT_ND
    out = {}
    T = 128
    B = rows // T
    N = B * T
    gen = torch.Generator(device=dev).manual_seed(7)
    W0 = torch.randn(K, D, device=dev, generator=gen)
    for lname, layout, shape in (('bdt', LAYOUT_BDT_AS_DTB, (B, D, T)), ('flat', LAYOUT_FLAT_ND, (N, D))):
        z = torch.randn(*shape, device=dev, generator=gen)
        g = torch.randn(*shape, device=dev, generator=gen)
        W = W0.clone()
        cs = torch.zeros(K, device=dev)
        ew = torch.randn(K, D, device=dev, generator=gen)
        ws = ops.vq_workspace(K, D, dev)
        idx = torch.empty(N, dtype=torch.int64, device=dev)
        stats = torch.empty(K * (D + 1), device=dev)
        q = torch.empty_like(z)
        gz = torch.empty_like(z)
        sc = torch.zeros(8, device=dev)
        one = torch.ones(1, device=dev)
        names = ['assign', 'ema_update', 'quantize', 'backward']
        acc = dict((n, 0.0) for n in names)

        def once(evs):
            def br(name, f):
                if evs is None:
                    f()
                    return
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                f()
                e1.record()
                evs.append((name, e0, e1))
            # the bottleneck as the captured training step runs it: search + statistics, EMA update, gather-only forward
            # (q = W_new[idx], z not read again), backward with the forward's losses formed in the same sweep over z
            br('assign', lambda: ops.vq_assign(z, W, layout, ws, idx=idx, stats=stats))
            br('ema_update', lambda: ops.vq_ema_update(cs, ew, W, stats, 0.99, 1e-5))
            br('quantize', lambda: ops.vq_gather(idx, W, layout, shape, out=q))
            br('backward', lambda: ops.vq_backward_loss(g, one, 2 * 0.25 / (N * D), z, idx, W, layout, ws, stats[:K], N, 0.25,
                                                        out=gz, scalars=sc))
        for _ in range(3):
            once(None)
        torch.cuda.synchronize()
        evs = []
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            once(evs)
        t1.record()
        torch.cuda.synchronize()
        total_ms = t0.elapsed_time(t1) / iters
        for n, e0, e1 in evs:
            acc[n] += e0.elapsed_time(e1) / iters
        alg_bytes = (20 * D + 16) * N                       # fwd 8D+8, bwd 12D+8 per row
        out[lname] = {'rows': N, 'rows_per_s': N / (total_ms * 1e-3), 'ms': total_ms,
                      'kernel_ms': dict((k, round(v, 4)) for k, v in acc.items()),
                      'algorithmic_gbs': alg_bytes / (total_ms * 1e-3) / 1e9,
                      'frac_of_hbm_peak': alg_bytes / (total_ms * 1e-3) / 1e9 / pk['hbm']}
        del z, g, q, gz
    out['peak_gbs'] = pk['hbm']
    out['peak_source'] = pk['src']
    out['bytes_per_row'] = 20 * D + 16
    out['K'], out['D'] = K, D
    out['passes'] = ('assign (reads z, writes idx + statistics) | ema_update | quantize = gather-only forward (reads idx, '
                     'writes q) | backward (reads g, z, idx; writes grad_z; forms the losses): z is read twice in total')
    out['search_engine'] = ('auto: streaming engine (row tiles as raw tf32 tcgen05 operands + exact fp32 settlement, '
                            'identical indices): flat rows arrive by TMA, (B,D,T) rows (B % 64 == 0) by a cp.async gather '
                            'into the same swizzled layout; element-wise kernels tiled with indices staged in smem')
    return out


def vq_large_bench(dev, K=4096, D=64, N=1 << 20, iters=10):
    """BASELINE configs[3] (mfcc39-codebook_sizes sweep): the nearest-code search at K = 4096 on the streamed tcgen05 distance GEMM
    (vq_search_large_kernel: search + index store + statistics), N = 2^20 flat rows.  'mma_tflops' counts what the tensor pipe
    executes (3 tf32 MMAs per product + the |e|^2 k-step); the ncu tensor-pipe counter is in profiles/r04k_ncu_k4096.txt."""
    import torch
    from vq_vae_speech_b200 import ops, LAYOUT_FLAT_ND
    gen = torch.Generator(device=dev).manual_seed(K)
    W = torch.randn(K, D, device=dev, generator=gen)
    z = torch.randn(N, D, device=dev, generator=gen)
    ws = ops.vq_workspace(K, D, dev)
    idx = torch.empty(N, dtype=torch.int64, device=dev)
    st = torch.empty(K * (D + 1), device=dev)
    for _ in range(3):
        ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.vq_assign(z, W, LAYOUT_FLAT_ND, ws, idx=idx, stats=st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = 2.0 * N * D * K
    return {'K': K, 'D': D, 'rows': N, 'ms': ms, 'rows_per_s': N / (ms * 1e-3), 'algorithmic_tflops': alg / ms / 1e9,
            'mma_tflops': alg * 3 * (D + 8) / D / ms / 1e9, 'counts_sum_equals_rows': bool(float(st[:K].sum()) == N),
            'engine': 'persistent tcgen05 3xTF32 search, exact fp32 settlement of near-ties (indices identical to the fp32 search)'}


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl b200 needs a CUDA device: the product has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        from vq_vae_speech_b200 import parallel
        parallel.init_nccl(dev, parallel.NCCL_MAX_CTAS if args.nccl_max_ctas is None else args.nccl_max_ctas)
    from vq_vae_speech_b200 import _lib, ops
    from vq_vae_speech_b200.convolutional_vq_vae import ConvolutionalVQVAE
    from vq_vae_speech_b200.trainer import FusedTrainStep
    pk = peaks()
    cfg = model_config(args)
    torch.manual_seed(1234)                 # same weights on every rank (replicated model)
    model = ConvolutionalVQVAE(cfg, dev).to(dev).train()
    B, T = args.batch, args.frames
    eng = FusedTrainStep(model, B, T, cfg['learning_rate'], precision=args.precision)
    args.dp_exchange = None if world == 1 else (
        'own kernels over NVLS multicast / peer memory (csrc/dp_nvls.cu): statistics summed through peer pointers, gradient '
        'allreduce folded into a sharded AMSGrad kernel (multimem.ld_reduce / multimem.st)' if eng.nvls is not None else
        'NCCL allreduces captured into the CUDA graph (4 gradient buckets + statistics)')
    gen = torch.Generator().manual_seed(1234 + rank)        # each rank trains on its own shard of the global batch
    host = [torch.randn(B, T, 39, generator=gen).pin_memory() for _ in range(8)]
    devb = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (>= 3): first step eager, later ones through the CUDA graph when single-GPU ----
    W = max(args.warmup, 3)
    for i in range(W):
        eng.load_batch(devb[i % 8])
        eng.step()
    eng.losses()
    # ---- timed region: K steps of the product path (FusedTrainStep.step(): one CUDA-graph replay per step; under data
    # parallelism the NCCL allreduces are captured into the same graph), inputs resident in HBM ----
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    lc0 = _lib.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    for i in range(args.steps):
        eng.load_batch(devb[i % 8])
        eng.step()
    t1.record()
    barrier()
    w1 = time.perf_counter()
    ms = t0.elapsed_time(t1)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = sampler.stop(w0, w1) if sampler else None
    ms_per_step = ms / args.steps
    value = B * world * args.steps / (ms * 1e-3)
    # kernels launched inside the timed region: a graph replay re-issues every recorded launch without passing through
    # the C ABI counter, so the count is the schedule length (+1 for the step-counter kernel of the optimizer)
    launches = args.steps * (eng.n_launch_calls + 1) if eng.use_graph else _lib.launch_count() - lc0

    # ---- instrumented region: the SAME K steps replayed launch by launch, every launch bracketed by a pair of CUDA
    # events on the launch stream -> per-kernel durations for the roofline (events cannot be placed inside a graph) ----
    barrier()
    events = []
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for i in range(args.steps):
        eng.load_batch(devb[i % 8])
        ops.replay(eng.schedule, events)
        eng.steps_done += 1
    i1.record()
    barrier()
    instr_ms = i0.elapsed_time(i1) / args.steps

    # ---- per-kernel accounting of the instrumented region ----
    fam = {}
    per_launch = {}
    for i, e0, e1 in events:
        entry = eng.schedule[i]
        name = entry[0].__name__
        dt = e0.elapsed_time(e1)
        f = fam.setdefault(name, [0.0, 0.0, 0])
        f[0] += dt
        f[1] += flops_of(entry)
        f[2] += 1
        if flops_of(entry) > 0:
            pl = per_launch.setdefault(i, [0.0, flops_of(entry), entry])
            pl[0] += dt
    kern_ms = sum(v[0] for v in fam.values())
    dom = max(fam.items(), key=lambda kv: kv[1][0])
    dname, (dms, dflops, dn) = dom
    achieved = dflops / (dms * 1e-3) / 1e12
    roofline = {'bound': 'tensor', 'kernel': dname, 'achieved': achieved, 'peak': pk['tensor'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tensor'], 'traffic': measured_traffic(dname),
                'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/traffic.json)',
                'peak_source': pk['src'] + ' bf16 sustained',
                'launches_per_step': dn // args.steps, 'avg_launch_ms': dms / dn,
                'share_of_step_kernel_time': dms / kern_ms, 'measured_in': 'instrumented replay of the same %d steps '
                '(%.3f ms/step with per-launch events vs %.3f ms/step in the timed region)' % (args.steps, instr_ms,
                                                                                              ms_per_step),
                'note': {'fp32': 'exact-fp32 CUDA-core implicit GEMM', '3xtf32': 'tcgen05 kind::tf32, 3 MMAs per product '
                         '(fp32-accurate split, 1e-5 parity)', 'tf32': 'tcgen05 kind::tf32 single pass'}[args.precision]
                + '; achieved counts ALGORITHMIC FLOPs = 2*M*Cred*k*B*L per launch (the split\'s extra MMAs are not counted)'}
    if args.precision == '3xtf32':
        # context for the fraction above (not part of the contract): what the tensor pipe executes is three kind::tf32 MMAs per
        # product, and kind::tf32 runs at half the bf16 rate; the ncu counters of the same binary are in profiles/
        roofline['context'] = {
            'executed_tf32_tflops': 3.0 * achieved, 'tf32_peak_tflops': pk['tensor'] / 2.0,
            'executed_frac_of_tf32_peak': 3.0 * achieved / (pk['tensor'] / 2.0),
            'ncu': 'profiles/r04r_ncu_gemm_family.txt: sm__pipe_tensor_cycles_active 75 % (768x768x3 conv) / 52 % (wgrad) of the '
                   'active cycles, SM clock 1.68-1.77 GHz under this load (power cap)'}
    breakdown = dict((k, {'ms_per_step': v[0] / args.steps, 'launches_per_step': v[2] // args.steps,
                          'tflops': (v[1] / (v[0] * 1e-3) / 1e12) if v[1] else None}) for k, v in fam.items())
    gemm_launches = []
    for i, (dt, fl, entry) in sorted(per_launch.items()):
        d = entry[2]
        if entry[0].__name__ == 'vqs_conv_gemm':
            shape = 'conv M=%d Cred=%d k=%d N=%d ldiv=%d' % (d.M, d.Cred, d.ksz, d.B * d.Lout, d.l_div)
        else:
            shape = 'wgrad M=%d Nw=%d Kred=%d' % (d.M, d.Cred * d.ksz, d.B * d.La)
        gemm_launches.append({'shape': shape, 'ms': round(dt / args.steps, 4),
                              'tflops': round(fl / (dt / args.steps * 1e-3) / 1e12, 1)})
    graph_ms = ms_per_step if eng.use_graph else None

    # ---- e2e: host (pinned) batches -> H2D -> step -> D2H of the losses, every step ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(args.steps):
        eng.step(host[i % 8])          # H2D copy from pinned memory + the step
        last = eng.losses()            # 32-byte D2H + stream sync
    e1.record()
    barrier()
    ems = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ems], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ems = float(tms.item())
    e2e = {'value': B * world * args.steps / (ems * 1e-3), 'unit': 'utterances/s',
           'h2d_bytes_per_step': B * T * 39 * 4, 'd2h_bytes_per_step': 32, 'ms_per_step': ems / args.steps,
           'last_losses': last}

    parity = dp_parity_check(eng, devb, world) if world > 1 else None
    if rank != 0:
        finish(world)
        return
    vq = None
    if not args.skip_vq and world == 1:
        vq = vq_bench(dev, pk, args.vq_rows, K=args.codes)
        vq['k4096'] = vq_large_bench(dev)
    eager = None
    if not args.skip_eager and not args.skip_cpu and world == 1:
        eager = gpu_eager_baseline(cfg, B, T, dev)
    cpu = None
    if not args.skip_cpu and world == 1:
        r = cpu_train_throughput(cfg, B, T, args.cpu_seconds)
        cpu = {'value': r['value'], 'unit': 'utterances/s', 'cores': r['threads'], 'kind': r['kind'],
               'sample': cpu_sample_text(r['kind'], r, B, T)}
        r1 = cpu_train_throughput(cfg, B, T, 0.0, warmup=1, fixed_steps=2, threads=1)     # also one thread
        cpu['single_thread'] = {'value': r1['value'], 'unit': 'utterances/s', 'cores': 1,
                                'sample': cpu_sample_text(r1['kind'], r1, B, T)}
    line = {
        'metric': 'train_utterances_per_sec', 'value': value, 'unit': 'utterances/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': W, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, world),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
        'ms_per_step_cuda_graph': graph_ms, 'ms_per_step_instrumented': instr_ms, 'kernel_breakdown': breakdown,
        'gemm_launches': gemm_launches, 'vq': vq, 'gpu_eager_baseline': eager, 'parity_check': parity,
        'params': eng.n_params, 'flops_per_step': sum(flops_of(e) for e in eng.schedule if e[0] is not None),
    }
    emit(line)
    finish(world)


def dp_parity_check(eng, devb, world):
    """Self-check of the data-parallel path that was just timed (N > 1; every rank takes part):
      * the replicated state -- flat parameters, AMSGrad moments, codebook, EMA state -- is BIT-identical on all ranks after
        the timed steps (DataParallelComm.assert_replicated: MIN- and MAX-allreduce agree);
      * one more step from the same state gives bit-identical parameters whether it is replayed from the captured CUDA graph
        (the exchange kernels / NCCL allreduces inside the graph) or issued launch by launch -- the launch-by-launch path is
        the one the 2-rank oracle tests pin to the per-shard contract (tests/test_parallel_gpu.py);
      * every rank's losses are finite and the EMA statistics of the step sum to world x rows (global counts).
    Raises on failure; returns the summary that goes into the JSON line."""
    import torch
    import torch.distributed as dist
    comm, vq = eng.comm, eng.model._vq
    replicated = [eng.flat_p, eng.opt_step]
    if eng.is_ema:
        replicated += [vq._embedding.weight.data, vq._ema_w.data, vq._ema_cluster_size]
    moments = [eng.flat_m, eng.flat_v, eng.flat_vmax]
    if getattr(eng, 'nvls', None) is None:
        replicated += moments                     # (with the NVLS exchange the AMSGrad moments are sharded over the ranks)
    state = replicated + (moments if getattr(eng, 'nvls', None) is not None else [])
    for i, t in enumerate(replicated):
        comm.assert_replicated(t.float() if t.dtype != torch.float32 else t, 'replicated state %d' % i)
    saved = [t.clone() for t in state]
    eng.load_batch(devb[0])
    eng.step()                                    # (a) the product path: graph replay when the graph is in use
    via_graph = eng.graph is not None
    a = [t.clone() for t in state]
    la = eng.losses()
    counts = eng.buf['stats'][:eng.dims['K']].sum().item()
    for t, sv in zip(state, saved):
        t.copy_(sv)
    eng.load_batch(devb[0])
    eng._run_schedule()                           # (b) the same schedule, launch by launch, eager NCCL
    lb = eng.losses()
    same = all(torch.equal(x, y) for x, y in zip(a, state))
    flag = torch.tensor([1.0 if same else 0.0, 1.0 if all(map(lambda v: v == v and abs(v) < 1e30, la.values())) else 0.0],
                        device=eng.dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_same, ok_finite = bool(flag[0].item() == 1.0), bool(flag[1].item() == 1.0)
    rows = eng.B * eng.dims['Tq'] * (world if eng.is_ema else 1)     # the statistics are allreduced for the EMA update only
    out = {'replicated_state_bit_identical': True, 'graph_replay_equals_eager_nccl': ok_same, 'via_cuda_graph': via_graph,
           'losses_finite_all_ranks': ok_finite, 'global_counts': counts, 'global_rows': rows,
           'rank0_losses_graph': la, 'rank0_losses_eager': lb}
    if not (ok_same and ok_finite and counts == rows):
        raise RuntimeError('data-parallel parity check failed: %r' % (out,))
    for i, t in enumerate(replicated):
        comm.assert_replicated(t.float() if t.dtype != torch.float32 else t, 'replicated state %d after the check' % i)
    out['exchange'] = 'nvls' if getattr(eng, 'nvls', None) is not None else 'nccl'
    return out


_REAL_STDOUT = None


def finish(world):
    """Multi-rank runs leave through os._exit once every rank is done: tearing down a process group whose collectives
    were captured into a CUDA graph can block in destroy_process_group (seen on 2 GPUs: the result was printed, then the
    processes hung until the launcher's timeout)."""
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stderr.flush()
        os._exit(0)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (e.g. NCCL's version banner) was
    redirected to stderr at start-up."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


if __name__ == '__main__':
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)
