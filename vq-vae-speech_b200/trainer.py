"""Fused, data-parallel training step: the B200-native replacement of ConvolutionalTrainer.iterate
(/root/reference/src/experiments/convolutional_trainer.py:44-74: zero_grad -> model -> MSE + vq_loss -> backward ->
Adam(lr, amsgrad=True).step()).

Instead of going through autograd, the step is a hand-scheduled sequence of libvqs_b200 launches over preallocated
buffers (activations, one flat parameter buffer, one flat gradient buffer, flat AMSGrad state).  The sequence is recorded
once (ops.set_recorder) and then either replayed launch by launch or captured into a CUDA graph, so a training step costs
one graph launch on the host.  Under data parallelism (one process per GPU, torch.distributed / NCCL) each rank runs the
step on its batch shard; the EMA statistics [counts | dw] are sum-allreduced between the assignment and the EMA update,
and the flat gradient buffer is allreduced in four buckets launched as the backward pass completes them (only the
last, smallest one -- encoder conv_1..3, 14 MB -- is not hidden under backward compute).

Parity contract: per rank, the step equals the reference step on that rank's shard with the same
codebook; EMA statistics are the sum over shards; gradients are the average over shards.
"""
import os

import numpy as np
import torch

from . import functional as F
from . import ops
from ._lib import LAYOUT_BDT_AS_DTB
from .modules import jitter_plan
from .parallel import DataParallelComm
from .ops import MASK_FLOAT, MASK_U8
from .vector_quantizer import VectorQuantizerEMA

_ALIGN = 64  # floats; every parameter starts 256-byte aligned inside the flat buffers


class FusedTrainStep(object):
    """Owns the flat parameter / gradient / optimizer buffers of `model` (a vq_vae_speech_b200 ConvolutionalVQVAE on a
    CUDA device; its parameters are re-pointed at views of the flat buffer, state_dict() keeps working) and runs training
    steps for a fixed (batch_size, T) shape."""

    def __init__(self, model, batch_size, num_frames, learning_rate, use_graph=True, process_group=None,
                 betas=(0.9, 0.999), eps=1e-8, precision='3xtf32', separate_target=None):
        """precision: GEMM engine of the conv / wgrad GEMMs -- '3xtf32' (tcgen05, fp32-accurate split; default), 'tf32'
        (tcgen05 single pass: cuDNN's default numerics for the reference on a GPU) or 'fp32' (exact FMA on CUDA cores).
        separate_target: the reference trains against data['output_features'] (convolutional_trainer.py:47), which every
        shipped experiment sets equal to the input features; True = step() takes its own target batch (B, T, F_out), False
        = the reconstruction target is the input batch (one copy and one layout pass fewer per step).  Default: True only
        when the model's output and input feature counts differ."""
        self.model = model
        self.precision = precision
        self.B, self.T = int(batch_size), int(num_frames)
        self.lr, self.betas, self.eps = float(learning_rate), betas, float(eps)
        self.comm = DataParallelComm(process_group)
        self.world = self.comm.world
        self.dev = next(model.parameters()).device
        if self.dev.type != 'cuda':
            raise RuntimeError('FusedTrainStep needs the model on a CUDA device (no CPU path)')
        enc, dec, vq = model._encoder, model._decoder, model._vq
        # use_speaker_conditioning (deconvolutional_decoder.py:108-111): the reference draws a fresh random speaker
        # embedding on the host RNG every forward (global_conditioning.py:34); step() draws the same one and uploads the
        # rows of this batch into a static buffer, like the jitter plan -- the captured schedule only reads that buffer
        self.use_speaker = bool(getattr(dec, '_use_speaker_conditioning', False))
        self.separate_target = separate_target
        self.is_ema = isinstance(vq, VectorQuantizerEMA)
        self.nl = enc._residual_stack._num_residual_layers
        if self.nl < 1:
            raise NotImplementedError('num_residual_layers must be >= 1')
        self.use_jitter = bool(dec._use_jitter)
        self._flatten_parameters()
        self._alloc_buffers()
        self.schedule = []
        prev = ops.set_recorder(self.schedule)
        prev_prec = ops.set_precision(precision)
        try:
            self._emit_step()
        finally:
            ops.set_recorder(prev)
            ops.set_precision(prev_prec)
        self.n_launch_calls = sum(1 for f, _, _ in self.schedule if f is not None)
        self.graph = None
        # data parallel: the NCCL allreduces are captured into the graph too (torch.distributed supports capture); set
        # VQS_DP_GRAPH=0 to replay launch by launch instead
        import os
        self.use_graph = bool(use_graph) and (self.world == 1 or os.environ.get('VQS_DP_GRAPH', '1') != '0')
        self.steps_done = 0

    # ------------------------------------------------------------------------------------------------
    def _trainable(self):
        seen, out = set(), []
        for name, p in self.model.named_parameters():      # named_parameters() de-duplicates the shared Residual
            if id(p) in seen:
                continue
            seen.add(id(p))
            if self.is_ema and name.startswith('_vq.'):
                continue       # EMA codebook / ema_w never receive a gradient; Adam skips them
            out.append((name, p))
        return out

    def _flatten_parameters(self):
        params = self._trainable()
        offs, total = [], 0
        for _, p in params:
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        dev = self.dev
        self.nvls = self.comm.nvls if self.world > 1 else None
        # Default: the whole exchange after the backward pass (vqs_dp_amsgrad_step).  VQS_DP_OVERLAP=1: bucket-wise exchange
        # BESIDE the backward pass -- as soon as a bucket's gradients are final on the main stream, a side stream runs barrier +
        # sharded AMSGrad with the reduce / broadcast through the switch for that bucket (vqs_dp_amsgrad_range: light kernels
        # that fit next to the GEMM CTAs); only the last, small bucket and the exit barrier remain after the backward pass.
        # Built, parity-green, and measured SLOWER at 2 GPUs (2.378 ms per step, 2.448 with the max-smem carve-out hint, against
        # 2.371 for the default and 2.227 for two independent replicas on the same box): what the exchange takes from the
        # GEMMs it runs beside costs more than hiding it gains.  Kept as an option.
        self.dp_overlap = self.nvls is not None and os.environ.get('VQS_DP_OVERLAP', '0') == '1'
        self.side = torch.cuda.Stream(device=self.dev, priority=-1) if self.dp_overlap else None   # high priority: its small blocks go first
        if self.nvls is not None:
            # symmetric buffers: the optimizer kernel reads the gradient SUM of all GPUs through the NVLS multicast mapping of
            # flat_g and broadcasts the new parameters through the one of flat_p (csrc/dp_nvls.cu)
            self.flat_p, self.mc_p, _ = self.nvls.symmetric_zeros(total)
            self.flat_g, self.mc_g, _ = self.nvls.symmetric_zeros(total)
        else:
            self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
            self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_vmax = torch.zeros(total, dtype=torch.float32, device=dev)
        self.opt_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.grads, self.param_names = {}, []
        self.param_offsets = {}          # name -> offset of the parameter in the flat buffers (p, g, m, v, vmax)
        first_decoder = None
        with torch.no_grad():
            for (name, p), off in zip(params, offs):
                view = self.flat_p[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                self.grads[name] = self.flat_g[off:off + p.numel()].view_as(p)
                self.param_names.append(name)
                self.param_offsets[name] = off
                if first_decoder is None and name.startswith('_decoder.'):
                    first_decoder = off
        # weight-normalised convs (use_kaiming_normal): the optimizer owns weight_g / weight_v; the effective
        # weight w = g v / ||v|| and its gradient dW live in scratch buffers the GEMMs use under the plain '.weight' name
        self.wn = {}
        self.weight_names = [n for n in self.param_names if n.endswith('.weight')]
        offs_by_name = dict((n, o) for (n, _), o in zip(params, offs))
        for name in self.param_names:
            if name.endswith('.weight_v'):
                base = name[:-2]
                v = self._p(name)
                self.wn[base] = dict(v=v, g=self._p(base + '_g'), w=torch.empty_like(v), dw=torch.zeros_like(v),
                                     norm=torch.empty(v.shape[0], dtype=torch.float32, device=dev), off=offs_by_name[name])
                self.grads[base] = self.wn[base]['dw']
                self.weight_names.append(base)
        self.bucket_split = first_decoder if first_decoder is not None else total
        # gradient allreduce buckets, in the order the backward pass completes them (flat order is encoder, pre_vq, [vq],
        # decoder): decoder transposed convs | rest of the decoder | encoder conv_4 .. pre_vq (+ codebook) | conv_1 .. conv_3.
        # VQS_DP_FINE=1 sends conv_3, conv_2 and conv_1 one by one instead (only conv_1's 0.4 MB then travel after the
        # backward pass has ended); measured at 2 GPUs that is 0.3 % SLOWER (3.435 vs 3.424 ms: two more collectives for
        # 17 MB that NVLink moves in ~25 us), at 4 / 8 GPUs it is unmeasured -- hence opt-in
        def first_of(prefix, default):
            o = [off for n, off in offs_by_name.items() if n.startswith(prefix)]
            return min(o) if o else default

        cut_t = first_of('_decoder._conv_trans_1.', total)
        cut_e = first_of('_encoder._conv_4.', 0)
        cut_3 = min(first_of('_encoder._conv_3.', 0), cut_e)
        cut_2 = min(first_of('_encoder._conv_2.', 0), cut_3)
        # (the bucket-wise NVLS exchange wants the fine split: whatever is in the LAST bucket travels after the backward pass)
        fine = os.environ.get('VQS_DP_FINE', '1' if self.dp_overlap else '0') == '1'
        if not fine:                                     # NCCL default: conv_1 .. conv_3 as ONE trailing bucket
            cut_3 = cut_2 = cut_e
        self.buckets = {'dec_convT': (cut_t, total), 'dec_rest': (self.bucket_split, cut_t),
                        'enc_hi': (cut_e, self.bucket_split), 'enc_c3': (cut_3, cut_e), 'enc_c2': (cut_2, cut_3),
                        'enc_c1': (0, cut_2)}
        self.n_params = sum(p.numel() for _, p in params)

    def _p(self, name):
        if name in getattr(self, 'wn', {}):
            return self.wn[name]['w']           # effective weight of a weight-normalised conv
        return dict(self.model.named_parameters())[name].data

    def _emit_wn_fold(self, lo, hi):
        """dW of the weight-normalised convs whose parameters lie in flat range [lo, hi) -> gradients of g and v."""
        for base, w in self.wn.items():
            if lo <= w['off'] < hi:
                ops.weight_norm_bwd(w['dw'], w['v'], w['g'], w['norm'], self.grads[base + '_v'], self.grads[base + '_g'])

    def _alloc_buffers(self):
        m = self.model
        B, T, dev = self.B, self.T, self.dev
        Fi = m._encoder._conv_1.in_channels
        C = m._encoder._conv_1.out_channels
        R_enc = m._encoder._residual_stack._layers[0]._block[1].out_channels
        R_dec = m._decoder._residual_stack._layers[0]._block[1].out_channels
        D = m._pre_vq_conv.out_channels
        K = m._vq._num_embeddings
        Fo = m._decoder._conv_trans_3.out_channels
        Tq = T // 2 + 1
        L2 = 2 * Tq
        self.dims = dict(B=B, T=T, Fi=Fi, C=C, R_enc=R_enc, R_dec=R_dec, D=D, K=K, Fo=Fo, Tq=Tq, L2=L2)
        if self.separate_target is None:
            self.separate_target = Fo != Fi
        if Fo != Fi and not self.separate_target:
            raise ValueError('output features (%d) differ from input features (%d): the step needs separate_target=True' % (Fo, Fi))
        Dd = m._decoder._conv_1.in_channels           # D, + 40 speaker features when use_speaker_conditioning
        self.dims['Dd'] = Dd
        if L2 + 3 < T:
            raise RuntimeError('decoder output shorter than the input')
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        u8 = lambda *s: torch.empty(*s, dtype=torch.uint8, device=dev)
        b = self.buf = {}
        b['x_in'] = f(B, T, Fi)                  # the (B, T, F) feature batch as the loader delivers it
        b['x'] = f(B, Fi, T)
        if self.separate_target:
            b['t_in'], b['target'] = f(B, T, Fo), f(B, Fo, T)
        if self.use_speaker:
            b['spk'] = torch.zeros(B, Dd - D, dtype=torch.float32, device=dev)
            b['qcat'], b['gqcat'] = f(B, Dd, Tq), f(B, Dd, Tq)
            self.host_spk = torch.zeros(B, Dd - D, dtype=torch.float32).pin_memory()
        b['a1'], b['h2'], b['m2'] = f(B, C, T), f(B, C, T), u8(B, C, T)
        b['a3'], b['h4'], b['m4'], b['h5'], b['m5'] = f(B, C, Tq), f(B, C, Tq), u8(B, C, Tq), f(B, C, Tq), u8(B, C, Tq)
        for i in range(self.nl):
            b['e_hh%d' % i] = f(B, R_enc, Tq)
            if i > 0:
                b['e_x%d' % i] = f(B, C, Tq)
        b['enc_out'], b['m_e'] = f(B, C, Tq), u8(B, C, Tq)
        b['z'], b['q'], b['qj'] = f(B, D, Tq), f(B, D, Tq), f(B, D, Tq)
        b['idx'] = torch.empty(B * Tq, dtype=torch.int64, device=dev)
        if self.nvls is not None and self.is_ema:
            # the local statistics live in a symmetric buffer (the peers read them), their sum over ranks in a local one
            b['stats_local'], _, self.stats_ptrs = self.nvls.symmetric_zeros((K * (D + 1) + 63) // 64 * 64)
            b['stats_local'] = b['stats_local'][:K * (D + 1)]
            b['stats'] = f(K * (D + 1))
        else:
            b['stats'] = f(K * (D + 1))
            b['stats_local'] = b['stats']
        b['vq_scalars'] = torch.zeros(8, dtype=torch.float32, device=dev)
        b['jitter_src'] = torch.arange(Tq, dtype=torch.int32, device=dev)
        b['d1'], b['u'] = f(B, C, Tq), f(B, C, L2)
        for i in range(self.nl):
            b['d_hh%d' % i] = f(B, R_dec, L2)
            if i > 0:
                b['d_x%d' % i] = f(B, C, L2)
        # the reconstruction is trimmed to T frames (vq_vae.py:133-137) and conv_trans_3 has kernel 2: only the first
        # min(L2 + 2, T) positions of conv_trans_2's output are ever read -- the rest is neither computed nor back-propagated
        # (N = 64 x 50 = 3200 columns were 150 tiles = two waves on 148 SMs; 64 x 47 are 144)
        Lt2 = min(L2 + 2, T)
        self.dims['Lt2'] = Lt2
        b['s'], b['t1'], b['t2'] = f(B, C, L2), f(B, C, L2), f(B, C, Lt2)
        b['recon'], b['g_recon'] = f(B, Fo, T), f(B, Fo, T)
        # slot 5 of the VQ scalars (its kernels write [0, 5)): all losses of a step sit in ONE 32-byte vector, read back by one copy
        b['recon_loss'] = b['vq_scalars'][5:6]
        b['one'] = torch.ones(1, dtype=torch.float32, device=dev)
        # backward scratch: two ping-pong gradient buffers per resolution + hidden-gradient buffers
        b['gA2'], b['gB2'] = f(B, C, L2 + 2), f(B, C, L2 + 2)
        b['gH2'] = f(B, max(R_dec, C), L2)
        b['gA1'], b['gB1'], b['gC1'] = f(B, C, Tq), f(B, C, Tq), f(B, C, Tq)
        b['gH1'] = f(B, max(R_enc, C), Tq)
        b['gT_a'], b['gT_b'] = f(B, C, T), f(B, C, T)
        b['gq'], b['gqj'], b['gz'] = f(B, D, Tq), f(B, D, Tq), f(B, D, Tq)
        # GEMM-ready weight operands, rebuilt at the top of every step by vqs_permute_weight: (name, role) ->
        # (buffer or None when the parameter is used as is, tap flag, permute mode)
        self.wperm = {}
        for name in self.weight_names:
            if name.startswith('_vq.'):
                continue
            p = self._p(name)
            roles = ('convT_fwd', 'convT_dgrad') if '_conv_trans_' in name else ('conv_fwd', 'conv_dgrad')
            for role in roles:
                if name == '_encoder._conv_1.weight' and role == 'conv_dgrad':
                    continue                       # the input features need no gradient
                tap, mode, numel = F.gemm_weight_layout(tuple(p.shape), role, self.precision)
                self.wperm[(name, role)] = (f(numel) if mode is not None else None, tap, mode)
        ws_bytes = 16
        for (M, Cr, k, La) in [(C, Fi, 3, T), (C, C, 3, T), (C, C, 4, Tq), (C, C, 3, Tq), (R_enc, C, 3, Tq),
                               (C, R_enc, 1, Tq), (D, C, 3, Tq), (C, Dd, 3, Tq), (R_dec, C, 3, L2), (C, R_dec, 1, L2),
                               (C, C, 3, L2), (C, C, 3, L2), (C, Fo, 2, L2 + 2)]:
            ws_bytes = max(ws_bytes, ops.wgrad_workspace_bytes(M, Cr, k, B, La))
        self.ws_wgrad = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        self.ws_vq = ops.vq_workspace(K, D, dev)
        self.ws_mse = ops.mse_workspace(dev)
        self.host_scalars = torch.zeros(8, dtype=torch.float32).pin_memory()

    # ------------------------------------------------------------------------------------------------
    # the step, emitted once through the recorder
    # ------------------------------------------------------------------------------------------------
    def _view(self, name, C, L):
        """(B, C, L) view at the head of a scratch buffer."""
        return self.buf[name].view(-1)[:self.B * C * L].view(self.B, C, L)

    def _allreduce_stats(self):
        self.comm.allreduce_stats(self.buf['stats'])

    def _allreduce_bucket(self, lo, hi):
        self.comm.start_bucket(self.flat_g, lo, hi)

    def _wait_buckets(self):
        self.comm.wait_buckets()

    def _nvls_bucket(self, name, first=False, last=False):
        """Host callable of the schedule: the gradients of bucket `name` are final on the current (main) stream -> fork to the
        side stream, which runs barrier + optimizer + exchange of that bucket beside the rest of the backward pass; the last
        bucket also runs the exit barrier and joins the side stream back.  Capturable (the fork / join become graph edges)."""
        lo, hi = self.buckets[name]
        if hi <= lo and not first and not last:
            return
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        ops.dp_amsgrad_range_on(self.side, self.nvls.ctx, self.mc_p, self.flat_p, self.mc_g, self.flat_m, self.flat_v,
                                self.flat_vmax, lo, hi, self.opt_step, self.lr, self.betas[0], self.betas[1], self.eps,
                                inc_step=first, ch_before=1, ch_after=2 if last else -1)
        if last:
            main.wait_stream(self.side)

    def _emit_step(self):
        m, b, d = self.model, self.buf, self.dims
        B, T, Tq, L2, C, D, K, Fi, Fo = d['B'], d['T'], d['Tq'], d['L2'], d['C'], d['D'], d['K'], d['Fi'], d['Fo']
        Lt2 = d['Lt2']
        E, DEC = '_encoder.', '_decoder.'
        RS1, RS2 = '_residual_stack._layers.0._block.1.weight', '_residual_stack._layers.0._block.3.weight'
        P, G, WP, ws = self._p, self.grads, self.wperm, self.ws_wgrad
        nl = self.nl
        # ---- 0. effective weights of weight-normalised convs, then the GEMM-ready operand images ----
        for base, w in self.wn.items():
            ops.weight_norm_fwd(w['v'], w['g'], w['w'], w['norm'])
        ops.permute_weights([(P(name), buf, mode) for (name, role), (buf, tap, mode) in WP.items() if buf is not None])

        def A(name, role):
            buf, tap, mode = WP[(name, role)]
            M, Cred, k = F._role_dims(tuple(P(name).shape), role)
            return F.GemmW(buf if buf is not None else P(name), tap, M, Cred, k)

        # every conv-like GEMM may use the wgrad scratch for split-K (few-tile and half-wave shapes, see launch_conv_tc)
        def cfwd(xin, name, bias, stride, pad, **kw):
            kw.setdefault('splitk_ws', ws)
            return F.conv1d_forward(xin, A(name, 'conv_fwd'), bias, stride, pad, **kw)

        def cdgrad(gy, name, Lx, stride, pad, **kw):
            kw.setdefault('splitk_ws', ws)
            return F.conv1d_dgrad(gy, A(name, 'conv_dgrad'), Lx, stride, pad, **kw)

        def tfwd(xin, name, bias, pad, **kw):
            kw.setdefault('splitk_ws', ws)
            return F.convT1d_forward(xin, A(name, 'convT_fwd'), bias, pad, **kw)

        def tdgrad(gy, name, Lx, pad, **kw):
            kw.setdefault('splitk_ws', ws)
            return F.convT1d_dgrad(gy, A(name, 'convT_dgrad'), Lx, pad, **kw)

        # ---- 1. encoder forward (convolutional_encoder.py:118-146) ----
        ops.blc_to_ncl(b['x_in'], b['x'])                                            # vq_vae.py:118
        if self.separate_target:
            ops.blc_to_ncl(b['t_in'], b['target'])                                   # trainer.py:47
        cfwd(b['x'], E + '_conv_1.weight', P(E + '_conv_1.bias'), 1, 1, out=b['a1'], relu=True)
        cfwd(b['a1'], E + '_conv_2.weight', P(E + '_conv_2.bias'), 1, 1, out=b['h2'], relu=True,
                         mask_out=b['m2'], add_post=b['a1'])
        cfwd(b['h2'], E + '_conv_3.weight', P(E + '_conv_3.bias'), 2, 2, out=b['a3'], relu=True)
        cfwd(b['a3'], E + '_conv_4.weight', P(E + '_conv_4.bias'), 1, 1, out=b['h4'], relu=True,
                         mask_out=b['m4'], add_post=b['a3'])
        cfwd(b['h4'], E + '_conv_5.weight', P(E + '_conv_5.bias'), 1, 1, out=b['h5'], relu=True,
                         mask_out=b['m5'], add_post=b['h4'])
        # residual stack: x_{i+1} = relu(x_i) + conv2(relu(conv1(relu(x_i)))), then relu, then + h5 (outer skip)
        xs = [b['h5']] + [b['e_x%d' % i] for i in range(1, nl)]
        for i in range(nl):
            cfwd(xs[i], E + RS1, None, 1, 1, out=b['e_hh%d' % i], x_relu=True, relu=True)
            if i + 1 < nl:
                cfwd(b['e_hh%d' % i], E + RS2, None, 1, 0, out=xs[i + 1], add_pre=xs[i],
                                 add_pre_relu=True)
            else:
                cfwd(b['e_hh%d' % i], E + RS2, None, 1, 0, out=b['enc_out'], add_pre=xs[i],
                                 add_pre_relu=True, relu=True, mask_out=b['m_e'], add_post=b['h5'])
        # (M = D = 64 output channels: a single row of tiles -> split-K over the wgrad scratch, see vqs_b200.h)
        cfwd(b['enc_out'], '_pre_vq_conv.weight', P('_pre_vq_conv.bias'), 1, 1, out=b['z'], splitk_ws=ws)

        # ---- 2. VQ bottleneck ----
        vq = m._vq
        cb = vq._embedding.weight.data
        ops.vq_assign(b['z'], cb, LAYOUT_BDT_AS_DTB, self.ws_vq, idx=b['idx'], stats=b['stats_local'])
        n_rows_total = B * Tq
        if self.is_ema:
            if self.world > 1:
                if self.nvls is not None:     # barrier + rank-ordered sum of the peers' vectors: one small kernel of ours
                    ops.dp_allreduce_small(self.nvls.ctx, self.stats_ptrs, b['stats'], channel=0)
                else:
                    ops.record_callable(self._allreduce_stats)
                n_rows_total = self.comm.total_rows(B * Tq)
            ops.vq_ema_update(vq._ema_cluster_size, vq._ema_w.data, cb, b['stats'], vq._decay, vq._epsilon)
        beta = float(vq._commitment_cost)
        # forward value of the bottleneck: q = W_new[idx] as a pure gather-write (z is not read again); the losses are formed by
        # the VQ backward kernel, which reads z, idx and the codebook anyway (vqs_b200.h: vqs_vq_backward_loss)
        ops.vq_gather(b['idx'], cb, LAYOUT_BDT_AS_DTB, (B, D, Tq), out=b['q'])
        self._vq_loss_args = (b['stats'][:K], n_rows_total, beta)

        # ---- 3. decoder forward (deconvolutional_decoder.py:100-137) ----
        dec_in = b['q']
        if self.use_jitter:
            ops.jitter_fwd(b['q'], b['jitter_src'], b['qj'])
            dec_in = b['qj']
        if self.use_speaker:                                                         # decoder.py:108-111
            ops.concat_channels(dec_in, b['spk'], out=b['qcat'])
            dec_in = b['qcat']
        cfwd(dec_in, DEC + '_conv_1.weight', P(DEC + '_conv_1.bias'), 1, 1, out=b['d1'])
        ops.upsample2_fwd(b['d1'], b['u'])
        xd = [b['u']] + [b['d_x%d' % i] for i in range(1, nl)]
        for i in range(nl):
            cfwd(xd[i], DEC + RS1, None, 1, 1, out=b['d_hh%d' % i], x_relu=True, relu=True)
            last = i + 1 == nl
            cfwd(b['d_hh%d' % i], DEC + RS2, None, 1, 0, out=b['s'] if last else xd[i + 1],
                             add_pre=xd[i], add_pre_relu=True, relu=last)
        tfwd(b['s'], DEC + '_conv_trans_1.weight', P(DEC + '_conv_trans_1.bias'), 1, out=b['t1'],
                          relu=True)
        tfwd(b['t1'], DEC + '_conv_trans_2.weight', P(DEC + '_conv_trans_2.bias'), 0, out_len=Lt2, out=b['t2'],
                          relu=True)
        tfwd(b['t2'], DEC + '_conv_trans_3.weight', P(DEC + '_conv_trans_3.bias'), 0, out_len=T,
                          out=b['recon'], splitk_ws=ws)                                          # trimmed to T (vq_vae.py:133-137)

        # ---- 4. loss (trainer.py:54-56): MSE against the input features, gradient in the same pass ----
        ops.mse_fwd_bwd(b['recon'], b['target'] if self.separate_target else b['x'], (Fo * T, T, 1), 1.0, b['recon_loss'],
                        b['g_recon'], self.ws_mse)

        # ---- 5. decoder backward ----
        gq2 = self._view('gA2', C, Lt2)
        F.convT1d_wgrad(b['g_recon'], b['t2'], G[DEC + '_conv_trans_3.weight'], 0, ws)
        ops.bias_grad(b['g_recon'], G[DEC + '_conv_trans_3.bias'])
        tdgrad(b['g_recon'], DEC + '_conv_trans_3.weight', Lt2, 0, out=gq2, mask=b['t2'],
                        mask_kind=MASK_FLOAT)
        gq1 = self._view('gB2', C, L2)
        F.convT1d_wgrad(gq2, b['t1'], G[DEC + '_conv_trans_2.weight'], 0, ws)
        ops.bias_grad(gq2, G[DEC + '_conv_trans_2.bias'])
        tdgrad(gq2, DEC + '_conv_trans_2.weight', L2, 0, out=gq1, mask=b['t1'], mask_kind=MASK_FLOAT)
        g = self._view('gA2', C, L2)
        F.convT1d_wgrad(gq1, b['s'], G[DEC + '_conv_trans_1.weight'], 1, ws)
        ops.bias_grad(gq1, G[DEC + '_conv_trans_1.bias'])
        tdgrad(gq1, DEC + '_conv_trans_1.weight', L2, 1, out=g, mask=b['s'], mask_kind=MASK_FLOAT)
        self._emit_wn_fold(*self.buckets['dec_convT'])
        if self.world > 1 and self.nvls is None:       # the three transposed convs are done: their gradients start travelling now
            ops.record_callable(lambda: self._allreduce_bucket(*self.buckets['dec_convT']))
        if self.dp_overlap:
            ops.record_callable(lambda: self._nvls_bucket('dec_convT', first=True))
        other = self._view('gB2', C, L2)
        R_dec = d['R_dec']
        gh = self._view('gH2', R_dec, L2)
        for n, i in enumerate(reversed(range(nl))):
            acc = n > 0                                      # the shared Residual: second application accumulates
            F.conv1d_wgrad(g, b['d_hh%d' % i], G[DEC + RS2], 1, 0, ws, accumulate=acc)
            cdgrad(g, DEC + RS2, L2, 1, 0, out=gh, mask=b['d_hh%d' % i], mask_kind=MASK_FLOAT)
            F.conv1d_wgrad(gh, xd[i], G[DEC + RS1], 1, 1, ws, x_relu=True, accumulate=acc)
            cdgrad(gh, DEC + RS1, L2, 1, 1, out=other, add_pre=g, mask=xd[i], mask_kind=MASK_FLOAT)
            g, other = other, g
        ops.upsample2_bwd(g, b['gA1'])
        gd1 = b['gA1']
        F.conv1d_wgrad(gd1, dec_in, G[DEC + '_conv_1.weight'], 1, 1, ws)
        ops.bias_grad(gd1, G[DEC + '_conv_1.bias'])
        gq = b['gq']
        g_dec_in = b['gqj'] if self.use_jitter else gq
        if self.use_speaker:       # gradient of the concatenation: its first D channels (the speaker features are constants)
            cdgrad(gd1, DEC + '_conv_1.weight', Tq, 1, 1, out=b['gqcat'], splitk_ws=ws)
            ops.slice_channels(b['gqcat'], D, out=g_dec_in)
        else:
            cdgrad(gd1, DEC + '_conv_1.weight', Tq, 1, 1, out=g_dec_in, splitk_ws=ws)
        if self.use_jitter:
            ops.jitter_bwd(b['gqj'], b['jitter_src'], gq)
        self._emit_wn_fold(*self.buckets['dec_rest'])
        if self.world > 1 and self.nvls is None:       # decoder gradients are complete: allreduce the rest of them under the encoder's backward
            ops.record_callable(lambda: self._allreduce_bucket(*self.buckets['dec_rest']))
        if self.dp_overlap:
            ops.record_callable(lambda: self._nvls_bucket('dec_rest'))

        # ---- 6. VQ backward (autograd of ema.py:165-169 / vector_quantizer.py:136-141), upstream d(loss)/d(vq_loss) = 1
        n_local = B * Tq
        counts, n_rows_total, _ = self._vq_loss_args
        ops.vq_backward_loss(gq, b['one'], 2.0 * beta / (n_local * D), b['z'], b['idx'], cb, LAYOUT_BDT_AS_DTB, self.ws_vq,
                             counts, n_rows_total, beta, out=b['gz'], scalars=b['vq_scalars'])
        if not self.is_ema:
            ops.vq_grad_codebook(b['stats'], cb, b['one'], 2.0 / (n_local * D), out=G['_vq._embedding.weight'])

        # ---- 7. encoder backward ----
        gz = b['gz']
        F.conv1d_wgrad(gz, b['enc_out'], G['_pre_vq_conv.weight'], 1, 1, ws)
        ops.bias_grad(gz, G['_pre_vq_conv.bias'])
        ge, g = b['gC1'], b['gA1']
        cdgrad(gz, '_pre_vq_conv.weight', Tq, 1, 1, out=ge, out2=g, mask2=b['m_e'], mask2_kind=MASK_U8)
        other = b['gB1']
        R_enc = d['R_enc']
        gh = self._view('gH1', R_enc, Tq)
        gp5 = None
        for n, i in enumerate(reversed(range(nl))):
            acc = n > 0
            F.conv1d_wgrad(g, b['e_hh%d' % i], G[E + RS2], 1, 0, ws, accumulate=acc)
            cdgrad(g, E + RS2, Tq, 1, 0, out=gh, mask=b['e_hh%d' % i], mask_kind=MASK_FLOAT)
            F.conv1d_wgrad(gh, xs[i], G[E + RS1], 1, 1, ws, x_relu=True, accumulate=acc)
            if i > 0:
                cdgrad(gh, E + RS1, Tq, 1, 1, out=other, add_pre=g, mask=xs[i], mask_kind=MASK_FLOAT)
                g, other = other, g
            else:
                # gh5 = ((g + dgrad) * (h5 > 0)) + ge ; gp5 = gh5 * m5
                cdgrad(gh, E + RS1, Tq, 1, 1, out=other, add_pre=g, mask=xs[0], mask_kind=MASK_FLOAT,
                               add_post=ge, out2=g, mask2=b['m5'], mask2_kind=MASK_U8)
                gh5, gp5 = other, g
        # conv_5: h5 = relu(p5) + h4
        F.conv1d_wgrad(gp5, b['h4'], G[E + '_conv_5.weight'], 1, 1, ws)
        ops.bias_grad(gp5, G[E + '_conv_5.bias'])
        gh4, gp4 = b['gC1'], self._view('gH1', C, Tq)      # ge (gC1) is dead after gh5 was formed
        cdgrad(gp5, E + '_conv_5.weight', Tq, 1, 1, out=gh4, add_pre=gh5, out2=gp4, mask2=b['m4'],
                       mask2_kind=MASK_U8)
        # conv_4: h4 = relu(p4) + a3 ; then a3 = relu(p3)
        F.conv1d_wgrad(gp4, b['a3'], G[E + '_conv_4.weight'], 1, 1, ws)
        ops.bias_grad(gp4, G[E + '_conv_4.bias'])
        gp3 = b['gA1']
        cdgrad(gp4, E + '_conv_4.weight', Tq, 1, 1, out=gp3, add_pre=gh4, mask=b['a3'],
                       mask_kind=MASK_FLOAT)
        self._emit_wn_fold(*self.buckets['enc_hi'])
        if self.world > 1 and self.nvls is None:       # conv_4 .. pre_vq (and the codebook gradient) are final
            ops.record_callable(lambda: self._allreduce_bucket(*self.buckets['enc_hi']))
        if self.dp_overlap:
            ops.record_callable(lambda: self._nvls_bucket('enc_hi'))
        # conv_3 (k4 s2 p2): a3 = relu(conv3(h2))
        F.conv1d_wgrad(gp3, b['h2'], G[E + '_conv_3.weight'], 2, 2, ws)
        ops.bias_grad(gp3, G[E + '_conv_3.bias'])
        gh2, gp2 = b['gT_a'], b['gT_b']
        self._emit_wn_fold(*self.buckets['enc_c3'])
        if self.world > 1 and self.nvls is None:
            ops.record_callable(lambda: self._allreduce_bucket(*self.buckets['enc_c3']))
        if self.dp_overlap:
            ops.record_callable(lambda: self._nvls_bucket('enc_c3'))
        cdgrad(gp3, E + '_conv_3.weight', T, 2, 2, out=gh2, out2=gp2, mask2=b['m2'], mask2_kind=MASK_U8)
        # conv_2: h2 = relu(p2) + a1 ; a1 = relu(p1)
        F.conv1d_wgrad(gp2, b['a1'], G[E + '_conv_2.weight'], 1, 1, ws)
        ops.bias_grad(gp2, G[E + '_conv_2.bias'])
        self._emit_wn_fold(*self.buckets['enc_c2'])
        if self.world > 1 and self.nvls is None:
            ops.record_callable(lambda: self._allreduce_bucket(*self.buckets['enc_c2']))
        if self.dp_overlap:
            ops.record_callable(lambda: self._nvls_bucket('enc_c2'))
        gp1 = self._view('gA2', C, T)
        cdgrad(gp2, E + '_conv_2.weight', T, 1, 1, out=gp1, add_pre=gh2, mask=b['a1'],
                       mask_kind=MASK_FLOAT)
        F.conv1d_wgrad(gp1, b['x'], G[E + '_conv_1.weight'], 1, 1, ws)
        ops.bias_grad(gp1, G[E + '_conv_1.bias'])

        # ---- 8. gradient allreduce (average) + fused AMSGrad over the flat buffers (trainer.py:41-42,68) ----
        g_scale = 1.0
        self._emit_wn_fold(*self.buckets['enc_c1'])
        if self.dp_overlap:       # the last bucket, the exit barrier, and the join of the side stream
            ops.record_callable(lambda: self._nvls_bucket('enc_c1', last=True))
            return
        if self.nvls is not None:
            # the gradient allreduce is folded into the optimizer: barrier, Adam on this rank's slice with
            # g = multimem.ld_reduce(gradients of all GPUs) / W, parameters broadcast by multimem.st, barrier.  Nothing ran
            # beside the backward pass; the optimizer state is sharded over the ranks.
            ops.dp_amsgrad_step(self.nvls.ctx, self.mc_p, self.flat_p, self.mc_g, self.flat_m, self.flat_v, self.flat_vmax,
                                self.opt_step, self.lr, self.betas[0], self.betas[1], self.eps, ch_before=1, ch_after=2)
            return
        if self.world > 1:
            ops.record_callable(lambda: self._allreduce_bucket(*self.buckets['enc_c1']))
            ops.record_callable(self._wait_buckets)
            g_scale = self.comm.grad_scale
        ops.amsgrad_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_vmax, self.opt_step, self.lr,
                         self.betas[0], self.betas[1], self.eps, g_scale=g_scale)

    # ------------------------------------------------------------------------------------------------
    def _run_schedule(self):
        ops.replay(self.schedule)

    def load_batch(self, x_btf, target_btf=None, non_blocking=True):
        """Stages one (B, T, F) feature batch (host pinned or device tensor) into the static input buffer, and the
        reconstruction target (B, T, F_out) when the step was built with separate_target."""
        self.buf['x_in'].copy_(x_btf, non_blocking=non_blocking)
        if self.separate_target:
            if target_btf is None:
                raise ValueError('this step was built with separate_target: pass the target batch (trainer.py:47)')
            self.buf['t_in'].copy_(target_btf, non_blocking=non_blocking)
        elif target_btf is not None:
            raise ValueError('this step reconstructs its input; build it with separate_target=True to pass a target')

    def set_speaker_features(self, speaker_dic, speaker_id):
        """Draws the reference's per-forward random speaker embedding on the host RNG -- the same nn.Embedding constructor
        + normal_(0, 0.1) sequence as global_conditioning.py:34,62-65, so the same torch seed gives the same features --
        and uploads the rows of this batch's speakers."""
        emb = torch.nn.Embedding(len(speaker_dic), self.dims['Dd'] - self.dims['D'], padding_idx=None)
        emb.weight.data.normal_(0, 0.1)
        rows = emb.weight.data[torch.as_tensor(speaker_id).view(self.B, -1)[:, 0].long().cpu()]
        self.host_spk.copy_(rows)
        self.buf['spk'].copy_(self.host_spk, non_blocking=True)

    def set_jitter_plan(self, src):
        self.buf['jitter_src'].copy_(torch.as_tensor(np.asarray(src, dtype=np.int32)), non_blocking=True)

    def step(self, x_btf=None, target_btf=None, speaker_dic=None, speaker_id=None):
        """One training step (convolutional_trainer.py:44-74) on the batch currently staged (or on x_btf; target_btf =
        data['output_features'] when built with separate_target; speaker_dic / speaker_id as ConvolutionalVQVAE.forward
        takes them, used only with use_speaker_conditioning).  Returns nothing; `losses()` reads results."""
        if x_btf is not None:
            self.load_batch(x_btf, target_btf)
        if self.use_speaker:
            if speaker_dic is None or speaker_id is None:
                raise ValueError('use_speaker_conditioning: step() needs speaker_dic and speaker_id')
            self.set_speaker_features(speaker_dic, speaker_id)
        if self.use_jitter:
            self.last_plan = jitter_plan(self.dims['Tq'], self.model._decoder._jitter._probability)
            self.set_jitter_plan(self.last_plan)
        if not self.use_graph or self.steps_done == 0:
            self._run_schedule()       # the first step always runs launch by launch (loads every kernel before capture)
        else:
            if self.graph is None:
                try:
                    self._capture()
                except Exception as exc:       # e.g. a collective that cannot be captured: stay on launch-by-launch replay
                    import warnings
                    warnings.warn('CUDA-graph capture of the training step failed (%s); replaying launch by launch' % exc)
                    self.use_graph = False
                    self.graph = None
                    torch.cuda.synchronize()            # the aborted capture has ended; nothing of it ran
                    self.comm.abandon()                 # Work handles created while capturing belong to no real launch
                    self._run_schedule()
                    self.steps_done += 1
                    return
            self.graph.replay()
        self.steps_done += 1

    def set_hyper_parameters(self, lr=None, betas=None, eps=None):
        """Changes the optimizer's hyper-parameters.  They are arguments of the recorded vqs_amsgrad_step launch, so the
        schedule entry is rewritten and a captured graph is dropped (re-captured on the next step)."""
        if lr is not None:
            self.lr = float(lr)
        if betas is not None:
            self.betas = (float(betas[0]), float(betas[1]))
        if eps is not None:
            self.eps = float(eps)
        for i, (fn, args, keep) in enumerate(self.schedule):
            if fn is not None and fn.__name__ == 'vqs_amsgrad_step':
                a = list(args)
                a[8:12] = [self.lr, self.betas[0], self.betas[1], self.eps]
                self.schedule[i] = (fn, tuple(a), keep)
            if fn is not None and fn.__name__ == 'vqs_dp_amsgrad_step':
                a = list(args)
                a[10:14] = [self.lr, self.betas[0], self.betas[1], self.eps]
                self.schedule[i] = (fn, tuple(a), keep)
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph = None

    def _capture(self):
        """Captures the recorded schedule into one CUDA graph.  The schedule is pure (no host sync, no allocation), and
        all state it touches lives in static buffers, so the capture pass itself performs no work."""
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            self._run_schedule()
        self.graph = g

    def losses(self):
        """{'loss', 'reconstruction_loss', 'vq_loss', 'perplexity'} of the last step: ONE 32-byte D2H copy of the scalar vector the
        step's kernels wrote + sync (the reference does three .item() syncs per step, trainer.py:57,58,70)."""
        self.host_scalars.copy_(self.buf['vq_scalars'], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = self.host_scalars
        r, v, p = float(h[5]), float(h[3] if self.is_ema else h[4]), float(h[2])
        return {'loss': r + v, 'reconstruction_loss': r, 'vq_loss': v, 'perplexity': p}

    def encoding_indices(self):
        return self.buf['idx'].view(-1, 1)

    def gradients(self):
        """name -> gradient (valid after a step; world > 1: summed over ranks, not yet divided).  With the NVLS exchange the
        flat buffer holds this rank's LOCAL gradient (the sum only ever exists inside the optimizer kernel), so the sum is
        formed here with a collective: every rank must call it (debugging / tests only)."""
        if self.nvls is None:
            return dict(self.grads)
        import torch.distributed as dist
        tot = self.flat_g.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.comm.pg)
        out = {}
        for name, g in self.grads.items():
            if name in self.param_offsets:
                off = self.param_offsets[name]
                out[name] = tot[off:off + g.numel()].view_as(g)
            else:
                out[name] = g
        return out


# ------------------------------------------------------------------------------------------------
# checkpoint wire format of the reference (convolutional_trainer.py:76-86, pipeline_factory.py:108-126)
# ------------------------------------------------------------------------------------------------
def _param_order(step):
    order, seen = [], set()
    for name, p in step.model.named_parameters():
        if id(p) not in seen:
            seen.add(id(p))
            order.append(name)
    return order


def optimizer_state_dict(step):
    """torch.optim.Adam(amsgrad=True)-compatible state_dict of a FusedTrainStep: the flat AMSGrad buffers are sliced back
    into per-parameter exp_avg / exp_avg_sq / max_exp_avg_sq, indexed in model.parameters() order like torch does, so
    `torch.optim.Adam(model.parameters(), lr, amsgrad=True).load_state_dict(...)` accepts it (the reference resumes
    exactly that way, pipeline_factory.py:118-120).  Parameters Adam never stepped (EMA codebook) have no state.  Only the
    named parameters the optimizer owns are listed (step.param_names: for weight-normalised convs weight_g / weight_v, not
    the effective-weight scratch buffers), at their offsets in the flat buffers."""
    order = _param_order(step)
    params = dict(step.model.named_parameters())
    state = {}
    nstep = int(step.opt_step.item())
    fm, fv, fx = step.flat_m, step.flat_v, step.flat_vmax
    if getattr(step, 'nvls', None) is not None:     # moments are sharded over the ranks: collect them (collective call)
        bk = list(step.buckets.values()) if getattr(step, 'dp_overlap', False) else None
        fm, fv, fx = (step.nvls.gather_sharded(t, bk) for t in (fm, fv, fx))
    if nstep > 0:
        for name in step.param_names:
            off, p = step.param_offsets[name], params[name]
            n = p.numel()
            state[order.index(name)] = {
                'step': torch.tensor(float(nstep)),
                'exp_avg': fm[off:off + n].view_as(p).clone(),
                'exp_avg_sq': fv[off:off + n].view_as(p).clone(),
                'max_exp_avg_sq': fx[off:off + n].view_as(p).clone(),
            }
    group = dict(lr=step.lr, betas=tuple(step.betas), eps=step.eps, weight_decay=0, amsgrad=True, maximize=False,
                 foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False,
                 params=list(range(len(order))))
    return {'state': state, 'param_groups': [group]}


def load_optimizer_state_dict(step, sd):
    """Inverse of optimizer_state_dict: accepts a torch Adam(amsgrad=True) state_dict (e.g. from a reference checkpoint),
    including its hyper-parameters -- optimizer.load_state_dict restores lr / betas / eps of the checkpoint
    (pipeline_factory.py:118-120), so does this."""
    order = _param_order(step)
    nstep = 0
    for idx, st in sd['state'].items():
        name = order[int(idx)]
        if name not in step.param_offsets:
            continue
        off, n = step.param_offsets[name], st['exp_avg'].numel()
        step.flat_m[off:off + n].copy_(st['exp_avg'].reshape(-1))
        step.flat_v[off:off + n].copy_(st['exp_avg_sq'].reshape(-1))
        step.flat_vmax[off:off + n].copy_(st['max_exp_avg_sq'].reshape(-1))
        nstep = max(nstep, int(float(st['step'])))
    step.opt_step.fill_(nstep)
    groups = sd.get('param_groups') or []
    if groups:
        g0 = groups[0]
        if not g0.get('amsgrad', True):
            import warnings
            warnings.warn('checkpoint optimizer has amsgrad=False; the fused step always runs Adam(amsgrad=True) '
                          '(convolutional_trainer.py:41-42)')
        step.set_hyper_parameters(lr=g0.get('lr'), betas=g0.get('betas'), eps=g0.get('eps'))


def save_checkpoint(step, path, experiment_name, epoch, train_res_recon_error=-1, train_res_perplexity=-1):
    """Writes the reference's per-epoch checkpoint dict (convolutional_trainer.py:76-86): same keys, same state_dict
    names, so either implementation can resume from the other's file."""
    torch.save({'experiment_name': experiment_name, 'epoch': epoch + 1,
                'model': {k: v.detach().clone() for k, v in step.model.state_dict().items()},
                'optimizer': optimizer_state_dict(step),
                'train_res_recon_error': train_res_recon_error, 'train_res_perplexity': train_res_perplexity}, path)


def load_checkpoint(step, path, map_location=None):
    """Restores model + optimizer state from a checkpoint written by save_checkpoint or by the reference trainer."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    with torch.no_grad():
        sd = step.model.state_dict()
        for k, v in ck['model'].items():
            sd[k].copy_(v)                 # in place: parameters are views of the flat buffer
    load_optimizer_state_dict(step, ck['optimizer'])
    return ck


def reference_config(**over):
    """The hot-path keys of configurations/vctk_features.yaml (:36-85) with the vq44-mfcc39 experiment's values."""
    cfg = dict(output_features_filters=13, augment_output_features=True, output_features_dim=47, verbose=False,
               input_features_dim=47, num_hiddens=768, num_residual_layers=2, use_kaiming_normal=False,
               input_features_type='mfcc', input_features_filters=13, augment_input_features=True,
               sampling_rate=16000, embedding_dim=64, decay=0.0, num_embeddings=44, commitment_cost=0.25,
               residual_channels=768, use_jitter=False, jitter_probability=0.12, use_speaker_conditioning=False,
               record_codebook_stats=False, learning_rate=2e-4, batch_size=2)
    cfg.update(over)
    return cfg
