"""Data-parallel plumbing of the training step: one process per GPU, torch.distributed (NCCL on B200, gloo in the CPU
tests).  The reference has no working multi-GPU path (its nn.DataParallel wrap is dead code, pipeline_factory.py:56-61);
the contract implemented here is SURVEY.md 8e:

  * the batch is sharded over ranks; every rank holds the full (bit-identical) weights, codebook and EMA state;
  * EMA statistics [counts | dw] are SUM-allreduced between the assignment and the EMA update, so every rank applies
    the same update and codebooks stay identical;
  * gradients are SUM-allreduced in buckets launched as the backward pass completes them (decoder first, so its
    allreduces overlap the encoder's backward) and divided by world_size inside the fused AMSGrad kernel (g_scale).

Nothing here touches CUDA directly, so the same class runs under gloo on CPU tensors (tests/test_parallel_cpu.py).
"""
import os

import torch
import torch.distributed as dist

# The conv / wgrad GEMMs of the step are planned as ONE wave of 144 CTAs with ~200 KB of shared memory each (one per SM);
# an NCCL kernel CTA cannot share an SM with them, so every SM NCCL takes beyond the 4 spare ones pushes a GEMM into a
# second wave (round 1, 8 GPUs: the conv_3 dgrad launched next to a gradient bucket took 0.205 ms instead of 0.111).
# 144 + 4 = 148: the communicator is capped at 4 CTAs.  65 MB of gradients per 3 ms step need ~25 GB/s -- far below what 4
# CTAs move over NVLink 5 / NVSwitch (NVLS reduces inside the switch).
NCCL_MAX_CTAS = 4


def init_nccl(device, max_ctas=NCCL_MAX_CTAS):
    """torch.distributed.init_process_group('nccl') for one process per GPU (RANK / WORLD_SIZE / MASTER_* from the
    environment, as torch.distributed.run sets them) with the communicator capped at `max_ctas` CTAs (see above; the
    environment variables NCCL_MAX_CTAS / NCCL_MIN_CTAS, when already set by the user, win)."""
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if max_ctas:
        os.environ.setdefault('NCCL_MAX_CTAS', str(int(max_ctas)))
        os.environ.setdefault('NCCL_MIN_CTAS', '1')
    dist.init_process_group('nccl', device_id=device)
    return DataParallelComm()


class NvlsExchange(object):
    """Symmetric buffers + signal pads for the library's own data-parallel kernels (csrc/dp_nvls.cu): every rank allocates
    the same buffers through torch.distributed._symmetric_memory (CUDA VMM + fabric handles; torch is used for the memory
    plumbing only), which maps each rank's copy into every process and the whole set behind one NVLS multicast address.
    Needs the NCCL backend, one process per GPU, <= 8 GPUs of one NVSwitch domain, multicast support."""

    def __init__(self, process_group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._symm = symm_mem
        self.pg = process_group if process_group is not None else dist.group.WORLD
        self.group_name = self.pg.group_name
        self.world, self.rank = dist.get_world_size(self.pg), dist.get_rank(self.pg)
        if self.world > _lib.DP_MAX_WORLD:
            raise RuntimeError('NvlsExchange supports at most %d ranks' % _lib.DP_MAX_WORLD)
        self.device = torch.device('cuda', torch.cuda.current_device())
        self._handles = []
        # a small buffer first: its handle carries the signal pads (one per rank, mapped everywhere)
        probe, hdl = self._alloc(64)
        if not hdl.multicast_ptr:
            raise RuntimeError('symmetric memory has no NVLS multicast mapping on this system')
        need = (_lib.DP_PAD_WORD0 + _lib.DP_CHANNELS * _lib.DP_MAX_WORLD) * 4
        if hdl.signal_pad_size < need:
            raise RuntimeError('signal pad of %d bytes < %d' % (hdl.signal_pad_size, need))
        # zero this library's words of the own pad, then a torch-side barrier: no kernel of ours has signalled yet
        pad = hdl.get_signal_pad(self.rank, (hdl.signal_pad_size // 4,), dtype=torch.int32)
        pad[_lib.DP_PAD_WORD0:_lib.DP_PAD_WORD0 + _lib.DP_CHANNELS * _lib.DP_MAX_WORLD].zero_()
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)
        self.epochs = torch.zeros(_lib.DP_EPOCH_WORDS, dtype=torch.int32, device=self.device)
        self.ctx = _lib.DpCtx()
        self.ctx.rank, self.ctx.world = self.rank, self.world
        for r in range(self.world):
            self.ctx.peer_pads[r] = int(hdl.signal_pad_ptrs[r])
        self.ctx.epochs = self.epochs.data_ptr()
        self._ctypes = ctypes

    def _alloc(self, numel):
        t = self._symm.empty(int(numel), dtype=torch.float32, device=self.device)
        hdl = self._symm.rendezvous(t, self.group_name)
        self._handles.append((t, hdl))
        return t, hdl

    def symmetric_zeros(self, numel):
        """(tensor, multicast address, _lib.DpPtrs of the peers' copies) of a zeroed symmetric fp32 buffer of `numel`
        elements.  Collective: every rank must call it in the same order with the same size."""
        from . import _lib
        t, hdl = self._alloc(numel)
        t.zero_()
        ptrs = _lib.DpPtrs()
        for r in range(self.world):
            ptrs.p[r] = int(hdl.buffer_ptrs[r])
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)
        return t, int(hdl.multicast_ptr), ptrs

    def slice_bounds(self, numel):
        """[lo, hi) of the flat buffers (in elements) that this rank's optimizer shard owns -- the split vqs_dp_amsgrad_step
        uses: float4 granules, ceil(n / 4 / W) per rank."""
        return shard_bounds(0, numel, self.world, self.rank)

    def range_slice(self, lo, hi):
        """[lo', hi') of bucket [lo, hi) that this rank's optimizer shard owns -- the split of vqs_dp_amsgrad_range."""
        return shard_bounds(lo, hi, self.world, self.rank)

    def gather_sharded(self, flat, buckets=None):
        """Full copy of a rank-sharded flat buffer (AMSGrad moments): every rank contributes its own slice -- of the whole
        buffer, or of every bucket in `buckets` ([(lo, hi), ...]) when the step exchanges bucket by bucket (collective)."""
        out = torch.zeros_like(flat)
        if buckets is None:
            owned = [self.slice_bounds(flat.numel())]
        else:
            owned = [self.range_slice(lo, hi) for lo, hi in buckets if hi > lo]
        for lo, hi in owned:
            out[lo:hi].copy_(flat[lo:hi])
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.pg)
        return out

    def peer_view(self, tensor, rank, numel=None):
        """Rank `rank`'s copy of a symmetric tensor of this exchange as a tensor of this process (P2P mapping)."""
        for t, hdl in self._handles:
            if t.data_ptr() == tensor.data_ptr():
                return hdl.get_buffer(rank, (numel if numel is not None else tensor.numel(),), torch.float32)
        raise KeyError('not a symmetric tensor of this exchange')


def shard_bounds(lo, hi, world, rank):
    """[lo', hi') of the element range [lo, hi) (multiples of 4) that rank `rank` of `world` owns in the sharded optimizer step:
    float4 granules, ceil(n / 4 / W) per rank, trailing ranks possibly empty -- exactly the split of vqs_dp_amsgrad_step
    (lo = 0, hi = n) and vqs_dp_amsgrad_range (csrc/dp_nvls.cu)."""
    b4, n4 = lo // 4, (hi - lo) // 4
    per = (n4 + world - 1) // world
    lo4 = b4 + per * rank
    hi4 = min(lo4 + per, b4 + n4)
    return 4 * lo4, 4 * max(hi4, lo4)


def nvls_available(process_group=None):
    """True when the library's own NVLS exchange can be used: NCCL process group with > 1 rank on CUDA, torch symmetric
    memory importable, not switched off by VQS_DP_NVLS=0 (the NCCL-allreduce path of round 1 stays as the fallback)."""
    if os.environ.get('VQS_DP_NVLS', '1') == '0':
        return False
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) < 2:
        return False
    if dist.get_backend(process_group) != 'nccl' or not torch.cuda.is_available():
        return False
    try:
        import importlib
        importlib.import_module('torch.distributed._symmetric_memory')
    except Exception:
        return False
    return True


class DataParallelComm(object):
    def __init__(self, process_group=None):
        self.pg = process_group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if self.enabled else 1
        self.rank = dist.get_rank(process_group) if self.enabled else 0
        self._works = []
        self.nvls = None
        if nvls_available(process_group):
            try:
                self.nvls = NvlsExchange(process_group)
            except Exception as exc:            # no multicast on this system: NCCL allreduces instead
                import warnings
                warnings.warn('NVLS exchange unavailable (%s); using NCCL allreduces' % (exc,))
                self.nvls = None
            # every rank must take the same path
            flag = torch.tensor([1 if self.nvls is not None else 0], device='cuda')
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=process_group)
            if int(flag.item()) == 0:
                self.nvls = None

    @property
    def grad_scale(self):
        """Factor the optimizer applies to the summed gradient: the average over ranks."""
        return 1.0 / self.world

    def allreduce_stats(self, stats):
        """In-place SUM of the packed [counts (K) | dw (K*D)] vector (blocking: the EMA update needs it)."""
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.pg)
        return stats

    def total_rows(self, local_rows):
        """Rows the (allreduced) counts were taken over: every rank contributes the same number of rows."""
        return local_rows * self.world

    def start_bucket(self, flat_grad, lo, hi):
        """Launches the asynchronous SUM-allreduce of flat_grad[lo:hi]."""
        if self.world > 1 and hi > lo:
            self._works.append(dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def wait_buckets(self):
        for w in self._works:
            w.wait()
        self._works = []

    def abandon(self):
        """Drops Work handles that were created inside an aborted CUDA-graph capture (they belong to no real launch)."""
        self._works = []

    def shard(self, global_batch):
        """This rank's contiguous slice [r*B/W, (r+1)*B/W) of a global batch (first dim)."""
        n = global_batch.shape[0]
        if n % self.world != 0:
            raise ValueError('global batch %d is not divisible by world size %d' % (n, self.world))
        per = n // self.world
        return global_batch[self.rank * per:(self.rank + 1) * per]

    def assert_replicated(self, tensor, what='tensor'):
        """Debug helper: raises unless `tensor` is bit-identical on every rank."""
        if self.world == 1:
            return
        lo = tensor.detach().clone()
        hi = tensor.detach().clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.pg)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.pg)
        if not torch.equal(lo, hi):
            raise RuntimeError('%s differs across data-parallel ranks' % what)
