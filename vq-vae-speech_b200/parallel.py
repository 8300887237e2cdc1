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


class DataParallelComm(object):
    def __init__(self, process_group=None):
        self.pg = process_group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if self.enabled else 1
        self.rank = dist.get_rank(process_group) if self.enabled else 0
        self._works = []

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
