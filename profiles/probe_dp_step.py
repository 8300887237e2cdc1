"""Data-parallel exchange alone (2+ GPUs under torch.distributed.run): CUDA-event time of the signal-pad barrier, of the statistics
allreduce and of vqs_dp_amsgrad_step on buffers of the benchmarked model's size (16.4 M parameters), with torch's own
multimem all-reduce of the same 65 MB as the yardstick for what the NVLS path delivers.  Knobs: VQS_DP_FUSED=0 (four launches),
VQS_DP_BLOCKS_PER_SM=1|2."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
from vq_vae_speech_b200 import ops
from vq_vae_speech_b200.parallel import NvlsExchange
ex = NvlsExchange()
n = 16_400_384
p, mc_p, _ = ex.symmetric_zeros(n)
g, mc_g, _ = ex.symmetric_zeros(n)
st_local, _, st_ptrs = ex.symmetric_zeros(2880)
st = torch.zeros(2880, device=dev)
m, v, vm = (torch.zeros(n, device=dev) for _ in range(3))
step = torch.zeros(1, dtype=torch.int64, device=dev)
g.normal_()
p.normal_()


def timed(name, f, iters=30, warm=5):
    for _ in range(warm):
        f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('%-44s %.4f ms per call (max over %d ranks)' % (name, float(t), world), flush=True)


timed('vqs_dp_barrier', lambda: ops.dp_barrier(ex.ctx, 3))
timed('vqs_dp_allreduce_small (2880 floats)', lambda: ops.dp_allreduce_small(ex.ctx, st_ptrs, st, channel=0))
timed('vqs_dp_amsgrad_step (16.4 M parameters)', lambda: ops.dp_amsgrad_step(ex.ctx, mc_p, p, mc_g, m, v, vm, step, 2e-4))
timed('vqs_amsgrad_step, no exchange', lambda: ops.amsgrad_step(p, g, m, v, vm, step, 2e-4))
try:
    timed('torch multimem_all_reduce_ of the 65 MB', lambda: torch.ops.symm_mem.multimem_all_reduce_(g, 'sum', dist.group.WORLD.group_name))
except Exception as e:
    if rank == 0:
        print('torch multimem_all_reduce_ failed: %r' % (e,))
timed('NCCL all_reduce of the 65 MB', lambda: dist.all_reduce(m))
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
