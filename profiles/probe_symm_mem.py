"""Feasibility probe (2+ GPUs under torch.distributed.run): torch symmetric memory with NVLS multicast on this box."""
import os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print('SYMM rank %d: buffer_ptrs %s signal_pad_ptrs %s multicast_ptr %s' % (
        rank, [hex(p) for p in hdl.buffer_ptrs], [hex(p) for p in hdl.signal_pad_ptrs], hex(hdl.multicast_ptr) if hdl.multicast_ptr else None), flush=True)
    print('SYMM attrs', [a for a in dir(hdl) if not a.startswith('_')], flush=True)
    t.fill_(float(rank + 1))
    hdl.barrier()
    try:
        torch.ops.symm_mem.multimem_all_reduce_(t, 'sum', dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print('SYMM rank %d multimem_all_reduce -> %s (expect %s)' % (rank, t[:3].tolist(), world * (world + 1) / 2), flush=True)
    except Exception as e:
        print('SYMM multimem_all_reduce failed: %r' % (e,), flush=True)
    # peer view through P2P pointers
    peer = hdl.get_buffer((rank + 1) % world, (8,), torch.float32)
    print('SYMM rank %d peer buffer head %s' % (rank, peer.tolist()), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print('SYMM failed: %r' % (e,), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
