"""Scratch: does torch symmetric memory (peer-mapped buffers across ranks) work on this box?  torchrun, 2+ ranks."""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
n = 1 << 24
t = symm.empty((world, n), dtype=torch.float64, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
peers = [hdl.get_buffer(r, (world, n), torch.float64) for r in range(world)]
t.zero_()
hdl.barrier()
src = torch.full((n,), float(rank + 1), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(world):
        peers[r][rank].copy_(src)          # push my block into every rank's buffer (peer stores / copy engine)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
hdl.barrier()
ok = all(float(t[r][0].item()) == r + 1 and float(t[r][-1].item()) == r + 1 for r in range(world))
print(f'rank {rank}: ok={ok} push of {world} x {n * 8 / 1e6:.0f} MB took {ms:.2f} ms -> {(world - 1) * n * 8 / ms / 1e6:.0f} GB/s remote; ptrs {[hex(p.data_ptr()) for p in peers]}', flush=True)
dist.destroy_process_group()
