#!/usr/bin/env python
"""Strong-scaling measurement of row-band sharding (BASELINE configs[2]): one 721x1440 near-global field,
hourly winds, pointwise x-boundary, particle rows split over the ranks with a 2-row recomputed halo, NCCL
gather of the finished bands.  Launch with torchrun; prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_rowbands.py [--nt 25]
"""
import argparse, json, os, sys
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine
from lagrangiancoherence_b200.rolling import band_ftle, gather_bands

ap = argparse.ArgumentParser()
ap.add_argument("--nt", type=int, default=73)       # BASELINE configs[2]: 72 h of hourly winds = 73 levels
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
a = ap.parse_args()
world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
lat, lon = S.grid_c3()
u, v = S.era5_like_winds(lat, lon, a.nt, noise=0.0)
eng = FtleEngine(lat, lon, -3600, SETTLS_order=4, xmode='pointwise', device=dev)
du, dv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def step():
    st = eng.stage(du, dv)                       # winds replicated: every rank prefilters the whole field
    band, _ = band_ftle(eng, st, world, rank)
    return gather_bands(band, lat.size) if world > 1 else band

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)

for _ in range(a.warmup):
    step()
ms = []
for _ in range(a.steps):
    flush.fill_(1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record()
    barrier()
    ms.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.sum(ms))], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    total = float(t.item())
    psteps = lat.size * lon.size * (a.nt - 1)
    print(json.dumps({'metric': 'particle-steps/s', 'value': psteps * a.steps / (total * 1e-3), 'n_gpus': world,
                      'ms_per_field': total / a.steps, 'fields_per_s': a.steps / (total * 1e-3), 'scaling': 'strong',
                      'config': {'workload': f'C3 721x1440 hourly, {a.nt - 1} intervals, S=4, cubic, f64, pointwise clamp, '
                                             f'row bands x{world} with 2-row halo, winds replicated, NCCL gather of bands'}}))
if world > 1:
    dist.destroy_process_group()
