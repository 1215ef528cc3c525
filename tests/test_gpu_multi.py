"""Two-rank NCCL run of the sharded paths on real GPUs (skipped on a single-GPU box): start-time sharding of a
rolling series (gathered with NCCL and with NVLink peer copies) and row-band sharding of one field, against the
single-GPU result."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from lagrangiancoherence_b200 import synthetic as S
    from lagrangiancoherence_b200.engine import FtleEngine
    from lagrangiancoherence_b200.rolling import ftle_row_bands, rolling_ftle, rolling_ftle_sharded
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    lat = np.linspace(-40.0, 0.0, 81)
    lon = np.linspace(-80.0, -30.0, 101)
    nt, nstarts = 5, 7
    u, v = S.era5_like_winds(lat, lon, nt + nstarts - 1)
    # start-time sharding, outer clamp
    eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='outer', device=dev)
    allf = rolling_ftle_sharded(u, v, lat, lon, nt, -21600, engine=eng)
    # same shards, finished chunks pushed into every rank's symmetric-memory buffer over NVLink (ragged: 4 + 3 windows)
    allp = rolling_ftle_sharded(u, v, lat, lon, nt, -21600, engine=eng, gather='p2p', chunk=2).clone()
    # row bands, pointwise clamp
    engp = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='pointwise', device=dev)
    st = engp.stage(u[:nt], v[:nt])
    full = ftle_row_bands(engp, u[:nt], v[:nt])
    if rank == 0:
        ref_all = rolling_ftle(u, v, lat, lon, nt, -21600, engine=eng, return_device=True)
        x, y = engp.advect(st)
        ref_full = engp.epilogue(x, y)[0]
        q.put((bool(torch.equal(allf, ref_all)) and bool(torch.equal(allp, ref_all)), bool(torch.equal(full, ref_full))))
    else:
        assert torch.equal(allp, allf)                       # every rank holds the whole series either way
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sharding(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    starts_ok, bands_ok = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert starts_ok and bands_ok
