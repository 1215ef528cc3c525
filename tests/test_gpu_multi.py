"""Two-rank NCCL run of the sharded paths on real GPUs (skipped on a single-GPU box): start-time sharding of a
rolling series (gathered with NCCL and with NVLink peer copies, onto every rank and onto rank 0 only) and row-band
sharding of one field -- pointwise clamp, and the as-executed outer-product clamp with its cross-rank exchange of the
column exit flags -- against the single-GPU result, bit for bit."""
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
    # gather onto rank 0 only (what the bench times): the other ranks receive nothing
    all0 = rolling_ftle_sharded(u, v, lat, lon, nt, -21600, engine=eng, gather='p2p', chunk=2, dst=0)
    assert (all0 is None) == (rank != 0)
    # row bands, pointwise clamp
    engp = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='pointwise', device=dev)
    st = engp.stage(u[:nt], v[:nt])
    full = ftle_row_bands(engp, u[:nt], v[:nt])
    # row bands under the as-executed outer-product clamp: column exit flags exchanged between the ranks inside the
    # persistent kernel after every sub-step (peer.ColumnFlagMail); strong winds so that both x-boundaries see exits;
    # one field, then three rolling windows in flight at once, then the same again (the mailboxes are reused)
    us, vs = 2.0 * u, 2.0 * v
    outer1 = ftle_row_bands(eng, us[:nt], vs[:nt])
    outer3 = ftle_row_bands(eng, us[:nt + 2], vs[:nt + 2], nwindows=3)
    outer3b = ftle_row_bands(eng, us[:nt + 2], vs[:nt + 2], nwindows=3)
    eng.check_finite()
    if rank == 0:
        ref_all = rolling_ftle(u, v, lat, lon, nt, -21600, engine=eng, return_device=True)
        x, y = engp.advect(st)
        ref_full = engp.epilogue(x, y)[0]
        sto = eng.stage(us[:nt + 2], vs[:nt + 2])
        xo, yo = eng.advect(sto, nsteps=nt - 1, nwindows=3)
        ref_outer = eng.epilogue(xo, yo)
        exits = bool((xo == lon.min()).any() and (xo == lon.max()).any())
        outer_ok = exits and bool(torch.equal(outer1, ref_outer[0]) and torch.equal(outer3, ref_outer) and torch.equal(outer3b, ref_outer))
        q.put((bool(torch.equal(allf, ref_all)) and bool(torch.equal(allp, ref_all)) and bool(torch.equal(all0.clone(), ref_all)),
               bool(torch.equal(full, ref_full)), outer_ok))
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
    starts_ok, bands_ok, outer_bands_ok = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert starts_ok and bands_ok and outer_bands_ok
