"""world_size-2 gloo run of the multi-GPU host logic (planning + gather) on the CPU.  The per-rank compute is
stubbed with the oracle restricted to the rank's shard, which also proves the sharding rules themselves:
  * start-time sharding: concatenating the ranks' windows reproduces the full rolling series;
  * row-band sharding with a 2-row recomputed halo (global row indices for the pole/one-sided rules)
    reproduces the single-process field exactly for the pointwise x-clamp."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.rolling import gather_bands, gather_fields, shard_rows, shard_starts

LAT = np.linspace(-30.0, 10.0, 33)
LON = np.linspace(-80.0, -24.0, 41)
DT, SORD, NT, NSTARTS = -21600, 2, 4, 5


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _band_sigma(u, v, rank, world):
    """What a rank computes under row-band sharding: integrate rows [in0, in1), epilogue on [out0, out1)."""
    out0, out1, in0, in1 = shard_rows(LAT.size, world, rank)
    x, y = O.parcel_propagation(u, v, LAT, LON, DT, SETTLS_order=SORD, xclamp='pointwise')
    # particles are independent under the pointwise clamp: a band run equals the band of the full run; NaN outside
    # the band proves the epilogue below never reads rows a rank did not integrate
    xb, yb = np.full_like(x, np.nan), np.full_like(y, np.nan)
    xb[in0:in1], yb[in0:in1] = x[in0:in1], y[in0:in1]
    jac = O.flowmap_gradient(xb, yb, LAT, LON)
    return O.sigma_max_closed_form(jac)[out0:out1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    u, v = S.era5_like_winds(LAT, LON, NT + NSTARTS - 1)
    # start-time sharding
    first, count = shard_starts(NSTARTS, world, rank)
    mine = np.stack([O.lcs_field(u[s:s + NT], v[s:s + NT], LAT, LON, DT, SETTLS_order=SORD) for s in range(first, first + count)])
    counts = [shard_starts(NSTARTS, world, r)[1] for r in range(world)]
    allf = gather_fields(torch.from_numpy(mine), counts)
    # row-band sharding of window 0
    band = torch.from_numpy(_band_sigma(u[:NT], v[:NT], rank, world))
    full = gather_bands(band, LAT.size)
    if rank == 0:
        q.put((allf.numpy(), full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_start_time_and_row_band_sharding_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    allf, full = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    u, v = S.era5_like_winds(LAT, LON, NT + NSTARTS - 1)
    ref = np.stack([O.lcs_field(u[s:s + NT], v[s:s + NT], LAT, LON, DT, SETTLS_order=SORD) for s in range(NSTARTS)])
    assert allf.shape == ref.shape and np.array_equal(allf, ref)
    x, y = O.parcel_propagation(u[:NT], v[:NT], LAT, LON, DT, SETTLS_order=SORD, xclamp='pointwise')
    ref_band = O.sigma_max_closed_form(O.flowmap_gradient(x, y, LAT, LON))
    assert full.shape == ref_band.shape and np.array_equal(full, ref_band)


def test_band_ftle_refuses_outer_clamp_across_ranks():
    from lagrangiancoherence_b200 import _lib
    from lagrangiancoherence_b200.rolling import band_ftle

    class FakeEngine:
        xmode = _lib.LCS_X_CLAMP_OUTER
        nlat = 33
    with pytest.raises(ValueError, match='outer-product'):
        band_ftle(FakeEngine(), None, world_size=2, rank=0)
