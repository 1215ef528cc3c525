"""world_size-2 gloo run of the multi-GPU host logic (planning + gather) on the CPU.  The per-rank compute is
stubbed with the oracle restricted to the rank's shard, which also proves the sharding rules themselves:
  * start-time sharding: concatenating the ranks' windows reproduces the full rolling series;
  * row-band sharding with a 2-row recomputed halo (global row indices for the pole/one-sided rules)
    reproduces the single-process field exactly for the pointwise x-clamp;
  * under the as-executed outer-product clamp the bands need exactly one thing from each other, the column exit
    flags of every sub-step."""
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


def _banded_outer_integration(u, v, world, S_order):
    """The rule the multi-GPU outer-clamp path implements (rolling.band_ftle + lcs_xrank), restated on the CPU: every
    rank advances ONLY its band of particle rows (NaN elsewhere: inert under every comparison), keeps its own row exit
    flags, and after each sub-step ORs the column exit flags of all ranks -- first for "< x_min", then, on the updated
    positions, for "> x_max" (trajectory.py:96-97).  Returns the rows each rank owns, stitched together."""
    cx, cy = O.conversions(LAT)
    cx = cx[:, None]
    y_min, y_max, x_min, x_max = LAT.min(), LAT.max(), LON.min(), LON.max()
    X, Y = np.meshgrid(LON, LAT)
    bands = [shard_rows(LAT.size, world, r) for r in range(world)]
    px, py = [], []
    for out0, out1, in0, in1 in bands:
        bx, by = np.full_like(X, np.nan), np.full_like(Y, np.nan)
        bx[in0:in1], by[in0:in1] = X[in0:in1], Y[in0:in1]
        px.append(bx); py.append(by)

    def clamp_all():
        for which, bound in ((lambda a: a < x_min, x_min), (lambda a: a > x_max, x_max)):
            with np.errstate(invalid='ignore'):
                hits = [which(bx) for bx in px]
            cols = np.zeros(LON.size, bool)
            for h in hits:
                cols |= h.any(axis=0)                              # the exchange: OR of the column flags over the ranks
            for r, h in enumerate(hits):
                rows = h.any(axis=1)                               # row flags never leave their rank
                px[r][np.ix_(rows, cols)] = np.where(np.isnan(px[r][np.ix_(rows, cols)]), np.nan, bound)

    def interp(F, r):
        with np.errstate(invalid='ignore'):
            return O.xr_map_coordinates(F, np.nan_to_num(px[r], nan=LON[0]), np.nan_to_num(py[r], nan=LAT[0]), LAT, LON, order=3)

    for t in range(u.shape[0] - 1):
        ua = [interp(u[t], r) for r in range(world)]
        va = [interp(v[t], r) for r in range(world)]
        for r in range(world):
            py[r] = O._clamp_y(py[r] + DT * cy * va[r], y_min, y_max) + 0 * px[r]      # "+ 0 * px": keep the NaN rows NaN
            px[r] = px[r] + DT * cx * ua[r]
        clamp_all()
        for _ in range(S_order):
            for r in range(world):
                v_t, v_tp = interp(v[t], r), interp(v[t + 1], r)
                u_t, u_tp = interp(u[t], r), interp(u[t + 1], r)
                py[r] = O._clamp_y(py[r] + 0.5 * DT * cy * (va[r] + 2 * v_t - v_tp), y_min, y_max) + 0 * px[r]
                px[r] = px[r] + 0.5 * DT * cx * (ua[r] + 2 * u_t - u_tp)
            clamp_all()
    outx = np.concatenate([px[r][b[0]:b[1]] for r, b in enumerate(bands)])
    outy = np.concatenate([py[r][b[0]:b[1]] for r, b in enumerate(bands)])
    return outx, outy


@pytest.mark.parametrize('world', [2, 3])
def test_row_bands_under_the_outer_clamp_need_only_the_column_flags(world):
    """Row-band sharding under the as-executed outer-product clamp: with the column exit flags OR-ed across the bands
    after every sub-step (and nothing else exchanged), the stitched bands equal the single-process integration bit for
    bit -- the halo rows a neighbour recomputes included."""
    u, v = S.era5_like_winds(LAT, LON, NT)
    u, v = 2.0 * u, 2.0 * v                                              # strong winds: exits through both x-boundaries
    rx, ry = O.parcel_propagation(u, v, LAT, LON, DT, SETTLS_order=SORD, xclamp='outer')
    assert (rx == LON.min()).mean() > 0.02 and (rx == LON.max()).any()   # the clamp is exercised on both sides
    bx, by = _banded_outer_integration(u, v, world, SORD)
    assert np.array_equal(bx, rx) and np.array_equal(by, ry)
    px, _ = O.parcel_propagation(u, v, LAT, LON, DT, SETTLS_order=SORD, xclamp='pointwise')
    assert not np.array_equal(px, rx)                                    # ... and it is not the pointwise clamp in disguise
