"""Rolling FTLE time series and multi-GPU sharding.

The reference has no library function for this: its application script loops one ``LCS`` call per
start time (area_of_influence.py:168-184) and its CLI is driven by an external per-start-time job
array (LCS.py:236-268).  Here the whole series is staged once (each level is prefiltered and
packed once, not once per window) and all start times of a chunk are integrated by one launch.

Sharding, one process per GPU, no data-path collective:
  * start times  -- rank r owns a contiguous block of windows (``shard_starts``);
  * row bands    -- rank r integrates particle rows [r0-2, r1+2) (2-row recomputed halo for the
                    +-2 y-stencil, tools.py:204-207) and writes rows [r0, r1) (``shard_rows``);
                    Under the as-executed outer-product clamp (quirk Q6) the bands are coupled through
                    the per-sub-step column exit flags: those are OR-ed across the ranks inside the
                    persistent kernel over NVLink mailboxes (peer.ColumnFlagMail), rows stay local.
NCCL is used only to gather the finished fields (``gather_fields`` / ``gather_bands``).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .engine import FtleEngine, precision_args


# ------------------------------------------------------------------ pure planning helpers (CPU-testable)
def shard_starts(nstarts, world_size, rank):
    """Contiguous block of start times for ``rank``: returns ``(first, count)``."""
    base, extra = divmod(nstarts, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def shard_rows(nrows, world_size, rank, halo=2):
    """Row band of ``rank``: ``(out0, out1, in0, in1)`` -- rows written and rows integrated."""
    base, extra = divmod(nrows, world_size)
    out0 = rank * base + min(rank, extra)
    out1 = out0 + base + (1 if rank < extra else 0)
    return out0, out1, max(0, out0 - halo), min(nrows, out1 + halo)


def chunk_starts(first, count, chunk):
    """Split a block of start times into launches of at most ``chunk`` windows."""
    return [(s, min(chunk, first + count - s)) for s in range(first, first + count, chunk)]


def chunk_schedule(count, chunk, ramp=False):
    """Launch sizes of the pipelined host path: ``chunk`` windows first and last -- the first upload and the last
    download cannot overlap anything, so they should be short -- and ``2*chunk`` in between, where the copies hide
    behind the kernels and bigger launches are cheaper per window (with ``chunk`` = 148 the middle launches are whole
    waves of one-CTA windows).  ``ramp``: long series start and end with ``chunk/2, chunk`` (half the exposed copy time
    again; the half-size launches run at four CTAs per window, a few per cent slower on a sixteenth of the work).
    Returns ``[(first, n), ...]`` covering ``0 .. count``."""
    if count <= 3 * chunk:
        return chunk_starts(0, count, chunk)
    if ramp and chunk >= 8 and count > 6 * chunk:
        head = [chunk // 6, chunk // 2, chunk] if int(ramp) >= 2 else [chunk // 2, chunk]
    else:
        head = [chunk]
    rem = count - 2 * sum(head)
    mid = []
    while rem > 2 * chunk:
        mid.append(2 * chunk)
        rem -= 2 * chunk
    if rem > 0:
        mid.append(rem)
    out, s = [], 0
    for n in head + mid + head[::-1]:
        out.append((s, n))
        s += n
    return out


# ------------------------------------------------------------------ single-GPU rolling series
def _pipeline(engine, in_shape, in_dtype, need_in):
    """Copy streams and double-buffered device staging for host-resident winds, cached on the engine
    (fresh streams/buffers per call would defeat the caching allocator)."""
    key = (tuple(in_shape), in_dtype, need_in)
    pipe = getattr(engine, '_pipe', None)
    if pipe is None or pipe['key'] != key:
        dev = engine.device
        pipe = {'key': key, 'up': torch.cuda.Stream(dev), 'down': torch.cuda.Stream(dev), 'in': [], 'in_free': []}
        if need_in:
            for _ in range(2):
                pipe['in'].append((torch.empty(in_shape, dtype=in_dtype, device=dev),
                                   torch.empty(in_shape, dtype=in_dtype, device=dev)))
                pipe['in_free'].append(torch.cuda.current_stream(dev).record_event())
        engine._pipe = pipe
    return pipe


def rolling_ftle(u, v, lat, lon, window_levels, timestep, SETTLS_order=4, interp_order=3, xclamp='outer',
                 cyclic_xboundary=False, precision='f64', device='cuda:0', starts=None, chunk=148,
                 log_scale=False, out=None, engine=None, return_device=False, on_chunk=None):
    """sigma_max fields for every start time of a wind series.

    ``u, v``: ``[nlev, nlat, nlon]`` (numpy, pinned or not, or device tensors).  Window ``s`` uses
    levels ``s .. s+window_levels-1`` exactly as ``LCS(...)(u.isel(time=slice(s, s+window_levels)))``
    would.  Returns ``[nstarts, nlat, nlon]`` (numpy unless ``return_device``; ``out`` = a pinned
    host tensor to fill).

    ``on_chunk(first, fields)``: called with every finished chunk (device tensor ``[n, nlat, nlon]``, ``first`` = index
    of its first window within this call) as soon as its kernels are queued -- the hook the multi-GPU driver uses to
    push chunks to its peers while the next chunk runs.

    Start times are processed in chunks of ``chunk`` windows (148 = one cluster pair per SM for the
    outer-clamp kernel).  With host-resident winds the chunks are pipelined over three streams:
    upload of chunk i+1's levels and download of chunk i-1's fields overlap chunk i's kernels; the
    middle chunks are twice as long as the first and the last (``chunk_schedule``).
    """
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    nlev = u.shape[0]
    nsteps = window_levels - 1
    first, count = (0, nlev - window_levels + 1) if starts is None else starts
    if count < 0 or first < 0 or first + count + nsteps > nlev:
        raise ValueError('start-time range runs past the wind series')
    if engine is None:
        engine = FtleEngine(lat, lon, timestep, SETTLS_order=SETTLS_order, interp_order=interp_order,
                            xmode='cyclic' if cyclic_xboundary else xclamp, device=device, **precision_args(precision))
    dev = engine.device
    on_device = isinstance(u, torch.Tensor) and u.is_cuda
    if not on_device:
        if not isinstance(u, torch.Tensor):
            u, v = torch.from_numpy(np.ascontiguousarray(u)), torch.from_numpy(np.ascontiguousarray(v))
        if u.dtype not in (torch.float32, torch.float64):
            u, v = u.double(), v.double()
    to_host = not return_device
    own_out = out is None
    if to_host and own_out:
        out = torch.empty((count, lat.size, lon.size), dtype=torch.float64).pin_memory()
    sigma = None if to_host else torch.empty((count, lat.size, lon.size), dtype=torch.float64, device=dev)
    if count == 0:                                       # an empty shard (more ranks than start times): nothing to integrate
        return sigma if return_device else (out.numpy() if own_out else out)
    chunks = chunk_starts(0, count, chunk) if on_device else chunk_schedule(count, chunk, ramp=int(os.environ.get('LCS_ROLLING_RAMP', '1')))
    engine.reset_status()                                # chunks OR into one flag, checked once after the loop
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        pipe = _pipeline(engine, (max(n for _, n in chunks) + nsteps,) + tuple(u.shape[1:]), u.dtype, need_in=not on_device)
        up, down = pipe['up'], pipe['down']
        for i, (s, n) in enumerate(chunks):
            lo, hi = first + s, first + s + n + nsteps          # levels this chunk touches
            if on_device:
                du, dv = u[lo:hi], v[lo:hi]
            else:
                bu, bv = pipe['in'][i % 2]
                up.wait_event(pipe['in_free'][i % 2])            # kernels of chunk i-2 are done with this buffer
                with torch.cuda.stream(up):
                    du, dv = bu[:hi - lo], bv[:hi - lo]
                    du.copy_(u[lo:hi], non_blocking=True)
                    dv.copy_(v[lo:hi], non_blocking=True)
                main.wait_event(up.record_event())
            staged = engine.stage(du, dv, reuse=True)            # one chunk's levels alive at a time (stream-ordered)
            x, y = engine.advect(staged, nsteps=nsteps, nwindows=n, level0=0, level_stride=1)
            if not on_device:
                pipe['in_free'][i % 2] = main.record_event()     # the integrator's pole rows read the raw levels too
            sig = engine.epilogue(x, y, log_scale=log_scale)
            if on_chunk is not None:
                on_chunk(s, sig)
            if to_host:
                down.wait_event(main.record_event())
                with torch.cuda.stream(down):
                    out[s:s + n].copy_(sig, non_blocking=True)
                sig.record_stream(down)
            else:
                sigma[s:s + n] = sig
        if to_host:
            down.synchronize()
    engine.check_finite()
    if return_device:
        return sigma
    return out.numpy() if own_out else out


# ------------------------------------------------------------------ collectives (gather only)
def gather_fields(local, counts, group=None, dst=None):
    """All ranks contribute ``[count_r, nlat, nlon]``; returns the concatenation on every rank
    (``dst=None``, all_gather) or on ``dst`` only.  Works on NCCL (device tensors) and gloo (CPU)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    maxc = max(counts)
    pad = local
    if local.shape[0] < maxc:
        pad = torch.zeros((maxc,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    if dst is not None and dist.get_rank(group) != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def gather_bands(local_band, nrows, group=None):
    """Row bands ``[..., rows_r, nlon]`` -> full ``[..., nrows, nlon]`` on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [shard_rows(nrows, world, r)[1] - shard_rows(nrows, world, r)[0] for r in range(world)]
    moved = local_band.movedim(-2, 0).contiguous()
    full = gather_fields(moved, counts, group=group)
    return full.movedim(0, -2).contiguous()


XRANK_MAX_GROUPS = 16        # windows in flight per rank in the cross-rank outer-clamp mode (same on every rank)


def band_xrank(engine, world_size, nwindows, group=None):
    """The cross-rank mailboxes a row-band integration under the outer-product clamp needs (``engine.advect(xrank=)``),
    cached on the engine; ``None`` for one rank and for the cyclic / pointwise x-boundary (no exchange).  Collective on
    first use: every rank of ``group`` must call it with the same arguments."""
    from . import _lib
    if engine.xmode != _lib.LCS_X_CLAMP_OUTER or world_size <= 1:
        return None
    ngroups = min(nwindows, XRANK_MAX_GROUPS)
    cache = getattr(engine, '_xrank', None)
    if cache is None or cache.ngroups != ngroups or cache.world != world_size:
        from .peer import ColumnFlagMail
        cache = engine._xrank = ColumnFlagMail(engine.part_lon.size, ngroups, group=group, device=engine.device)
    return cache


def band_ftle(engine, staged, world_size, rank, nsteps=None, nwindows=1, level0=0, log_scale=False, group=None):
    """Row-band shard of one (or several) fields: integrate own rows + 2-row halo, run the epilogue
    on own rows.  Returns ``(sigma_band [nwindows, rows, nlon], (out0, out1))``.

    Cyclic / pointwise x-boundary: particles are independent, no exchange.  As-executed outer-product clamp (quirk Q6):
    the row flags of a band are local, the column flags are OR-ed over all bands after every sub-step through NVLink
    mailboxes inside the persistent kernel (``peer.ColumnFlagMail``; needs a process group on CUDA devices)."""
    out0, out1, in0, in1 = shard_rows(engine.nlat, world_size, rank)
    xrank = band_xrank(engine, world_size, nwindows, group)
    x, y = engine.advect(staged, nsteps=nsteps, nwindows=nwindows, level0=level0, rows=(in0, in1), xrank=xrank)
    sigma = engine.epilogue(x, y, log_scale=log_scale, in_row0=in0, out_rows=(out0, out1))
    return sigma, (out0, out1)


# ------------------------------------------------------------------ sharded drivers (one process per GPU)
def rolling_ftle_sharded(u, v, lat, lon, window_levels, timestep, group=None, dst=None, gather='nccl', **kw):
    """Start-time sharding of a rolling series over the ranks of ``group``: every rank integrates its contiguous
    block of windows (staging only the levels that block touches) and the finished fields are gathered on every rank:
    ``gather='nccl'`` one all_gather at the end; ``gather='p2p'`` each finished chunk is pushed into every rank's
    symmetric-memory buffer over NVLink while the next chunk is integrated (``peer.PeerFields``: copy engines, no SM
    taken from the integrator; CUDA + NCCL groups only).  ``kw`` as for :func:`rolling_ftle`.
    Returns ``[nstarts, nlat, nlon]`` on the device."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nstarts = u.shape[0] - window_levels + 1
    first, count = shard_starts(nstarts, world, rank)
    counts = [shard_starts(nstarts, world, r)[1] for r in range(world)]
    kw.pop('return_device', None)
    if gather == 'p2p':
        from .peer import PeerFields
        engine = kw.get('engine')
        device = engine.device if engine is not None else kw.get('device', 'cuda:0')
        peer = PeerFields(counts, len(lat), len(lon), group=group, device=device, dst=dst)
        rolling_ftle(u, v, lat, lon, window_levels, timestep, starts=(first, count), return_device=True,
                     on_chunk=peer.push, **kw)
        peer.finish()
        return peer.result() if dst is None or rank == dst else None
    if gather != 'nccl':
        raise ValueError("gather must be 'nccl' or 'p2p'")
    mine = rolling_ftle(u, v, lat, lon, window_levels, timestep, starts=(first, count), return_device=True, **kw)
    return gather_fields(mine, counts, group=group, dst=dst)


def ftle_row_bands(engine, u, v, group=None, log_scale=False, nwindows=None):
    """Row-band sharding of ONE field (or, ``nwindows=n``, of the ``n`` rolling windows of the series) over the ranks
    of ``group``: winds replicated, 2-row recomputed halo; under the outer-product clamp the column exit flags are
    exchanged between the ranks after every sub-step (``band_ftle``).  Returns the full ``[nlat, nlon]`` field
    (``[n, nlat, nlon]`` with ``nwindows``) on every rank."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = 1 if nwindows is None else int(nwindows)
    staged = engine.stage(u, v)
    band, _ = band_ftle(engine, staged, world, rank, nsteps=staged.nlev - n, nwindows=n, log_scale=log_scale, group=group)
    full = gather_bands(band, engine.nlat, group=group)
    return full[0] if nwindows is None else full
