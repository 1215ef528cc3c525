"""``parcel_propagation`` with the reference's signature (trajectory.py:8-18), executed by the
CUDA integrator ``lcs_advect``.  Labelled-array bookkeeping happens here on the host; every
per-particle operation runs on the device."""
from __future__ import annotations

import numpy as np
import pandas as pd

from ..engine import FtleEngine, precision_args
from ..labelled import coord_values, make_like

XCLAMP_DEFAULT = 'outer'     # what the reference executes for cyclic_xboundary=False (quirk Q6)


def _sorted_winds(U, V, propdim):
    """sortby longitude then latitude (trajectory.py:49-52) and a (time, lat, lon) view."""
    U = U.sortby('longitude').sortby('latitude')
    V = V.sortby('longitude').sortby('latitude')
    return U.transpose(propdim, 'latitude', 'longitude'), V.transpose(propdim, 'latitude', 'longitude')


_ENGINES = {}          # a handful of engines keyed by grid + recipe: repeated calls on the same grid (a rolling series driven
_ENGINES_MAX = 4       # through the reference-shaped API) reuse the device tables, workspaces and streams


def _cached_engine(lat, lon, timestep, SETTLS_order, interp_order, xmode, device, precision):
    key = (lat.tobytes(), lon.tobytes(), str(lat.dtype), str(lon.dtype), repr(timestep), type(timestep).__name__,
           int(SETTLS_order), int(interp_order), xmode, str(device), precision)
    eng = _ENGINES.pop(key, None)
    if eng is None:
        eng = FtleEngine(lat, lon, timestep, SETTLS_order=SETTLS_order, interp_order=interp_order,
                         xmode=xmode, device=device, **precision_args(precision))
        while len(_ENGINES) >= _ENGINES_MAX:
            _ENGINES.pop(next(iter(_ENGINES)))
    _ENGINES[key] = eng                                 # most recently used last
    return eng


def propagate(U, V, timestep, propdim, return_traj, SETTLS_order, interp_order, cyclic_xboundary,
              xclamp=XCLAMP_DEFAULT, device='cuda:0', precision='f64', engine=None, resample=None, rows=None):
    """Shared by parcel_propagation and LCS.__call__: returns device tensors plus the metadata the
    callers need to label them.  ``rows(lat) -> (r0, r1) or None``: restrict the integration to a band of particle rows of
    the SORTED grid (subdomain work skipping; only where particles are independent, i.e. not under the outer clamp)."""
    U, V = _sorted_winds(U, V, propdim)
    lat = coord_values(U, 'latitude')
    lon = coord_values(U, 'longitude')
    tcoord = coord_values(U, propdim) if resample is None else resample[0]
    times = tcoord.tolist()                                        # trajectory.py:58
    if timestep < 0:
        times.reverse()                                            # labels only (quirk Q2), :59-60
    xmode = 'cyclic' if cyclic_xboundary else xclamp
    if engine is None:
        engine = _cached_engine(lat, lon, timestep, SETTLS_order, interp_order, xmode, device, precision)
    # winds the global path already regridded / truncated on the device stay there (labelled.DeviceArray)
    uu = getattr(U, '_device_values', None)
    vv = getattr(V, '_device_values', None)
    if uu is None or vv is None:
        uu, vv = np.asarray(U.values), np.asarray(V.values)
    # resample= (LCS.py:88-90) is applied on the device inside the staging: coarse levels are prefiltered once, winds and
    # coefficients are refined linearly (engine.stage)
    staged = engine.stage(uu, vv, resample=None if resample is None else tuple(resample[1:]))
    band = rows(lat) if (rows is not None and xmode != 'outer') else None
    out = engine.advect(staged, return_traj=return_traj, rows=band)
    return engine, out, U, lat, lon, times, band


def parcel_propagation(U, V, timestep=1, propdim='time', verbose=True, return_traj=False,
                       SETTLS_order=0, copy=False, interp_order=3, cyclic_xboundary=False,
                       *, xclamp=XCLAMP_DEFAULT, device='cuda:0', precision='f64'):
    """Lagrangian 2-time-level advection (drop-in for trajectory.py:8-144).

    Same arguments, defaults and return convention as the reference: final
    ``(positions_x, positions_y)`` on the arrival grid carrying the scalar coordinate
    ``propdim = times[-1]``, or with ``return_traj`` the ``(propdim, latitude, longitude)`` stacks
    (level 0 = the start grid) labelled ``pd.to_datetime(times)`` -- ``times`` reversed when
    ``timestep < 0`` (trajectory.py:58-60,138-142).  ``copy`` is accepted for compatibility: inputs
    are never mutated.  Keyword-only extras select the engine: ``xclamp`` ('outer' = as executed,
    or 'pointwise'), ``device``, ``precision`` ('f64' | 'f32' packed-wind storage).
    """
    verboseprint = print if verbose else (lambda *a, **k: None)
    _, out, Us, lat, lon, times, _ = propagate(U, V, timestep, propdim, return_traj, SETTLS_order, interp_order,
                                               cyclic_xboundary, xclamp, device, precision)
    for t in times[:-1]:
        verboseprint(f'Propagating time {t}')                       # trajectory.py:81
    if return_traj:
        xt, yt = out[2][0].cpu().numpy(), out[3][0].cpu().numpy()
        tcoord = np.asarray(pd.to_datetime(times))                  # trajectory.py:138
        coords = {propdim: tcoord, 'latitude': lat, 'longitude': lon}
        dims = (propdim, 'latitude', 'longitude')
        return make_like(U, xt, dims, coords), make_like(U, yt, dims, coords)
    x, y = out[0][0].cpu().numpy(), out[1][0].cpu().numpy()
    coords = {'latitude': lat, 'longitude': lon, propdim: np.asarray(times[-1])}   # trajectory.py:141-142
    dims = ('latitude', 'longitude')
    return make_like(U, x, dims, coords), make_like(U, y, dims, coords)
