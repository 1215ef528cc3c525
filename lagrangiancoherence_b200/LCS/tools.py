"""Array-level seams with the reference's names (tools.py:11-48,190-267), on the device."""
from __future__ import annotations

import numpy as np

from .. import engine as _engine
from ..labelled import coord_values, make_like

EARTH_RADIUS = 6371000      # tools.py:249


def xr_map_coordinates(da, new_x, new_y, isglobal=True, order=1, *, device='cuda:0'):
    """Sample ``da(latitude, longitude)`` at positions in degrees (tools.py:11-41).

    Only the ``isglobal=True`` branch exists: the reference's ``else`` branch references an
    undefined name (tools.py:47) and is never taken."""
    if not isglobal:
        raise NameError("name 'da_except_poles' is not defined")     # what the reference raises, tools.py:47
    da = da.transpose('latitude', 'longitude')
    lat, lon = coord_values(da, 'latitude'), coord_values(da, 'longitude')
    nx = np.asarray(new_x.values if hasattr(new_x, 'values') else new_x, dtype=np.float64)
    ny = np.asarray(new_y.values if hasattr(new_y, 'values') else new_y, dtype=np.float64)
    out = _engine.map_coordinates_device(np.asarray(da.values), nx, ny, lat, lon, order=order, device=device)
    out = out.cpu().numpy().astype(np.asarray(da.values).dtype, copy=False)   # scipy answers in the input dtype
    return make_like(da, out, ('latitude', 'longitude'), {'latitude': lat, 'longitude': lon})


def fourth_order_derivative(arr, dim=0, isglobal=True, *, device='cuda:0'):
    """4th-order centred / halved one-sided stencil on an f32 ``[lat, lon]`` array (tools.py:190-245)."""
    return _engine.fourth_order_derivative_device(arr, dim=dim, isglobal=isglobal, device=device).cpu().numpy()


def derivative_spherical_coords(da, dim=0, isglobal=True, *, device='cuda:0'):
    """Stencil in index space on f32, then divide by the metric spacing (tools.py:248-267)."""
    if dim not in (0, 1):
        raise ValueError('Dim must be either 0 or 1.')
    da = da.sortby('latitude').sortby('longitude').transpose('latitude', 'longitude')
    lat, lon = coord_values(da, 'latitude'), coord_values(da, 'longitude')
    y = lat * np.pi / 180
    dx = (np.pi / 180) * (lon[1] - lon[0]) * EARTH_RADIUS * np.cos(y)
    dy = (np.pi / 180) * (lat[1] - lat[0]) * EARTH_RADIUS
    import torch
    deriv = _engine.fourth_order_derivative_device(np.asarray(da.values).astype('float32'), dim=dim,
                                                   isglobal=isglobal, device=device).double()   # f32 stencil, tools.py:258
    if dim == 0:
        deriv = deriv / torch.tensor(float(dy), dtype=torch.float64, device=deriv.device)      # tools.py:262 (true division)
    else:
        deriv = deriv / torch.from_numpy(np.ascontiguousarray(dx, dtype=np.float64)).to(deriv.device)[:, None]  # :264
    return make_like(da, deriv.cpu().numpy(), ('latitude', 'longitude'), {'latitude': lat, 'longitude': lon})
