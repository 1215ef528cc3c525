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


def _dsc_device(t, lat, lon, dim, isglobal, device):
    """derivative_spherical_coords on a device tensor (lat, lon): f32 cast, stencil kernel, f64 division."""
    import torch
    lib = _engine._lib.load()
    a = t.float().contiguous()                                                    # tools.py:258
    out = torch.empty_like(a)
    _engine._lib.check(lib.lcs_fourth_order_derivative(_engine._ptr(a), a.shape[0], a.shape[1], int(dim),
                                                       int(bool(isglobal)), _engine._ptr(out),
                                                       _engine._stream(a.device)), 'lcs_fourth_order_derivative')
    y = lat * np.pi / 180
    if dim == 0:
        dy = (np.pi / 180) * (lat[1] - lat[0]) * EARTH_RADIUS
        return out.double() / torch.tensor(float(dy), dtype=torch.float64, device=a.device)
    dx = (np.pi / 180) * (lon[1] - lon[0]) * EARTH_RADIUS * np.cos(y)
    return out.double() / torch.from_numpy(np.ascontiguousarray(dx, dtype=np.float64)).to(a.device)[:, None]


def find_ridges_spherical_hessian(da, sigma=.5, scheme='first_order', tolerance_threshold=0.0005e-3,
                                  return_eigvectors=False, isglobal=True, *, device='cuda:0'):
    """Hessian ridge filter of an FTLE field in spherical coordinates (tools.py:52-155), as executed upstream:
    Gaussian smoothing on the (longitude, latitude) transpose, five stencil passes that each re-cast to f32, the
    Hessian's inf/NaN zeroed, and a per-point 2x2 eigen-decomposition whose eigenvector ROW (sic, tools.py:108) is
    dotted with the gradient.  Returns ``(dt_prod, eigmin)`` in the input's dimension order; ``scheme`` is accepted
    and ignored, as upstream.  ``return_eigvectors=True`` returns the six arrays of tools.py:148-152:
    ``dt_prod, eigmin, dt_prod_`` (the unthresholded dot product), ``eigvectors`` (dim ``eigvectors``; zero where
    eigmin >= 0), ``gradient`` (dim ``elements``) and ``angle = 180/pi * arctan(e0/e1)``."""
    import torch
    dims_in = tuple(da.dims)
    d2 = da.sortby('latitude').sortby('longitude').transpose('latitude', 'longitude')
    lat, lon = coord_values(d2, 'latitude'), coord_values(d2, 'longitude')
    dev = torch.device(device)
    eng = _engine.FtleEngine(lat, lon, 1, device=dev)
    with torch.cuda.device(dev):
        f = torch.from_numpy(np.ascontiguousarray(d2.values, dtype=np.float64)).to(dev)
        if isinstance(sigma, (float, int)):                                           # tools.py:75-76, (lon, lat) layout
            f = eng.gaussian(f.t().contiguous(), sigma).t().contiguous()
        ddadx = _dsc_device(f, lat, lon, 1, isglobal, dev)                            # tools.py:78-82
        ddady = _dsc_device(f, lat, lon, 0, isglobal, dev)
        d2x2 = _dsc_device(ddadx, lat, lon, 1, isglobal, dev)
        d2y2 = _dsc_device(ddady, lat, lon, 0, isglobal, dev)
        d2xy = _dsc_device(ddadx, lat, lon, 0, isglobal, dev)
        dt_prod, eigmin = torch.empty_like(f), torch.empty_like(f)
        dt_raw = ev0 = ev1 = None
        if return_eigvectors:
            dt_raw, ev0, ev1 = torch.empty_like(f), torch.empty_like(f), torch.empty_like(f)
        lib = _engine._lib.load()
        ddadx, ddady = ddadx.contiguous(), ddady.contiguous()
        _engine._lib.check(lib.lcs_ridge_classify(*[_engine._ptr(t.contiguous()) for t in (d2x2, d2xy, d2y2, ddadx, ddady)],
                                                  f.numel(), float(tolerance_threshold), _engine._ptr(dt_prod),
                                                  _engine._ptr(eigmin), _engine._ptr(dt_raw), _engine._ptr(ev0),
                                                  _engine._ptr(ev1), _engine._stream(dev)), 'lcs_ridge_classify')
    coords = {'latitude': lat, 'longitude': lon}
    out = []
    for t in (dt_prod, eigmin):
        r = make_like(da, t.cpu().numpy(), ('latitude', 'longitude'), coords)
        out.append(r.transpose(*dims_in))
    if not return_eigvectors:
        return tuple(out)
    out.append(make_like(da, dt_raw.cpu().numpy(), ('latitude', 'longitude'), coords).transpose(*dims_in))   # dt_prod_, :129
    e0, e1, em = ev0.cpu().numpy(), ev1.cpu().numpy(), eigmin.cpu().numpy()
    with np.errstate(all='ignore'):
        angle = 180 / np.pi * np.arctan(e0 / e1)                                      # tools.py:125 (before the zeroing)
    evec = np.where(em < 0, np.stack([e0, e1]), 0.0)                                  # tools.py:132
    # labels as upstream leaves them: isel(elements=[1, 2]) of the Hessian renamed to 'eigvectors' (tools.py:123-124)
    ecoords = dict(coords, eigvectors=np.array(['d2dadxdy', 'd2dadydx']))
    out.append(make_like(da, evec, ('eigvectors', 'latitude', 'longitude'), ecoords).transpose('eigvectors', *dims_in))
    gcoords = dict(coords, elements=np.array(['ddadx', 'ddady']))
    grad = np.stack([ddadx.cpu().numpy(), ddady.cpu().numpy()])
    out.append(make_like(da, grad, ('elements', 'latitude', 'longitude'), gcoords).transpose('elements', *dims_in))
    out.append(make_like(da, angle, ('latitude', 'longitude'), coords).transpose(*dims_in))
    return tuple(out)


def latlonsel(array, lat, lon, latname='lat', lonname='lon'):
    """Crop to the open intervals ``lat[0] < latitude < lat[-1]``, ``lon[0] < longitude < lon[-1]`` (slice or list), the
    in-tree analogue of the ``xr_tools.latlonsel`` that LCS.py:143-144 calls (tools.py:158-188: two strict-inequality
    masks applied with ``where(mask, drop=True)``).  Pure label selection, done on the host."""
    assert latname in array.coords, f"Coord. {latname} not present in array"          # tools.py:168
    assert lonname in array.coords, f"Coord. {lonname} not present in array"          # tools.py:169
    lat1, lat2 = (lat.start, lat.stop) if isinstance(lat, slice) else (lat[0], lat[-1])
    lon1, lon2 = (lon.start, lon.stop) if isinstance(lon, slice) else (lon[0], lon[-1])
    lonv, latv = coord_values(array, lonname), coord_values(array, latname)
    lonmask = (lonv < lon2) & (lonv > lon1)                                           # tools.py:184
    latmask = (latv < lat2) & (latv > lat1)                                           # tools.py:185
    return array.isel({lonname: np.flatnonzero(lonmask)}).isel({latname: np.flatnonzero(latmask)})
