"""``LCS`` and ``flowmap_gradient`` with the reference's signatures (LCS.py:19-225), executed by the
CUDA engine: ``lcs_prefilter`` + ``lcs_pack_pairs`` + ``lcs_advect`` + ``lcs_ftle_epilogue``."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import FtleEngine
from ..labelled import coord_values, is_dataset, is_xarray, make_like
from .trajectory import propagate, XCLAMP_DEFAULT

DERIVATIVE_NAMES = ['dxdx', 'dxdy', 'dydx', 'dydy', 'dzdx', 'dzdy', 'dxdr', 'dydr', 'dzdr']   # LCS.py:210-218


def _crop_index(coord, sl):
    """Strict-inequality crop (tools.py:184-186); the ``xr_tools.latlonsel`` the reference imports
    (LCS.py:13) is not in its tree, its in-tree analogue is the model."""
    keep = np.ones(coord.shape, bool)
    if sl is not None:
        if sl.start is not None:
            keep &= coord > sl.start
        if sl.stop is not None:
            keep &= coord < sl.stop
    return keep


class LCS:
    """API to compute the finite-time Lyapunov exponent field of 2-D winds (drop-in for LCS.py:19-168).

    ``LCS(timestep, timedim, SETTLS_order, subdomain, return_dpts, gauss_sigma)`` then
    ``lcs(ds)`` or ``lcs(u=u, v=v)``.  Returns the largest singular value of the as-executed
    3x3 (quirk Q7) as a ``(timedim=1, latitude, longitude)`` array; callers apply
    ``0.5*np.log(.)`` themselves, as with the reference (examples/ideal_vortex.py:282).
    """
    earth_r = 6371000  # metres

    def __init__(self, timestep=1, timedim='time', SETTLS_order=0, subdomain=None, return_dpts=False,
                 gauss_sigma=None):
        self.timestep = timestep
        self.SETTLS_order = SETTLS_order
        self.timedim = timedim
        self.subdomain = subdomain
        self.gauss_sigma = gauss_sigma
        self.return_dpts = return_dpts

    def __call__(self, ds=None, u=None, v=None, verbose=True, s=None, resample=None, s_is_error=False,
                 isglobal=False, return_traj=False, interp_to_common_grid=True, traj_interp_order=3,
                 truncation=20, *, xclamp=XCLAMP_DEFAULT, device='cuda:0', precision='f64'):
        print('!' * 100)                                             # LCS.py:74 (unconditional upstream)
        verboseprint = print if verbose else (lambda *a, **k: None)
        timestep, timedim = self.timestep, self.timedim
        self.verbose = verbose
        if is_dataset(ds):                                           # LCS.py:81-83
            u, v = ds.u.copy(), ds.v.copy()
        elif isinstance(ds, str):                                    # LCS.py:84-87
            from ..ncio import open_dataset                            # xarray when installed, else classic NetCDF-3 via scipy
            ds = open_dataset(ds)
            u, v = ds.u.copy(), ds.v.copy()
        resample_plan_ = None
        if isinstance(resample, str):                                # LCS.py:88-91: linear refinement in time
            from ..timeaxis import resample_plan
            resample_plan_ = resample_plan(coord_values(u, timedim), resample)
            new_t = resample_plan_[0]
            timestep = np.sign(timestep) * (new_t[1] - new_t[0]).astype('timedelta64[s]').astype('float')   # LCS.py:91
        assert set(u.dims) == set(v.dims), "u and v dims are different"                      # LCS.py:95
        assert set(u.dims) == {'latitude', 'longitude', timedim}, \
            'array dims should be latitude and longitude only'                               # LCS.py:96
        if isglobal:                                                 # LCS.py:105-120
            if interp_to_common_grid:                                # LCS.py:106-114
                u, v = (_to_common_grid(a, timedim, device) for a in (u, v))
            if truncation is not None:                               # LCS.py:115-118: VectorWind(u, v).truncate(., truncation)
                u, v = (_truncate(a, timedim, truncation, device) for a in (u, v))
            cyclic_xboundary = True
            self.subdomain = None
        else:
            cyclic_xboundary = False
        if s is None:                                                # LCS.py:124-126 (printed, never used)
            u0 = np.asarray(u.isel({timedim: 0}).values)
            s = int(10 * u0.size * u0.std())
            print('using s = ' + str(s / 1e6) + '1e6')
        verboseprint("*---- Parcel propagation ----*")
        # Subdomain work skipping: the crop keeps rows [o0, o1) of the sorted grid and the y-stencil reaches
        # two rows beyond them, so where particles are independent (pointwise clamp; the as-executed outer clamp couples all
        # rows) and neither the departure points nor their Gaussian smoothing are asked for, only rows [o0-2, o1+2) are
        # integrated.  The band is decided inside propagate(), on the sorted latitudes.
        sub = self.subdomain if isinstance(self.subdomain, dict) else None
        skip_ok = sub is not None and not (self.return_dpts or return_traj) and not isinstance(self.gauss_sigma, (float, int))

        def band_of(lat_sorted):
            keep = np.flatnonzero(_crop_index(lat_sorted, sub.get('latitude')))
            if not keep.size:
                return None
            return max(0, int(keep[0]) - 2), min(lat_sorted.size, int(keep[-1]) + 3)
        engine, out, Us, lat, lon, times, band = propagate(u, v, timestep, timedim, return_traj, self.SETTLS_order,
                                                           traj_interp_order, cyclic_xboundary, xclamp, device, precision,
                                                           resample=resample_plan_, rows=band_of if skip_ok else None)
        x_dep, y_dep = out[0], out[1]
        in_row0 = band[0] if band is not None else 0
        verboseprint("*---- Computing deformation tensor ----*")
        xs, ys = x_dep, y_dep
        if isinstance(self.gauss_sigma, (float, int)):               # LCS.py:187-190 (inside flowmap_gradient upstream)
            xs, ys = engine.gaussian(x_dep, self.gauss_sigma), engine.gaussian(y_dep, self.gauss_sigma)
        latkeep = lonkeep = None
        out_rows = mask = None
        if isinstance(self.subdomain, dict):                         # LCS.py:143-144 (crop after the derivatives)
            latkeep = _crop_index(lat, self.subdomain.get('latitude'))
            lonkeep = _crop_index(lon, self.subdomain.get('longitude'))
            rows = np.flatnonzero(latkeep)
            if rows.size:
                out_rows = (int(rows[0]), int(rows[-1]) + 1)         # only these rows are computed ...
                # ... and only the kept points are checked for inf: upstream crops before dropna / norm (LCS.py:143-154)
                mask = np.outer(latkeep[out_rows[0]:out_rows[1]], lonkeep)
        verboseprint("*---- Computing eigenvalues ----*")
        engine.reset_status()
        sigma = engine.epilogue(xs, ys, out_rows=out_rows, mask=mask, in_row0=in_row0)
        engine.check_finite()                                        # ValueError on inf, as scipy.linalg.norm (LCS.py:154)
        sigma = sigma[0].cpu().numpy()
        verboseprint("*---- Done eigenvalues ----*")
        olat, olon = lat, lon
        if latkeep is not None:
            r0 = out_rows[0] if out_rows else 0
            sigma = sigma[latkeep[r0:r0 + sigma.shape[0]]][:, lonkeep] if out_rows else sigma[:0, :0]
            olat, olon = lat[latkeep], lon[lonkeep]
        sigma, olat, olon = drop_unused_levels(sigma, olat, olon)    # LCS.py:146,157: dropna('points') ... unstack('points')
        tvals = coord_values(Us, timedim) if resample_plan_ is None else resample_plan_[0]
        timestamp = tvals[-1] if np.sign(timestep) == 1 else tvals[0]                         # LCS.py:158
        coords = {'latitude': olat, 'longitude': olon, 'time': np.asarray(timestamp)}        # LCS.py:159
        if timedim == 'time':
            coords['time'] = np.asarray(timestamp)[None]
        eigenvalues = make_like(u, sigma[None], (timedim, 'latitude', 'longitude'), coords)  # LCS.py:160
        dims2 = ('latitude', 'longitude')
        c2 = {'latitude': lat, 'longitude': lon, timedim: np.asarray(times[-1])}
        xd = make_like(u, x_dep[0].cpu().numpy(), dims2, c2) if (self.return_dpts or return_traj) else None
        yd = make_like(u, y_dep[0].cpu().numpy(), dims2, c2) if (self.return_dpts or return_traj) else None
        if return_traj:
            import pandas as pd
            tc = {timedim: np.asarray(pd.to_datetime(times)), 'latitude': lat, 'longitude': lon}
            d3 = (timedim, 'latitude', 'longitude')
            x_trajs = make_like(u, out[2][0].cpu().numpy(), d3, tc)
            y_trajs = make_like(u, out[3][0].cpu().numpy(), d3, tc)
        if self.return_dpts and return_traj:                         # LCS.py:161-168
            return eigenvalues, xd, yd, x_trajs, y_trajs
        elif self.return_dpts:
            return eigenvalues, xd, yd
        elif return_traj:
            return eigenvalues, x_trajs, y_trajs
        return eigenvalues


def _to_common_grid(da, timedim, device):
    """LCS.py:101-114 for one component: sort, then ``interp(linear)`` to the 360 x 721 grid with the NaNs filled from
    ``reindex(nearest)`` -- one device kernel (engine.regrid_device).  The result stays on the device (DeviceArray)."""
    from ..engine import regrid_device
    from ..labelled import DeviceArray
    from ..regrid import common_grid
    d = da.sortby('latitude').sortby('longitude').transpose(timedim, 'latitude', 'longitude')
    lats, lons = common_grid()
    src = getattr(d, '_device_values', None)
    out = regrid_device(d.values if src is None else src, coord_values(d, 'latitude'), coord_values(d, 'longitude'), lats, lons, device=device)
    coords = {timedim: coord_values(d, timedim), 'latitude': lats, 'longitude': lons}
    if is_xarray(da):                    # pragma: no cover - needs xarray: xarray in, xarray out
        return make_like(da, out.cpu().numpy(), (timedim, 'latitude', 'longitude'), coords)
    return DeviceArray(out, (timedim, 'latitude', 'longitude'), coords)


def _truncate(da, timedim, truncation, device):
    """LCS.py:115-118 for one component: windspharm's triangular truncation (pyspharm / SPHEREPACK regular grid) as
    one device operator (engine.spectral_truncate_device; spectral.py describes it and why its parity is unpinned).
    windspharm refuses grids that are not global and equally spaced with a ValueError; so does this."""
    from ..engine import spectral_truncate_device
    from ..labelled import DeviceArray
    from ..spectral import check_regular_global_grid
    d = da.sortby('latitude').sortby('longitude').transpose(timedim, 'latitude', 'longitude')
    lat, lon = coord_values(d, 'latitude'), coord_values(d, 'longitude')
    check_regular_global_grid(lat, lon)
    src = getattr(d, '_device_values', None)
    out = spectral_truncate_device(d.values if src is None else src, int(truncation), device=device)
    coords = {timedim: coord_values(d, timedim), 'latitude': lat, 'longitude': lon}
    if is_xarray(da):                    # pragma: no cover
        return make_like(da, out.cpu().numpy(), (timedim, 'latitude', 'longitude'), coords)
    return DeviceArray(out, (timedim, 'latitude', 'longitude'), coords)


def drop_unused_levels(sigma, lat, lon):
    """``def_tensor.dropna('points')`` followed by ``unstack('points')`` (LCS.py:146,157): points with a NaN
    derivative come back as NaN, but a latitude (longitude) none of whose points survived is no longer a level of
    the stacked index and disappears from the result altogether (xarray removes unused levels when unstacking)."""
    nan = np.isnan(sigma)
    if not nan.any():
        return sigma, lat, lon
    keep_r, keep_c = ~nan.all(axis=1), ~nan.all(axis=0)
    if keep_r.all() and keep_c.all():
        return sigma, lat, lon
    return sigma[keep_r][:, keep_c], np.asarray(lat)[keep_r], np.asarray(lon)[keep_c]


def flowmap_gradient(x_departure, y_departure, sigma=None, *, device='cuda:0'):
    """The nine stacked 'derivatives' ``(derivatives, latitude, longitude)`` of LCS.py:171-225."""
    # derivative_spherical_coords sorts by latitude, then longitude (tools.py:251-252)
    xd = x_departure.sortby('latitude').sortby('longitude').transpose('latitude', 'longitude')
    yd = y_departure.sortby('latitude').sortby('longitude').transpose('latitude', 'longitude')
    lat, lon = coord_values(xd, 'latitude'), coord_values(xd, 'longitude')
    engine = FtleEngine(lat, lon, 1, device=device)
    dev = torch.device(device)
    tx = torch.from_numpy(np.ascontiguousarray(xd.values, dtype=np.float64)).to(dev)
    ty = torch.from_numpy(np.ascontiguousarray(yd.values, dtype=np.float64)).to(dev)
    if isinstance(sigma, (float, int)):                              # LCS.py:187-190
        tx, ty = engine.gaussian(tx, sigma), engine.gaussian(ty, sigma)
    _, jac = engine.epilogue(tx, ty, return_jac=True)
    jac = jac[0]
    full = torch.cat([jac, torch.zeros((3,) + tuple(jac.shape[1:]), dtype=jac.dtype, device=dev)])   # LCS.py:206-208
    coords = {'derivatives': np.array(DERIVATIVE_NAMES), 'latitude': lat, 'longitude': lon}
    return make_like(x_departure, full.cpu().numpy(), ('derivatives', 'latitude', 'longitude'), coords)


def main(argv):
    """The reference's command line (LCS.py:236-265): ``timestep timedim SETTLS_order subdomain ds_path outpath return_traj``.

    As upstream: the subdomain argument (``lon0/lon1/lat0/lat1``) is parsed and then NOT used (``subdomain=None``,
    LCS.py:246-247); the call is the global one (regrid to 360 x 721, T20 truncation, cubic); with ``return_traj`` the
    trajectories are saved next to the field under the names upstream derives (``SL_attracting`` -> ``x_departure`` /
    ``y_departure``); and the INPUT FILE IS REMOVED afterwards (LCS.py:265) -- the caller upstream is a job script that
    writes one partial input per task.  An eighth argument ``keep`` (extension) leaves the input in place."""
    import os
    print('*----- ARGS ------*')
    print(argv)
    coords = str(argv[4]).split('/')
    subdomain = {'longitude': slice(float(coords[0]), float(coords[1])),            # parsed, unused: as upstream
                 'latitude': slice(float(coords[2]), float(coords[3]))}
    del subdomain
    lcs = LCS(timestep=float(argv[1]), timedim=str(argv[2]), SETTLS_order=int(argv[3]), subdomain=None)
    input_path, outpath = str(argv[5]), str(argv[6])
    return_traj = argv[7] == 'True'
    if return_traj:
        out, x_departure, y_departure = lcs(ds=input_path, isglobal=True, interp_to_common_grid=True, truncation=20,
                                            traj_interp_order=3, return_traj=return_traj)
        print('Saving to ' + outpath)
        out.to_netcdf(outpath)
        x_departure.to_netcdf(outpath.replace('SL_attracting', 'x_departure'))
        y_departure.to_netcdf(outpath.replace('SL_attracting', 'y_departure'))
    else:
        out = lcs(ds=input_path, isglobal=True, interp_to_common_grid=True, truncation=20,
                  traj_interp_order=3, return_traj=return_traj)
        print('Saving to ' + outpath)
        out.to_netcdf(outpath)
    if not (len(argv) > 8 and argv[8] == 'keep'):
        os.remove(input_path)                                                        # LCS.py:265: subprocess.call(['rm', input_path])


if __name__ == '__main__':
    import sys
    main(sys.argv)
