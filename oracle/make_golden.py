#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference/LCS/*.py).

TEST INFRASTRUCTURE.  Runs only in the build container (the GPU box has no /root/reference); the
fixtures it writes are committed and are what tests/test_oracle_golden.py and the -m gpu parity
tests read.

How the reference is made importable here (nothing under /root/reference is modified or copied):
  * its modules import ``LagrangianCoherence.LCS.*`` (LCS.py:12,15; trajectory.py:4), so a temporary
    directory holding a symlink ``LagrangianCoherence -> /root/reference`` is put on sys.path;
  * the imports this image lacks (xarray, dask, xr_tools, IPython, windspharm, cftime) resolve to
    the stand-ins in oracle/refshim/ -- see oracle/refshim/xarray/__init__.py for what that means
    for the strength of the pin.

Inputs are regenerated from seeds by lagrangiancoherence_b200.synthetic, so a fixture stores only the
case description and the reference's outputs.

    python oracle/make_golden.py                      # rewrites tests/golden/
    python oracle/make_golden.py --only a,b,seams     # only these fixtures (the others keep their bytes)
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f'{REF} is not present: goldens can only be regenerated in the build container')
    tmp = tempfile.mkdtemp(prefix='lcs_ref_')
    os.symlink(REF, os.path.join(tmp, 'LagrangianCoherence'))
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'refshim'))
    sys.path.insert(0, tmp)
    sys.path.insert(0, ROOT)
    import xarray as xr                                  # the stand-in
    assert xr.__version__.endswith('refshim')
    from LagrangianCoherence.LCS import LCS as ref_lcs, trajectory as ref_traj, tools as ref_tools
    return xr, ref_lcs, ref_traj, ref_tools


# case name -> description; winds come from synthetic.era5_like_winds / ideal_vortex with these arguments
CASES = {
    'regional_outer_p3': dict(kind='era5', nlat=41, nlon=57, lat=(-30.0, 10.0), lon=(-80.0, -24.0), nt=5, seed=0,
                              timestep=-21600, S=4, order=3, cyclic=False),
    'regional_outer_p1': dict(kind='era5', nlat=41, nlon=57, lat=(-30.0, 10.0), lon=(-80.0, -24.0), nt=5, seed=1,
                              timestep=-21600, S=4, order=1, cyclic=False),
    'regional_forward_S2': dict(kind='era5', nlat=33, nlon=45, lat=(-20.0, 12.0), lon=(-70.0, -26.0), nt=4, seed=2,
                                timestep=10800, S=2, order=3, cyclic=False),
    'regional_contained': dict(kind='era5', nlat=41, nlon=57, lat=(-30.0, 10.0), lon=(-80.0, -24.0), nt=5, seed=3,
                               timestep=-21600, S=4, order=3, cyclic=False, contained=True, scale=0.3),
    'regional_S0': dict(kind='era5', nlat=25, nlon=31, lat=(-10.0, 14.0), lon=(-60.0, -30.0), nt=3, seed=4,
                        timestep=-3600, S=0, order=3, cyclic=False),
    'cyclic_vortex_backward': dict(kind='vortex', timestep=-21600, S=4, order=3, cyclic=True),
    'cyclic_vortex_forward': dict(kind='vortex', timestep=21600, S=2, order=3, cyclic=True),
    # traj_interp_order other than the default 3 and 1 (any order scipy accepts works upstream)
    'regional_outer_p2': dict(kind='era5', nlat=33, nlon=45, lat=(-20.0, 12.0), lon=(-70.0, -26.0), nt=4, seed=6,
                              timestep=-21600, S=3, order=2, cyclic=False),
    'regional_outer_p4': dict(kind='era5', nlat=33, nlon=45, lat=(-20.0, 12.0), lon=(-70.0, -26.0), nt=4, seed=7,
                              timestep=-21600, S=2, order=4, cyclic=False),
    'regional_outer_p5': dict(kind='era5', nlat=33, nlon=45, lat=(-20.0, 12.0), lon=(-70.0, -26.0), nt=4, seed=8,
                              timestep=10800, S=3, order=5, cyclic=False),
    # f32 winds on f64 coordinates (how ERA5 is stored): scipy answers in the input dtype and numpy's promotion rules
    # (NEP 50) then decide the precision of every increment (tools.py:26-30, trajectory.py:86-87,110-112)
    'regional_f32_winds': dict(kind='era5', nlat=41, nlon=57, lat=(-30.0, 10.0), lon=(-80.0, -24.0), nt=5, seed=9,
                               timestep=-21600, S=4, order=3, cyclic=False, dtype='float32'),
    'regional_f32_winds_p1': dict(kind='era5', nlat=33, nlon=45, lat=(-20.0, 12.0), lon=(-70.0, -26.0), nt=4, seed=10,
                                  timestep=10800, S=2, order=1, cyclic=False, dtype='float32'),
    'descending_lat_dims_shuffled': dict(kind='era5', nlat=29, nlon=37, lat=(-28.0, 0.0), lon=(-72.0, -36.0), nt=4, seed=5,
                                         timestep=-21600, S=3, order=3, cyclic=False, flip_lat=True,
                                         dims=['latitude', 'time', 'longitude']),
}


def make_inputs(case):
    from lagrangiancoherence_b200 import synthetic as S
    if case['kind'] == 'vortex':
        u, v, lat, lon = S.ideal_vortex(**S.vortex_config_subtropical)
    else:
        lat = np.linspace(*case['lat'], case['nlat'])
        lon = np.linspace(*case['lon'], case['nlon'])
        u, v = S.era5_like_winds(lat, lon, case['nt'], seed=case['seed'], contained=case.get('contained', False))
        u, v = u * case.get('scale', 1.0), v * case.get('scale', 1.0)
        if case.get('dtype'):
            u, v = u.astype(case['dtype']), v.astype(case['dtype'])
    time = (np.datetime64('2000-01-01T00') + np.arange(u.shape[0]) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
    return u, v, lat, lon, time


def to_xr(xr, case, u, v, lat, lon, time):
    coords = {'time': time, 'latitude': lat, 'longitude': lon}
    du = xr.DataArray(u, coords, ('time', 'latitude', 'longitude'), name='u')
    dv = xr.DataArray(v, coords, ('time', 'latitude', 'longitude'), name='v')
    if case.get('flip_lat'):
        idx = np.arange(lat.size)[::-1]
        du, dv = du.isel(latitude=idx), dv.isel(latitude=idx)
    if case.get('dims'):
        du, dv = du.transpose(*case['dims']), dv.transpose(*case['dims'])
    return du, dv


def main():
    xr, ref_lcs, ref_traj, ref_tools = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    quiet = io.StringIO()
    manifest = {}
    only = None
    if '--only' in sys.argv:
        only = set(sys.argv[sys.argv.index('--only') + 1].split(','))
        manifest = json.load(open(os.path.join(GOLDEN, 'manifest.json')))
    for name, case in CASES.items():
        if only is not None and name not in only:
            continue
        u, v, lat, lon, time = make_inputs(case)
        du, dv = to_xr(xr, case, u, v, lat, lon, time)
        out = {}
        with contextlib.redirect_stdout(quiet):
            # a3: parcel_propagation, trajectories (trajectory.py:8-144).  The reference indexes `time`
            # positionally after isel(time=0) (trajectory.py:76), so it needs (time, latitude, longitude) order.
            if not case.get('dims'):
                xt, yt = ref_traj.parcel_propagation(du, dv, timestep=case['timestep'], propdim='time', verbose=False,
                                                     return_traj=True, SETTLS_order=case['S'], copy=True,
                                                     interp_order=case['order'], cyclic_xboundary=case['cyclic'])
                levels = np.arange(xt.shape[0]) if case['kind'] != 'vortex' else np.array([0, 3, xt.shape[0] - 1])
                out['traj_levels'] = levels                       # the big case keeps three levels to stay small
                out['x_traj'], out['y_traj'] = xt.values[levels], yt.values[levels]
                out['traj_time'] = np.asarray(xt.coords['time']).astype('datetime64[ns]').astype('int64')
            # a2: LCS.__call__ (LCS.py:48-168); isglobal only flips cyclic_xboundary when regrid/truncation are off
            if True:
                lcs = ref_lcs.LCS(timestep=case['timestep'], timedim='time', SETTLS_order=case['S'], return_dpts=True)
                eig, xd, yd = lcs(u=du, v=dv, verbose=False, isglobal=case['cyclic'], interp_to_common_grid=False,
                                  truncation=None, traj_interp_order=case['order'])
                out['sigma'] = eig.values
                out['sigma_time'] = np.asarray(eig.coords['time']).astype('datetime64[ns]').astype('int64')
                out['sigma_lat'], out['sigma_lon'] = eig.coords['latitude'], eig.coords['longitude']
                out['x_dep'], out['y_dep'] = xd.values, yd.values
                # a6: flowmap_gradient on the reference's own departure points (LCS.py:171-225)
                if case['kind'] != 'vortex':
                    dt = ref_lcs.flowmap_gradient(xd, yd)
                    out['def_tensor'] = dt.values
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        manifest[name] = case
        print(f'{name}: ' + ', '.join(f'{k}{tuple(np.shape(v))}' for k, v in out.items()))

    if only is not None and 'seams' not in only:
        with open(os.path.join(GOLDEN, 'manifest.json'), 'w') as f:
            json.dump(manifest, f, indent=1, sort_keys=True)
        return
    # seams: xr_map_coordinates, derivative_spherical_coords, fourth_order_derivative, subdomain crop
    case = CASES['regional_outer_p3']
    u, v, lat, lon, time = make_inputs(case)
    du, dv = to_xr(xr, case, u, v, lat, lon, time)
    rng = np.random.default_rng(11)
    X, Y = np.meshgrid(lon, lat)
    px, py = X + rng.normal(0, 3.0, X.shape), Y + rng.normal(0, 3.0, Y.shape)
    seams = {'px': px, 'py': py}
    for order in (1, 2, 3, 4, 5):
        seams[f'map_coordinates_p{order}'] = ref_tools.xr_map_coordinates(du.isel(time=0), px, py, order=order).values
    c2 = {'latitude': lat, 'longitude': lon}
    Xs = xr.DataArray(6371000 * np.sin((py - 90) * np.pi / 180) * np.cos(px * np.pi / 180), c2, ('latitude', 'longitude'))
    seams['X'] = Xs.values
    for dim in (0, 1):
        seams[f'derivative_spherical_dim{dim}'] = ref_tools.derivative_spherical_coords(Xs, dim=dim).values
        for isglobal in (True, False):
            seams[f'fourth_order_dim{dim}_global{int(isglobal)}'] = ref_tools.fourth_order_derivative(
                Xs.values.astype('float32'), dim=dim, isglobal=isglobal)
    sub = {'latitude': slice(-20, 0), 'longitude': slice(-70, -40)}
    with contextlib.redirect_stdout(quiet):
        eig = ref_lcs.LCS(timestep=case['timestep'], SETTLS_order=case['S'], subdomain=sub)(u=du, v=dv, verbose=False)
        # resample='3H' (LCS.py:88-91, as area_of_influence.py:181 uses it) and gauss_sigma (LCS.py:187-190)
        rs = ref_lcs.LCS(timestep=case['timestep'], SETTLS_order=2, return_dpts=True)(u=du, v=dv, verbose=False, resample='3h')
        gs = ref_lcs.LCS(timestep=case['timestep'], SETTLS_order=case['S'], gauss_sigma=1.5)(u=du, v=dv, verbose=False)
    seams['resample_sigma'], seams['resample_x_dep'], seams['resample_y_dep'] = rs[0].values, rs[1].values, rs[2].values
    seams['resample_time'] = np.asarray(rs[0].coords['time']).astype('datetime64[ns]').astype('int64')
    seams['gauss_sigma_field'] = gs.values
    # find_ridges_spherical_hessian (tools.py:52-155) on 0.5*log(sigma) of the gauss-smoothed field
    ftle = 0.5 * np.log(np.where(gs.values[0] > 0, gs.values[0], np.nan))
    ftle = np.where(np.isfinite(ftle), ftle, 0.0)
    fda = xr.DataArray(ftle, {'latitude': lat, 'longitude': lon}, ('latitude', 'longitude'))
    ridges, eigmin = ref_tools.find_ridges_spherical_hessian(fda, sigma=1.2, tolerance_threshold=0.002e-3)
    seams['ridge_input'], seams['ridge_dt_prod'], seams['ridge_eigmin'] = ftle, ridges.values, eigmin.values
    six = ref_tools.find_ridges_spherical_hessian(fda, sigma=1.2, tolerance_threshold=0.002e-3, return_eigvectors=True)
    assert np.array_equal(six[0].values, ridges.values) and np.array_equal(six[1].values, eigmin.values)
    seams['ridge_dt_raw'], seams['ridge_eigvectors'] = six[2].values, six[3].values            # tools.py:148-152
    seams['ridge_gradient'], seams['ridge_angle'] = six[4].values, six[5].values
    assert six[3].dims == ('eigvectors', 'latitude', 'longitude') and six[4].dims == ('elements', 'latitude', 'longitude')
    # isglobal=True with the 360 x 721 regrid (LCS.py:105-114), truncation=None; coarse 2-degree
    # vortex winds, so the target rows beyond +-88 degrees and the last longitudes take the nearest-label branch.
    # The fields are stored on a stride-5 subgrid to keep the fixture small.
    from lagrangiancoherence_b200 import synthetic as S
    gu, gv, glat, glon = S.ideal_vortex(**S.vortex_config_subtropical)
    gu, gv = gu[:4], gv[:4]
    gt = (np.datetime64('2000-01-01T00') + np.arange(4) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
    gc = {'time': gt, 'latitude': glat, 'longitude': glon}
    gdu = xr.DataArray(gu, gc, ('time', 'latitude', 'longitude'))
    gdv = xr.DataArray(gv, gc, ('time', 'latitude', 'longitude'))
    with contextlib.redirect_stdout(quiet):
        geig, gxd, gyd = ref_lcs.LCS(timestep=-21600, timedim='time', SETTLS_order=2, return_dpts=True)(
            u=gdu, v=gdv, verbose=False, isglobal=True, interp_to_common_grid=True, truncation=None)
    lats, lons = np.linspace(-89.75, 89.75, 360), np.linspace(-180, 179.5, 721)
    ur = gdu.interp(latitude=lats, longitude=lons, method='linear')
    ur = ur.where(~xr.ufuncs.isnan(ur), gdu.reindex(latitude=lats, longitude=lons, method='nearest'))   # LCS.py:108-112
    seams['regrid_u'] = ur.values[:, ::5, ::5]
    seams['regrid_sigma'], seams['regrid_x_dep'], seams['regrid_y_dep'] = (a[..., ::5, ::5] for a in (geig.values, gxd.values, gyd.values))
    seams['regrid_lat'], seams['regrid_lon'] = np.asarray(geig.coords['latitude']), np.asarray(geig.coords['longitude'])
    seams['subdomain_sigma'] = eig.values
    seams['subdomain_lat'], seams['subdomain_lon'] = eig.coords['latitude'], eig.coords['longitude']
    np.savez_compressed(os.path.join(GOLDEN, 'seams.npz'), **seams)
    print('seams: ' + ', '.join(seams))
    import scipy, numba
    manifest['_meta'] = {'generator': 'oracle/make_golden.py', 'reference': REF, 'xarray': 'oracle/refshim stand-in',
                         'numpy': np.__version__, 'scipy': scipy.__version__, 'numba': numba.__version__}
    with open(os.path.join(GOLDEN, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
