"""The reference-shaped Python API (LCS, parcel_propagation, flowmap_gradient, tools seams) on the GPU
against the oracle, written the way a test of the reference would read."""
import numpy as np
import pytest

from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import DataArray, Dataset, synthetic as S

pytestmark = pytest.mark.gpu

FTLE_REL = 1e-5     # north_star tolerance on the FTLE field (f32 noise floor of the reference's own stencil)


def winds(nt=5, nlat=41, nlon=57, flip=False, dims=('time', 'latitude', 'longitude')):
    lat = np.linspace(-30.0, 10.0, nlat)
    lon = np.linspace(-80.0, -24.0, nlon)
    u, v = S.era5_like_winds(lat, lon, nt)
    time = np.datetime64('2000-01-01T00') + np.arange(nt) * np.timedelta64(6, 'h')
    coords = {'time': time.astype('datetime64[ns]'), 'latitude': lat, 'longitude': lon}
    du = DataArray(u, ('time', 'latitude', 'longitude'), coords)
    dv = DataArray(v, ('time', 'latitude', 'longitude'), coords)
    if flip:                       # descending latitude, as ERA5 ships it: the API must sort (LCS.py:101-104)
        du = du.isel(latitude=np.arange(nlat)[::-1])
        dv = dv.isel(latitude=np.arange(nlat)[::-1])
    du, dv = du.transpose(*dims), dv.transpose(*dims)
    return du, dv, u, v, lat, lon, time


def close_fraction(a, b, rel):
    return (np.abs(a - b) <= rel * np.abs(b) + 1e-12).mean()


@pytest.mark.parametrize('flip,dims', [(False, ('time', 'latitude', 'longitude')),
                                       (True, ('latitude', 'time', 'longitude'))])
def test_lcs_call_matches_oracle(cuda_device, flip, dims, capsys):
    from lagrangiancoherence_b200.LCS.LCS import LCS
    du, dv, u, v, lat, lon, time = winds(flip=flip, dims=dims)
    out = LCS(timestep=-6 * 3600, timedim='time', SETTLS_order=4)(u=du, v=dv, verbose=False)
    assert '!' * 100 in capsys.readouterr().out                       # LCS.py:74
    assert out.dims == ('time', 'latitude', 'longitude') and out.shape == (1, lat.size, lon.size)
    assert np.array_equal(out.coords['latitude'], lat)                # ascending arrival grid
    assert out.coords['time'][0] == time[0].astype('datetime64[ns]')  # backward: stamped with time[0], LCS.py:158
    ref = O.lcs_field(u, v, lat, lon, -21600, SETTLS_order=4)
    assert close_fraction(out.values[0], ref, FTLE_REL) >= 0.995


def test_lcs_forward_dataset_return_tuples(cuda_device):
    from lagrangiancoherence_b200.LCS.LCS import LCS
    du, dv, u, v, lat, lon, time = winds()
    ds = Dataset({'u': du, 'v': dv})
    res = LCS(timestep=6 * 3600, SETTLS_order=2, return_dpts=True)(ds, verbose=False, return_traj=True)
    assert len(res) == 5                                              # LCS.py:161-162
    eig, xd, yd, xt, yt = res
    assert eig.coords['time'][0] == time[-1].astype('datetime64[ns]')  # forward: time[-1]
    rx, ry = O.parcel_propagation(u, v, lat, lon, 21600, SETTLS_order=2, return_traj=True)
    assert xt.shape == rx.shape and xt.dims == ('time', 'latitude', 'longitude')
    assert np.abs(xt.values - rx).max() <= 1e-10 * np.abs(lon).max()
    assert np.abs(yt.values - ry).max() <= 1e-10 * np.abs(lat).max()
    assert np.array_equal(xd.values, xt.values[-1])
    assert len(LCS(timestep=21600, SETTLS_order=2, return_dpts=True)(ds, verbose=False)) == 3
    assert len(LCS(timestep=21600, SETTLS_order=2)(ds, verbose=False, return_traj=True)) == 3


def test_lcs_subdomain_crops_like_latlonsel(cuda_device):
    from lagrangiancoherence_b200.LCS.LCS import LCS
    du, dv, u, v, lat, lon, _ = winds()
    sub = {'latitude': slice(-20, 0), 'longitude': slice(-70, -40)}
    out = LCS(timestep=-21600, SETTLS_order=4, subdomain=sub)(u=du, v=dv, verbose=False)
    keep_lat = (lat > -20) & (lat < 0)
    keep_lon = (lon > -70) & (lon < -40)
    assert np.array_equal(out.coords['latitude'], lat[keep_lat]) and np.array_equal(out.coords['longitude'], lon[keep_lon])
    ref = O.lcs_field(u, v, lat, lon, -21600, SETTLS_order=4)[keep_lat][:, keep_lon]
    assert close_fraction(out.values[0], ref, FTLE_REL) >= 0.995


@pytest.mark.parametrize('flip', [False, True])
def test_subdomain_skips_rows_under_the_pointwise_clamp_without_changing_a_bit(cuda_device, flip, monkeypatch):
    """with a subdomain and independent particles (xclamp='pointwise') only the kept rows +-2 are integrated
    (the y-stencil's reach); the result must equal the crop of the full field bit for bit -- pole-row and one-sided-stencil
    rules are keyed on global row indices.  Crops touching the first / last rows, descending input latitudes, and the
    cases where skipping must NOT happen (departure points returned, Gaussian smoothing) included."""
    from lagrangiancoherence_b200.LCS.LCS import LCS
    from lagrangiancoherence_b200.engine import FtleEngine
    du, dv, u, v, lat, lon, _ = winds(flip=flip)
    full = LCS(timestep=-21600, SETTLS_order=4)(u=du, v=dv, verbose=False, xclamp='pointwise')
    seen = []
    orig = FtleEngine.advect

    def spy(self, staged, *a, **k):
        seen.append(k.get('rows'))
        return orig(self, staged, *a, **k)
    monkeypatch.setattr(FtleEngine, 'advect', spy)
    for sub in ({'latitude': slice(-20, 0), 'longitude': slice(-70, -40)}, {'latitude': slice(-31, -25)}, {'latitude': slice(5, 11)},
                {'longitude': slice(-60, -50)}):
        out = LCS(timestep=-21600, SETTLS_order=4, subdomain=sub)(u=du, v=dv, verbose=False, xclamp='pointwise')
        keep_lat = _keep(lat, sub.get('latitude'))
        keep_lon = _keep(lon, sub.get('longitude'))
        assert np.array_equal(out.coords['latitude'], lat[keep_lat]) and np.array_equal(out.coords['longitude'], lon[keep_lon])
        assert np.array_equal(out.values[0], full.values[0][keep_lat][:, keep_lon], equal_nan=True), sub
        r = np.flatnonzero(keep_lat)
        assert seen[-1] == (max(0, r[0] - 2), min(lat.size, r[-1] + 3))
    sub = {'latitude': slice(-20, 0)}
    res = LCS(timestep=-21600, SETTLS_order=4, subdomain=sub, return_dpts=True)(u=du, v=dv, verbose=False, xclamp='pointwise')
    assert seen[-1] is None and res[1].shape == (lat.size, lon.size)                 # departure points asked for: no skipping
    LCS(timestep=-21600, SETTLS_order=4, subdomain=sub, gauss_sigma=1.0)(u=du, v=dv, verbose=False, xclamp='pointwise')
    assert seen[-1] is None                                                          # smoothing reaches beyond the band
    LCS(timestep=-21600, SETTLS_order=4, subdomain=sub)(u=du, v=dv, verbose=False)   # as-executed outer clamp couples all rows
    assert seen[-1] is None


def _keep(coord, sl):
    k = np.ones(coord.shape, bool)
    if sl is not None:
        k &= (coord > sl.start) & (coord < sl.stop)
    return k


def test_lcs_asserts_on_bad_dims(cuda_device):
    from lagrangiancoherence_b200.LCS.LCS import LCS
    du, dv, *_ = winds()
    bad = DataArray(du.values, ('time', 'lat', 'lon'))
    with pytest.raises(AssertionError):
        LCS(timestep=-21600)(u=bad, v=bad, verbose=False)


def test_parcel_propagation_cyclic_ideal_vortex(cuda_device):
    """configs[0]: the example's idealised vortex, backward S=4 and forward S=2, cyclic (ideal_vortex.py:262-279)."""
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation
    u, v, lat, lon = S.ideal_vortex(**S.vortex_config_subtropical)
    time = (np.datetime64('2000-01-01T00') + np.arange(u.shape[0]) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
    coords = {'time': time, 'latitude': lat, 'longitude': lon}
    du, dv = DataArray(u, ('time', 'latitude', 'longitude'), coords), DataArray(v, ('time', 'latitude', 'longitude'), coords)
    for dt, s_order in ((-21600, 4), (21600, 2)):
        x, y = parcel_propagation(du, dv, timestep=dt, propdim='time', SETTLS_order=s_order, copy=True,
                                  return_traj=True, cyclic_xboundary=True, verbose=False)
        rx, ry = O.parcel_propagation(u, v, lat, lon, dt, SETTLS_order=s_order, cyclic_xboundary=True, return_traj=True)
        ex = np.abs(x.values - rx) / 180.0
        ey = np.abs(y.values - ry) / 90.0
        # the vortex core winds are not smooth (|u| jumps across the centre): allow a handful of particles that sit
        # on the wrap / pole-row discontinuities to differ, everything else within 1e-10
        assert (ex > 1e-10).mean() <= 1e-3 and (ey > 1e-10).mean() <= 1e-3, ((ex > 1e-10).mean(), ex.max())
        assert x.dims == ('time', 'latitude', 'longitude')
        expect_t = time[::-1] if dt < 0 else time                     # labels reversed when backward, trajectory.py:59-60
        assert np.array_equal(x.coords['time'], expect_t)
    xf, yf = parcel_propagation(du, dv, timestep=-21600, SETTLS_order=4, cyclic_xboundary=True, verbose=False)
    assert xf.dims == ('latitude', 'longitude') and np.asarray(xf.coords['time']).ndim == 0


def test_flowmap_gradient_and_tools(cuda_device):
    from lagrangiancoherence_b200.LCS.LCS import flowmap_gradient
    from lagrangiancoherence_b200.LCS import tools
    du, dv, u, v, lat, lon, _ = winds()
    rx, ry = O.parcel_propagation(u, v, lat, lon, -21600, SETTLS_order=4)
    c2 = {'latitude': lat, 'longitude': lon}
    xd, yd = DataArray(rx, ('latitude', 'longitude'), c2), DataArray(ry, ('latitude', 'longitude'), c2)
    dt = flowmap_gradient(xd, yd)
    ref = O.flowmap_gradient(rx, ry, lat, lon)
    assert dt.dims == ('derivatives', 'latitude', 'longitude') and dt.shape == ref.shape
    assert list(dt.coords['derivatives'][:2]) == ['dxdx', 'dxdy']
    assert (dt.values != ref).mean() <= 1e-3 and np.all(dt.values[6:] == 0)
    # seams
    X = (O.EARTH_R * np.sin((ry - 90) * np.pi / 180) * np.cos(rx * np.pi / 180))
    for dim in (0, 1):
        got = tools.derivative_spherical_coords(DataArray(X, ('latitude', 'longitude'), c2), dim=dim)
        assert np.array_equal(got.values, O.derivative_spherical_coords(X, lat, lon, dim=dim))
        a32 = X.astype('float32')
        for isglobal in (True, False):
            assert np.array_equal(tools.fourth_order_derivative(a32, dim=dim, isglobal=isglobal),
                                  O.fourth_order_derivative(a32, dim=dim, isglobal=isglobal))
    got = tools.xr_map_coordinates(du.isel(time=0), xd, yd, order=3)
    assert np.abs(got.values - O.xr_map_coordinates(u[0], rx, ry, lat, lon, order=3)).max() <= 1e-12 * np.abs(u).max()


def test_spectral_norm_seam(cuda_device):
    from scipy.linalg import norm
    from lagrangiancoherence_b200.engine import spectral_norm_3x3_device
    rng = np.random.default_rng(0)
    vals = rng.normal(size=(3, 3, 500))
    got = spectral_norm_3x3_device(vals).cpu().numpy()
    assert np.abs(got - norm(vals, axis=(0, 1), ord=2)).max() <= 1e-12
    vals[2] = 0                                                       # the hot path's layout (LCS.py:206-208)
    got = spectral_norm_3x3_device(vals).cpu().numpy()
    assert np.abs(got - norm(vals, axis=(0, 1), ord=2)).max() <= 1e-13


def test_rolling_matches_per_window_calls(cuda_device):
    from lagrangiancoherence_b200.rolling import rolling_ftle
    lat = np.linspace(-30.0, 10.0, 41)
    lon = np.linspace(-80.0, -24.0, 57)
    u, v = S.era5_like_winds(lat, lon, 9)
    fields = rolling_ftle(u, v, lat, lon, 5, -21600, SETTLS_order=4, chunk=2)
    assert fields.shape == (5, lat.size, lon.size)
    for s in (0, 3, 4):
        ref = O.lcs_field(u[s:s + 5], v[s:s + 5], lat, lon, -21600, SETTLS_order=4)
        assert close_fraction(fields[s], ref, FTLE_REL) >= 0.995


def test_rolling_inf_in_an_early_chunk_still_raises(cuda_device):
    """scipy.linalg.norm(check_finite=True) raises on any inf derivative (LCS.py:154).  The rolling driver runs one
    epilogue per chunk and checks once at the end: an inf met by the FIRST chunk must survive the later launches."""
    from lagrangiancoherence_b200 import rolling
    from lagrangiancoherence_b200.engine import FtleEngine
    lat = np.linspace(-30.0, 10.0, 41)
    lon = np.linspace(-80.0, -24.0, 57)
    u, v = S.era5_like_winds(lat, lon, 12)
    eng = FtleEngine(lat, lon, -21600, SETTLS_order=1, xmode='pointwise', device=cuda_device)
    dy = eng.dy
    eng.dy = 0.0                                     # zero metric spacing -> +-inf derivatives (first chunk only)

    def restore(first, fields):
        eng.dy = dy
    with pytest.raises(ValueError, match='infs or NaNs'):
        rolling.rolling_ftle(u, v, lat, lon, 3, -21600, SETTLS_order=1, xclamp='pointwise', engine=eng, chunk=3, on_chunk=restore)
    out = rolling.rolling_ftle(u, v, lat, lon, 3, -21600, SETTLS_order=1, xclamp='pointwise', engine=eng, chunk=3)   # flag was cleared
    assert np.isfinite(out).any()


def test_bench_line_has_the_contract_keys(cuda_device):
    """`python bench.py` prints ONE JSON line carrying the driver's contract (value, e2e with the copied bytes, kernel
    launch count, clocks, roofline with the live-measured ceiling, cpu_baseline unless skipped)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--steps', '3', '--warmup', '3', '--batch', '296',
                          '--no-cpu-baseline'], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline'):
        assert key in d, key
    assert d['metric'] == 'particle-steps/s' and d['n_gpus'] == 1 and d['steps'] == 3 and d['dtype'] == 'f64'
    assert d['value'] > 1e9 and 0 < d['e2e']['value'] < d['value']                 # copies inside the timed region cost something
    assert d['e2e']['h2d_bytes_per_step'] == 2 * 304 * 281 * 321 * 8 and d['e2e']['d2h_bytes_per_step'] == 296 * 281 * 321 * 8
    assert d['gpu_launches'] == 3 * 5                                            # prefilter x2, pack, integrator, epilogue per step
    r = d['roofline']
    assert set(('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic')) <= set(r)
    assert r['unit'] == 'GB/s' and 0.3 < r['frac'] < 1.3 and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert 'workload' in d['config'] and 'sm_mhz' in d['clocks']
