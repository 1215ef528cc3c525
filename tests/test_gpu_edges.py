"""Edge cases of the drop-in API on the GPU: degenerate series, orders, dtypes, error behaviour."""
import numpy as np
import pytest
import torch

from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import DataArray, synthetic as S

pytestmark = pytest.mark.gpu


def arrays(u, v, lat, lon):
    time = (np.datetime64('2001-03-01T00') + np.arange(u.shape[0]) * np.timedelta64(1, 'h')).astype('datetime64[ns]')
    coords = {'time': time, 'latitude': lat, 'longitude': lon}
    return DataArray(u, ('time', 'latitude', 'longitude'), coords), DataArray(v, ('time', 'latitude', 'longitude'), coords), time


@pytest.fixture(scope='module')
def case():
    lat = np.linspace(-30.0, 10.0, 41)
    lon = np.linspace(-80.0, -24.0, 57)
    u, v = S.era5_like_winds(lat, lon, 4)
    return u, v, lat, lon


def test_single_level_series_returns_the_start_grid(cuda_device, case):
    """nt = 1: the loop of trajectory.py:80 does not run; positions are the meshgrid, sigma is the identity map's."""
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation
    from lagrangiancoherence_b200.LCS.LCS import LCS
    u, v, lat, lon = case
    du, dv, _ = arrays(u[:1], v[:1], lat, lon)
    x, y = parcel_propagation(du, dv, timestep=-3600, SETTLS_order=4, verbose=False)
    X, Y = np.meshgrid(lon, lat)
    assert np.array_equal(x.values, X) and np.array_equal(y.values, Y)
    sig = LCS(timestep=-3600, SETTLS_order=4)(u=du, v=dv, verbose=False)
    ref = O.spectral_norm_field(O.flowmap_gradient(X, Y, lat, lon))
    assert np.abs(sig.values[0] - ref).max() <= 1e-5 * ref.max()


@pytest.mark.parametrize('S_order', [0, 1, 7])
@pytest.mark.parametrize('order', [1, 3])
def test_settls_orders_and_interp_orders(cuda_device, case, S_order, order):
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation
    u, v, lat, lon = case
    du, dv, _ = arrays(u, v, lat, lon)
    x, y = parcel_propagation(du, dv, timestep=3600, SETTLS_order=S_order, interp_order=order, verbose=False)
    rx, ry = O.parcel_propagation(u, v, lat, lon, 3600, SETTLS_order=S_order, interp_order=order)
    assert np.abs(x.values - rx).max() <= 1e-10 * np.abs(lon).max()
    assert np.abs(y.values - ry).max() <= 1e-10 * np.abs(lat).max()


def test_unsupported_interp_order_is_loud(cuda_device, case):
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation
    u, v, lat, lon = case
    du, dv, _ = arrays(u, v, lat, lon)
    with pytest.raises(RuntimeError, match='spline order not supported'):      # scipy's error and message upstream
        parcel_propagation(du, dv, timestep=3600, interp_order=6, verbose=False)
    with pytest.raises(RuntimeError):                                           # order 0: empty slices upstream
        parcel_propagation(du, dv, timestep=3600, interp_order=0, verbose=False)


@pytest.mark.parametrize('order,xclamp', [(3, 'outer'), (1, 'outer'), (3, 'pointwise'), (2, 'outer')])
def test_float32_winds_follow_the_reference_dtype_propagation(cuda_device, case, order, xclamp):
    """f32 winds on f64 coordinates (how ERA5 is stored): scipy returns the samples in the input dtype (tools.py:26-30)
    and numpy's promotion rules (NEP 50) form the y increments in f32 (trajectory.py:86,110).  The engine reproduces
    that (lcs_advect_opts.round32): departure points within 1e-10 relative of the oracle run on the SAME f32 inputs,
    which tests/test_oracle_golden.py pins bit for bit to the unmodified reference (regional_f32_winds*)."""
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation
    from lagrangiancoherence_b200.LCS.LCS import LCS
    u, v, lat, lon = case
    u32, v32 = u.astype(np.float32), v.astype(np.float32)
    du, dv, _ = arrays(u32, v32, lat, lon)
    sx, sy = np.abs(lon).max(), np.abs(lat).max()
    for timestep in (-3600, np.float64(-3600.0)):          # Python scalar: weak (f32 products); numpy scalar: f64 products
        rx, ry = O.parcel_propagation(u32, v32, lat, lon, timestep, SETTLS_order=3, interp_order=order, xclamp=xclamp)
        assert rx.dtype == np.float64
        x, y = parcel_propagation(du, dv, timestep=timestep, SETTLS_order=3, interp_order=order, verbose=False, xclamp=xclamp)
        ex, ey = np.abs(x.values - rx) / sx, np.abs(y.values - ry) / sy
        assert (ex > 1e-10).mean() <= 1e-3 and (ey > 1e-10).mean() <= 1e-3, (order, xclamp, timestep, ex.max(), ey.max())
        assert np.median(ex) <= 1e-13 and np.median(ey) <= 1e-13
    # the promoted-f64 evaluation of the same values is NOT what the reference computes (~1e-8 .. 1e-7 away)
    p64 = O.parcel_propagation(u32.astype(np.float64), v32.astype(np.float64), lat, lon, -3600, SETTLS_order=3,
                               interp_order=order, xclamp=xclamp)
    r32 = O.parcel_propagation(u32, v32, lat, lon, -3600, SETTLS_order=3, interp_order=order, xclamp=xclamp)
    assert np.abs(p64[0] - r32[0]).max() / sx > 1e-9
    # whole path: sigma against the oracle on the same f32 inputs
    if order == 3 and xclamp == 'outer':
        ref = O.lcs_field(u32, v32, lat, lon, -3600, SETTLS_order=3)
        got = LCS(timestep=-3600, SETTLS_order=3)(u=du, v=dv, verbose=False).values[0]
        ok = np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-12
        assert ok.mean() >= 0.995, ok.mean()


def test_infinite_derivative_raises_like_scipy_norm(cuda_device, case):
    """A +-inf derivative makes scipy.linalg.norm(check_finite=True) raise ValueError at LCS.py:154; a NaN one is
    silently dropped (LCS.py:146).  An infinite position gives NaN, not inf (cos(inf) = nan), upstream as here, so the
    inf branch is reached through a zero metric spacing."""
    from lagrangiancoherence_b200.engine import FtleEngine
    u, v, lat, lon = case
    eng = FtleEngine(lat, lon, -3600, device=cuda_device)
    X, Y = np.meshgrid(lon, lat)
    tx, ty = torch.from_numpy(X).to(cuda_device), torch.from_numpy(Y).to(cuda_device)
    eng.epilogue(tx, ty)
    eng.check_finite()
    eng.dy = 0.0                            # d/dy of a varying field divided by 0 -> +-inf
    eng.epilogue(tx, ty)
    with pytest.raises(ValueError):
        eng.check_finite()
    eng = FtleEngine(lat, lon, -3600, device=cuda_device)
    X[20, 20] = np.inf                      # -> NaN coordinates -> NaN derivatives around it, no error
    sig = eng.epilogue(torch.from_numpy(X).to(cuda_device), ty)
    eng.check_finite()
    s = sig[0].cpu().numpy()
    assert np.isnan(s[20, 21]) and np.isfinite(s[20, 20]) and np.isfinite(s[5, 5])


def test_inputs_are_never_mutated_and_engine_is_reusable(cuda_device, case):
    from lagrangiancoherence_b200.LCS.LCS import LCS
    u, v, lat, lon = case
    du, dv, _ = arrays(u, v, lat, lon)
    u0, v0 = du.values.copy(), dv.values.copy()
    lcs = LCS(timestep=-3600, SETTLS_order=4, return_dpts=True)
    a = lcs(u=du, v=dv, verbose=False)
    b = lcs(u=du, v=dv, verbose=False)
    assert np.array_equal(du.values, u0) and np.array_equal(dv.values, v0)
    assert all(np.array_equal(p.values, q.values, equal_nan=True) for p, q in zip(a, b))


def test_capi_rejects_bad_arguments_on_device(cuda_device, case):
    from lagrangiancoherence_b200 import _lib
    from lagrangiancoherence_b200.engine import FtleEngine
    u, v, lat, lon = case
    eng = FtleEngine(lat, lon, -3600, SETTLS_order=4, xmode='pointwise', device=cuda_device)
    st = eng.stage(u, v)
    with pytest.raises(ValueError, match='past the staged'):
        eng.advect(st, nsteps=3, nwindows=2)
    with pytest.raises(_lib.LcsError, match='halo'):
        x, y = eng.advect(st, rows=(10, 20))
        eng.epilogue(x, y, in_row0=10, out_rows=(10, 20))       # band without its 2-row halo


def test_float32_inputs_agree_with_reference_at_its_own_f32_noise(cuda_device, case):
    """With f32 winds AND f32 coordinates (how xarray decodes ERA5 NetCDF) numpy's dtype propagation makes the
    reference integrate entirely in f32 (meshgrid of f32 coordinates, f32 samples, f32 updates: trajectory.py:68-70,
    86-87).  The engine promotes to f64; the two agree to the reference's own f32 rounding noise, not to 1e-10.
    Stated tolerance: departure points within 2e-5 relative (a few hundred f32 ulps accumulated over the sub-steps)."""
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation
    u, v, lat, lon = case
    u32, v32, lat32, lon32 = u.astype(np.float32), v.astype(np.float32), lat.astype(np.float32), lon.astype(np.float32)
    rx, ry = O.parcel_propagation(u32, v32, lat32, lon32, -3600, SETTLS_order=4, xclamp='outer')
    assert rx.dtype == np.float32 and ry.dtype == np.float32            # the reference's result is f32 here
    du, dv, _ = arrays(u32, v32, lat32, lon32)
    x, y = parcel_propagation(du, dv, timestep=-3600, SETTLS_order=4, verbose=False)
    assert x.values.dtype == np.float64
    assert np.abs(x.values - rx).max() <= 2e-5 * np.abs(lon).max()
    assert np.abs(y.values - ry).max() <= 2e-5 * np.abs(lat).max()
