"""Known-answer tests that follow analytically from the reference code: they pin
the oracle's reading of trajectory.py / LCS.py independently of any implementation."""
import numpy as np
import pytest

from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import synthetic as S


def grid(nlat=31, nlon=41):
    return np.linspace(-30.0, 0.0, nlat), np.linspace(-60.0, -20.0, nlon)


def test_zero_wind_is_the_identity_map():
    lat, lon = grid()
    z = np.zeros((4, lat.size, lon.size))
    x, y = O.parcel_propagation(z, z, lat, lon, -21600, SETTLS_order=4)
    X, Y = np.meshgrid(lon, lat)
    assert np.array_equal(x, X) and np.array_equal(y, Y)


def test_uniform_zonal_wind_displaces_by_one_plus_S_per_step():
    """Q1: SETTLS iterations accumulate, so steady uniform flow moves (1+S)*dt*U0*conversion_x(row) per interval
    (before clamping).  Pole rows (first/last `order`) sample with order 1/'constant'; the last row's index is
    nlat > nlat-1 and reads 0, so it never moves (Q3, Q4)."""
    lat, lon = np.linspace(-30.0, 0.0, 31), np.linspace(-180.0, 180.0, 73)     # wide: nobody reaches the clamp
    nt, U0, dt, S_ = 3, 5.0, 3600.0, 4
    u = np.full((nt, lat.size, lon.size), U0)
    v = np.zeros_like(u)
    x, y = O.parcel_propagation(u, v, lat, lon, dt, SETTLS_order=S_, xclamp='pointwise')
    cx, _ = O.conversions(lat)
    X, Y = np.meshgrid(lon, lat)
    assert np.array_equal(y, Y)
    expect = X + (1 + S_) * (nt - 1) * dt * U0 * cx[:, None]
    inner = np.s_[3:-3, 2:-12]           # interior rows, columns that stay clear of the wrap fold/clamp
    assert np.abs(x[inner] - expect[inner]).max() <= 1e-9
    assert np.array_equal(x[-1], X[-1])                                        # last row samples 0: frozen


def test_wind_levels_ascend_whatever_the_sign_of_timestep():
    """Q2: a negative timestep only flips the displacement; levels are still consumed 0,1,2,..."""
    lat, lon = grid()
    u, v = S.era5_like_winds(lat, lon, 4, contained=True)
    xf, yf = O.parcel_propagation(u, v, lat, lon, 3600, SETTLS_order=0, interp_order=1)
    xb, yb = O.parcel_propagation(-u, -v, lat, lon, -3600, SETTLS_order=0, interp_order=1)
    assert np.array_equal(xf, xb) and np.array_equal(yf, yb)


def test_outer_clamp_equals_pointwise_when_nobody_exits_and_differs_when_some_do():
    lat, lon = grid()
    u, v = S.era5_like_winds(lat, lon, 4, contained=True)
    a = O.parcel_propagation(u * 0.2, v * 0.2, lat, lon, -3600, SETTLS_order=2, xclamp='outer')
    b = O.parcel_propagation(u * 0.2, v * 0.2, lat, lon, -3600, SETTLS_order=2, xclamp='pointwise')
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    u, v = S.era5_like_winds(lat, lon, 4)
    a = O.parcel_propagation(u, v, lat, lon, -21600, SETTLS_order=2, xclamp='outer')
    b = O.parcel_propagation(u, v, lat, lon, -21600, SETTLS_order=2, xclamp='pointwise')
    assert (a[0] != b[0]).mean() > 0.01
    assert a[0].min() >= lon.min() and a[0].max() <= lon.max()


def test_outer_clamp_is_orthogonal_assignment():
    """Q6 on a hand-made array: rows {0, 2} and columns {1, 3} hold an exit -> the 2x2 product is set."""
    px = np.array([[0., -9., 0., 0.], [0., 0., 0., 0.], [0., 0., 0., -9.]])
    got = O._clamp_x(px, -5.0, 5.0, 'outer')
    expect = px.copy()
    expect[np.ix_([0, 2], [1, 3])] = -5.0
    assert np.array_equal(got, expect)
    assert np.array_equal(O._clamp_x(px, -5.0, 5.0, 'pointwise'), np.where(px < -5, -5.0, px))


def test_cyclic_wrap_uses_python_sign_mod_180():
    px = np.array([[-181.0, -180.0, 179.0, 180.0, 190.0, 361.0]])
    got = O._wrap_x_cyclic(px)
    assert np.array_equal(got, np.array([[179.0, 0.0, 179.0, -180.0, -170.0, -179.0]]))    # note -180 -> 0, trajectory.py:93


def test_identity_map_sigma_closed_form():
    """Zero wind: X,Y,Z of the grid itself.  Away from the one-sided rows the 4th-order stencil of a smooth
    function is accurate, so the scrambled 3x3 (Q7) can be written down: M = [[Xx,Xy,Yx],[Yy,Zx,Zy]] with
    Xx = -sin(LAT) sin(LON)/cos(lat) etc.; sigma_max must match its closed form to the f32 noise floor."""
    lat, lon = np.linspace(-40.0, -10.0, 61), np.linspace(-70.0, -30.0, 81)
    X, Y = np.meshgrid(lon, lat)
    sig = O.spectral_norm_field(O.flowmap_gradient(X, Y, lat, lon))
    LON, LAT = np.deg2rad(X), np.deg2rad(Y - 90)
    cl = np.cos(np.deg2rad(Y))
    a = -np.sin(LAT) * np.sin(LON) / cl      # dX/dx = (1/(R cos lat)) dX/dlon
    b = np.cos(LAT) * np.cos(LON)            # dX/dy = (1/R) dX/dlat
    c = np.sin(LAT) * np.cos(LON) / cl       # dY/dx
    d = np.cos(LAT) * np.sin(LON)            # dY/dy
    e = np.zeros_like(a)                     # dZ/dx
    f = -np.sin(LAT)                         # dZ/dy
    ref = O.sigma_max_closed_form(np.stack([a, b, c, d, e, f]))
    inner = np.s_[2:-2, 2:-2]                # x-stencil is periodic across the regional edge: skip 2 columns too
    assert np.abs(sig[inner] - ref[inner]).max() <= 2e-4 * ref[inner].max()


def test_nan_derivative_drops_the_point_and_inf_raises():
    lat, lon = grid()
    X, Y = np.meshgrid(lon, lat)
    Xn = X.copy()
    Xn[10, 10] = np.nan
    sig = O.spectral_norm_field(O.flowmap_gradient(Xn, Y, lat, lon))
    # centred stencils skip their own centre: the NaN poisons its +-1/+-2 neighbours, not its own point
    assert np.isfinite(sig[10, 10]) and np.isfinite(sig[20, 20])
    assert all(np.isnan(sig[i, j]) for i, j in [(10, 11), (10, 12), (10, 8), (11, 10), (12, 10), (8, 10)])
    dt = O.flowmap_gradient(X, Y, lat, lon)
    dt[0, 5, 5] = np.inf
    with pytest.raises(ValueError):
        O.spectral_norm_field(dt)


@pytest.mark.parametrize('order', [1, 2, 3, 4, 5])
def test_displacement_law_holds_for_every_spline_order(order):
    """B-spline interpolation of any order reproduces a constant field exactly (weights sum to one, prefilter has unit
    DC gain), so the (1+S) displacement law of uniform flow does not depend on `traj_interp_order`; what does is the
    number of pole rows (first/last `order` arrival rows sample with order 1/'constant', tools.py:31-39)."""
    lat, lon = np.linspace(-30.0, 0.0, 31), np.linspace(-180.0, 180.0, 73)
    nt, U0, dt, S_ = 3, 5.0, 3600.0, 2
    u = np.full((nt, lat.size, lon.size), U0)
    x, y = O.parcel_propagation(u, np.zeros_like(u), lat, lon, dt, SETTLS_order=S_, interp_order=order, xclamp='pointwise')
    cx, _ = O.conversions(lat)
    X, Y = np.meshgrid(lon, lat)
    expect = X + (1 + S_) * (nt - 1) * dt * U0 * cx[:, None]
    inner = np.s_[order:-order, 2:-12]
    assert np.array_equal(y, Y)
    assert np.abs(x[inner] - expect[inner]).max() <= 1e-9
    assert np.array_equal(x[-1], X[-1])                    # the last row's index is nlat > nlat-1: samples 0, never moves (Q4)
