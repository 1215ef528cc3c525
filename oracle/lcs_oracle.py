"""CPU oracle: numpy/scipy/numba restatement of the reference's FTLE hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the reference
lines it follows (paths are relative to /root/reference).  The arithmetic lives in the same
three third-party calls the reference makes -- ``scipy.ndimage.map_coordinates``
(tools.py:26-30,35-39), a 4th-order stencil (tools.py:190-245, restated below) and
``scipy.linalg.norm(ord=2)`` (LCS.py:154) -- and the expressions are written in the
reference's order of evaluation so dtype promotion and rounding follow numpy exactly.

Pinned third-party versions (the reference pins none, requirements.txt:1-5):
scipy 1.18.1 (``mode='wrap'`` / prefilter boundary behaviour changed in 1.6), numpy 2.3.5
(NEP 50 promotion: ``f32_array / np.float64_scalar`` is f64), numba 0.65.0.

xarray semantics restated by hand: ``sortby`` (inputs here are already ascending),
broadcasting of a ``(latitude,)`` vector against ``(latitude, longitude)``,
``where(cond, other)``, Python-sign ``%`` and ORTHOGONAL indexing for
``da[np.where(cond)] = c`` (xarray treats a tuple of unlabelled 1-D arrays as an outer
indexer) -- selectable with ``xclamp='outer'`` (as executed) or ``'pointwise'``.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import map_coordinates, gaussian_filter
from scipy.linalg import norm

EARTH_R = 6371000  # trajectory.py:54, LCS.py:193, tools.py:249


# --------------------------------------------------------------------------------------
# a4: xr_map_coordinates (tools.py:11-48, always the isglobal=True branch)
# --------------------------------------------------------------------------------------
def index_map(pos, coord):
    """Degrees -> fractional index, tools.py:19-22 (n points, not n-1: quirk Q4)."""
    n = coord.shape[0]
    return n * (pos - coord.min()) / (coord.max() - coord.min())


def xr_map_coordinates(field, new_x, new_y, lat, lon, order=1):
    """Sample ``field[nlat, nlon]`` at positions (degrees) -- tools.py:11-41.

    Rows ``order .. n-order-1`` of the POSITION array (arrival-row index, quirk Q3) use
    ``map_coordinates(order=order, mode='wrap')``; the first/last ``order`` rows use
    ``order=1, mode='constant'``.  The result has ``field``'s dtype (scipy allocates the
    output with the input dtype).
    """
    field = np.asarray(field)
    ix = index_map(np.asarray(new_x), lon)
    iy = index_map(np.asarray(new_y), lat)
    nrow = ix.shape[0]
    out = np.empty(ix.shape, dtype=field.dtype)
    idxs = np.arange(order, nrow - order)                       # tools.py:25
    out[idxs, :] = map_coordinates(
        field, np.array([iy[idxs, :].ravel(), ix[idxs, :].ravel()]),
        order=order, mode='wrap').reshape(len(idxs), ix.shape[1])   # tools.py:26-30
    pole_idxs = np.hstack([np.arange(0, order), np.arange(-order, 0)])  # tools.py:31-33
    out[pole_idxs, :] = map_coordinates(
        field, np.array([iy[pole_idxs, :].ravel(), ix[pole_idxs, :].ravel()]),
        order=1, mode='constant').reshape(len(pole_idxs), ix.shape[1])  # tools.py:35-39
    return out


# --------------------------------------------------------------------------------------
# a3: parcel_propagation (trajectory.py:8-144)
# --------------------------------------------------------------------------------------
def conversions(lat):
    """deg/m factors, trajectory.py:54-57 (arrival-grid latitude, quirk Q5)."""
    conversion_y = 180 / (EARTH_R * np.pi)
    conversion_x = 180 / (np.pi * EARTH_R * np.abs(np.cos(lat * np.pi / 180)))
    return conversion_x, conversion_y


def _clamp_y(py, y_min, y_max):
    # trajectory.py:89-90 / 115-116 ; DataArray.where(cond, other): keep where cond else other
    py = np.where(py > y_min, py, y_min)
    py = np.where(py < y_max, py, y_max)
    return py


def _wrap_x_cyclic(px):
    # trajectory.py:93-94 / 119-120 ; numpy % has the sign of the divisor
    px = np.where(px > -180, px, px % 180)
    px = np.where(px < 180, px, -180 + (px % 180))
    return px


def _clamp_x(px, x_min, x_max, xclamp):
    # trajectory.py:96-97 / 122-123
    px = px.copy()
    if xclamp == 'outer':          # as executed: orthogonal (outer-product) assignment, quirk Q6
        r, c = np.where(px < x_min)
        if r.size:
            px[np.ix_(np.unique(r), np.unique(c))] = x_min
        r, c = np.where(px > x_max)
        if r.size:
            px[np.ix_(np.unique(r), np.unique(c))] = x_max
    elif xclamp == 'pointwise':
        px[px < x_min] = x_min
        px[px > x_max] = x_max
    else:
        raise ValueError(xclamp)
    return px


def parcel_propagation(U, V, lat, lon, timestep=1, SETTLS_order=0, interp_order=3,
                       cyclic_xboundary=False, xclamp='outer', return_traj=False):
    """Two-time-level advection with accumulating SETTLS sub-iterations.

    ``U, V``: ``(nt, nlat, nlon)`` on ascending ``lat``/``lon`` (the reference sorts first,
    trajectory.py:49-52).  Wind levels are always consumed in ascending index order,
    whatever the sign of ``timestep`` (quirk Q2, trajectory.py:80-84).  Returns final
    ``(positions_x, positions_y)`` or, with ``return_traj``, the ``(nt, nlat, nlon)`` stacks
    including the t0 grid (trajectory.py:76-77,125-126,138-139).
    """
    U = np.asarray(U)
    V = np.asarray(V)
    nt = U.shape[0]
    conversion_x, conversion_y = conversions(lat)
    conversion_x = conversion_x[:, None]                          # xr.broadcast, trajectory.py:57
    y_min, y_max = lat.min(), lat.max()                            # trajectory.py:63-66
    x_min, x_max = lon.min(), lon.max()
    positions_x, positions_y = np.meshgrid(lon, lat)               # trajectory.py:68-70
    pos_list_x = [positions_x]
    pos_list_y = [positions_y]

    def interp(F, px, py):
        return xr_map_coordinates(F, px, py, lat, lon, order=interp_order)

    def bounds(px, py):
        py = _clamp_y(py, y_min, y_max)
        if cyclic_xboundary:
            px = _wrap_x_cyclic(px)
        else:
            px = _clamp_x(px, x_min, x_max, xclamp)
        return px, py

    for time_idx in range(nt - 1):                                 # trajectory.py:80
        va = interp(V[time_idx], positions_x, positions_y)         # :82
        ua = interp(U[time_idx], positions_x, positions_y)         # :84
        positions_y = positions_y + timestep * conversion_y * va   # :86
        positions_x = positions_x + timestep * conversion_x * ua   # :87
        positions_x, positions_y = bounds(positions_x, positions_y)    # :89-97
        k = 0
        while k < SETTLS_order:                                    # :100 (accumulates, quirk Q1)
            v_t = interp(V[time_idx], positions_x, positions_y)            # :105
            v_tp = interp(V[time_idx + 1], positions_x, positions_y)       # :106
            u_t = interp(U[time_idx], positions_x, positions_y)            # :107
            u_tp = interp(U[time_idx + 1], positions_x, positions_y)       # :108
            positions_y = positions_y + 0.5 * timestep * conversion_y * (va + 2 * v_t - v_tp)  # :110
            positions_x = positions_x + 0.5 * timestep * conversion_x * (ua + 2 * u_t - u_tp)  # :112
            positions_x, positions_y = bounds(positions_x, positions_y)    # :115-123
            k += 1
        pos_list_x.append(positions_x)
        pos_list_y.append(positions_y)
    if return_traj:
        return np.stack(pos_list_x), np.stack(pos_list_y)
    return pos_list_x[-1], pos_list_y[-1]


# --------------------------------------------------------------------------------------
# a7: fourth_order_derivative / derivative_spherical_coords (tools.py:190-267)
# --------------------------------------------------------------------------------------
def fourth_order_derivative(arr, dim=0, isglobal=True):
    """Index-space stencil on an f32 array, restating tools.py:190-245.

    numba typing of the reference expression (arr is f32): the differences are f32, the
    ``(4/3)`` / ``(1/3)`` factors promote to f64, the result is rounded to f32 on store.
    """
    a = np.asarray(arr)
    out = np.zeros_like(a)
    n0, n1 = a.shape
    c43, c13 = 4 / 3, 1 / 3
    if dim == 0:
        if n0 > 4:
            d1 = (a[3:n0 - 1] - a[1:n0 - 3]).astype(np.float64)
            d2 = (a[4:n0] - a[0:n0 - 4]).astype(np.float64)
            out[2:n0 - 2] = c43 * d1 / 2 - c13 * d2 / 4             # tools.py:202-207
        for i in (0, 1):
            out[i] = (a[i + 1] - a[i]).astype(np.float64) / 2       # tools.py:210-213
        for i in (-1, -2):
            out[i] = (a[i] - a[i - 1]).astype(np.float64) / 2       # tools.py:214-217
    elif dim == 1:
        if isglobal:
            j = np.arange(n1)
            d1 = (a[:, (j + 1) % n1] - a[:, (j - 1) % n1]).astype(np.float64)
            d2 = (a[:, (j + 2) % n1] - a[:, (j - 2) % n1]).astype(np.float64)
            out[:] = c43 * d1 / 2 - c13 * d2 / 4                    # tools.py:221-228
        else:
            d1 = (a[:, 3:n1 - 1] - a[:, 1:n1 - 3]).astype(np.float64)
            d2 = (a[:, 4:n1] - a[:, 0:n1 - 4]).astype(np.float64)
            out[:, 2:n1 - 2] = c43 * d1 / 2 - c13 * d2 / 4          # tools.py:230-235
            for j in (0, 1):
                out[:, j] = (a[:, j + 1] - a[:, j]).astype(np.float64) / 2   # :237-240
            for j in (-1, -2):
                out[:, j] = (a[:, j] - a[:, j - 1]).astype(np.float64) / 2   # :241-244
    else:
        raise ValueError('Dim must be either 0 or 1.')
    return out


def _numba_fourth_order_derivative():
    """A loop-form twin of the stencil, jitted like the reference's (tools.py:190) -- used by
    tests to confirm the vectorised form above has numba's mixed f32/f64 rounding."""
    from numba import jit

    @jit(nopython=True)
    def kernel(arr, dim, isglobal):
        out = np.zeros_like(arr)
        n0 = arr.shape[0]
        n1 = arr.shape[1]
        if dim == 0:
            for i in range(2, n0 - 2):
                for j in range(n1):
                    out[i, j] = (4 / 3) * (arr[i + 1, j] - arr[i - 1, j]) / 2 \
                        - (1 / 3) * (arr[i + 2, j] - arr[i - 2, j]) / 4
            for i in (0, 1):
                for j in range(n1):
                    out[i, j] = (arr[i + 1, j] - arr[i, j]) / 2
            for i in (-1, -2):
                for j in range(n1):
                    out[i, j] = (arr[i, j] - arr[i - 1, j]) / 2
        else:
            for i in range(n0):
                for j in range(n1):
                    if isglobal:
                        out[i, j] = (4 / 3) * (arr[i, (j + 1) % n1] - arr[i, (j - 1) % n1]) / 2 \
                            - (1 / 3) * (arr[i, (j + 2) % n1] - arr[i, (j - 2) % n1]) / 4
        return out
    return kernel


def derivative_spherical_coords(field, lat, lon, dim=0, isglobal=True):
    """tools.py:248-267: f32 stencil, then divide by the metric spacing in f64."""
    y = lat * np.pi / 180                                          # tools.py:254
    dx = (np.pi / 180) * (lon[1] - lon[0]) * EARTH_R * np.cos(y)   # :255
    dy = (np.pi / 180) * (lat[1] - lat[0]) * EARTH_R               # :256
    deriv = fourth_order_derivative(np.asarray(field).astype('float32'), dim=dim, isglobal=isglobal)  # :258
    if dim == 0:
        return deriv / dy                                          # :262
    elif dim == 1:
        return deriv / dx[:, None]                                 # :264
    raise ValueError('Dim must be either 0 or 1.')


# --------------------------------------------------------------------------------------
# a6: flowmap_gradient (LCS.py:171-225)  and  a8: spectral norm (LCS.py:145-157)
# --------------------------------------------------------------------------------------
def flowmap_gradient(x_departure, y_departure, lat, lon, sigma=None):
    """Returns the 9 'derivatives' stacked ``(9, nlat, nlon)`` in the reference's order
    dxdx,dxdy,dydx,dydy,dzdx,dzdy,dxdr,dydr,dzdr (LCS.py:210-223, quirk Q7)."""
    if isinstance(sigma, (float, int)):                            # LCS.py:187-190
        x_departure = gaussian_filter(x_departure, sigma=sigma)
        y_departure = gaussian_filter(y_departure, sigma=sigma)
    LON = x_departure * np.pi / 180                                # :195
    LAT = (y_departure - 90) * np.pi / 180                         # :196 (negative colatitude)
    X = EARTH_R * np.sin(LAT) * np.cos(LON)                        # :197
    Y = EARTH_R * np.sin(LAT) * np.sin(LON)                        # :198
    Z = EARTH_R * np.cos(LAT)                                      # :199
    dXdx = derivative_spherical_coords(X, lat, lon, dim=1)         # :200-205
    dXdy = derivative_spherical_coords(X, lat, lon, dim=0)
    dYdx = derivative_spherical_coords(Y, lat, lon, dim=1)
    dYdy = derivative_spherical_coords(Y, lat, lon, dim=0)
    dZdx = derivative_spherical_coords(Z, lat, lon, dim=1)
    dZdy = derivative_spherical_coords(Z, lat, lon, dim=0)
    zero = np.zeros_like(dXdx)                                     # :206-208
    return np.stack([dXdx, dXdy, dYdx, dYdy, dZdx, dZdy, zero, zero, zero])


def spectral_norm_field(def_tensor, mask=None):
    """LCS.py:145-157: drop points with any NaN (or outside ``mask``), reshape the 9 stacked
    components row-major to 3x3 per point, largest singular value; dropped points -> NaN.
    ``scipy.linalg.norm`` raises ValueError on inf (check_finite), as in the reference."""
    nine, nlat, nlon = def_tensor.shape
    flat = def_tensor.reshape(nine, nlat * nlon)
    keep = ~np.isnan(flat).any(axis=0)                             # dropna('points'), :146
    if mask is not None:
        keep &= mask.ravel()
    vals = flat[:, keep].reshape([3, 3, int(keep.sum())])          # :152-153
    out = np.full(nlat * nlon, np.nan)
    if vals.shape[-1]:
        out[keep] = norm(vals, axis=(0, 1), ord=2)                 # :154
    return out.reshape(nlat, nlon)


def sigma_max_closed_form(def_tensor):
    """Closed form for sigma_max of [[a,b,c],[d,e,f],[0,0,0]] -- used to
    document that the SVD and the closed form agree; not a reference function."""
    a, b, c, d, e, f = def_tensor[:6]
    g11 = a * a + b * b + c * c
    g22 = d * d + e * e + f * f
    g12 = a * d + b * e + c * f
    return np.sqrt(0.5 * (g11 + g22 + np.sqrt((g11 - g22) ** 2 + 4 * g12 * g12)))


def subdomain_mask(lat, lon, subdomain):
    """Strict-inequality crop of tools.py:184-186 applied as a point mask (LCS.py:143-144;
    the ``xr_tools.latlonsel`` the reference imports is not in its tree -- SURVEY.md a9)."""
    la = subdomain['latitude']
    lo = subdomain['longitude']
    latmask = np.ones(lat.shape, bool)
    lonmask = np.ones(lon.shape, bool)
    if la.start is not None:
        latmask &= lat > la.start
    if la.stop is not None:
        latmask &= lat < la.stop
    if lo.start is not None:
        lonmask &= lon > lo.start
    if lo.stop is not None:
        lonmask &= lon < lo.stop
    return latmask[:, None] & lonmask[None, :]


# --------------------------------------------------------------------------------------
# find_ridges_spherical_hessian (tools.py:52-155)
# --------------------------------------------------------------------------------------
def find_ridges_spherical_hessian(field, lat, lon, sigma=.5, tolerance_threshold=0.0005e-3, isglobal=True,
                                  return_eigvectors=False):
    """Hessian ridge filter of an FTLE field ``(nlat, nlon)``; returns ``(dt_prod, eigmin)``, or with
    ``return_eigvectors`` the six arrays of tools.py:148-152: ``dt_prod, eigmin, dt (unthresholded), eigvectors
    [2, nlat, nlon] (zeroed where eigmin >= 0, :132), gradient [2, nlat, nlon], angle = 180/pi*arctan(e0/e1)`` (:125,
    from the eigenvectors BEFORE the zeroing).

    As executed: Gaussian smoothing on the (longitude, latitude) transpose (tools.py:70-76), five
    derivative_spherical_coords passes, each re-casting its input to f32 (:78-82), inf/NaN of the Hessian set to 0
    (:93-94), then per point ``np.linalg.eig`` of [[xx, xy], [xy, yy]] with the upstream quirk that a ROW of the
    eigenvector matrix is taken (:108): ``dt = dot(eig[1][argmin(eig[0])], gradient)``;
    ``eigmin = eig[0][argmax(|eig[0]|)]`` (:119); ``dt_prod = 1`` where not(|dt| > tol) and eigmin < 0, else 0 (:134-136)."""
    f = np.asarray(field, dtype=np.float64)
    if isinstance(sigma, (float, int)):
        f = gaussian_filter(f.T, sigma=sigma).T                                         # :75-76
    ddadx = derivative_spherical_coords(f, lat, lon, dim=1, isglobal=isglobal)          # :78
    ddady = derivative_spherical_coords(f, lat, lon, dim=0, isglobal=isglobal)          # :79
    d2dadx2 = derivative_spherical_coords(ddadx, lat, lon, dim=1, isglobal=isglobal)    # :80
    d2dady2 = derivative_spherical_coords(ddady, lat, lon, dim=0, isglobal=isglobal)    # :81
    d2dadxdy = derivative_spherical_coords(ddadx, lat, lon, dim=0, isglobal=isglobal)   # :82
    hess = np.stack([d2dadx2, d2dadxdy, d2dadxdy, d2dady2]).reshape(4, -1)
    hess = np.where(np.abs(hess) != np.inf, hess, 0)                                    # :93
    hess = np.where(~np.isnan(hess), hess, 0)                                           # :94
    grad = np.stack([ddadx, ddady]).reshape(2, -1)
    H = hess.T.reshape(-1, 2, 2)
    w, v = np.linalg.eig(H)                                                             # :107, one LAPACK call per point
    n = np.arange(H.shape[0])
    row = v[n, np.argmin(w, axis=1)]                                                    # :108 (a row, not a column)
    gT = np.ascontiguousarray(grad.T)
    dt = np.array([np.dot(row[i], gT[i]) for i in range(row.shape[0])])                # :116 np.dot of two 2-vectors (BLAS ddot: fused)
    eigmin = w[n, np.argmax(np.abs(w), axis=1)]                                         # :119
    dt_prod = np.where(np.abs(dt) > tolerance_threshold, 0.0, 1.0)                      # :134-135 (NaN -> 1)
    dt_prod = np.where(np.sign(eigmin) == -1, dt_prod, 0.0)                             # :136
    if return_eigvectors:
        with np.errstate(all='ignore'):
            angle = 180 / np.pi * np.arctan(row[:, 0] / row[:, 1])                      # :125
        evec = np.where(eigmin < 0, row.T, 0.0)                                         # :132
        return (dt_prod.reshape(f.shape), eigmin.reshape(f.shape), dt.reshape(f.shape),
                evec.reshape((2,) + f.shape), grad.reshape((2,) + f.shape), angle.reshape(f.shape))
    return dt_prod.reshape(f.shape), eigmin.reshape(f.shape)


def eig_sym2x2_lapack(a, b, d):
    """What ``np.linalg.eig`` (LAPACK dgeev -> dlanv2) returns for [[a, b], [b, d]]: eigenvalues in the order
    (rt1, rt2) and the rotation (cs, sn) with eigenvector matrix [[cs, -sn], [sn, cs]] (before dgeev's final
    renormalisation, which moves entries by at most one ulp).  The spec of the CUDA ridge kernel."""
    a, b, d = (np.asarray(t, dtype=np.float64) for t in (a, b, d))
    with np.errstate(all='ignore'):
        p = 0.5 * (a - d)
        bc = np.abs(b)
        scale = np.maximum(np.abs(p), bc)
        z = p / scale * p + bc / scale * bc
        z = p + np.copysign(np.sqrt(scale) * np.sqrt(z), p)
        rt1 = d + z
        rt2 = d - bc / z * bc
        tau = np.hypot(b, z)
        cs, sn = z / tau, b / tau
    diag = b == 0
    return (np.where(diag, a, rt1), np.where(diag, d, rt2), np.where(diag, 1.0, cs), np.where(diag, 0.0, sn))


# --------------------------------------------------------------------------------------
# resample= (LCS.py:88-91): u.resample({timedim: freq}).interpolate('linear')
# --------------------------------------------------------------------------------------
def resample_linear(U, times, freq):
    """Linear refinement in time as xarray executes it: new index = pandas resample bin labels, values from
    scipy.interpolate.interp1d(kind='linear') (scipy 1.18.1) over float64 ns offsets:
    y = ((x_new - x_lo)/(x_hi - x_lo)) * y_hi + ((x_hi - x_new)/(x_hi - x_lo)) * y_lo, bracket from
    searchsorted(x, x_new, 'left') clipped to [1, n-1].  Returns ``(U_new, new_times)``."""
    import pandas as pd
    t = np.asarray(times).astype('datetime64[ns]')
    new = pd.Series(0.0, index=pd.DatetimeIndex(t)).resample(freq).asfreq().index.values.astype('datetime64[ns]')
    x = (t - t.min()).astype('int64').astype(np.float64)
    xn = (new - t.min()).astype('int64').astype(np.float64)
    hi = np.clip(np.searchsorted(x, xn, side='left'), 1, x.size - 1)
    lo = hi - 1
    U = np.asarray(U)
    shp = (-1,) + (1,) * (U.ndim - 1)
    w_hi = ((xn - x[lo]) / (x[hi] - x[lo])).reshape(shp)
    w_lo = ((x[hi] - xn) / (x[hi] - x[lo])).reshape(shp)
    return w_hi * U[hi] + w_lo * U[lo], new


# --------------------------------------------------------------------------------------
# SURVEY 8f rank 4 (the feasible half): the fixed common grid of the global path, LCS.py:105-114
# --------------------------------------------------------------------------------------
def regrid_to_common_grid(U, lat, lon):
    """``u.interp(latitude=lats, longitude=lons, method='linear')`` with NaNs filled from
    ``u.reindex(..., method='nearest')`` (LCS.py:106-113), through the third-party calls xarray makes: successive
    ``scipy.interpolate.interp1d`` along latitude then longitude (xarray decomposes orthogonal linear interpolation
    in indexer order) and ``pandas.Index.get_indexer(method='nearest')``.  Returns ``(U_new, lats, lons)``."""
    from scipy.interpolate import interp1d
    import pandas as pd
    lats = np.linspace(-89.75, 89.75, 180 * 2)                                         # :106
    lons = np.linspace(-180, 179.5, 360 * 2 + 1)                                       # :107
    U = np.asarray(U)
    near = U[:, pd.Index(lat).get_indexer(lats, method='nearest')][:, :, pd.Index(lon).get_indexer(lons, method='nearest')]
    it = interp1d(np.asarray(lat, dtype=np.float64), U, kind='linear', axis=1, bounds_error=False, fill_value=np.nan,
                  assume_sorted=True)(lats)
    it = interp1d(np.asarray(lon, dtype=np.float64), it, kind='linear', axis=2, bounds_error=False, fill_value=np.nan,
                  assume_sorted=True)(lons)
    return np.where(~np.isnan(it), it, near), lats, lons                               # :112-113


# --------------------------------------------------------------------------------------
# a2: LCS.__call__ (LCS.py:48-168), regional path and the cheap isglobal/truncation=None path
# --------------------------------------------------------------------------------------
def lcs_field(U, V, lat, lon, timestep, SETTLS_order=0, traj_interp_order=3,
              cyclic_xboundary=False, xclamp='outer', gauss_sigma=None, subdomain=None,
              return_dpts=False, return_traj=False, resample=None):
    """sigma_max field ``(nlat, nlon)`` for one window of winds ``(nt, nlat, nlon)``.

    The caller-side FTLE scaling ``0.5*log(sigma)`` (examples/ideal_vortex.py:282,288) is NOT
    applied here, as in the reference.
    """
    if resample is not None:                                                   # LCS.py:88-91: (times, freq)
        times, freq = resample
        U, new_t = resample_linear(U, times, freq)
        V, _ = resample_linear(V, times, freq)
        timestep = np.sign(timestep) * (new_t[1] - new_t[0]).astype('timedelta64[s]').astype('float')
    res = parcel_propagation(U, V, lat, lon, timestep, SETTLS_order=SETTLS_order,
                             interp_order=traj_interp_order, cyclic_xboundary=cyclic_xboundary,
                             xclamp=xclamp, return_traj=return_traj)          # LCS.py:129-134
    if return_traj:
        x_trajs, y_trajs = res
        x_dep, y_dep = x_trajs[-1], y_trajs[-1]                                # :135-139
    else:
        x_dep, y_dep = res
    def_tensor = flowmap_gradient(x_dep, y_dep, lat, lon, sigma=gauss_sigma)   # :142
    mask = subdomain_mask(lat, lon, subdomain) if isinstance(subdomain, dict) else None
    sigma = spectral_norm_field(def_tensor, mask)                              # :145-157
    out = (sigma,)
    if return_dpts:
        out += (x_dep, y_dep)
    if return_traj:
        out += (x_trajs, y_trajs)
    return out[0] if len(out) == 1 else out


# --------------------------------------------------------------------------------------
# Explicit restatement of the published scipy.ndimage algorithm (scipy 1.18.1, not in
# /root/reference: it is the third-party dependency behind tools.py:26,35).  It is the
# SPEC for the CUDA gather/prefilter kernels and is itself checked against scipy in
# tests/test_oracle_scipy_spec.py (gather: bit-exact; prefilter: a few ulp).
# --------------------------------------------------------------------------------------
SPLINE_POLE = np.sqrt(3.0) - 2.0


def spline_poles(order):
    """Poles of the B-spline prefilter, scipy ni_splines.c:get_filter_poles (orders 2..5)."""
    if order == 2:
        return [np.sqrt(8.0) - 3.0]
    if order == 3:
        return [np.sqrt(3.0) - 2.0]
    if order == 4:
        return [np.sqrt(664.0 - np.sqrt(438976.0)) + np.sqrt(304.0) - 19.0,
                np.sqrt(664.0 + np.sqrt(438976.0)) - np.sqrt(304.0) - 19.0]
    if order == 5:
        return [np.sqrt(67.5 - np.sqrt(4436.25)) + np.sqrt(26.25) - 6.5,
                np.sqrt(67.5 + np.sqrt(4436.25)) - np.sqrt(26.25) - 6.5]
    raise RuntimeError('spline order not supported')          # scipy's message for order outside 0..5


def prefilter_line_mirror(c, z=None, order=3):
    """B-spline prefilter of one line, mirror boundary, exact causal initialisation: the gain of all poles
    first, then per pole a causal and an anticausal recursion (scipy ni_splines.c:apply_filter)."""
    c = np.array(c, dtype=np.float64)
    n = c.shape[0]
    if n < 2:
        return c
    poles = [z] if z is not None else spline_poles(order)
    gain = 1.0
    for z in poles:
        gain *= (1.0 - z) * (1.0 - 1.0 / z)
    c *= gain
    for z in poles:
        z_n_1 = z ** (n - 1)
        z_i = z
        c0 = c[0] + z_n_1 * c[n - 1]
        for i in range(1, n - 1):
            c0 += z_i * (c[i] + z_n_1 * c[n - 1 - i])
            z_i *= z
        c[0] = c0 / (1 - z_n_1 * z_n_1)
        for i in range(1, n):
            c[i] += z * c[i - 1]
        c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1)
        for i in range(n - 2, -1, -1):
            c[i] = z * (c[i + 1] - c[i])
    return c


def prefilter_2d(field, order=3):
    """Axis 0 then axis 1, f64 (what map_coordinates does on every call)."""
    c = np.array(field, dtype=np.float64)
    for j in range(c.shape[1]):
        c[:, j] = prefilter_line_mirror(c[:, j], order=order)
    for i in range(c.shape[0]):
        c[i, :] = prefilter_line_mirror(c[i, :], order=order)
    return c


def fold_wrap(c, n):
    """scipy 'wrap' coordinate fold: period n-1."""
    sz = n - 1
    if c < 0:
        c += sz * (int(-c / sz) + 1)
    elif c > n - 1:
        c -= sz * int(c / sz)
    return c


def mirror_index(i, n):
    if i < 0:
        sz2 = 2 * n - 2
        i = sz2 * int(-i / sz2) + i
        i = i + sz2 if i <= 1 - n else -i
    elif i > n - 1:
        sz2 = 2 * n - 2
        i -= sz2 * int(i / sz2)
        if i >= n:
            i = sz2 - i
    return i


def spline_weights(c, order):
    """First tap index and the order+1 interpolation weights at coordinate ``c``
    (scipy ni_splines.c:get_spline_interpolation_weights and the `start` rule of NI_GeometricTransform),
    in scipy's order of operations: the restatement is bit-exact against scipy (tests/test_oracle_scipy_spec.py)."""
    if order & 1:
        start = int(np.floor(c)) - order // 2
        x = c - np.floor(c)
    else:
        start = int(np.floor(c + 0.5)) - order // 2
        x = c - np.floor(c + 0.5)
    y = x
    z = 1.0 - x
    w = [0.0] * (order + 1)
    if order == 1:
        w[0] = 1.0 - x
    elif order == 2:
        w[1] = 0.75 - x * x
        y = 0.5 - x
        w[0] = 0.5 * y * y
    elif order == 3:
        w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
        w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
        w[0] = z * z * z / 6.0
    elif order == 4:
        t = x * x
        w[2] = t * (t * 0.25 - 0.625) + 115.0 / 192.0
        y = 1.0 + x
        w[1] = y * (y * (y * (5.0 - y) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0
        w[3] = z * (z * (z * (5.0 - z) / 6.0 - 1.25) + 5.0 / 24.0) + 55.0 / 96.0
        y = 0.5 - x
        t = y * y
        w[0] = t * t / 24.0
    elif order == 5:
        t = y * y
        w[2] = t * (t * (0.25 - y / 12.0) - 0.5) + 0.55
        t = z * z
        w[3] = t * (t * (0.25 - z / 12.0) - 0.5) + 0.55
        y = y + 1.0
        w[1] = y * (y * (y * (y * (y / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425
        z = z + 1.0
        w[4] = z * (z * (z * (z * (z / 24.0 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425
        y = 1.0 - x
        t = y * y
        w[0] = y * t * t / 120.0
    else:
        raise RuntimeError('spline order not supported')
    last = 1.0
    for i in range(order):
        last -= w[i]
    w[order] = last
    return start, tuple(w)


def cubic_weights(x):
    return spline_weights(x, 3)


def gather_spline_wrap(coef, cy, cx, order=3):
    """map_coordinates(order, mode='wrap', prefilter=False) at one point: fold (period n-1), (order+1)^2 taps with
    mirrored indices, accumulated in scipy's order ((c*wy)*wx, axis-0 index outer)."""
    ny, nx = coef.shape
    cy = fold_wrap(cy, ny)
    cx = fold_wrap(cx, nx)
    sy, wy = spline_weights(cy, order)
    sx, wx = spline_weights(cx, order)
    t = 0.0
    for i in range(order + 1):
        for j in range(order + 1):
            v = coef[mirror_index(sy + i, ny), mirror_index(sx + j, nx)]
            v = v * wy[i]
            v = v * wx[j]
            t += v
    return t


def gather_cubic_wrap(coef, cy, cx):
    return gather_spline_wrap(coef, cy, cx, 3)


def gather_linear_constant(field, cy, cx):
    ny, nx = field.shape
    if cy < 0 or cy > ny - 1 or cx < 0 or cx > nx - 1:
        return 0.0
    fy = np.floor(cy)
    fx = np.floor(cx)
    y = cy - fy
    x = cx - fx
    wy = (1.0 - y, 1.0 - (1.0 - y))      # scipy: weights[order] = 1 - sum(others)
    wx = (1.0 - x, 1.0 - (1.0 - x))
    t = 0.0
    for i in range(2):
        for j in range(2):
            v = field[mirror_index(int(fy) + i, ny), mirror_index(int(fx) + j, nx)]
            v = v * wy[i]
            v = v * wx[j]
            t += v
    return t


def gather_linear_wrap(field, cy, cx):
    ny, nx = field.shape
    cy = fold_wrap(cy, ny)
    cx = fold_wrap(cx, nx)
    fy = np.floor(cy)
    fx = np.floor(cx)
    y = cy - fy
    x = cx - fx
    wy = (1.0 - y, 1.0 - (1.0 - y))      # scipy: weights[order] = 1 - sum(others)
    wx = (1.0 - x, 1.0 - (1.0 - x))
    t = 0.0
    for i in range(2):
        for j in range(2):
            v = field[mirror_index(int(fy) + i, ny), mirror_index(int(fx) + j, nx)]
            v = v * wy[i]
            v = v * wx[j]
            t += v
    return t
