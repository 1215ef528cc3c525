"""Synthetic wind fields and grids for the bench, the tests and the examples.

Pure numpy; shapes and formulas are the named configurations C1..C5 of SURVEY.md section 8(d).
The four analytic generators restate the test-input recipes of the reference's example
(examples/ideal_vortex.py:11-208) in vectorised form and return plain
``(u, v, lat, lon)`` with ``u, v`` shaped ``(nt, nlat, nlon)``.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------ grids
def grid_c1():
    """Ideal-vortex grid: arange(-88, 89, 2) x arange(-180, 180, 2) (ideal_vortex.py:158-159,220)."""
    return np.arange(-88, 89, 2).astype(np.float64), np.arange(-180, 180, 2).astype(np.float64)


def grid_c2():
    """0.25 deg South-America domain, lat -55..15, lon -100..-20 (281 x 321)."""
    return np.linspace(-55.0, 15.0, 281), np.linspace(-100.0, -20.0, 321)


def grid_c3():
    """0.25 deg near-global limited domain, lat -90..90, lon -180..179.75 (721 x 1440)."""
    return np.linspace(-90.0, 90.0, 721), np.linspace(-180.0, 179.75, 1440)


def era5_like_winds(lat, lon, nt, t0=0, seed=0, noise=0.5, contained=False, dtype=np.float64):
    """Smooth multi-mode winds (m/s) of SURVEY.md 8(d); level index ``t0 + k`` enters the phase.

    ``contained=True`` multiplies by a cos^2 taper that vanishes on the lateral boundaries so
    that no particle leaves the domain (outer-product and pointwise x-clamps then agree).
    The noise term is a fixed spatial pattern per level drawn from ``default_rng(seed + level)``
    so that rolling windows cut from a long series see identical levels.
    """
    phi = np.deg2rad(lat)[:, None]
    lam = np.deg2rad(lon)[None, :]
    u = np.empty((nt, lat.size, lon.size), dtype=np.float64)
    v = np.empty_like(u)
    for k in range(nt):
        t = float(t0 + k)
        u[k] = 10 + 8 * np.sin(3 * phi + 0.1 * t) * np.cos(2 * lam) + 3 * np.sin(7 * lam + 5 * phi - 0.1 * t)
        v[k] = 6 * np.cos(4 * lam - 0.1 * t) * np.sin(3 * phi) + 3 * np.cos(6 * phi - 5 * lam + 0.1 * t)
        if noise:
            rng = np.random.default_rng(seed + t0 + k)
            u[k] += noise * _smooth(rng.normal(size=u[k].shape))
            v[k] += noise * _smooth(rng.normal(size=v[k].shape))
    if contained:
        ty = np.sin(np.pi * (lat - lat[0]) / (lat[-1] - lat[0])) ** 2
        tx = np.sin(np.pi * (lon - lon[0]) / (lon[-1] - lon[0])) ** 2
        taper = ty[:, None] * tx[None, :]
        u *= taper
        v *= taper
    return u.astype(dtype), v.astype(dtype)


def _smooth(a, passes=2):
    """Cheap separable 1-2-1 smoothing (keeps the noise band-limited like analysed winds)."""
    for _ in range(passes):
        a = 0.25 * (np.roll(a, 1, 0) + np.roll(a, -1, 0)) + 0.5 * a
        a = 0.25 * (np.roll(a, 1, 1) + np.roll(a, -1, 1)) + 0.5 * a
    return a


# ------------------------------------------------------------------ analytic generators
def _axes(lat_min, lat_max, lon_min, lon_max, dx, dy):
    return np.arange(lat_min, lat_max, dy).astype(np.float64), np.arange(lon_min, lon_max, dx).astype(np.float64)


def ideal_saddle(lat_min, lat_max, lon_min, lon_max, dx, dy, nt, max_intensity=10):
    """u grows linearly with the row index, v with the column index (ideal_vortex.py:11-42)."""
    lat, lon = _axes(lat_min, lat_max, lon_min, lon_max, dx, dy)
    ny, nx = lat.size, lon.size
    u2 = (max_intensity * np.arange(ny) / ny - .5 * max_intensity)[:, None] * np.ones((1, nx))
    v2 = np.ones((ny, 1)) * (max_intensity * np.arange(nx) / nx - .5 * max_intensity)[None, :]
    return np.repeat(u2[None], nt, 0), np.repeat(v2[None], nt, 0), lat, lon


def rotating_saddle(lat_min, lat_max, lon_min, lon_max, dx, dy, nt, max_intensity=10, center=(0, 0), **_):
    """ideal_vortex.py:45-86."""
    lat, lon = _axes(lat_min, lat_max, lon_min, lon_max, dx, dy)
    new_x = ((lon - center[0]) / 180)[None, :]
    new_y = ((lat - center[1]) / 90)[:, None]
    u = np.empty((nt, lat.size, lon.size))
    v = np.empty_like(u)
    for t in range(nt):
        s, c = np.sin(4 * t / nt), np.cos(4 * t / nt)
        u[t] = np.sqrt(2) * max_intensity * (s * new_x + (2 + c) * new_y)
        v[t] = np.sqrt(2) * max_intensity * ((-2 * c) * new_x - s * new_y)
    return u, v, lat, lon


def shear_flow(lat_min, lat_max, lon_min, lon_max, dx, dy, nt, max_intensity=10, **_):
    """Uniform zonal wind (ideal_vortex.py:89-129)."""
    lat, lon = _axes(lat_min, lat_max, lon_min, lon_max, dx, dy)
    u = np.full((nt, lat.size, lon.size), float(max_intensity))
    return u, np.zeros_like(u), lat, lon


def ideal_vortex(lat_min, lat_max, lon_min, lon_max, dx, dy, nt, max_intensity=10, radius=5,
                 center=(0, 0), u_c=0, v_c=0, basic_zonal=2, k=0, **_):
    """Rankine-like vortex with optional translation (ideal_vortex.py:132-208)."""
    lat, lon = _axes(lat_min, lat_max, lon_min, lon_max, dx, dy)
    u = np.empty((nt, lat.size, lon.size))
    v = np.empty_like(u)
    for t in range(nt):
        new_x = (lon - center[0] - u_c * t)[None, :] * np.ones((lat.size, 1))
        if k > 0:
            new_y = lat - center[1] - v_c * np.sin(k * 2 * np.pi * t / nt)
        elif k == 0:
            new_y = lat - center[1] - v_c * t
        else:
            raise ValueError('Meridional wavenumber k must be greater than zero.')
        new_y = new_y[:, None] * np.ones((1, lon.size))
        distance = np.sqrt(new_x ** 2 + new_y ** 2)
        theta = np.arccos(new_y / (distance + 1e-8))
        mag = np.where(distance > radius, max_intensity * radius ** 2 / (2 * np.where(distance > 0, distance, 1)),
                       max_intensity * 0.5 * distance)
        u[t] = np.cos(theta) * mag + basic_zonal
        v[t] = np.where(new_x < 0, np.sin(theta) * mag, np.sin(theta + np.pi) * mag)
    return u, v, lat, lon


vortex_config_subtropical = {'lat_min': -88, 'lat_max': 89, 'lon_min': -180, 'lon_max': 180, 'dx': 2,
                             'dy': 2, 'u_c': 0, 'k': 0, 'v_c': 0, 'nt': 8, 'radius': 2,
                             'max_intensity': 60, 'center': [-55, -20], 'basic_zonal': 0}
saddle_config = {'lat_min': -70, 'lat_max': -10, 'lon_min': -70, 'lon_max': -10, 'dx': 1, 'dy': 1, 'nt': 10,
                 'max_intensity': 10}
shear_flow_config = {'lat_min': -40, 'lat_max': 40, 'lon_min': -60, 'lon_max': 20, 'dx': 1, 'dy': 1,
                     'nt': 30, 'max_intensity': 1, 'center': [-20, 0]}
