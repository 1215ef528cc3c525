"""Triangular spectral truncation of the global path (LCS.py:115-118): host-side tables.

Upstream: ``VectorWind(u, v).truncate(u, truncation=20)`` -- windspharm hands each field to pyspharm's
``grdtospec(field, ntrunc)`` / ``spectogrd`` pair, i.e. SPHEREPACK's ``shaes`` analysis followed by ``shses`` synthesis
on its "regular" grid.  SPHEREPACK places the ``nlat`` rows it is given at colatitudes ``theta_i = i*pi/(nlat-1)``, poles
included, whatever the latitudes of the data were (the 360 rows of the reference's common grid lie at +-89.75 ... and are
treated as if they reached the poles), and the ``nlon`` columns at ``2*pi*j/nlon``.  Its analysis is Swarztrauber's:
the m-th zonal Fourier coefficient, as a function of colatitude, is replaced by its trigonometric interpolant through
the grid values (a cosine series for even m, a sine series for odd m) and the integral of that interpolant against the
normalised associated Legendre function ``Pbar_n^m(theta) sin(theta)`` is taken exactly.  Truncation at T keeps
``m <= n <= T``.  The whole operation is linear and separable:

    out = sum_m  A_m . ( field . F_m ) . F_m^+          A_m = S_m W_m  (nlat x nlat, rank T - m + 1)

with ``F_m`` the cos / sin columns of the longitude DFT, ``W_m`` the analysis (quadrature) rows and ``S_m`` the
synthesis columns.  This module builds ``A_m`` and the two DFT matrices in f64; ``lcs_spectral_truncate`` applies them
on the device.

PARITY UNPINNED against the reference: windspharm / pyspharm / SPHEREPACK are absent from this image (no source, no
wheel), pyspharm computes in single precision, and nothing in the reference tree holds an output of this step.  What is
pinned: the operator is checked against an independent restatement in ``oracle/spectral_oracle.py`` (different
interpolation and quadrature code) and against analytic known answers -- harmonics of degree <= T pass unchanged,
degree > T vanish, the operator is idempotent (tests/test_oracle_spectral.py).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np


def legendre_normalised(m, nmax, theta):
    """``Pbar_n^m(theta)`` for ``n = m .. nmax`` (rows), orthonormal on [0, pi] for the weight sin(theta):
    ``Pbar = sqrt((2n+1)/2 (n-m)!/(n+m)!) P_n^m(cos theta)``, by the standard three-term recurrence in n."""
    theta = np.asarray(theta, dtype=np.float64)
    x, s = np.cos(theta), np.sin(theta)
    out = np.zeros((nmax - m + 1,) + theta.shape)
    # Pbar_m^m = sqrt((2m+1)!! / (2 (2m)!!)) sin^m
    pmm = np.full_like(theta, np.sqrt(0.5))
    for k in range(1, m + 1):
        pmm = pmm * np.sqrt((2.0 * k + 1.0) / (2.0 * k)) * s
    out[0] = pmm
    if nmax > m:
        out[1] = np.sqrt(2.0 * m + 3.0) * x * pmm
    for n in range(m + 2, nmax + 1):
        a = np.sqrt((4.0 * n * n - 1.0) / (n * n - m * m))
        b = np.sqrt(((n - 1.0) ** 2 - m * m) / (4.0 * (n - 1.0) ** 2 - 1.0))
        out[n - m] = a * (x * out[n - m - 1] - b * out[n - m - 2])
    return out


def _bary_matrix(nodes, weights, targets):
    """Barycentric interpolation matrix from ``nodes`` (barycentric ``weights``) to ``targets``."""
    d = targets[:, None] - nodes[None, :]
    hit = d == 0.0
    d[hit] = 1.0
    c = weights[None, :] / d
    mat = c / c.sum(axis=1, keepdims=True)
    rows = hit.any(axis=1)
    mat[rows] = hit[rows].astype(np.float64)
    return mat


@lru_cache(maxsize=8)
def truncation_tables(nlat, nlon, ntrunc):
    """``(A [T+1, nlat, nlat], Fc [nlon, 2T+1], Fi [2T+1, nlon])`` of the truncation operator described above.

    Column c of ``Fc``: c = 0 -> the zonal mean, c = 2m-1 / 2m -> the cos / sin coefficient of wavenumber m; ``Fi`` maps
    them back to the grid.  Rows of ``A_m`` act on a coefficient as a function of the (SPHEREPACK) colatitude."""
    T = int(ntrunc)
    if nlat < 3 or nlon < 4:
        raise ValueError('spectral truncation needs at least 3 latitudes and 4 longitudes')
    if T < 0 or T > nlat - 1 or 2 * T + 1 > nlon:
        raise ValueError(f'truncation {T} is not resolved by a {nlat} x {nlon} grid')
    N = nlat - 1
    i = np.arange(nlat)
    theta = i * np.pi / N                                           # SPHEREPACK's colatitudes, poles included
    xg = np.cos(theta)
    # exact quadrature in x = cos(theta): the integrands are polynomials of degree <= N + T
    nq = (N + T) // 2 + 2
    xq, wq = np.polynomial.legendre.leggauss(nq)
    thq = np.arccos(xq)
    # even m: interpolant = polynomial of degree N in x through the Chebyshev-Lobatto points (a cosine series in theta)
    w_even = (-1.0) ** i
    w_even[0] *= 0.5
    w_even[-1] *= 0.5
    E_even = _bary_matrix(xg, w_even, xq)
    # odd m: interpolant = sin(theta) * polynomial of degree N-2 through f_i / sin(theta_i) at the interior points
    # (a sine series in theta; the pole rows do not enter)
    ii = i[1:-1]
    w_odd = (-1.0) ** ii * np.sin(theta[1:-1]) ** 2
    E_odd = np.zeros((nq, nlat))
    E_odd[:, 1:-1] = np.sin(thq)[:, None] * _bary_matrix(xg[1:-1], w_odd, xq) / np.sin(theta[1:-1])[None, :]
    A = np.zeros((T + 1, nlat, nlat))
    for m in range(T + 1):
        Pq = legendre_normalised(m, T, thq)                         # [T-m+1, nq]
        Pg = legendre_normalised(m, T, theta)                       # [T-m+1, nlat]
        W = (Pq * wq[None, :]) @ (E_even if m % 2 == 0 else E_odd)  # analysis rows
        A[m] = Pg.T @ W
    phi = 2.0 * np.pi * np.arange(nlon) / nlon
    Fc = np.zeros((nlon, 2 * T + 1))
    Fi = np.zeros((2 * T + 1, nlon))
    Fc[:, 0] = 1.0 / nlon
    Fi[0] = 1.0
    for m in range(1, T + 1):
        Fc[:, 2 * m - 1] = 2.0 * np.cos(m * phi) / nlon
        Fc[:, 2 * m] = 2.0 * np.sin(m * phi) / nlon
        Fi[2 * m - 1] = np.cos(m * phi)
        Fi[2 * m] = np.sin(m * phi)
    return A, Fc, Fi


def check_regular_global_grid(lat, lon):
    """windspharm refuses grids that are not global and equally spaced (``ValueError``); same here.  Accepted latitudes:
    equally spaced, either pole to pole (odd or even count) or cell-centred (+-(90 - d/2)), in either direction."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    d = np.diff(lat)
    if lat.size < 3 or np.abs(d - d[0]).max() > 1e-3 * abs(d[0]):
        raise ValueError('Invalid equally-spaced latitudes (they may be non-global)')
    step, top = abs(d[0]), max(abs(lat[0]), abs(lat[-1]))
    if not (abs(top - 90.0) < 1e-3 or abs(top - (90.0 - 0.5 * step)) < 1e-3) or abs(lat[0] + lat[-1]) > 1e-3:
        raise ValueError('Invalid equally-spaced latitudes (they may be non-global)')
    dl = np.diff(lon)
    if lon.size < 4 or np.abs(dl - dl[0]).max() > 1e-3 * abs(dl[0]):
        raise ValueError('longitudes must be equally spaced to be truncated spectrally')
