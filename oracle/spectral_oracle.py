"""CPU restatement of the triangular spectral truncation of the global path (LCS.py:115-118) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product's tables live in
lagrangiancoherence_b200/spectral.py and are built by different code (barycentric interpolation + Gauss-Legendre
quadrature there; explicit DCT-I / DST-I sums + Fejer's second rule here), so agreement between the two is a
check of both.

PARITY UNPINNED against the reference: ``VectorWind(u, v).truncate(field, truncation)`` is windspharm -> pyspharm ->
SPHEREPACK (Fortran, single precision), none of which is in this image or in the reference tree, and the reference holds
no output of this step.  What is restated here is the published algorithm of that call chain:
  * windspharm.standard.VectorWind.truncate: ``spectogrd(grdtospec(field, ntrunc=truncation))``;
  * pyspharm Spharmt(nlon, nlat, gridtype='regular'): SPHEREPACK shaes / shses, rows at colatitudes
    ``theta_i = i*pi/(nlat-1)`` (poles included, whatever the data's latitudes), columns at ``2*pi*j/nlon``;
  * SPHEREPACK's analysis on that grid (Swarztrauber 1979; Adams & Swarztrauber 1999): the zonal Fourier coefficient of
    wavenumber m, as a function of colatitude, is replaced by its trigonometric interpolant -- cosine series for even m,
    sine series for odd m -- whose integral against ``Pbar_n^m(theta) sin(theta)`` is taken exactly; triangular
    truncation keeps ``m <= n <= ntrunc``; synthesis sums ``a_n^m Pbar_n^m(theta_i)`` back.
Known-answer tests (tests/test_oracle_spectral.py): spherical harmonics of degree <= ntrunc are reproduced, degree >
ntrunc are annihilated, the operator is idempotent and linear.
"""
from __future__ import annotations

from math import factorial

import numpy as np
from scipy.special import lpmv


def pbar(m, n, theta):
    """Associated Legendre function of cos(theta), orthonormal on [0, pi] for the weight sin(theta)."""
    norm = np.sqrt((2 * n + 1) / 2.0 * factorial(n - m) / factorial(n + m))
    return norm * lpmv(m, n, np.cos(theta))


def _interpolant(f, m, theta_dense):
    """Trigonometric interpolant of the grid values ``f[..., nlat]`` (rows at i*pi/(nlat-1)) evaluated at ``theta_dense``."""
    nlat = f.shape[-1]
    N = nlat - 1
    th = np.arange(nlat) * np.pi / N
    if m % 2 == 0:                                    # DCT-I: F = sum'' c_k cos(k theta)
        k = np.arange(N + 1)
        wi = np.ones(nlat); wi[0] = wi[-1] = 0.5
        c = (2.0 / N) * (f * wi) @ np.cos(np.outer(th, k))           # [..., N+1]
        wk = np.ones(N + 1); wk[0] = wk[-1] = 0.5
        return (c * wk) @ np.cos(np.outer(k, theta_dense))
    k = np.arange(1, N)                               # DST-I on the interior rows: F = sum s_k sin(k theta)
    s = (2.0 / N) * f[..., 1:-1] @ np.sin(np.outer(th[1:-1], k))
    return s @ np.sin(np.outer(k, theta_dense))


def truncate_field(g, ntrunc):
    """One field ``g[nlat, nlon]`` -> its triangular truncation at ``ntrunc`` on SPHEREPACK's regular grid."""
    g = np.asarray(g, dtype=np.float64)
    nlat, nlon = g.shape
    N = nlat - 1
    theta = np.arange(nlat) * np.pi / N
    phi = 2.0 * np.pi * np.arange(nlon) / nlon
    # Exact integration over [0, pi]: every integrand here, F(theta) Pbar_n^m(theta) sin(theta), is an ODD trigonometric
    # polynomial (a sine series of degree <= N + ntrunc + 1).  Its sine coefficients follow from its values at M - 1
    # interior equispaced points (DST-I, M > degree) and int_0^pi sin(k t) dt = 2/k for odd k, 0 for even k: that gives
    # the weights wt_j below (Fejer's second rule).
    M = 2
    while M <= N + ntrunc + 2:
        M *= 2
    j = np.arange(1, M)
    td = j * np.pi / M
    kodd = np.arange(1, M, 2)
    wt = (2.0 / M) * (np.sin(np.outer(td, kodd)) @ (2.0 / kodd))
    out = np.zeros_like(g)
    for m in range(ntrunc + 1):
        parts = [g @ np.cos(m * phi) * ((1.0 if m == 0 else 2.0) / nlon)]
        if m > 0:
            parts.append(g @ np.sin(m * phi) * (2.0 / nlon))
        for which, f in enumerate(parts):             # f[nlat]: the coefficient as a function of colatitude
            F = _interpolant(f[None, :], m, td)[0]
            rec = np.zeros(nlat)
            for n in range(m, ntrunc + 1):
                a = np.sum(wt * F * pbar(m, n, td) * np.sin(td))
                rec += a * pbar(m, n, theta)
            out += np.outer(rec, np.cos(m * phi) if which == 0 else np.sin(m * phi))
    return out


def truncate_series(series, ntrunc):
    """``[nlev, nlat, nlon]`` -> truncated series (LCS.py:116-118 applies this to u and to v separately)."""
    return np.stack([truncate_field(f, ntrunc) for f in np.asarray(series)])
