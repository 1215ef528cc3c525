"""The spectral truncation of the global path (LCS.py:115-118): the oracle restatement against analytic known answers,
and the product's host tables (lagrangiancoherence_b200/spectral.py) against the oracle.  Parity against the reference
itself is UNPINNED for this step (windspharm / pyspharm / SPHEREPACK are not available): see the module headers."""
import numpy as np
import pytest

from oracle import spectral_oracle as SO
from lagrangiancoherence_b200 import spectral


def harmonic(n, m, nlat, nlon, kind='cos'):
    theta = np.arange(nlat) * np.pi / (nlat - 1)
    phi = 2 * np.pi * np.arange(nlon) / nlon
    return np.outer(SO.pbar(m, n, theta), np.cos(m * phi) if kind == 'cos' else np.sin(m * phi))


@pytest.mark.parametrize('nlat,nlon,T', [(37, 72, 10), (40, 81, 12)])
def test_oracle_keeps_resolved_harmonics_and_removes_the_rest(nlat, nlon, T):
    rng = np.random.default_rng(0)
    keep = sum(rng.normal() * harmonic(n, m, nlat, nlon, k) for n, m, k in
               [(0, 0, 'cos'), (1, 0, 'cos'), (3, 2, 'sin'), (T, T, 'cos'), (T, 1, 'sin'), (T - 1, T - 2, 'cos'), (5, 5, 'sin')])
    drop = sum(rng.normal() * harmonic(n, m, nlat, nlon, k) for n, m, k in
               [(T + 1, 0, 'cos'), (T + 1, T + 1, 'sin'), (T + 3, 2, 'cos'), (nlat // 2, 1, 'sin'), (T + 2, T, 'cos')])
    out = SO.truncate_field(keep + drop, T)
    assert np.abs(out - keep).max() <= 1e-12 * np.abs(keep).max()
    again = SO.truncate_field(out, T)
    assert np.abs(again - out).max() <= 1e-12 * np.abs(out).max()           # idempotent


def test_oracle_is_linear_and_real():
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=(19, 40)), rng.normal(size=(19, 40))
    lhs = SO.truncate_field(2.0 * a - 3.0 * b, 6)
    rhs = 2.0 * SO.truncate_field(a, 6) - 3.0 * SO.truncate_field(b, 6)
    assert np.abs(lhs - rhs).max() <= 1e-12 * np.abs(lhs).max()


@pytest.mark.parametrize('nlat,nlon,T', [(37, 72, 10), (40, 81, 12), (31, 64, 20)])
def test_product_tables_equal_the_oracle_operator(nlat, nlon, T):
    """A_m, Fc, Fi of the product (barycentric interpolation + Gauss-Legendre) applied with plain matrix products give
    the oracle's result (DCT/DST sums + Fejer quadrature) on arbitrary, non-band-limited fields."""
    rng = np.random.default_rng(2)
    g = rng.normal(size=(nlat, nlon)) + 5.0
    A, Fc, Fi = spectral.truncation_tables(nlat, nlon, T)
    Y = g @ Fc
    Z = np.empty_like(Y)
    Z[:, 0] = A[0] @ Y[:, 0]
    for m in range(1, T + 1):
        Z[:, 2 * m - 1] = A[m] @ Y[:, 2 * m - 1]
        Z[:, 2 * m] = A[m] @ Y[:, 2 * m]
    out = Z @ Fi
    ref = SO.truncate_field(g, T)
    assert np.abs(out - ref).max() <= 1e-11 * np.abs(ref).max()


def test_reference_common_grid_is_accepted_and_other_grids_are_refused():
    from lagrangiancoherence_b200.regrid import common_grid
    lats, lons = common_grid()
    spectral.check_regular_global_grid(lats, lons)                       # 360 cell-centred rows, 721 columns
    spectral.check_regular_global_grid(np.linspace(90, -90, 73), np.arange(0, 360, 2.5))
    with pytest.raises(ValueError, match='non-global'):
        spectral.check_regular_global_grid(np.linspace(-55, 15, 281), np.linspace(-100, -20, 321))
    with pytest.raises(ValueError):
        spectral.truncation_tables(10, 16, 12)
