"""Pin the explicit restatement of scipy.ndimage's algorithm (the SPEC of the CUDA gather/prefilter
kernels, oracle/lcs_oracle.py) against scipy 1.18.1 itself -- the third-party call behind
tools.py:26-30,35-39 of the reference."""
import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import lcs_oracle as O


@pytest.fixture(scope='module')
def field():
    return np.random.default_rng(5).normal(size=(23, 31)) * 10


def coords(n, ny, nx, seed):
    rng = np.random.default_rng(seed)
    cy, cx = rng.uniform(-30, 60, n), rng.uniform(-40, 80, n)
    # edge cases: grid points, last index, one past (Q4 maps the last row/col there), exact periods, tiny offsets
    special = [(0, 0), (ny - 1, nx - 1), (ny, nx), (2 * (ny - 1), 2 * (nx - 1)), (-1e-9, nx - 1 + 1e-9),
               (ny - 1 - 1e-12, 0.5), (-(ny - 1), -(nx - 1)), (0.5, nx - 1.5)]
    for i, (a, b) in enumerate(special):
        cy[i], cx[i] = a, b
    return cy, cx


def test_cubic_wrap_gather_is_bit_exact(field):
    C = ndi.spline_filter(field, order=3, output=np.float64, mode='mirror')
    cy, cx = coords(20000, *field.shape, seed=1)
    ref = ndi.map_coordinates(C, np.array([cy, cx]), order=3, mode='wrap', prefilter=False)
    mine = np.array([O.gather_cubic_wrap(C, a, b) for a, b in zip(cy, cx)])
    assert np.array_equal(ref, mine)


@pytest.mark.parametrize('order', [2, 4, 5])
def test_other_spline_orders_gather_is_bit_exact(field, order):
    """traj_interp_order is free upstream (any order scipy accepts); orders 2, 4, 5: weights, first-tap rule
    (floor(x + 0.5) for even orders) and mirrored taps, bit for bit against scipy."""
    C = ndi.spline_filter(field, order=order, output=np.float64, mode='mirror')
    cy, cx = coords(6000, *field.shape, seed=10 + order)
    ref = ndi.map_coordinates(C, np.array([cy, cx]), order=order, mode='wrap', prefilter=False)
    mine = np.array([O.gather_spline_wrap(C, a, b, order) for a, b in zip(cy, cx)])
    assert np.array_equal(ref, mine)


@pytest.mark.parametrize('order', [2, 4, 5])
def test_other_spline_orders_prefilter_and_composition(field, order):
    """Poles and gain of orders 2, 4, 5 (two poles for 4 and 5: conditioning costs a digit or two, stated 1e-13)."""
    ref = ndi.spline_filter(field, order=order, output=np.float64, mode='mirror')
    C = O.prefilter_2d(field, order=order)
    assert np.abs(C - ref).max() <= 1e-13 * np.abs(ref).max()
    cy, cx = coords(2000, *field.shape, seed=20 + order)
    full = ndi.map_coordinates(field, np.array([cy, cx]), order=order, mode='wrap')
    mine = np.array([O.gather_spline_wrap(C, a, b, order) for a, b in zip(cy, cx)])
    assert np.abs(full - mine).max() <= 1e-12 * np.abs(field).max()
    with pytest.raises(RuntimeError):
        O.spline_poles(6)


def test_linear_wrap_and_constant_gathers_are_bit_exact(field):
    cy, cx = coords(20000, *field.shape, seed=2)
    ref = ndi.map_coordinates(field, np.array([cy, cx]), order=1, mode='wrap')
    assert np.array_equal(ref, np.array([O.gather_linear_wrap(field, a, b) for a, b in zip(cy, cx)]))
    cy, cx = np.clip(cy, -3, field.shape[0] + 2), np.clip(cx, -3, field.shape[1] + 2)
    ref = ndi.map_coordinates(field, np.array([cy, cx]), order=1, mode='constant')
    assert np.array_equal(ref, np.array([O.gather_linear_constant(field, a, b) for a, b in zip(cy, cx)]))


@pytest.mark.parametrize('shape', [(5, 7), (23, 31), (64, 33)])
def test_prefilter_restatement_within_a_few_ulp(shape):
    f = np.random.default_rng(3).normal(size=shape) * 10
    ref = ndi.spline_filter(f, order=3, output=np.float64, mode='mirror')
    assert np.abs(O.prefilter_2d(f) - ref).max() <= 1e-14 * np.abs(ref).max()
    # mode='wrap' prefilters with the mirror initialisation too (scipy >= 1.6): the coefficients are identical
    assert np.array_equal(ref, ndi.spline_filter(f, order=3, output=np.float64, mode='wrap'))


def test_two_sided_fir_form_of_the_prefilter():
    """The CUDA prefilter evaluates c[i] = sum_k sqrt(3) z^|k| s[mirror(i+k)], |k| <= 32 (prefilter.cu)."""
    rng = np.random.default_rng(4)
    z = np.sqrt(3.0) - 2.0
    h0 = (1 - z) * (1 - 1 / z) * (-z) / (1 - z * z)
    assert abs(h0 - np.sqrt(3.0)) < 1e-15
    for n in (5, 9, 40, 321):
        s = rng.normal(size=n)
        ref = ndi.spline_filter1d(s, order=3, output=np.float64, mode='mirror')
        got = np.array([sum(h0 * z ** abs(k) * s[O.mirror_index(i + k, n)] for k in range(-32, 33)) for i in range(n)])
        assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()


def test_full_map_coordinates_composition(field):
    """prefilter (restated) + gather (restated) == map_coordinates(order=3, mode='wrap') to a few ulp."""
    C = O.prefilter_2d(field)
    cy, cx = coords(3000, *field.shape, seed=6)
    ref = ndi.map_coordinates(field, np.array([cy, cx]), order=3, mode='wrap')
    mine = np.array([O.gather_cubic_wrap(C, a, b) for a, b in zip(cy, cx)])
    assert np.abs(ref - mine).max() <= 1e-13 * np.abs(field).max()


def test_stencil_vectorised_form_has_numba_rounding():
    """The oracle's numpy stencil equals a numba-jitted loop twin of tools.py:190-245 bit for bit
    (f32 differences, f64 combination, f32 store)."""
    k = O._numba_fourth_order_derivative()
    a = (np.random.default_rng(0).normal(size=(40, 50)) * 6e6).astype('float32')
    for dim in (0, 1):
        assert np.array_equal(k(a, dim, True), O.fourth_order_derivative(a, dim, True))


def test_sigma_closed_form_matches_lapack_svd():
    rng = np.random.default_rng(1)
    dt = np.concatenate([rng.normal(size=(6, 30, 40)) * np.array([1, 1e-3, 10, 1, 1e2, 1])[:, None, None],
                         np.zeros((3, 30, 40))])
    ref = O.spectral_norm_field(dt)
    assert np.abs(O.sigma_max_closed_form(dt) - ref).max() <= 1e-13 * ref.max()
