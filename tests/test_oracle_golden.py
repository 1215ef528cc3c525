"""The oracle against outputs of the UNMODIFIED reference (run under oracle/refshim by oracle/make_golden.py,
fixtures committed under tests/golden/).  The oracle calls the same scipy/numba arithmetic in the same order, so
agreement is demanded bit for bit."""
import json
import os

import numpy as np
import pytest

from oracle import lcs_oracle as O
from oracle.make_golden import CASES, make_inputs
from lagrangiancoherence_b200 import synthetic as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def test_manifest_matches_the_generator():
    manifest = json.load(open(os.path.join(GOLDEN, 'manifest.json')))
    assert {k: v for k, v in manifest.items() if k != '_meta'} == json.loads(json.dumps(CASES))
    assert manifest['_meta']['scipy'] == '1.18.1'        # the oracle's pinned third-party versions


@pytest.mark.parametrize('name', list(CASES))
def test_oracle_reproduces_reference_bitwise(name):
    case, g = CASES[name], load(name)
    u, v, lat, lon, time = make_inputs(case)
    x, y = O.parcel_propagation(u, v, lat, lon, case['timestep'], SETTLS_order=case['S'], interp_order=case['order'],
                                cyclic_xboundary=case['cyclic'], xclamp='outer', return_traj=True)
    if 'x_traj' in g:
        lv = g['traj_levels']
        assert np.array_equal(x[lv], g['x_traj']) and np.array_equal(y[lv], g['y_traj'])
        labels = time[::-1] if case['timestep'] < 0 else time              # trajectory.py:59-60
        assert np.array_equal(g['traj_time'], labels.astype('int64'))
    assert np.array_equal(x[-1], g['x_dep']) and np.array_equal(y[-1], g['y_dep'])
    sigma = O.lcs_field(u, v, lat, lon, case['timestep'], SETTLS_order=case['S'], traj_interp_order=case['order'],
                        cyclic_xboundary=case['cyclic'])
    assert np.array_equal(sigma, g['sigma'][0], equal_nan=True)
    assert np.array_equal(g['sigma_lat'], lat) and np.array_equal(g['sigma_lon'], lon)     # ascending, whatever came in
    stamp = time[-1] if case['timestep'] > 0 else time[0]                    # LCS.py:158
    assert g['sigma_time'][0] == stamp.astype('int64')
    if 'def_tensor' in g:
        assert np.array_equal(O.flowmap_gradient(x[-1], y[-1], lat, lon), g['def_tensor'])


def test_seams_reproduce_reference_bitwise():
    g = load('seams')
    case = CASES['regional_outer_p3']
    u, v, lat, lon, _ = make_inputs(case)
    for order in (1, 2, 3, 4, 5):
        assert np.array_equal(O.xr_map_coordinates(u[0], g['px'], g['py'], lat, lon, order=order), g[f'map_coordinates_p{order}'])
    for dim in (0, 1):
        assert np.array_equal(O.derivative_spherical_coords(g['X'], lat, lon, dim=dim), g[f'derivative_spherical_dim{dim}'])
        for isglobal in (True, False):
            ref = g[f'fourth_order_dim{dim}_global{int(isglobal)}']      # the reference's own numba function
            assert ref.dtype == np.float32
            assert np.array_equal(O.fourth_order_derivative(g['X'].astype('float32'), dim=dim, isglobal=isglobal), ref)
    sub = {'latitude': slice(-20, 0), 'longitude': slice(-70, -40)}
    sigma = O.lcs_field(u, v, lat, lon, case['timestep'], SETTLS_order=case['S'], subdomain=sub)
    keep = O.subdomain_mask(lat, lon, sub)
    rows, cols = keep.any(1), keep.any(0)
    assert np.array_equal(g['subdomain_lat'], lat[rows]) and np.array_equal(g['subdomain_lon'], lon[cols])
    assert np.array_equal(sigma[rows][:, cols], g['subdomain_sigma'][0])


def test_outer_clamp_is_what_the_reference_executes():
    """On the exiting-particles case the pointwise clamp does NOT reproduce the reference, the outer-product one does."""
    case, g = CASES['regional_outer_p3'], load('regional_outer_p3')
    u, v, lat, lon, _ = make_inputs(case)
    xp, _ = O.parcel_propagation(u, v, lat, lon, case['timestep'], SETTLS_order=case['S'], xclamp='pointwise')
    assert (xp != g['x_dep']).mean() > 0.01
    case, g = CASES['regional_contained'], load('regional_contained')
    u, v, lat, lon, _ = make_inputs(case)
    xp, _ = O.parcel_propagation(u, v, lat, lon, case['timestep'], SETTLS_order=case['S'], xclamp='pointwise')
    assert np.array_equal(xp, g['x_dep'])


def test_resample_and_gauss_sigma_reproduce_reference_bitwise():
    """resample='3h' (LCS.py:88-91: pandas bin labels + scipy interp1d in time, timestep re-derived at :91) and
    gauss_sigma (LCS.py:187-190: scipy gaussian_filter of the departure points) through the unmodified reference."""
    g = load('seams')
    case = CASES['regional_outer_p3']
    u, v, lat, lon, time = make_inputs(case)
    sig, xd, yd = O.lcs_field(u, v, lat, lon, case['timestep'], SETTLS_order=2, resample=(time, '3h'), return_dpts=True)
    assert np.array_equal(xd, g['resample_x_dep']) and np.array_equal(yd, g['resample_y_dep'])
    assert np.array_equal(sig, g['resample_sigma'][0])
    assert g['resample_time'][0] == time[0].astype('int64')
    sig = O.lcs_field(u, v, lat, lon, case['timestep'], SETTLS_order=case['S'], gauss_sigma=1.5)
    assert np.array_equal(sig, g['gauss_sigma_field'][0])


def test_resample_plan_matches_oracle_weights():
    from lagrangiancoherence_b200.timeaxis import resample_plan
    _, _, _, _, time = make_inputs(CASES['regional_outer_p3'])
    u = np.random.default_rng(0).normal(size=(time.size, 4, 5))
    new, lo, w_hi, w_lo = resample_plan(time, '3h')
    ref, new_ref = O.resample_linear(u, time, '3h')
    assert np.array_equal(new, new_ref) and new.size == 2 * time.size - 1
    got = w_hi[:, None, None] * u[lo + 1] + w_lo[:, None, None] * u[lo]
    assert np.array_equal(got, ref)
    assert np.array_equal(got[::2], u)              # levels that coincide with input levels come back exactly


def test_ridge_filter_reproduces_reference_bitwise():
    """find_ridges_spherical_hessian (tools.py:52-155) through the unmodified reference, incl. the per-point
    np.linalg.eig loop and its row-of-the-eigenvector-matrix quirk."""
    g = load('seams')
    _, _, lat, lon, _ = make_inputs(CASES['regional_outer_p3'])
    dt_prod, eigmin = O.find_ridges_spherical_hessian(g['ridge_input'], lat, lon, sigma=1.2, tolerance_threshold=0.002e-3)
    assert np.array_equal(dt_prod, g['ridge_dt_prod']) and np.array_equal(eigmin, g['ridge_eigmin'])
    six = O.find_ridges_spherical_hessian(g['ridge_input'], lat, lon, sigma=1.2, tolerance_threshold=0.002e-3,
                                          return_eigvectors=True)                       # tools.py:148-152
    for got, key in zip(six, ('ridge_dt_prod', 'ridge_eigmin', 'ridge_dt_raw', 'ridge_eigvectors', 'ridge_gradient',
                              'ridge_angle')):
        assert np.array_equal(got, g[key], equal_nan=True), key
    assert 0.02 < dt_prod.mean() < 0.98                           # a non-trivial mask


def test_lapack_2x2_conventions_closed_form():
    """The closed form the CUDA ridge kernel uses equals np.linalg.eig (dgeev/dlanv2) on symmetric 2x2 input:
    eigenvalues bit for bit in LAPACK's order, eigenvectors to one ulp, incl. b = 0 and a = d."""
    rng = np.random.default_rng(0)
    n = 50000
    a, d = rng.normal(size=n), rng.normal(size=n)
    b = rng.normal(size=n) * rng.choice([0, 1, 1e-3, 1e3], size=n)
    a[:10] = d[:10]
    b[10:20] = 0
    a[10:15] = d[10:15]
    w, v = np.linalg.eig(np.stack([np.stack([a, b], -1), np.stack([b, d], -1)], -2))
    rt1, rt2, cs, sn = O.eig_sym2x2_lapack(a, b, d)
    assert np.array_equal(w[:, 0], rt1) and np.array_equal(w[:, 1], rt2)
    for got, ref in ((v[:, 0, 0], cs), (v[:, 1, 0], sn), (v[:, 0, 1], -sn), (v[:, 1, 1], cs)):
        assert np.abs(got - ref).max() <= 2.3e-16


def test_global_regrid_path_reproduces_reference_bitwise():
    """SURVEY 8f rank 4 (feasible half): LCS(...)(isglobal=True, interp_to_common_grid=True, truncation=None) through the
    unmodified reference (LCS.py:105-114: interp(linear) + reindex(nearest) fill to the 360 x 721 grid, cyclic
    boundary); fixture on a stride-5 subgrid."""
    g = load('seams')
    u, v, lat, lon = S.ideal_vortex(**S.vortex_config_subtropical)
    U, lats, lons = O.regrid_to_common_grid(u[:4], lat, lon)
    V, _, _ = O.regrid_to_common_grid(v[:4], lat, lon)
    assert np.array_equal(lats, g['regrid_lat']) and np.array_equal(lons, g['regrid_lon'])
    assert np.array_equal(U[:, ::5, ::5], g['regrid_u'])
    sig, xd, yd = O.lcs_field(U, V, lats, lons, -21600, SETTLS_order=2, cyclic_xboundary=True, return_dpts=True)
    assert np.array_equal(sig[::5, ::5], g['regrid_sigma'][0], equal_nan=True)
    assert np.array_equal(xd[::5, ::5], g['regrid_x_dep']) and np.array_equal(yd[::5, ::5], g['regrid_y_dep'])
