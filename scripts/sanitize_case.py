"""Small cases through every kernel family, for `compute-sanitizer --tool memcheck python scripts/sanitize_case.py`
(one tool per gpurun call): prefilter (both forms, orders 2..5), packing, the fused / group-persistent (state in global
memory and in registers) / phased integrators with trajectories and row bands, epilogue with mask and Jacobian, filters,
regrid, spectral truncation, ridge classification.  Ragged sizes on purpose."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S, engine as E
from lagrangiancoherence_b200.engine import FtleEngine, precision_args

dev = 'cuda:0'
lat = np.linspace(-30.0, 10.0, 37)
lon = np.linspace(-80.0, -24.0, 53)
u, v = S.era5_like_winds(lat, lon, 7)
for form in ('1', '2'):
    os.environ['LCS_PREFILTER_FORM'] = form
    for order in (2, 3, 4, 5):
        E.prefilter_device(u, v, dev, order=order)
os.environ['LCS_PREFILTER_FORM'] = '0'
for prec in ('f64', 'f32', 'f32fast'):
    for xmode in ('pointwise', 'outer'):
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev, **precision_args(prec))
        st = eng.stage(u, v, reuse=True)
        for state, groups, mode in (('-1', '0', '0'), ('0', '2', '0'), ('0', '0', '1')):
            os.environ.update(LCS_OUTER_STATE=state, LCS_OUTER_GROUPS=groups, LCS_OUTER_MODE=mode)
            eng._ws = None
            x, y, xt, yt = eng.advect(st, nsteps=3, nwindows=3, return_traj=True)
            xb, yb = eng.advect(st, nsteps=3, nwindows=2, rows=(5, 21))
        sig, jac = eng.epilogue(x, y, return_jac=True, mask=np.ones((lat.size, lon.size), bool))
        eng.epilogue(xb, yb, in_row0=5, out_rows=(7, 19))
        eng.gaussian(x, 1.5)
        eng.check_finite()
u32, v32 = u.astype(np.float32), v.astype(np.float32)
for order in (1, 3):
    eng = FtleEngine(lat, lon, -21600, SETTLS_order=2, interp_order=order, xmode='outer', device=dev)
    eng.advect(eng.stage(u32, v32), nwindows=2, nsteps=3)                     # round32 kernels
eng = FtleEngine(lat, lon, -21600, SETTLS_order=1, xmode='pointwise', strict=True, device=dev)
eng.advect(eng.stage(u, v))                                                     # PAIR4 / strict
glat = np.arange(-88.0, 89.0, 8.0); glon = np.arange(-180.0, 180.0, 8.0)
gu, _ = S.era5_like_winds(glat, glon, 2)
from lagrangiancoherence_b200.regrid import common_grid
E.regrid_device(gu, glat, glon, *common_grid(), device=dev)
E.spectral_truncate_device(np.random.default_rng(0).normal(size=(2, 19, 36)), 5, device=dev)
E.map_coordinates_device(u[0], *np.meshgrid(lon, lat), lat, lon, order=3, device=dev)
E.fourth_order_derivative_device(u[0].astype(np.float32), dim=1, device=dev)
E.spectral_norm_3x3_device(np.random.default_rng(1).normal(size=(3, 3, 100)), device=dev)
from lagrangiancoherence_b200 import DataArray
from lagrangiancoherence_b200.LCS.tools import find_ridges_spherical_hessian
find_ridges_spherical_hessian(DataArray(np.abs(u[0]), ('latitude', 'longitude'), {'latitude': lat, 'longitude': lon}))
torch.cuda.synchronize()
print('sanitize_case: all kernels ran')
