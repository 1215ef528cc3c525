"""Locate the first trajectories that leave the oracle in the f32-wind (round32) mode."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine

lat = np.linspace(-30.0, 10.0, 41)
lon = np.linspace(-80.0, -24.0, 57)
u, v = S.era5_like_winds(lat, lon, 4)
u32, v32 = u.astype(np.float32), v.astype(np.float32)
for order, xmode, Sord in ((1, 'pointwise', 0), (1, 'pointwise', 1), (1, 'outer', 3), (3, 'outer', 3)):
    rx, ry = O.parcel_propagation(u32, v32, lat, lon, -3600, SETTLS_order=Sord, interp_order=order, xclamp=xmode, return_traj=True)
    eng = FtleEngine(lat, lon, -3600, SETTLS_order=Sord, interp_order=order, xmode=xmode)
    st = eng.stage(u32, v32)
    print('round32 =', st.round32)
    x, y, xt, yt = eng.advect(st, return_traj=True)
    xt, yt = xt[0].cpu().numpy(), yt[0].cpu().numpy()
    for lev in range(xt.shape[0]):
        ex = np.abs(xt[lev] - rx[lev]) / 80.0
        ey = np.abs(yt[lev] - ry[lev]) / 30.0
        bad = (ex > 1e-10) | (ey > 1e-10)
        print(f'order {order} {xmode} S{Sord} level {lev}: {bad.sum()} bad, max ex {ex.max():.2e} ey {ey.max():.2e}')
        if bad.any() and lev <= 1:
            rr, cc = np.nonzero(bad)
            for r, c in list(zip(rr, cc))[:6]:
                print('   row', r, 'col', c, 'gpu', repr(xt[lev, r, c]), repr(yt[lev, r, c]), 'ref', repr(rx[lev, r, c]), repr(ry[lev, r, c]),
                      'prev', repr(rx[lev - 1, r, c]), repr(ry[lev - 1, r, c]))
