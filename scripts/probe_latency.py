"""Scratch: wall-clock latency of one drop-in LCS call on the C2 grid (host numpy in, host numpy out)."""
import os, sys, time, io, contextlib
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import DataArray, synthetic as S
from lagrangiancoherence_b200.LCS.LCS import LCS

lat, lon = S.grid_c2()
u, v = S.era5_like_winds(lat, lon, 9)
t = (np.datetime64('2000-01-01T00') + np.arange(9) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
coords = {'time': t, 'latitude': lat, 'longitude': lon}
du, dv = DataArray(u, ('time', 'latitude', 'longitude'), coords), DataArray(v, ('time', 'latitude', 'longitude'), coords)
for clamp in ('outer', 'pointwise'):
    lcs = LCS(timestep=-21600, SETTLS_order=4)
    ts = []
    for i in range(12):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            out = lcs(u=du, v=dv, verbose=False, xclamp=clamp)
        ts.append(time.perf_counter() - t0)
    print(clamp, 'LCS.__call__ wall ms: first %.1f, median of rest %.2f, min %.2f' % (ts[0] * 1e3, np.median(ts[2:]) * 1e3, np.min(ts[2:]) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(5): lcs(u=du, v=dv, verbose=False)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
