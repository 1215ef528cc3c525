"""Scratch: staging time (prefilter + pack) of a C2 series, for A/B runs of library variants."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine
lat, lon = S.grid_c2()
u, v = S.era5_like_winds(lat, lon, 1192, noise=0.0)
eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='outer')
du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); st = eng.stage(du, dv); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(os.path.basename(os.environ.get('LCS_B200_LIB', 'main')), 'stage ms (1192 levels):', round(float(np.median(ts[2:])), 3), 'chk', float(st.coef_a.double().sum().item()))
