"""One C2 field through the engine, pointwise and outer clamp, three times each: the command the single-field ncu launch
list (profiles/r02b_launches_single_field.csv) is taken from."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine
lat, lon = S.grid_c2()
u, v = S.era5_like_winds(lat, lon, 9)
du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
for xmode in ('pointwise', 'outer'):
    eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode)
    for _ in range(3):
        sig = eng.ftle(du, dv)
    torch.cuda.synchronize()
    print(xmode, float(sig.sum()))
