import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import synthetic as S, engine as E
lat = np.linspace(-30.0, 10.0, 41); lon = np.linspace(-80.0, -24.0, 57)
u, v = S.era5_like_winds(lat, lon, 2)
rng = np.random.default_rng(3)
px = np.meshgrid(lon, lat)[0] + rng.normal(0, 3.0, (lat.size, lon.size))
py = np.meshgrid(lon, lat)[1] + rng.normal(0, 3.0, (lat.size, lon.size))
ref = O.xr_map_coordinates(u[0], px, py, lat, lon, order=1)
got = E.map_coordinates_device(u[0], px, py, lat, lon, order=1).cpu().numpy()
bad = np.argwhere(got != ref)
print('mismatch', len(bad), 'of', got.size, 'max', np.abs(got-ref).max())
for r, c in bad[:8]:
    iy = O.index_map(py[r, c], lat); ix = O.index_map(px[r, c], lon)
    print(r, c, repr(iy), repr(ix), repr(got[r, c]), repr(ref[r, c]), repr(O.gather_linear_wrap(u[0], iy, ix)))
