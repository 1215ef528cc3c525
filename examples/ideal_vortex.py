#!/usr/bin/env python
"""The reference's idealised-vortex example (examples/ideal_vortex.py:211-288) on the B200 engine:
departure points backward (S=4) and forward (S=2) on the cyclic 2-degree grid, then the attracting and
repelling FTLE fields on the 360x721 common grid of the global path, `0.5*log(sigma)` applied by the caller exactly
as upstream does -- with upstream's defaults: the regrid to the common grid and the T20 spectral truncation of the winds
(windspharm upstream; lcs_spectral_truncate here, see lagrangiancoherence_b200/spectral.py for what is and is not pinned
about that step).  No plotting: prints summary statistics of what the reference's figures show.

    python examples/ideal_vortex.py            # needs a B200 and the built liblcs_b200.so
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import DataArray, Dataset, synthetic as S          # noqa: E402
from lagrangiancoherence_b200.LCS import LCS, trajectory                         # noqa: E402


def main():
    u, v, lat, lon = S.ideal_vortex(**S.vortex_config_subtropical)
    times = (np.datetime64('2000-01-01T00') + np.arange(u.shape[0]) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
    coords = {'time': times, 'latitude': lat, 'longitude': lon}
    ds = Dataset({'u': DataArray(u, ('time', 'latitude', 'longitude'), coords, name='u'),
                  'v': DataArray(v, ('time', 'latitude', 'longitude'), coords, name='v')})

    t0 = time.perf_counter()
    x_dye, y_dye = trajectory.parcel_propagation(ds.u, ds.v, timestep=-6 * 3600, propdim='time', SETTLS_order=4,
                                                 copy=True, return_traj=True, cyclic_xboundary=True, verbose=False)
    x, y = trajectory.parcel_propagation(ds.u, ds.v, timestep=6 * 3600, propdim='time', SETTLS_order=2,
                                         copy=True, return_traj=True, cyclic_xboundary=True, verbose=False)
    # upstream calls isglobal=True with its defaults (examples/ideal_vortex.py:280-287): the 360x721 regrid followed by
    # a T20 spherical-harmonic truncation of the winds, both on the device here
    rcs = LCS.LCS(timestep=6 * 3600, timedim='time', SETTLS_order=4)
    ftle_r = np.log(rcs(ds.copy(), isglobal=True, verbose=False)) / 2
    acs = LCS.LCS(timestep=-6 * 3600, timedim='time', SETTLS_order=4)
    ftle_a = np.log(acs(ds.copy(), isglobal=True, verbose=False)) / 2
    dt = time.perf_counter() - t0
    lat, lon = ftle_a.coords['latitude'], ftle_a.coords['longitude']           # the common grid now

    origin = y_dye.isel(time=0).values - y_dye.isel(time=-1).values
    print(f'winds {u.shape[1]}x{u.shape[2]}, FTLE grid {len(lat)}x{len(lon)}, {u.shape[0]} levels; four calls in {dt * 1e3:.1f} ms')
    print(f'latitude displacement of the dye (backward, 42 h): {np.nanmin(origin):+.2f} .. {np.nanmax(origin):+.2f} deg')
    for name, f in (('attracting', ftle_a), ('repelling', ftle_r)):
        vals = f.values[np.isfinite(f.values)]
        j, i = np.unravel_index(np.nanargmax(np.where(np.isfinite(f.values[0]), f.values[0], -np.inf)), f.values[0].shape)
        print(f'{name} FTLE 0.5*log(sigma): min {vals.min():.3f}  median {np.median(vals):.3f}  max {vals.max():.3f} '
              f'at lat {lat[j]:.0f} lon {lon[i]:.0f}')


if __name__ == '__main__':
    main()
