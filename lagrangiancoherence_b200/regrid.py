"""The fixed 360 x 721 "common grid" of the reference's global path (LCS.py:105-114): host-side planning.

Upstream: ``u.interp(latitude=lats, longitude=lons, method='linear')`` with the NaNs (targets outside the source
coordinates) filled from ``u.reindex(..., method='nearest')``.  xarray decomposes the orthogonal linear interpolation
into successive 1-D ``scipy.interpolate.interp1d`` calls in the order the indexers are given (latitude, then
longitude); scipy 1.18.1 evaluates ``w_hi * y_hi + w_lo * y_lo`` with the bracket from ``searchsorted(x, x_new)``
clipped to ``[1, n-1]``; the nearest labels come from ``pandas.Index.get_indexer(target, method='nearest')``.
The per-axis tables built here feed the device kernel ``lcs_regrid_linear_nearest``."""
from __future__ import annotations

import numpy as np
import pandas as pd


def common_grid():
    """LCS.py:106-107 (note the 721 longitudes from -180 to 179.5: spacing 359.5/720 degrees, as written upstream)."""
    return np.linspace(-89.75, 89.75, 180 * 2), np.linspace(-180, 179.5, 360 * 2 + 1)


def axis_plan(src, dst):
    """``(lo int32, w_hi, w_lo, valid uint8, nearest int32)`` for one axis: new value = w_hi*y[lo+1] + w_lo*y[lo]
    where ``valid`` (inside the source range; NaN otherwise, interp1d's fill_value), else y[nearest]."""
    x = np.asarray(src, dtype=np.float64)
    xn = np.asarray(dst, dtype=np.float64)
    if x.size < 2 or np.any(np.diff(x) <= 0):
        raise ValueError('source coordinates must be ascending with at least two points')
    hi = np.clip(np.searchsorted(x, xn, side='left'), 1, x.size - 1)
    lo = hi - 1
    w_hi = (xn - x[lo]) / (x[hi] - x[lo])
    w_lo = (x[hi] - xn) / (x[hi] - x[lo])
    valid = (xn >= x[0]) & (xn <= x[-1])
    nearest = pd.Index(x).get_indexer(xn, method='nearest')
    return lo.astype(np.int32), w_hi, w_lo, valid.astype(np.uint8), nearest.astype(np.int32)
