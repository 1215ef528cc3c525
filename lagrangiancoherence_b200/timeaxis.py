"""Host bookkeeping for `resample=` (LCS.py:88-91): the new time index pandas/xarray would build and, for
every new level, the bracketing input levels and the numerator/denominator of the linear interpolation."""
from __future__ import annotations

import numpy as np
import pandas as pd


def resample_plan(times, freq):
    """``times``: datetime64 coordinate of the wind series (at least two levels).  Returns
    ``(new_times, lo, w_hi, w_lo)``: new level k = w_hi[k] * u[lo[k]+1] + w_lo[k] * u[lo[k]] -- the form
    scipy 1.18.1's interp1d(kind='linear') evaluates, which xarray's ``resample(...).interpolate('linear')`` calls
    over float64 nanosecond offsets (bracket = searchsorted(..., 'left') clipped to [1, n-1])."""
    t = np.asarray(times).astype('datetime64[ns]')
    if t.size < 2:
        raise ValueError('resample needs at least two time levels')
    new = pd.Series(0.0, index=pd.DatetimeIndex(t)).resample(freq).asfreq().index.values.astype('datetime64[ns]')
    x = (t - t.min()).astype('int64').astype(np.float64)           # xarray's _floatize_x: ns offsets from the minimum
    xn = (new - t.min()).astype('int64').astype(np.float64)
    hi = np.clip(np.searchsorted(x, xn, side='left'), 1, x.size - 1)
    lo = hi - 1
    return new, lo.astype(np.int32), (xn - x[lo]) / (x[hi] - x[lo]), (x[hi] - xn) / (x[hi] - x[lo])
