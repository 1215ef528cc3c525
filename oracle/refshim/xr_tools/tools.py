"""refshim: latlonsel as LCS.py:144 calls it (keyword args = the subdomain dict's keys).  Semantics taken from the
in-tree analogue tools.py:158-187: strict inequalities, rows/columns outside are dropped."""
import numpy as np


def latlonsel(array, latitude=None, longitude=None, latname='latitude', lonname='longitude'):
    def keep(coord, sl):
        m = np.ones(coord.shape, bool)
        if sl is not None and sl.start is not None:
            m &= coord > sl.start
        if sl is not None and sl.stop is not None:
            m &= coord < sl.stop
        return np.flatnonzero(m)
    return array.isel({latname: keep(array.coords[latname], latitude), lonname: keep(array.coords[lonname], longitude)})
