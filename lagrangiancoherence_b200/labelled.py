"""Minimal labelled arrays for the reference-shaped API.

The reference's public callables take and return ``xarray.DataArray`` (LCS.py:48-51,
trajectory.py:8-18).  xarray is not installed in this image, so the API layer is duck-typed:
it accepts anything with ``.dims``, ``.values`` and coordinate lookup by name -- a real
``xarray.DataArray`` when xarray is importable, or the small :class:`DataArray` below -- and
answers with the same kind of object it was given.  Only what the hot path's callers use is
implemented (``dims/coords/values/shape/size/isel/sortby/transpose/copy/expand_dims/
assign_coords`` and elementwise numpy arithmetic); this is not an xarray replacement.
"""
from __future__ import annotations

import numpy as np

try:                                    # pragma: no cover - xarray is absent in the build image
    import xarray as _xr
except Exception:                       # noqa: BLE001
    _xr = None


class DataArray:
    """N-d array with named dimensions and 1-D coordinates (plus scalar coordinates)."""

    __array_priority__ = 50

    def __init__(self, data, dims, coords=None, name=None):
        self.values = np.asarray(data)
        self.dims = tuple(dims)
        if len(self.dims) != self.values.ndim:
            raise ValueError(f'{len(self.dims)} dims for a {self.values.ndim}-d array')
        self.coords = {}
        for k, v in (coords or {}).items():
            v = v.values if isinstance(v, DataArray) else np.asarray(v)
            if k in self.dims and v.shape != (self.values.shape[self.dims.index(k)],):
                raise ValueError(f'coordinate {k!r} has shape {v.shape}')
            self.coords[k] = v
        self.name = name

    # ---- basic protocol
    @property
    def shape(self):
        return self.values.shape

    @property
    def size(self):
        return self.values.size

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def ndim(self):
        return self.values.ndim

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __len__(self):
        return len(self.values)

    def __repr__(self):
        cs = ', '.join(f'{k}[{np.size(v)}]' for k, v in self.coords.items())
        return f'<lcs_b200.DataArray {self.name or ""} dims={self.dims} shape={self.shape} coords=({cs})>'

    def __getattr__(self, item):                     # da.latitude -> coordinate as a DataArray
        coords = self.__dict__.get('coords', {})
        if item in coords:
            return self[item]
        raise AttributeError(item)

    def __getitem__(self, key):
        if isinstance(key, str):
            v = self.coords[key]
            return DataArray(v, (key,) if v.ndim == 1 else (), {key: v} if v.ndim == 1 else {}, name=key)
        return DataArray(self.values[key], self.dims) if np.ndim(self.values[key]) == self.values.ndim \
            else self.values[key]

    # ---- structure
    def copy(self, deep=True, data=None):
        vals = np.array(self.values, copy=True) if data is None else np.asarray(data)
        if data is not None and vals.shape != self.values.shape:
            raise ValueError('copy(data=...) must keep the shape')
        return DataArray(vals, self.dims, {k: np.array(v, copy=True) for k, v in self.coords.items()}, self.name)

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        if Ellipsis in dims:
            i = dims.index(Ellipsis)
            rest = [d for d in self.dims if d not in dims]
            dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
        order = [self.dims.index(d) for d in dims]
        return DataArray(self.values.transpose(order), dims, self.coords, self.name)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        vals, dims, coords = self.values, list(self.dims), dict(self.coords)
        for d, idx in indexers.items():
            ax = dims.index(d)
            vals = np.take(vals, idx, axis=ax) if not isinstance(idx, slice) else vals[(slice(None),) * ax + (idx,)]
            if d in coords:
                coords[d] = coords[d][idx]
            if np.ndim(idx) == 0 and not isinstance(idx, slice):
                dims.pop(ax)
        return DataArray(vals, dims, coords, self.name)

    def sortby(self, dim, ascending=True):
        c = self.coords[dim]
        d = np.diff(c)
        if ascending and (d > 0).all():                 # already sorted: no copy (13 MB per wind array at C2)
            return self
        if ascending and (d < 0).all():                 # ERA5-style descending latitude: a reversed view
            return self.isel({dim: slice(None, None, -1)})
        order = np.argsort(c, kind='stable')
        if not ascending:
            order = order[::-1]
        return self.isel({dim: order})

    def expand_dims(self, dim):
        coords = dict(self.coords)
        if dim in coords and np.ndim(coords[dim]) == 0:
            coords[dim] = np.asarray(coords[dim])[None]
        return DataArray(self.values[None], (dim,) + self.dims, coords, self.name)

    def assign_coords(self, coords=None, **kw):
        new = dict(self.coords)
        new.update(dict(coords or {}, **kw))
        return DataArray(self.values, self.dims, new, self.name)

    def drop(self, name):
        return DataArray(self.values, self.dims, {k: v for k, v in self.coords.items() if k != name}, self.name)

    def where(self, cond, other=np.nan):
        return self.copy(data=np.where(np.asarray(cond), self.values, other))

    def to_netcdf(self, path):                      # LCS.py:250-262 (the command line saves its results this way)
        from .ncio import to_netcdf
        return to_netcdf(self, path)

    # ---- elementwise arithmetic (enough for callers' ``np.log(out) / 2``)
    def _wrap(self, vals):
        return DataArray(vals, self.dims, self.coords, self.name) if np.shape(vals) == self.shape else vals

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__':
            return NotImplemented
        args = [x.values if isinstance(x, DataArray) else x for x in inputs]
        return self._wrap(getattr(ufunc, method)(*args, **kwargs))

    def _bin(self, other, op):
        o = other.values if isinstance(other, DataArray) else other
        return self._wrap(op(self.values, o))

    __add__ = lambda s, o: s._bin(o, np.add)
    __radd__ = lambda s, o: s._bin(o, lambda a, b: np.add(b, a))
    __sub__ = lambda s, o: s._bin(o, np.subtract)
    __rsub__ = lambda s, o: s._bin(o, lambda a, b: np.subtract(b, a))
    __mul__ = lambda s, o: s._bin(o, np.multiply)
    __rmul__ = lambda s, o: s._bin(o, lambda a, b: np.multiply(b, a))
    __truediv__ = lambda s, o: s._bin(o, np.true_divide)
    __rtruediv__ = lambda s, o: s._bin(o, lambda a, b: np.true_divide(b, a))
    __neg__ = lambda s: s._wrap(-s.values)
    __lt__ = lambda s, o: s._bin(o, np.less)
    __gt__ = lambda s, o: s._bin(o, np.greater)
    __le__ = lambda s, o: s._bin(o, np.less_equal)
    __ge__ = lambda s, o: s._bin(o, np.greater_equal)


class DeviceArray(DataArray):
    """A DataArray whose values live on the GPU until somebody asks for them.  The global path of ``LCS.__call__`` regrids
    and truncates the winds on the device (LCS.py:105-118) and hands them straight to the integrator; the host copy that
    ``.values`` promises is made on first access only (a 360 x 721 series is 19 MB per component: copying every
    intermediate down and up again was 25 of the 34 ms of a default global call)."""

    def __init__(self, tensor, dims, coords=None, name=None):
        self._device_values = tensor
        self._host_values = None
        self.dims = tuple(dims)
        if len(self.dims) != tensor.dim():
            raise ValueError(f'{len(self.dims)} dims for a {tensor.dim()}-d array')
        self.coords = {}
        for k, v in (coords or {}).items():
            v = v.values if isinstance(v, DataArray) else np.asarray(v)
            if k in self.dims and v.shape != (tensor.shape[self.dims.index(k)],):
                raise ValueError(f'coordinate {k!r} has shape {v.shape}')
            self.coords[k] = v
        self.name = name

    @property
    def values(self):
        if self._host_values is None:
            self._host_values = self._device_values.cpu().numpy()
        return self._host_values

    @property
    def shape(self):
        return tuple(self._device_values.shape)

    @property
    def dtype(self):
        return np.dtype(str(self._device_values.dtype).replace('torch.', ''))

    @property
    def size(self):
        return self._device_values.numel()

    @property
    def ndim(self):
        return self._device_values.dim()

    def transpose(self, *dims):
        if tuple(dims) == self.dims:                    # already in that order: stay on the device
            return self
        return DataArray(self.values, self.dims, self.coords, self.name).transpose(*dims)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        if len(indexers) == 1 and self._host_values is None:
            (d, idx), = indexers.items()
            if d == self.dims[0] and np.ndim(idx) == 0 and not isinstance(idx, slice):     # one level: download that level only
                coords = dict(self.coords)
                if d in coords:
                    coords[d] = coords[d][idx]
                return DataArray(self._device_values[int(idx)].cpu().numpy(), self.dims[1:], coords, self.name)
        return DataArray(self.values, self.dims, self.coords, self.name).isel(indexers)

    def copy(self, deep=True, data=None):
        return DataArray(self.values, self.dims, self.coords, self.name).copy(deep, data)


class Dataset:
    """Bag of named DataArrays: ``ds.u`` / ``ds.v`` as LCS.__call__ expects (LCS.py:81-83)."""

    def __init__(self, data_vars):
        self.data_vars = dict(data_vars)

    def __getattr__(self, item):
        dv = self.__dict__.get('data_vars', {})
        if item in dv:
            return dv[item]
        raise AttributeError(item)

    def __getitem__(self, item):
        return self.data_vars[item]

    def copy(self):
        return Dataset({k: v.copy() for k, v in self.data_vars.items()})


# ------------------------------------------------------------------ duck-typing helpers
def is_xarray(obj):
    return _xr is not None and isinstance(obj, (_xr.DataArray, _xr.Dataset))


def is_dataset(obj):
    return isinstance(obj, Dataset) or (_xr is not None and isinstance(obj, _xr.Dataset))


def coord_values(da, name):
    c = da[name]
    return np.asarray(c.values if hasattr(c, 'values') else c)


def make_like(template, data, dims, coords, name=None):
    """Build an output of the same family as ``template`` (xarray in -> xarray out)."""
    if is_xarray(template):             # pragma: no cover - needs xarray
        return _xr.DataArray(data, dims=dims, coords=coords, name=name)
    return DataArray(data, dims, coords, name)
