"""Stand-in for the slice of xarray that the reference's hot path touches -- TEST INFRASTRUCTURE.

Purpose: execute the UNMODIFIED reference sources (/root/reference/LCS/{LCS,trajectory,tools}.py) in
the build container, where xarray is not installed, to generate golden vectors for the oracle
(oracle/make_golden.py).  Only the operations those files perform are implemented, with xarray's
documented semantics:

* arithmetic / comparisons / numpy ufuncs broadcast BY DIMENSION NAME; a bare ndarray operand
  broadcasts positionally (numpy rules) against the labelled operand's data;
* ``where(cond, other)`` keeps values where cond holds;
* ``da[key] = value`` with a tuple of unlabelled 1-D integer arrays is ORTHOGONAL (outer) indexing
  (xarray: "unlabelled array indexers are orthogonal") -- the behaviour behind quirk Q6;
* ``sortby``, ``isel`` (ints, slices, integer arrays), ``transpose``, ``copy(data=)``, ``concat`` along an
  existing or a new (pandas.Index-named) dimension, ``broadcast``, ``apply_ufunc`` (plain call on .values),
  ``merge``/``to_array``/``rename``, ``stack``/``dropna``/``unstack`` over two dims, ``expand_dims``,
  ``assign_coords``, ``drop``.

This is this repository's reading of xarray, not xarray: the resulting pin is "reference code +
emulated xarray" (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

__version__ = '0.0-refshim'


def _is_da(x):
    return isinstance(x, DataArray)


class DataArray:
    __array_priority__ = 60

    def __init__(self, data=None, coords=None, dims=None, name=None, attrs=None):
        self.values = np.asarray(data)
        if dims is None:
            dims = tuple(coords.keys()) if isinstance(coords, dict) and len(coords) == self.values.ndim else ()
        self.dims = tuple(dims)
        assert len(self.dims) == self.values.ndim, (self.dims, self.values.shape)
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = np.asarray(v.values if _is_da(v) else v)
        self.name = name
        self.attrs = dict(attrs or {})
        self._stack = None          # (dim, (a, b), full_coord_a, full_coord_b)

    # ------------------------------------------------------------- basics
    @property
    def shape(self):
        return self.values.shape

    @property
    def size(self):
        return self.values.size

    @property
    def ndim(self):
        return self.values.ndim

    @property
    def dtype(self):
        return self.values.dtype

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __len__(self):
        return self.values.shape[0]

    def __int__(self):
        return int(self.values)

    def __float__(self):
        return float(self.values)

    def __repr__(self):
        return f'<refshim.DataArray {self.name} dims={self.dims} shape={self.shape}>'

    def __getattr__(self, item):
        coords = self.__dict__.get('coords', {})
        if item in coords:
            return self[item]
        stack = self.__dict__.get('_stack')
        if stack is not None and item == stack[0]:        # hessian.points: the stacked index (tools.py:96)
            return self
        raise AttributeError(item)

    def sel(self, **kw):
        (dim, index), = kw.items()
        assert self._stack is not None and dim == self._stack[0], 'refshim: sel is only implemented on a stacked dim'
        a, b = self._stack[1]
        mine = list(zip(self.coords['_' + a].tolist(), self.coords['_' + b].tolist()))
        want = list(zip(index.coords['_' + a].tolist(), index.coords['_' + b].tolist()))
        if mine == want:
            return self
        pos = {k: i for i, k in enumerate(mine)}
        idx = np.array([pos[k] for k in want])
        ax = self.dims.index(dim)
        coords = {k: (v[idx] if k.startswith('_') else v) for k, v in self.coords.items()}
        out = DataArray(np.take(self.values, idx, axis=ax), coords, self.dims, self.name)
        out._stack = self._stack
        return out

    # ------------------------------------------------------------- reindex / interp (LCS.py:108-111, the global regrid)
    def reindex(self, indexers=None, method=None, **kw):
        """xarray: label lookup through pandas ``Index.get_indexer(target, method=...)``; with method='nearest' and no
        tolerance every new label finds a source label (ties go to the larger one on an increasing index)."""
        assert method == 'nearest', 'refshim: reindex is only implemented for method="nearest"'
        out = self
        for dim, target in dict(indexers or {}, **kw).items():
            idx = pd.Index(out.coords[dim]).get_indexer(np.asarray(target), method='nearest')
            out = out.isel({dim: idx})
            out.coords[dim] = np.asarray(target)
        return out

    def interp(self, coords=None, method='linear', assume_sorted=False, **kw):
        """xarray: orthogonal linear interpolation over several dimensions is decomposed into successive 1-D
        interpolations in the order the indexers are given (xarray.core.missing.decompose_interp), each one
        ``scipy.interpolate.interp1d(x, y, kind='linear', axis=..., bounds_error=False, fill_value=nan)``."""
        from scipy.interpolate import interp1d
        assert method == 'linear', 'refshim: interp is only implemented for method="linear"'
        out = self
        for dim, target in dict(coords or {}, **kw).items():
            if not assume_sorted:
                out = out.sortby(dim)
            ax = out.dims.index(dim)
            x = np.asarray(out.coords[dim], dtype=np.float64)
            f = interp1d(x, out.values, kind='linear', axis=ax, bounds_error=False, fill_value=np.nan, assume_sorted=True)
            newc = dict(out.coords)
            newc[dim] = np.asarray(target)
            out = DataArray(f(np.asarray(target, dtype=np.float64)), newc, out.dims, out.name, out.attrs)
        return out

    def _new(self, values, dims=None, coords=None):
        dims = self.dims if dims is None else tuple(dims)
        coords = self.coords if coords is None else coords
        keep = {k: v for k, v in coords.items() if v.ndim == 0 or k in dims or k.startswith('_')}
        out = DataArray(values, keep, dims, self.name, self.attrs)
        out._stack = self._stack
        return out

    # ------------------------------------------------------------- indexing
    def __getitem__(self, key):
        if isinstance(key, str):
            v = self.coords[key]
            if v.ndim == 0:
                return DataArray(v, {key: v}, (), key)
            return DataArray(v, {key: v}, (key,), key)
        raise NotImplementedError('positional __getitem__ is not used by the reference hot path')

    def __setitem__(self, key, value):
        if isinstance(key, str):                       # da['time'] = timestamp  (LCS.py:159)
            self.coords[key] = np.asarray(value.values if _is_da(value) else value)
            return
        if not isinstance(key, tuple):
            key = (key,)
        value = value.values if _is_da(value) else value
        if all(isinstance(k, np.ndarray) and k.ndim == 1 for k in key) and len(key) == self.ndim:
            if any(k.size == 0 for k in key):
                return
            self.values[np.ix_(*key)] = value          # ORTHOGONAL assignment (trajectory.py:96-97)
            return
        raise NotImplementedError(f'__setitem__ with key {key!r}')

    def isel(self, indexers=None, drop=False, **kw):
        indexers = dict(indexers or {}, **kw)
        vals, dims, coords = self.values, list(self.dims), dict(self.coords)
        for d, idx in indexers.items():
            ax = dims.index(d)
            if isinstance(idx, slice):
                vals = vals[(slice(None),) * ax + (idx,)]
                if d in coords:
                    coords[d] = coords[d][idx]
            elif np.ndim(idx) == 0:
                vals = np.take(vals, int(idx), axis=ax)
                if d in coords:
                    coords[d] = np.asarray(coords[d][int(idx)])
                dims.pop(ax)
            else:
                idx = np.asarray(idx)
                vals = np.take(vals, idx, axis=ax)
                if d in coords:
                    coords[d] = coords[d][idx]
        out = DataArray(vals, coords, dims, self.name, self.attrs)
        out._stack = self._stack
        return out

    def sortby(self, name, ascending=True):
        order = np.argsort(self.coords[name], kind='stable')
        if not ascending:
            order = order[::-1]
        return self.isel({name: order})

    def transpose(self, *dims):
        if not dims:
            dims = self.dims[::-1]
        dims = list(dims)
        if Ellipsis in dims:
            i = dims.index(Ellipsis)
            rest = [d for d in self.dims if d not in dims]
            dims = dims[:i] + rest + dims[i + 1:]
        out = self._new(self.values.transpose([self.dims.index(d) for d in dims]), dims)
        return out

    @property
    def T(self):
        return self.transpose()

    def copy(self, deep=True, data=None):
        if data is None:
            vals = np.array(self.values, copy=True)
        else:
            vals = np.asarray(data.values if _is_da(data) else data)
            assert vals.shape == self.shape, f'copy(data=) shape {vals.shape} != {self.shape}'
        out = DataArray(vals, {k: np.array(v, copy=True) for k, v in self.coords.items()}, self.dims, self.name, self.attrs)
        out._stack = self._stack
        return out

    def drop(self, name, **_):
        names = [name] if isinstance(name, str) else list(name)
        for n in names:
            if n not in self.coords:
                raise ValueError(f'{n!r} is not a coordinate')     # what fix_time_coord catches, trajectory.py:133-136
        out = DataArray(self.values, {k: v for k, v in self.coords.items() if k not in names}, self.dims, self.name,
                        self.attrs)
        out._stack = self._stack
        return out

    drop_vars = drop

    def assign_coords(self, coords=None, **kw):
        new = dict(self.coords)
        for k, v in dict(coords or {}, **kw).items():
            new[k] = np.asarray(v.values if _is_da(v) else v)
        out = DataArray(self.values, new, self.dims, self.name, self.attrs)
        out._stack = self._stack
        return out

    def expand_dims(self, dim):
        coords = dict(self.coords)
        if dim in coords and coords[dim].ndim == 0:
            coords[dim] = coords[dim][None]
        return DataArray(self.values[None], coords, (dim,) + self.dims, self.name, self.attrs)

    def rename(self, mapping=None, **kw):
        if isinstance(mapping, str):
            out = self.copy()
            out.name = mapping
            return out
        m = dict(mapping or {}, **kw)
        dims = tuple(m.get(d, d) for d in self.dims)
        coords = {m.get(k, k): v for k, v in self.coords.items()}
        out = DataArray(self.values, coords, dims, self.name, self.attrs)
        if self._stack is not None:                    # renaming another dimension keeps the stacked index (tools.py:123)
            new, (a, b) = self._stack
            if not {new, a, b} & set(m):
                out._stack = self._stack
        return out

    # ------------------------------------------------------------- resample(...).interpolate('linear')  (LCS.py:89-90)
    def resample(self, indexer=None, **kw):
        (dim, freq), = dict(indexer or {}, **kw).items()
        return _Resampler(self, dim, freq)

    # ------------------------------------------------------------- reductions
    def _reduce(self, fn):
        return DataArray(fn(self.values), {k: v for k, v in self.coords.items() if v.ndim == 0}, (), self.name)

    def min(self):
        return self._reduce(np.min)

    def max(self):
        return self._reduce(np.max)

    def std(self):
        return self._reduce(np.std)

    def mean(self):
        return self._reduce(np.mean)

    # ------------------------------------------------------------- arithmetic by dimension name
    def _binary(self, other, op, reflected=False):
        if _is_da(other):
            dims = list(self.dims) + [d for d in other.dims if d not in self.dims]

            def expand(a):
                v = a.values.transpose([a.dims.index(d) for d in dims if d in a.dims])
                shape = [a.values.shape[a.dims.index(d)] if d in a.dims else 1 for d in dims]
                return v.reshape(shape)
            x, y = expand(self), expand(other)
            coords = dict(other.coords)
            coords.update(self.coords)
        else:
            dims, x, y, coords = list(self.dims), self.values, other, self.coords
        res = op(y, x) if reflected else op(x, y)
        out = DataArray(res, {k: v for k, v in coords.items() if v.ndim == 0 or k in dims or k.startswith('_')}, dims, self.name)
        out._stack = self._stack if self._stack is not None else (other._stack if _is_da(other) else None)
        return out

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != '__call__' or kwargs.get('out') is not None:
            return NotImplemented
        if len(inputs) == 1:
            return self._new(ufunc(self.values, **kwargs))
        a, b = inputs
        if _is_da(a):
            return a._binary(b, lambda p, q: ufunc(p, q, **kwargs))
        return b._binary(a, lambda p, q: ufunc(p, q, **kwargs), reflected=True)

    __add__ = lambda s, o: s._binary(o, np.add)
    __radd__ = lambda s, o: s._binary(o, np.add, True)
    __sub__ = lambda s, o: s._binary(o, np.subtract)
    __rsub__ = lambda s, o: s._binary(o, np.subtract, True)
    __mul__ = lambda s, o: s._binary(o, np.multiply)
    __rmul__ = lambda s, o: s._binary(o, np.multiply, True)
    __truediv__ = lambda s, o: s._binary(o, np.true_divide)
    __rtruediv__ = lambda s, o: s._binary(o, np.true_divide, True)
    __mod__ = lambda s, o: s._binary(o, np.mod)
    __pow__ = lambda s, o: s._binary(o, np.power)
    __eq__ = lambda s, o: s._binary(o, np.equal)
    __ne__ = lambda s, o: s._binary(o, np.not_equal)
    __hash__ = None
    __lt__ = lambda s, o: s._binary(o, np.less)
    __gt__ = lambda s, o: s._binary(o, np.greater)
    __le__ = lambda s, o: s._binary(o, np.less_equal)
    __ge__ = lambda s, o: s._binary(o, np.greater_equal)
    __neg__ = lambda s: s._new(-s.values)
    __invert__ = lambda s: s._new(~s.values)

    def where(self, cond, other=np.nan, drop=False):
        assert not drop, 'where(drop=True) is only used by latlonsel, which the shim implements directly'
        objs = [self, cond if _is_da(cond) else DataArray(np.asarray(cond), {}, self.dims)]
        if _is_da(other):
            objs.append(other)
        b = broadcast(*objs)
        oth = b[2].values if _is_da(other) else other
        return b[0]._new(np.where(b[1].values, b[0].values, oth))

    # ------------------------------------------------------------- stack / dropna / unstack (two dims)
    def stack(self, mapping=None, **kw):
        (new, (a, b)), = dict(mapping or {}, **kw).items()
        rest = [d for d in self.dims if d not in (a, b)]
        v = self.transpose(*rest, a, b).values
        na, nb = v.shape[-2], v.shape[-1]
        A, B = np.meshgrid(self.coords[a], self.coords[b], indexing='ij')
        coords = {k: val for k, val in self.coords.items() if k not in (a, b)}
        coords['_' + a], coords['_' + b] = A.ravel(), B.ravel()
        out = DataArray(v.reshape(v.shape[:-2] + (na * nb,)), coords, rest + [new], self.name)
        out._stack = (new, (a, b))
        return out

    def dropna(self, dim, how='any'):
        ax = self.dims.index(dim)
        other = tuple(i for i in range(self.ndim) if i != ax)
        bad = np.isnan(self.values).any(axis=other) if how == 'any' else np.isnan(self.values).all(axis=other)
        keep = np.flatnonzero(~bad)
        coords = {}
        for k, v in self.coords.items():
            coords[k] = v[keep] if (v.ndim == 1 and v.shape[0] == self.shape[ax] and k.startswith('_')) else v
        out = DataArray(np.take(self.values, keep, axis=ax), coords, self.dims, self.name)
        out._stack = self._stack
        return out

    def unstack(self, dim=None):
        if self._stack is None or self._stack[0] not in self.dims:
            return self                                   # nothing stacked (tools.py:147 unstacks twice)
        new, (a, b) = self._stack
        ca, cb = self.coords['_' + a], self.coords['_' + b]
        ua, ub = np.unique(ca), np.unique(cb)            # unused index levels are dropped, as xarray does
        ax = self.dims.index(new)
        lead = [d for d in self.dims if d != new]
        v = np.moveaxis(self.values, ax, -1)
        out = np.full(v.shape[:-1] + (ua.size, ub.size), np.nan)
        out[..., np.searchsorted(ua, ca), np.searchsorted(ub, cb)] = v
        coords = {k: val for k, val in self.coords.items() if not k.startswith('_')}
        coords[a], coords[b] = ua, ub
        return DataArray(out, coords, lead + [a, b], self.name)      # _stack is None again


class _Resampler:
    """xarray: DataArrayResample.interpolate(kind) == obj.interp({dim: new pandas bin labels}, method=kind,
    kwargs={'bounds_error': False}); for one dimension that is scipy.interpolate.interp1d over the time axis
    converted to float64 nanosecond offsets from its minimum (xarray.core.missing._floatize_x)."""

    def __init__(self, obj, dim, freq):
        self.obj, self.dim, self.freq = obj, dim, freq

    def interpolate(self, kind='linear'):
        import pandas as pd
        from scipy.interpolate import interp1d
        t = np.asarray(self.obj.coords[self.dim]).astype('datetime64[ns]')
        new = pd.Series(0.0, index=pd.DatetimeIndex(t)).resample(self.freq).asfreq().index.values.astype('datetime64[ns]')
        x = (t - t.min()).astype('int64').astype(np.float64)
        xn = (new - t.min()).astype('int64').astype(np.float64)
        ax = self.obj.dims.index(self.dim)
        f = interp1d(x, self.obj.values, kind=kind, axis=ax, bounds_error=False, fill_value=np.nan, assume_sorted=True, copy=False)
        coords = dict(self.obj.coords)
        coords[self.dim] = new
        return DataArray(f(xn), coords, self.obj.dims, self.obj.name)


class Dataset:
    def __init__(self, data_vars=None):
        self.data_vars = dict(data_vars or {})

    def __getattr__(self, item):
        dv = self.__dict__.get('data_vars', {})
        if item in dv:
            return dv[item]
        raise AttributeError(item)

    def __getitem__(self, k):
        return self.data_vars[k]

    def copy(self):
        return Dataset({k: v.copy() for k, v in self.data_vars.items()})

    def to_array(self, dim='variable'):
        names = list(self.data_vars)
        first = self.data_vars[names[0]]
        vals = np.stack([self.data_vars[n].transpose(*first.dims).values for n in names])
        coords = dict(first.coords)
        coords[dim] = np.array(names)
        return DataArray(vals, coords, (dim,) + first.dims)


# ----------------------------------------------------------------- module-level functions
def merge(objs):
    return Dataset({o.name: o for o in objs})


def zeros_like(da):
    return da.copy(data=np.zeros_like(da.values))


def apply_ufunc(func, *args, **kw):
    first = next(a for a in args if _is_da(a))
    return first._new(func(*[a.values if _is_da(a) else a for a in args]))


def broadcast(*args):
    dims = []
    for a in args:
        dims += [d for d in a.dims if d not in dims]
    shape = {}
    coords = {}
    for a in args:
        for d, n in zip(a.dims, a.shape):
            shape[d] = n
        coords.update({k: v for k, v in a.coords.items()})
    out = []
    for a in args:
        v = a.values.transpose([a.dims.index(d) for d in dims if d in a.dims])
        v = v.reshape([shape[d] if d in a.dims else 1 for d in dims])
        v = np.broadcast_to(v, [shape[d] for d in dims]).copy()
        o = DataArray(v, {k: c for k, c in coords.items() if c.ndim == 0 or k in dims or k.startswith('_')}, dims, a.name)
        o._stack = next((x._stack for x in args if x._stack is not None), None)
        out.append(o)
    return tuple(out)


def concat(objs, dim):
    import pandas as pd
    objs = list(objs)
    if isinstance(dim, str):                                  # along an existing dimension (tools.py:41)
        first = objs[0]
        ax = first.dims.index(dim)
        vals = np.concatenate([o.transpose(*first.dims).values for o in objs], axis=ax)
        coords = dict(first.coords)
        coords[dim] = np.concatenate([o.coords[dim] for o in objs])
        return DataArray(vals, coords, first.dims, first.name)
    assert isinstance(dim, pd.Index)                          # new dimension (trajectory.py:138-139)
    first = objs[0]
    vals = np.stack([o.transpose(*first.dims).values if _is_da(o) else np.asarray(o) for o in objs])
    coords = {k: v for k, v in first.coords.items() if k in first.dims}
    coords[dim.name] = np.asarray(dim.values)
    return DataArray(vals, coords, (dim.name,) + first.dims, first.name)


def open_dataset(*a, **k):
    raise NotImplementedError('refshim: no NetCDF I/O')


open_dataarray = open_dataset


class _Ufuncs:
    @staticmethod
    def isnan(x):
        return np.isnan(x)


ufuncs = _Ufuncs()
