"""NetCDF in / out for the path-taking call ``LCS(...)(ds='winds.nc')`` and the command line (LCS.py:84-87, 236-265).

Upstream this is ``xr.open_dataset`` / ``DataArray.to_netcdf``.  With xarray installed those are used as they are.
Without it (this image) classic NetCDF-3 files (CDF-1 / CDF-2, what ``scipy.io.netcdf_file`` reads and writes) are
handled here with the part of xarray's CF decoding that a wind file exercises:

* packed variables: ``scale_factor`` / ``add_offset`` applied after masking ``_FillValue`` / ``missing_value`` with NaN.
  Result dtype as current xarray chooses it (``_choose_float_dtype``): float32 only when the stored data are f32, or
  integers of at most 16 bits, and neither attribute is f64; else float64 (ERA5's int16 + f64 attributes decode to f64).
  This matters: f32 winds take the reference's f32 dtype propagation, f64 winds the f64 path;
* time coordinates: ``units = '<unit> since <date>'`` on a standard / gregorian / proleptic_gregorian calendar decoded to
  ``datetime64[ns]``; other calendars raise (upstream would hand back cftime objects, which LCS.py:91 cannot subtract
  into a timedelta64 either).

NetCDF-4 / HDF5 files need netCDF4 or h5py, neither of which is installed: they are refused with a message that says so.
"""
from __future__ import annotations

import re

import numpy as np

from .labelled import DataArray, Dataset, _xr

_UNITS = {'day': 86400.0, 'days': 86400.0, 'd': 86400.0, 'hour': 3600.0, 'hours': 3600.0, 'hr': 3600.0, 'hrs': 3600.0, 'h': 3600.0,
          'minute': 60.0, 'minutes': 60.0, 'min': 60.0, 'mins': 60.0, 'second': 1.0, 'seconds': 1.0, 'sec': 1.0, 'secs': 1.0, 's': 1.0,
          'millisecond': 1e-3, 'milliseconds': 1e-3, 'ms': 1e-3}
_STANDARD_CALENDARS = ('standard', 'gregorian', 'proleptic_gregorian')


def _attr(var, name, default=None):
    a = getattr(var, name, default)
    if isinstance(a, bytes):
        a = a.decode()
    return a


def decode_time(values, units, calendar='standard'):
    """CF time -> datetime64[ns] (xarray's decode_cf_datetime for the standard calendars)."""
    m = re.match(r'\s*(\w+)\s+since\s+(.+?)\s*$', str(units))
    if m is None:
        raise ValueError(f'cannot decode time units {units!r}')
    unit, ref = m.group(1).lower(), m.group(2)
    if unit not in _UNITS:
        raise ValueError(f'unsupported time unit {unit!r}')
    if str(calendar).lower() not in _STANDARD_CALENDARS:
        raise NotImplementedError(f'calendar {calendar!r}: only {_STANDARD_CALENDARS} decode to datetime64')
    ref = re.sub(r'\s*(UTC|Z|\+00:?00)\s*$', '', ref.strip())
    ref = re.sub(r'^(\d{1,4})-(\d{1,2})-(\d{1,2})', lambda g: f'{int(g.group(1)):04d}-{int(g.group(2)):02d}-{int(g.group(3)):02d}', ref)
    ref = re.sub(r'[ T](\d{1,2}):(\d{1,2})(?::(\d{1,2}(?:\.\d+)?))?$',
                 lambda g: f'T{int(g.group(1)):02d}:{int(g.group(2)):02d}:' + (f'{float(g.group(3)):09.6f}' if g.group(3) else '00'), ref)
    t0 = np.datetime64(ref, 'ns')
    v = np.asarray(values)
    if np.issubdtype(v.dtype, np.integer):
        ns = v.astype(np.int64) * np.int64(round(_UNITS[unit] * 1e9))
    else:
        ns = np.round(v.astype(np.float64) * (_UNITS[unit] * 1e9)).astype(np.int64)
    return t0 + ns.astype('timedelta64[ns]')


def _choose_float_dtype(raw_dtype, scale, offset):
    """xarray.coding.variables._choose_float_dtype: f32 only when nothing asks for more."""
    raw_dtype = np.dtype(raw_dtype)
    if raw_dtype.kind == 'f':
        if raw_dtype.itemsize <= 4 and all(a is None or np.asarray(a).dtype.itemsize <= 4 for a in (scale, offset)):
            return np.float32
        return np.float64 if raw_dtype.itemsize > 4 or scale is not None or offset is not None else np.float32
    if raw_dtype.kind in 'iu' and raw_dtype.itemsize <= 2:
        kinds = [np.asarray(a).dtype for a in (scale, offset) if a is not None]
        if all(k.kind == 'f' and k.itemsize <= 4 for k in kinds):
            return np.float32
    return np.float64


def decode_variable(var):
    """Masked + unpacked values of a scipy ``netcdf_variable`` (mask_and_scale of xarray)."""
    raw = np.array(var[...])                       # a copy: the mmap goes away with the file
    if raw.dtype.byteorder == '>':
        raw = raw.astype(raw.dtype.newbyteorder('='))
    scale, offset = getattr(var, 'scale_factor', None), getattr(var, 'add_offset', None)
    fills = [getattr(var, n) for n in ('_FillValue', 'missing_value') if hasattr(var, n)]
    if scale is None and offset is None and not fills:
        return raw
    dt = _choose_float_dtype(raw.dtype, scale, offset)
    out = raw.astype(dt)
    for f in fills:
        out[np.isin(raw, np.atleast_1d(f))] = np.nan
    if scale is not None:
        out *= np.asarray(scale).astype(dt).reshape(-1)[0]
    if offset is not None:
        out += np.asarray(offset).astype(dt).reshape(-1)[0]
    return out


def _read_netcdf3(path):
    from scipy.io import netcdf_file
    with open(path, 'rb') as fh:
        magic = fh.read(4)
    if magic[:3] != b'CDF':
        kind = 'NetCDF-4 / HDF5' if magic[1:4] == b'HDF' else 'not a NetCDF'
        raise NotImplementedError(f'{path}: {kind} file; only classic NetCDF-3 can be read without xarray + netCDF4 '
                                  '(convert with `nccopy -k classic`, or pass u= and v= arrays)')
    f = netcdf_file(path, 'r', mmap=False)
    try:
        coords = {}
        for name, var in f.variables.items():
            if var.dimensions == (name,):                          # coordinate variable
                vals = decode_variable(var)
                units = _attr(var, 'units', '')
                if isinstance(units, str) and ' since ' in units:
                    vals = decode_time(vals, units, _attr(var, 'calendar', 'standard'))
                coords[name] = vals
        data = {}
        for name, var in f.variables.items():
            if name in coords:
                continue
            dims = tuple(var.dimensions)
            c = {d: coords[d] for d in dims if d in coords}
            data[name] = DataArray(decode_variable(var), dims, c, name=name)
    finally:
        f.close()
    return Dataset(data)


def open_dataset(path):
    """``xr.open_dataset(path)`` (LCS.py:85): xarray when it is installed, else the NetCDF-3 reader above."""
    if _xr is not None:                                            # pragma: no cover - xarray is absent in the build image
        return _xr.open_dataset(path)
    return _read_netcdf3(path)


def to_netcdf(da, path):
    """``DataArray.to_netcdf(path)`` (LCS.py:250-262) as a classic NetCDF-3 file: the array under its name (or
    ``__xarray_dataarray_variable__``, as xarray calls an unnamed one), its dimension coordinates, datetime64 coordinates
    encoded as integer hours / seconds since the first stamp."""
    if _xr is not None and isinstance(da, _xr.DataArray):         # pragma: no cover
        return da.to_netcdf(path)
    from scipy.io import netcdf_file
    f = netcdf_file(path, 'w', version=2)
    try:
        for d, n in zip(da.dims, da.shape):
            f.createDimension(d, n)
        for d in da.dims:
            if d not in da.coords:
                continue
            c = np.asarray(da.coords[d])
            if np.issubdtype(c.dtype, np.datetime64):
                c = c.astype('datetime64[ns]')
                t0 = c.min() if c.size else np.datetime64('1970-01-01', 'ns')
                secs = (c - t0).astype('timedelta64[s]').astype(np.int64)
                exact_hours = bool(((c - t0).astype(np.int64) % 3_600_000_000_000 == 0).all())
                v = f.createVariable(d, 'i', (d,))
                v[:] = (secs // 3600 if exact_hours else secs).astype(np.int32)
                v.units = ('hours' if exact_hours else 'seconds') + ' since ' + str(t0.astype('datetime64[s]')).replace('T', ' ')
                v.calendar = 'proleptic_gregorian'
            else:
                c = c.astype(np.float64) if c.dtype.kind in 'fiu' else c
                v = f.createVariable(d, c.dtype.char, (d,))
                v[:] = c
        vals = np.asarray(da.values)
        if vals.dtype.kind not in 'fiu':
            vals = vals.astype(np.float64)
        v = f.createVariable(da.name or '__xarray_dataarray_variable__', vals.dtype.char, tuple(da.dims))
        v[...] = vals
    finally:
        f.close()
    return path
