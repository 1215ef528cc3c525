"""NetCDF-3 in / out behind ``LCS(...)(ds='path.nc')`` and the command line (LCS.py:84-87, 236-265): CF decoding on
the CPU; the path-taking call and the CLI on the GPU."""
import os
import sys

import numpy as np
import pytest
from scipy.io import netcdf_file

from lagrangiancoherence_b200 import DataArray, ncio, synthetic as S


def write_winds(path, u, v, lat, lon, hours0=876576, step_h=6, packed=False):
    f = netcdf_file(path, 'w')
    nt = u.shape[0]
    f.createDimension('time', nt); f.createDimension('latitude', lat.size); f.createDimension('longitude', lon.size)
    t = f.createVariable('time', 'i', ('time',)); t[:] = hours0 + step_h * np.arange(nt)
    t.units = 'hours since 1900-01-01 00:00:00.0'; t.calendar = 'gregorian'
    la = f.createVariable('latitude', 'f', ('latitude',)); la[:] = lat
    lo = f.createVariable('longitude', 'f', ('longitude',)); lo[:] = lon
    for name, a in (('u', u), ('v', v)):
        if packed:                                   # ERA5 style: int16 + f64 scale_factor / add_offset
            off, sc = float(a.max() + a.min()) / 2, float(np.ptp(a)) / 65000.0
            var = f.createVariable(name, 'h', ('time', 'latitude', 'longitude'))
            var[:] = np.round((a - off) / sc).astype(np.int16)
            var.scale_factor = np.float64(sc); var.add_offset = np.float64(off); var._FillValue = np.int16(-32767)   # (scipy stores Python floats as f32)
        else:
            var = f.createVariable(name, 'd', ('time', 'latitude', 'longitude'))
            var[:] = a
    f.close()


def test_cf_decoding(tmp_path):
    lat = np.linspace(10.0, -10.0, 9)                 # descending, as ERA5 ships it
    lon = np.linspace(0.0, 20.0, 11)
    u, v = S.era5_like_winds(lat[::-1], lon, 3)
    p = str(tmp_path / 'w.nc')
    write_winds(p, u, v, lat, lon)
    ds = ncio.open_dataset(p)
    assert ds.u.dims == ('time', 'latitude', 'longitude') and ds.u.dtype == np.float64
    assert np.array_equal(ds.u.values, u) and np.array_equal(ds.v.values, v)
    assert ds.u.coords['time'].dtype == np.dtype('datetime64[ns]')
    assert ds.u.coords['time'][0] == np.datetime64('2000-01-01T00', 'ns') and ds.u.coords['time'][2] == np.datetime64('2000-01-01T12', 'ns')
    assert np.allclose(ds.u.coords['latitude'], lat)
    write_winds(p, u, v, lat, lon, packed=True)
    ds = ncio.open_dataset(p)
    assert ds.u.dtype == np.float64                   # int16 with f64 attributes: float64, as xarray decodes ERA5
    assert np.abs(ds.u.values - u).max() <= np.ptp(u) / 65000.0


def test_packed_f32_attributes_and_fill_values(tmp_path):
    p = str(tmp_path / 'p.nc')
    f = netcdf_file(p, 'w')
    f.createDimension('x', 4)
    a = f.createVariable('a', 'h', ('x',)); a[:] = [1, 2, 7, 4]
    a.scale_factor = np.float32(0.5); a.add_offset = np.float32(2.0); a._FillValue = np.int16(7)
    b = f.createVariable('b', 'f', ('x',)); b[:] = [1, 2, 3, 4]
    f.close()
    ds = ncio.open_dataset(p)
    assert ds['a'].dtype == np.float32 and ds['b'].dtype == np.float32
    assert np.array_equal(ds['a'].values, np.array([2.5, 3.0, np.nan, 4.0], np.float32), equal_nan=True)


@pytest.mark.parametrize('units,first', [('days since 2000-1-1', '2000-01-03T00'), ('seconds since 1970-01-01T00:00:00Z', '1970-01-01T00:00:02'),
                                         ('minutes since 2001-02-03 04:05', '2001-02-03T04:07')])
def test_time_units(units, first):
    assert ncio.decode_time(np.array([2]), units)[0] == np.datetime64(first, 'ns')
    with pytest.raises(NotImplementedError):
        ncio.decode_time(np.array([2]), units, calendar='360_day')


def test_round_trip_of_a_result_and_refusal_of_hdf5(tmp_path):
    t = np.array(['2000-01-01T00'], dtype='datetime64[ns]')
    out = DataArray(np.random.default_rng(0).random((1, 4, 5)), ('time', 'latitude', 'longitude'),
                    {'time': t, 'latitude': np.arange(4.0), 'longitude': np.arange(5.0)})
    p = str(tmp_path / 'o.nc')
    out.to_netcdf(p)
    back = ncio.open_dataset(p)['__xarray_dataarray_variable__']
    assert np.array_equal(back.values, out.values) and back.coords['time'][0] == t[0] and back.dims == out.dims
    h = str(tmp_path / 'h.nc')
    open(h, 'wb').write(b'\x89HDF\r\n\x1a\n' + b'\0' * 64)
    with pytest.raises(NotImplementedError, match='NetCDF-4'):
        ncio.open_dataset(h)


@pytest.mark.gpu
def test_lcs_call_with_a_path_equals_the_array_call(cuda_device, tmp_path):
    from lagrangiancoherence_b200.LCS.LCS import LCS
    lat = np.linspace(10.0, -30.0, 41)
    lon = np.linspace(-80.0, -24.0, 57)
    u, v = S.era5_like_winds(lat[::-1], lon, 5)
    u, v = u[:, ::-1], v[:, ::-1]
    p = str(tmp_path / 'w.nc')
    write_winds(p, u, v, lat, lon)
    lcs = LCS(timestep=-6 * 3600, timedim='time', SETTLS_order=4)
    a = lcs(ds=p, verbose=False)
    ds = ncio.open_dataset(p)
    b = lcs(u=ds.u, v=ds.v, verbose=False)
    assert np.array_equal(a.values, b.values) and a.coords['time'][0] == np.datetime64('2000-01-01T00', 'ns')


@pytest.mark.gpu
def test_command_line(cuda_device, tmp_path, capsys):
    """LCS.py:236-265: global call (regrid + T20), result saved, the input removed afterwards (as upstream), kept with `keep`."""
    from lagrangiancoherence_b200.LCS import LCS as M
    lat = np.arange(-88.0, 89.0, 4.0)
    lon = np.arange(-180.0, 180.0, 4.0)
    u, v = S.era5_like_winds(lat, lon, 3)
    p, o = str(tmp_path / 'input_partial.nc'), str(tmp_path / 'SL_attracting_x.nc')
    write_winds(p, u, v, lat, lon)
    M.main(['LCS.py', '-21600', 'time', '2', '-80/-30/-40/0', p, o, 'True', 'keep'])
    assert os.path.exists(p)
    out = ncio.open_dataset(o)['__xarray_dataarray_variable__']
    assert out.shape == (1, 360, 721) and np.isfinite(out.values).all()
    xt = ncio.open_dataset(o.replace('SL_attracting', 'x_departure'))['__xarray_dataarray_variable__']
    assert xt.shape == (3, 360, 721) and os.path.exists(o.replace('SL_attracting', 'y_departure'))
    ref = M.LCS(timestep=-21600.0, timedim='time', SETTLS_order=2)(ds=p, isglobal=True, return_traj=False, verbose=False)
    assert np.array_equal(ref.values, out.values)
    M.main(['LCS.py', '-21600', 'time', '2', '-80/-30/-40/0', p, o, 'False'])
    assert not os.path.exists(p)                                     # subprocess.call(['rm', input_path]), LCS.py:265
    assert 'Saving to' in capsys.readouterr().out
