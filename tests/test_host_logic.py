"""CPU-only checks: the C-ABI library builds, loads and exports every symbol include/lcs_b200.h declares
(no compute calls), the labelled-array shim, the sharding planners, and loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    from lagrangiancoherence_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    header = open(os.path.join(ROOT, 'include', 'lcs_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(lcs_[a-z0-9_]+)\s*\(', header))
    assert {'lcs_prefilter', 'lcs_pack_pairs', 'lcs_pack_es', 'lcs_advect', 'lcs_ftle_epilogue', 'lcs_map_coordinates',
            'lcs_fourth_order_derivative', 'lcs_spectral_norm_3x3', 'lcs_gather_peak'} <= declared
    handle = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(handle, name), f'{name} declared in lcs_b200.h but not exported'
    from lagrangiancoherence_b200 import _lib
    assert set(_lib.SIGNATURES) == declared          # the ctypes binding covers the whole header
    handle.lcs_abi_version.restype = ctypes.c_int
    assert handle.lcs_abi_version() == _lib.ABI_VERSION


def test_ctypes_structs_match_the_header_layout(lib_path):
    from lagrangiancoherence_b200 import _lib
    assert ctypes.sizeof(_lib.Grid) == 8 + 4 * 8
    assert ctypes.sizeof(_lib.Particles) == 16 + 4 * 8 + 2 * 8
    assert ctypes.sizeof(_lib.AdvectOpts) == 10 * 4 + 8 and ctypes.sizeof(_lib.XRank) == 16 + 8 + 8
    assert ctypes.sizeof(_lib.Winds) == 8 + 4 * 8 + 8


def test_argument_validation_without_a_gpu(lib_path):
    """Entry points reject bad arguments before touching the device (safe to call on a CPU-only box)."""
    from lagrangiancoherence_b200 import _lib
    lib = _lib.load()
    assert lib.lcs_prefilter(None, None, 0, None, None, None, 0, 1, 8, 8, 3, None) == -1
    assert b'null' in lib.lcs_last_error()
    assert lib.lcs_pack_pairs(None, None, 0, None, 0, 2, 8, 8, None) == -1
    assert lib.lcs_ftle_epilogue(None, None, 1, 8, 8, 0, 8, 0, 8, None, 1.0, None, 0, None, None, None, None) == -1
    assert lib.lcs_prefilter_scratch_bytes(3, 10, 20) == 2 * 3 * 10 * 20 * 8
    part = _lib.Particles(10, 20, 0, 10, None, None, None, None, 1.0, 0.5)
    opts = _lib.AdvectOpts(8, 4, 3, _lib.LCS_X_CLAMP_POINTWISE, 0, 1, 0, 1, 0)
    assert lib.lcs_advect_workspace_bytes(ctypes.byref(part), ctypes.byref(opts)) == 0
    opts.xmode = _lib.LCS_X_CLAMP_OUTER
    assert lib.lcs_advect_workspace_bytes(ctypes.byref(part), ctypes.byref(opts)) > 10 * 20 * 32


def test_engine_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from lagrangiancoherence_b200 import _lib
    from lagrangiancoherence_b200.engine import FtleEngine
    with pytest.raises(_lib.LcsError, match='no CPU path'):
        FtleEngine(np.linspace(0, 1, 8), np.linspace(0, 1, 8), 1.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'lagrangiancoherence_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{f} imports the oracle'


# ------------------------------------------------------------------ labelled arrays
def test_dataarray_shim_semantics():
    from lagrangiancoherence_b200 import DataArray, Dataset
    lat, lon, t = np.array([3., 1., 2.]), np.array([10., 20.]), np.arange(4)
    a = DataArray(np.arange(24.).reshape(4, 3, 2), ('time', 'latitude', 'longitude'),
                  {'time': t, 'latitude': lat, 'longitude': lon})
    s = a.sortby('latitude')
    assert np.array_equal(s.coords['latitude'], [1., 2., 3.]) and np.array_equal(s.values[0, :, 0], [2., 4., 0.])
    assert a.transpose('latitude', 'time', 'longitude').shape == (3, 4, 2)
    assert a.isel(time=0).dims == ('latitude', 'longitude') and a.isel({'time': slice(1, 3)}).shape == (2, 3, 2)
    assert np.array_equal(a.latitude.values, lat) and a['time'].shape == (4,)
    half_log = np.log(a + 1) / 2
    assert isinstance(half_log, DataArray) and half_log.dims == a.dims
    c = a.copy(data=np.zeros_like(a.values))
    assert c.values.sum() == 0 and a.values.sum() != 0
    assert a.expand_dims('x').shape == (1, 4, 3, 2)
    ds = Dataset({'u': a, 'v': a})
    assert ds.u is a and ds.copy().v is not a
    with pytest.raises(ValueError):
        DataArray(np.zeros((2, 2)), ('a',))


# ------------------------------------------------------------------ sharding planners
def test_shard_planners_cover_everything_exactly_once():
    from lagrangiancoherence_b200.rolling import shard_rows, shard_starts, chunk_starts
    for n in (1, 7, 8, 281, 721, 8760):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_starts(n, world, r) for r in range(world)]
            assert sum(c for _, c in blocks) == n
            assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
            rows = [shard_rows(n, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n
            for o0, o1, i0, i1 in rows:
                assert i0 == max(0, o0 - 2) and i1 == min(n, o1 + 2)        # 2-row halo: y-stencil spans +-2 rows
    assert chunk_starts(5, 10, 4) == [(5, 4), (9, 4), (13, 2)]


def test_dropna_unstack_removes_fully_dropped_rows_and_columns():
    """LCS.py:146,157: a point with a NaN derivative is dropped from the stacked index and comes back as NaN after
    unstack; a latitude/longitude with no surviving point is not a level any more and vanishes from the result."""
    from lagrangiancoherence_b200.LCS.LCS import drop_unused_levels
    lat, lon = np.arange(5.0), np.arange(6.0)
    s = np.arange(30.0).reshape(5, 6)
    out, la, lo = drop_unused_levels(s, lat, lon)
    assert out is s and la is lat and lo is lon
    s[1, 2] = np.nan                       # isolated NaN: shape kept
    out, la, lo = drop_unused_levels(s, lat, lon)
    assert out.shape == (5, 6) and np.isnan(out[1, 2])
    s[3, :] = np.nan                       # whole row dropped
    s[:, 0] = np.nan                       # whole column dropped
    out, la, lo = drop_unused_levels(s, lat, lon)
    assert out.shape == (4, 5) and np.array_equal(la, [0, 1, 2, 4]) and np.array_equal(lo, [1, 2, 3, 4, 5])
    assert np.isnan(out[1, 1]) and np.isnan(out).sum() == 1


def test_numa_binding_helper_is_best_effort():
    """affinity.bind_host_to_gpu: parses sysfs cpulists, and is a silent no-op where NVML / sysfs give no answer
    (this container has no GPU)."""
    from lagrangiancoherence_b200 import affinity
    assert affinity._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist('') == set()
    before = os.sched_getaffinity(0)
    res = affinity.bind_host_to_gpu(0)
    assert res is None or set(res) == {'node', 'cpus'}
    if res is None:
        assert os.sched_getaffinity(0) == before


def test_chunk_schedule_covers_the_series_with_short_ends():
    from lagrangiancoherence_b200.rolling import chunk_schedule
    assert chunk_schedule(1184, 148) == [(0, 148), (148, 296), (444, 296), (740, 296), (1036, 148)]
    for count in (1, 7, 148, 149, 296, 443, 444, 445, 600, 1095, 8760):
        for chunk in (24, 148):
            sched = chunk_schedule(count, chunk)
            assert sched[0][0] == 0 and sum(n for _, n in sched) == count
            assert all(a + n == b for (a, n), (b, _) in zip(sched, sched[1:]))
            assert all(0 < n <= 2 * chunk for _, n in sched)
            assert sched[0][1] <= chunk and sched[-1][1] <= chunk
            ramp = chunk_schedule(count, chunk, ramp=True)
            assert ramp[0][0] == 0 and sum(n for _, n in ramp) == count and all(0 < n <= 2 * chunk for _, n in ramp)
            assert all(a + n == b for (a, n), (b, _) in zip(ramp, ramp[1:]))
    assert chunk_schedule(1184, 148, ramp=True) == [(0, 74), (74, 148), (222, 296), (518, 296), (814, 148), (962, 148), (1110, 74)]


def test_latlonsel_strict_open_intervals():
    """tools.py:158-188: both bounds are strict, slices and lists are accepted, either coordinate name works."""
    from lagrangiancoherence_b200 import DataArray
    from lagrangiancoherence_b200.LCS.tools import latlonsel
    lat, lon = np.arange(-4.0, 5.0), np.arange(10.0, 20.0)
    da = DataArray(np.arange(90.0).reshape(9, 10), ('lat', 'lon'), {'lat': lat, 'lon': lon})
    out = latlonsel(da, slice(-2, 2), [12, 15])
    assert np.array_equal(out.coords['lat'], [-1, 0, 1]) and np.array_equal(out.coords['lon'], [13, 14])
    assert np.array_equal(out.values, da.values[3:6, 3:5])
    da2 = DataArray(da.values, ('latitude', 'longitude'), {'latitude': lat, 'longitude': lon})
    assert latlonsel(da2, slice(-2, 2), slice(12, 15), latname='latitude', lonname='longitude').shape == (3, 2)
    with pytest.raises(AssertionError):
        latlonsel(da2, slice(-2, 2), slice(12, 15))


def test_bench_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the reference's CPU arithmetic on the host cores) prints ONE JSON line with the keys
    the driver reads; it needs no GPU, so the contract is checked here."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['metric'] == 'particle-steps/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['vs_baseline'] is None


def test_regrid_axis_plan_reproduces_interp1d_and_nearest():
    """regrid.axis_plan feeds lcs_regrid_linear_nearest; applied in numpy with the kernel's arithmetic (w_hi*y_hi + w_lo*y_lo,
    NaN outside the source range, nearest-label fill) it must equal scipy's interp1d + pandas' nearest reindex bit for bit,
    on irregular ascending coordinates, targets outside the range, and targets that hit source points exactly."""
    import pandas as pd
    from scipy.interpolate import interp1d
    from lagrangiancoherence_b200.regrid import axis_plan, common_grid
    rng = np.random.default_rng(7)
    for n_src, n_dst in ((5, 40), (89, 360), (180, 721), (2, 9)):
        src = np.sort(rng.uniform(-80, 80, n_src))
        src[1:] += np.arange(1, n_src) * 1e-3                      # strictly ascending
        dst = np.sort(np.concatenate([rng.uniform(-95, 95, n_dst - 3), src[:2], [src[-1]]]))
        y = rng.normal(size=(3, n_src))
        lo, w_hi, w_lo, valid, near = axis_plan(src, dst)
        got = w_hi * y[:, lo + 1] + w_lo * y[:, lo]
        got = np.where(valid.astype(bool), got, np.nan)
        got = np.where(np.isnan(got), y[:, near], got)
        ref = interp1d(src, y, kind='linear', axis=1, bounds_error=False, fill_value=np.nan, assume_sorted=True)(dst)
        ref = np.where(np.isnan(ref), y[:, pd.Index(src).get_indexer(dst, method='nearest')], ref)
        assert np.array_equal(got, ref)
    lats, lons = common_grid()
    assert lats.size == 360 and lons.size == 721 and lats[0] == -89.75 and lons[-1] == 179.5     # LCS.py:106-107
    with pytest.raises(ValueError):
        axis_plan([0.0, 0.0, 1.0], [0.5])


def test_device_array_stays_lazy_until_values_are_asked_for():
    """labelled.DeviceArray (the winds of the global path after the device regrid / truncation): the operations LCS.__call__
    and propagate() apply -- sortby on sorted coordinates, transpose to the order it already has, coordinate lookup, isel of
    one level -- must not download the series; everything else falls back to a host DataArray."""
    import torch
    from lagrangiancoherence_b200.labelled import DeviceArray, DataArray, coord_values
    t = torch.arange(24, dtype=torch.float64).reshape(2, 3, 4)
    d = DeviceArray(t, ('time', 'latitude', 'longitude'), {'time': np.arange(2), 'latitude': np.arange(3.0), 'longitude': np.arange(4.0)})
    assert d.shape == (2, 3, 4) and d.dtype == np.float64 and d.size == 24 and d.ndim == 3
    assert d.sortby('longitude').sortby('latitude') is d and d.transpose('time', 'latitude', 'longitude') is d
    assert np.array_equal(coord_values(d, 'latitude'), np.arange(3.0))
    lvl = d.isel({'time': 0})
    assert isinstance(lvl, DataArray) and lvl.dims == ('latitude', 'longitude') and np.array_equal(lvl.values, t[0].numpy())
    assert d._host_values is None                                   # nothing above touched the whole series
    assert d.transpose('latitude', 'time', 'longitude').shape == (3, 2, 4)
    assert np.array_equal(d.values, t.numpy()) and np.array_equal((d * 2).values, 2 * t.numpy())
    assert np.array_equal(d.copy().values, t.numpy()) and np.array_equal(d.isel(latitude=slice(1, 3)).values, t.numpy()[:, 1:3])
