#!/usr/bin/env python
"""One measured line per BASELINE.json config on ONE GPU (bench.py holds the headline metric on configs[1]; the
multi-GPU lines come from `bench.py --gpus N` and scripts/bench_rowbands.py).  Device times are CUDA events after
warm-up with a 256 MiB L2 flush before every timed call; `wall` is the host clock around the public API call.

    python scripts/bench_configs.py [C1 C2 C3 C4 C5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine
from lagrangiancoherence_b200.labelled import DataArray
from lagrangiancoherence_b200.LCS.LCS import LCS

dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def wall(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter(); fn(); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts))


def arrays(u, v, lat, lon, hours):
    t = (np.datetime64('2000-01-01T00') + np.arange(u.shape[0]) * np.timedelta64(hours, 'h')).astype('datetime64[ns]')
    c = {'time': t, 'latitude': lat, 'longitude': lon}
    return DataArray(u, ('time', 'latitude', 'longitude'), c), DataArray(v, ('time', 'latitude', 'longitude'), c)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c1():
    """configs[0]: examples/ideal_vortex.py, 89x180 2-degree grid, nt=8, cyclic, through the public LCS call."""
    u, v, lat, lon = S.ideal_vortex(**S.vortex_config_subtropical)
    du, dv = arrays(u, v, lat, lon, 6)
    lcs = LCS(timestep=-21600, timedim='time', SETTLS_order=4)
    devnull = open(os.devnull, 'w')

    def call():
        so = sys.stdout
        sys.stdout = devnull
        try:
            return lcs(u=du, v=dv, verbose=False, isglobal=True, interp_to_common_grid=False, truncation=None)
        finally:
            sys.stdout = so
    ms = wall(call)
    eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='cyclic', device=dev)
    tu, tv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
    dms = timed(lambda: eng.ftle(tu, tv))
    psteps = lat.size * lon.size * (u.shape[0] - 1)
    emit(config='C1 ideal vortex 89x180, nt=8, S=4, cubic, cyclic (examples/ideal_vortex.py)', particle_steps=psteps,
         api_wall_ms=ms, device_ms=dms, particle_steps_per_s_device=psteps / dms * 1e3, fields_per_s_api=1e3 / ms)


def c2():
    """configs[1] as ONE drop-in call (the batched headline is bench.py)."""
    lat, lon = S.grid_c2()
    u, v = S.era5_like_winds(lat, lon, 9)
    du, dv = arrays(u, v, lat, lon, 6)
    for xclamp in ('outer', 'pointwise'):
        lcs = LCS(timestep=-21600, timedim='time', SETTLS_order=4)
        devnull = open(os.devnull, 'w')

        def call():
            so = sys.stdout
            sys.stdout = devnull
            try:
                return lcs(u=du, v=dv, verbose=False, xclamp=xclamp)
            finally:
                sys.stdout = so
        ms = wall(call)
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xclamp, device=dev)
        tu, tv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
        dms = timed(lambda: eng.ftle(tu, tv))
        psteps = lat.size * lon.size * 8
        emit(config=f'C2 281x321, nt=9, S=4, cubic, f64, xclamp={xclamp}: ONE field per call', particle_steps=psteps,
             api_wall_ms=ms, device_ms=dms, particle_steps_per_s_device=psteps / dms * 1e3, fields_per_s_api=1e3 / ms)


def c3():
    """configs[2]: 721x1440 hourly, 72 h backward (nt=73), one field and a batch of 8 start times, one GPU."""
    lat, lon = S.grid_c3()
    for B, xmode in ((1, 'pointwise'), (1, 'outer'), (8, 'pointwise'), (8, 'outer')):
        nt = 73
        u, v = S.era5_like_winds(lat, lon, nt - 1 + B, noise=0.0)
        eng = FtleEngine(lat, lon, -3600, SETTLS_order=4, xmode=xmode, device=dev)
        tu, tv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
        st = eng.stage(tu, tv)
        x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device=dev); y = torch.empty_like(x)
        t_stage = timed(lambda: eng.stage(tu, tv), n=3, warm=1)
        t_adv = timed(lambda: eng.advect(st, nsteps=nt - 1, nwindows=B, out=(x, y)), n=3, warm=1)
        t_epi = timed(lambda: eng.epilogue(x, y), n=3, warm=1)
        psteps = B * lat.size * lon.size * (nt - 1)
        emit(config=f'C3 721x1440 hourly, 72 h backward (nt=73), S=4, cubic, f64, xclamp={xmode}, {B} start time(s) per launch',
             particle_steps=psteps, stage_ms=t_stage, advect_ms=t_adv, epilogue_ms=t_epi,
             particle_steps_per_s_advect=psteps / t_adv * 1e3,
             particle_steps_per_s_whole=psteps / (t_stage + t_adv + t_epi) * 1e3, fields_per_s=B / (t_stage + t_adv + t_epi) * 1e3)
        del st, tu, tv, x, y
        torch.cuda.empty_cache()


def c4():
    """configs[3]: hourly rolling series on the C2 grid, 48 h windows (nt=49); 1095 start times = one GPU's share of a
    year sharded 8-way, through the public pipelined rolling API (pinned host winds in, pinned host fields out)."""
    from lagrangiancoherence_b200.rolling import rolling_ftle
    lat, lon = S.grid_c2()
    B, nt = 1095, 49
    u, v = S.era5_like_winds(lat, lon, B + nt - 1, noise=0.0)
    hu, hv = torch.from_numpy(u).pin_memory(), torch.from_numpy(v).pin_memory()
    out = torch.empty((B, lat.size, lon.size), dtype=torch.float64).pin_memory()
    for xclamp in ('outer', 'pointwise'):
        eng = FtleEngine(lat, lon, -3600, SETTLS_order=4, xmode=xclamp, device=dev)
        ms = timed(lambda: rolling_ftle(hu, hv, lat, lon, nt, -3600, engine=eng, out=out, chunk=148), n=3, warm=1)
        psteps = B * lat.size * lon.size * (nt - 1)
        emit(config=f'C4 rolling series, C2 grid, hourly, 48 h windows (nt=49), 1095 start times (1/8 of a year), xclamp={xclamp}, '
                    'host in -> host out', particle_steps=psteps, e2e_ms=ms, particle_steps_per_s_e2e=psteps / ms * 1e3,
             fields_per_s_e2e=B / ms * 1e3, year_on_8_gpus_s=ms * 1e-3)


def c5():
    """configs[4]: particle grid refined 4x per dimension (1121x1281) over the C2 winds, trajectories stored."""
    lat, lon = S.grid_c2()
    u, v = S.era5_like_winds(lat, lon, 9)
    fl = np.linspace(lat[0], lat[-1], 4 * (lat.size - 1) + 1)
    fo = np.linspace(lon[0], lon[-1], 4 * (lon.size - 1) + 1)
    for xmode in ('pointwise', 'outer'):
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev, part_lat=fl, part_lon=fo)
        st = eng.stage(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev))
        ms = timed(lambda: eng.advect(st, return_traj=True), n=5, warm=2)
        _, _, xt, yt = eng.advect(st, return_traj=True)
        hx = torch.empty(xt.shape, dtype=xt.dtype).pin_memory(); hy = torch.empty_like(hx).pin_memory()
        d2h = timed(lambda: (hx.copy_(xt, non_blocking=True), hy.copy_(yt, non_blocking=True)), n=3, warm=1)
        psteps = fl.size * fo.size * 8
        emit(config=f'C5 1121x1281 particles (4x refined) over C2 winds, nt=9, trajectories stored, xclamp={xmode}',
             particle_steps=psteps, advect_ms=ms, particle_steps_per_s_advect=psteps / ms * 1e3,
             trajectory_bytes=int(2 * xt.numel() * 8), trajectory_write_GBs=2 * xt.numel() * 8 / ms / 1e6,
             d2h_ms=d2h, d2h_GBs=2 * xt.numel() * 8 / d2h / 1e6)


if __name__ == '__main__':
    which = sys.argv[1:] or ['C1', 'C2', 'C3', 'C4', 'C5']
    for name in which:
        {'C1': c1, 'C2': c2, 'C3': c3, 'C4': c4, 'C5': c5}[name]()
