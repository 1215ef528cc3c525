"""Scratch probe (round 2, second session): f32 fast-path accuracy / speed, and where a single C2 call spends its time.
    python scripts/probe_r2b.py [f32] [single] [profile]"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine, precision_args

dev = torch.device('cuda', 0)


def timeit(fn, n=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def f32():
    lat = np.linspace(-40.0, 0.0, 161)
    lon = np.linspace(-80.0, -30.0, 201)
    u, v = S.era5_like_winds(lat, lon, 9)
    for xmode in ('pointwise', 'outer'):
        res = {}
        for prec in ('f64', 'f32', 'f32fast'):
            eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev, **precision_args(prec))
            x, y = eng.advect(eng.stage(u, v))
            sig = eng.epilogue(x, y)[0].cpu().numpy()
            res[prec] = (x[0].cpu().numpy(), y[0].cpu().numpy(), sig)
        ref = res['f64'][2]
        good = ref > 1e-6
        fref = 0.5 * np.log(ref[good])
        for prec in ('f32', 'f32fast'):
            frel = np.abs(0.5 * np.log(res[prec][2][good]) - fref) / np.maximum(np.abs(fref), 1e-3)
            print(json.dumps(dict(case='tolerance', xmode=xmode, precision=prec,
                                  pos_rel_y=float(np.abs(res[prec][1] - res['f64'][1]).max() / 40),
                                  pos_rel_x=float(np.abs(res[prec][0] - res['f64'][0]).max() / 80),
                                  within_1e5=float((frel <= 1e-5).mean()), within_1e4=float((frel <= 1e-4).mean()))), flush=True)
    lat, lon = S.grid_c2()
    B = 296
    u, v = S.era5_like_winds(lat, lon, 8 + B, noise=0.0)
    du, dv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
    for xmode in ('pointwise', 'outer'):
        base = None
        for prec in ('f64', 'f32', 'f32fast'):
            eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev, **precision_args(prec))
            st = eng.stage(du, dv)
            x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device=dev); y = torch.empty_like(x)
            ms = timeit(lambda: eng.advect(st, nsteps=8, nwindows=B, out=(x, y)))
            base = base or ms
            print(json.dumps(dict(case='speed C2 x296', xmode=xmode, precision=prec, advect_ms=ms, speedup_vs_f64=base / ms,
                                  Gpsteps=B * lat.size * lon.size * 8 / ms / 1e6)), flush=True)
            del eng, st


def single():
    lat, lon = S.grid_c2()
    u, v = S.era5_like_winds(lat, lon, 9)
    du, dv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
    for xmode in ('pointwise', 'outer'):
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev)
        st = eng.stage(du, dv)
        x, y = eng.advect(st)
        out = dict(case='single C2 field', xmode=xmode,
                   stage_ms=timeit(lambda: eng.stage(du, dv)),
                   advect_ms=timeit(lambda: eng.advect(st, out=(x, y))),
                   epilogue_ms=timeit(lambda: eng.epilogue(x, y)),
                   ftle_ms=timeit(lambda: eng.ftle(du, dv)))
        if hasattr(eng, 'ftle_graph'):
            out['ftle_graph_ms'] = timeit(lambda: eng.ftle_graph(du, dv))
        print(json.dumps(out), flush=True)


def kernels():
    # true kernel times of ONE C2 field: 30 launches back to back between two events (the CPU runs ahead of the GPU),
    # against the one-launch-per-event-pair figure that includes the host's launch path
    lat, lon = S.grid_c2()
    u, v = S.era5_like_winds(lat, lon, 9)
    du, dv = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
    for xmode in ('pointwise', 'outer'):
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev)
        st = eng.stage(du, dv)
        x, y = eng.advect(st)
        out = dict(case='one C2 field, back-to-back launches', xmode=xmode)
        for name, fn in (('stage', lambda: eng.stage(du, dv, reuse=True)), ('advect', lambda: eng.advect(st, out=(x, y))),
                         ('epilogue', lambda: eng.epilogue(x, y)), ('ftle', lambda: eng.ftle(du, dv))):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            for _ in range(30):
                fn()
            b.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            out[name + '_ms'] = a.elapsed_time(b) / 30
            out[name + '_host_issue_ms'] = (t1 - t0) * 1e3 / 30
        print(json.dumps(out), flush=True)


def stage():
    for name, (lat, lon), nlev in (('C2', S.grid_c2(), 1192), ('C3', S.grid_c3(), 73), ('C2 single', S.grid_c2(), 9)):
        u, v = S.era5_like_winds(lat, lon, min(nlev, 40), noise=0.0)
        reps = (nlev + u.shape[0] - 1) // u.shape[0]
        du = torch.from_numpy(u).to(dev).repeat(reps, 1, 1)[:nlev].contiguous(); dv = torch.from_numpy(v).to(dev).repeat(reps, 1, 1)[:nlev].contiguous()
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='pointwise', device=dev)
        from lagrangiancoherence_b200 import _lib
        from lagrangiancoherence_b200.engine import _ptr, _stream
        cu, cv = torch.empty_like(du), torch.empty_like(dv)
        scratch = torch.empty((2,) + tuple(du.shape), dtype=torch.float64, device=dev)
        ms_pre = timeit(lambda: _lib.check(eng.lib.lcs_prefilter(_ptr(du), _ptr(dv), _lib.LCS_F64, _ptr(cu), _ptr(cv), _ptr(scratch), scratch.numel() * 8,
                                                                 nlev, lat.size, lon.size, 3, _stream(dev)), 'prefilter'))
        del cu, cv, scratch
        ms_all = timeit(lambda: eng.stage(du, dv))
        vals = 2 * nlev * lat.size * lon.size
        print(json.dumps(dict(case='staging ' + name, nlev=nlev, prefilter_ms=ms_pre, stage_ms=ms_all, pack_ms=ms_all - ms_pre,
                              prefilter_GBs_compulsory=vals * 32 / ms_pre / 1e6)), flush=True)
        del du, dv, eng


def blocks():
    for name, (lat, lon), nt, xmode in (('C2', S.grid_c2(), 9, 'pointwise'), ('C1', S.grid_c1(), 8, 'cyclic')):
        for B in (1, 2, 4, 16):
            u, v = S.era5_like_winds(lat, lon, nt - 1 + B, noise=0.0)
            eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, device=dev)
            st = eng.stage(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev))
            x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device=dev); y = torch.empty_like(x)
            res = {}
            os.environ['LCS_ADVECT_SMALL'] = '0'
            for bt in ('256', '128', '64', '0'):
                os.environ['LCS_ADVECT_BLOCK'] = bt
                res[bt] = timeit(lambda: eng.advect(st, nsteps=nt - 1, nwindows=B, out=(x, y)), n=15)
            os.environ['LCS_ADVECT_BLOCK'] = '0'
            for sm in ('1', '-1'):
                os.environ['LCS_ADVECT_SMALL'] = sm
                res['small=' + sm] = timeit(lambda: eng.advect(st, nsteps=nt - 1, nwindows=B, out=(x, y)), n=15)
            print(json.dumps(dict(case='block size ' + name, windows=B, advect_ms=res)), flush=True)


def timing():
    # with LCS_B200_LIB = a -DLCS_OUTER_TIMING build and LCS_OUTER_TIMING_PRINT=1: cycles per sub-step by phase
    lat, lon = S.grid_c2()
    for B in (1, 4, 296):
        u, v = S.era5_like_winds(lat, lon, 8 + B, noise=0.0)
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode='outer', device=dev)
        st = eng.stage(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev))
        x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device=dev); y = torch.empty_like(x)
        os.environ['LCS_OUTER_TIMING_PRINT'] = '0'
        for _ in range(3):
            eng.advect(st, nsteps=8, nwindows=B, out=(x, y))
        torch.cuda.synchronize()
        os.environ['LCS_OUTER_TIMING_PRINT'] = '1'
        eng.advect(st, nsteps=8, nwindows=B, out=(x, y))
        torch.cuda.synchronize()
        os.environ['LCS_OUTER_TIMING_PRINT'] = '0'
        print(json.dumps(dict(case='outer timing', windows=B, advect_ms=timeit(lambda: eng.advect(st, nsteps=8, nwindows=B, out=(x, y))))), flush=True)


def global_call():
    # the reference's default global call: 2-degree winds -> 360x721 regrid -> T20 truncation -> cyclic integration -> FTLE
    from lagrangiancoherence_b200.labelled import DataArray, Dataset
    from lagrangiancoherence_b200.LCS.LCS import LCS
    u, v, lat, lon = S.ideal_vortex(**S.vortex_config_subtropical)
    t = (np.datetime64('2000-01-01T00') + np.arange(u.shape[0]) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
    c = {'time': t, 'latitude': lat, 'longitude': lon}
    ds = Dataset({'u': DataArray(u, ('time', 'latitude', 'longitude'), c), 'v': DataArray(v, ('time', 'latitude', 'longitude'), c)})
    lcs = LCS(timestep=-21600, timedim='time', SETTLS_order=4)
    devnull = open(os.devnull, 'w')

    def call():
        so = sys.stdout
        sys.stdout = devnull
        try:
            return lcs(ds, isglobal=True, verbose=False)
        finally:
            sys.stdout = so
    for _ in range(3):
        out = call()
    ts = []
    for _ in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter(); call(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps(dict(case='global default call (regrid + T20), 89x180 -> 360x721, nt=%d' % u.shape[0], ms_median=float(np.median(ts)),
                          ms_min=float(np.min(ts)), checksum=float(np.nansum(out.values)))), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        call()
    pr.disable()
    s_ = io.StringIO()
    pstats.Stats(pr, stream=s_).sort_stats('cumulative').print_stats(22)
    print(s_.getvalue()[:5000])


def profile():
    from lagrangiancoherence_b200.labelled import DataArray
    from lagrangiancoherence_b200.LCS.LCS import LCS
    lat, lon = S.grid_c2()
    u, v = S.era5_like_winds(lat, lon, 9)
    t = (np.datetime64('2000-01-01T00') + np.arange(9) * np.timedelta64(6, 'h')).astype('datetime64[ns]')
    c = {'time': t, 'latitude': lat, 'longitude': lon}
    du, dv = DataArray(u, ('time', 'latitude', 'longitude'), c), DataArray(v, ('time', 'latitude', 'longitude'), c)
    lcs = LCS(timestep=-21600, timedim='time', SETTLS_order=4)
    devnull = open(os.devnull, 'w')

    def call():
        so = sys.stdout
        sys.stdout = devnull
        try:
            return lcs(u=du, v=dv, verbose=False)
        finally:
            sys.stdout = so
    for _ in range(3):
        call()
    ts = []
    for _ in range(9):
        torch.cuda.synchronize(); t0 = time.perf_counter(); call(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps(dict(case='LCS call wall', ms_median=float(np.median(ts)), ms_min=float(np.min(ts)))), flush=True)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        call()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
    print(s.getvalue()[:9000])


if __name__ == '__main__':
    what = sys.argv[1:] or ['f32', 'single', 'profile']
    for w in what:
        globals()[w]()
