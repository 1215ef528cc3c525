"""Scratch timing probe (not the bench): per-kernel timings of the C2 / C3 configurations."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S, _lib
from lagrangiancoherence_b200.engine import FtleEngine, _ptr, _stream

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

def run(cfg, B, xmode, order=3, pair='f64', strict=False, layout='es'):
    if cfg == 'C2':
        lat, lon = S.grid_c2(); nt = 9; dt = -21600
    else:
        lat, lon = S.grid_c3(); nt = 13; dt = -3600
    nlev = nt - 1 + B
    u, v = S.era5_like_winds(lat, lon, nlev, noise=0.0)
    eng = FtleEngine(lat, lon, dt, SETTLS_order=4, interp_order=order, xmode=xmode, pair_dtype=pair, strict=strict, layout=layout)
    du = torch.from_numpy(u).cuda(); dv = torch.from_numpy(v).cuda()
    st = eng.stage(du, dv)
    t_stage = timeit(lambda: eng.stage(du, dv))
    x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
    t_adv = timeit(lambda: eng.advect(st, nsteps=nt-1, nwindows=B, out=(x, y)))
    t_epi = timeit(lambda: eng.epilogue(x, y))
    psteps = B * lat.size * lon.size * (nt - 1)
    bytes_ps = (2 + 4 * 4) * (order + 1) ** 2 * (8 if pair == 'f64' else 4)
    print(json.dumps(dict(cfg=cfg, B=B, layout=layout, xmode=xmode, order=order, pair=pair, strict=strict, band=os.environ.get("LCS_ADVECT_BAND_LOG2", "2"),
          stage_ms=t_stage, advect_ms=t_adv, epi_ms=t_epi,
          Mpsteps_per_s=psteps / t_adv[0] / 1e3, gather_GBs=psteps * bytes_ps / t_adv[0] / 1e6)), flush=True)
    return eng, st

if __name__ == '__main__':
    lib = _lib.load()
    for B in (1, 64):
        for xmode in ('pointwise', 'outer'):
            run('C2', B, xmode)
    run('C2', 64, 'pointwise', strict=True)
    run('C2', 64, 'pointwise', layout='pair4')
    run('C2', 64, 'outer', layout='pair4')
    run('C2', 64, 'pointwise', pair='f32')
    run('C2', 64, 'pointwise', order=1)
    run('C3', 1, 'pointwise')
    # gather peak on an L2-resident C2 level
    eng, st = run('C2', 64, 'pointwise')
    sink = torch.zeros(1, dtype=torch.float64, device='cuda')
    buf = torch.zeros(281 * 321 * 32 + 1024, dtype=torch.uint8, device='cuda')
    for vec in (4, 2):
        for jit in (0.0, 2.0, 8.0, 40.0):
            fn = lambda: _lib.check(lib.lcs_gather_peak(_ptr(buf), 0, vec, 281, 321, 281, 321, 64, 4, jit, 40, _ptr(sink), _stream(eng.device)), 'gp')
            t = timeit(fn)
            print('gather_peak vec', vec, 'jitter', jit, 'ms', t, 'GB/s', 64 * 281 * 321 * 40 * 16 * vec * 8 / t[0] / 1e6, flush=True)
