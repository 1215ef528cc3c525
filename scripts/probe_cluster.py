"""Scratch timing probe (not the bench): integrator time of one configuration, for A/B runs of library variants
(LCS_B200_LIB=variants/x.so python scripts/probe_cluster.py C2 296 outer f64)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S, _lib
from lagrangiancoherence_b200.engine import FtleEngine

cfg, B, xmode, prec = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
extra = dict(a.split('=') for a in sys.argv[5:])
if cfg == 'C2':
    lat, lon = S.grid_c2(); nt, dt = 9, -21600
elif cfg == 'C4':
    lat, lon = S.grid_c2(); nt, dt = 49, -3600
else:
    lat, lon = S.grid_c3(); nt, dt = 13, -3600
u, v = S.era5_like_winds(lat, lon, nt - 1 + B, noise=0.0)
eng = FtleEngine(lat, lon, dt, SETTLS_order=4, interp_order=int(extra.get('order', 3)), xmode=xmode, pair_dtype=prec,
                 **({'arith': extra['arith']} if 'arith' in extra else {}))
du, dv = torch.from_numpy(u).cuda(), torch.from_numpy(v).cuda()
st = eng.stage(du, dv)
x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(fn, n=7, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        flush.fill_(1); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))
t_adv = timeit(lambda: eng.advect(st, nsteps=nt - 1, nwindows=B, out=(x, y)))
t_stage = timeit(lambda: eng.stage(du, dv))
t_epi = timeit(lambda: eng.epilogue(x, y))
print(json.dumps(dict(lib=os.path.basename(os.environ.get('LCS_B200_LIB', 'main')), cfg=cfg, B=B, xmode=xmode, prec=prec, extra=extra,
                      advect_ms=t_adv, stage_ms=t_stage, epi_ms=t_epi,
                      Gpsteps=B * lat.size * lon.size * (nt - 1) / t_adv[0] / 1e6,
                      chk=float(x.double().sum().item()))), flush=True)
