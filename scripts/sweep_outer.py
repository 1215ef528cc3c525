"""A/B sweep of the outer-clamp integrator variants in one process (the library reads its LCS_* knobs per call).
Usage: python scripts/sweep_outer.py [--batch 1184] [--workload C2] CONFIG ...   with CONFIG = "K=V,K=V" (LCS_ prefix implied)
Prints one JSON line per configuration: advect ms (best / median of --reps) and whether the positions are
bit-identical to the first configuration."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S          # noqa: E402
from lagrangiancoherence_b200.engine import FtleEngine, precision_args      # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1184)
ap.add_argument('--workload', default='C2')
ap.add_argument('--xclamp', default='outer')
ap.add_argument('--precision', default='f64')
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('configs', nargs='*')
a = ap.parse_args()

if a.workload == 'C3':
    lat, lon = S.grid_c3(); nt, dt = 73, -3600
elif a.workload == 'C4':
    lat, lon = S.grid_c2(); nt, dt = 49, -3600
else:
    lat, lon = S.grid_c2(); nt, dt = 9, -21600
dev = torch.device('cuda:0')
B = a.batch
u, v = S.era5_like_winds(lat, lon, B + nt - 1)
if a.precision != 'f64':
    u, v = u.astype(np.float32), v.astype(np.float32)
eng = FtleEngine(lat, lon, dt, SETTLS_order=4, interp_order=3, xmode=a.xclamp, device=dev, **precision_args(a.precision))
st = eng.stage(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev))
x = torch.empty((B, lat.size, lon.size), dtype=torch.float64, device=dev)
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
keys = set()
for cfg in a.configs or ['']:
    for k in keys:
        os.environ.pop(k, None)
    keys = set()
    for kv in filter(None, cfg.split(',')):
        k, val = kv.split('=')
        os.environ['LCS_' + k] = val
        keys.add('LCS_' + k)
    eng._ws = None
    times = []
    try:
        for i in range(a.reps + 2):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.advect(st, nsteps=nt - 1, nwindows=B, out=(x, y))
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1))
        eng.check_finite()
    except Exception as exc:                                   # noqa: BLE001
        print(json.dumps({'config': cfg, 'error': repr(exc)}), flush=True)
        continue
    same = None
    if ref is None:
        ref = (x.clone(), y.clone())
    else:
        same = bool(torch.equal(ref[0], x) and torch.equal(ref[1], y))
    psteps = B * lat.size * lon.size * (nt - 1)
    print(json.dumps({'config': cfg, 'ms_best': min(times), 'ms_median': float(np.median(times)),
                      'Gpsteps_per_s': psteps / (min(times) * 1e-3) / 1e9, 'bit_identical_to_first': same}), flush=True)
