"""Scratch: accuracy of the f32-storage fast path against the oracle (sigma and FTLE=0.5*log(sigma))."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lcs_oracle as O
from lagrangiancoherence_b200 import synthetic as S
from lagrangiancoherence_b200.engine import FtleEngine

lat = np.linspace(-40.0, 0.0, 161); lon = np.linspace(-80.0, -30.0, 201)   # 0.25 deg
u, v = S.era5_like_winds(lat, lon, 9)
for xmode, xclamp in (('pointwise', 'pointwise'), ('outer', 'outer')):
    rx, ry = O.parcel_propagation(u, v, lat, lon, -21600, SETTLS_order=4, xclamp=xclamp)
    ref = O.spectral_norm_field(O.flowmap_gradient(rx, ry, lat, lon))
    for pair, arith in (('f64', 'f64'), ('f32', 'f64'), ('f32', 'f32')):
        eng = FtleEngine(lat, lon, -21600, SETTLS_order=4, xmode=xmode, pair_dtype=pair, arith=arith)
        st = eng.stage(u, v)
        x, y = eng.advect(st)
        sig = eng.epilogue(x, y)[0].cpu().numpy()
        x, y = x[0].cpu().numpy(), y[0].cpu().numpy()
        ex = np.abs(x - rx) / np.abs(lon).max(); ey = np.abs(y - ry) / np.abs(lat).max()
        good = ref > 1e-6
        rel = np.abs(sig - ref)[good] / ref[good]
        ftle, fref = 0.5 * np.log(sig[good]), 0.5 * np.log(ref[good])
        frel = np.abs(ftle - fref) / np.maximum(np.abs(fref), 1e-3)
        print(xmode, pair, 'arith', arith, 'pos max %.2e med %.2e | sigma rel: med %.2e p99 %.2e max %.2e frac<1e-5 %.4f frac<1e-4 %.4f | ftle rel frac<1e-5 %.4f frac<1e-4 %.4f' % (
            max(ex.max(), ey.max()), np.median(np.maximum(ex, ey)), np.median(rel), np.percentile(rel, 99), rel.max(), (rel < 1e-5).mean(), (rel < 1e-4).mean(), (frel < 1e-5).mean(), (frel < 1e-4).mean()))
