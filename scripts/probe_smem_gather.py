"""Scratch: measured ceiling of a shared-memory-tile gather (best case: compact block lattice) next to the direct
L1/L2 gather the integrator uses.  Same element size (16 B), same 4x4 taps, same block tiling, C2 grid."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lagrangiancoherence_b200 import synthetic as S, _lib
from lagrangiancoherence_b200.engine import _ptr, _stream
lib = _lib.load()
lat, lon = S.grid_c2()
dev = torch.device('cuda', 0)
B, iters = 296, 40
buf = torch.zeros((lat.size * lon.size + 8) * 16, dtype=torch.uint8, device=dev)
sink = torch.zeros(1, dtype=torch.float64, device=dev)
def run(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
direct = run(lambda: _lib.check(lib.lcs_gather_peak(_ptr(buf), _lib.LCS_F64, 2, lat.size, lon.size, lat.size, lon.size, B, 4, 0.0, iters, _ptr(sink), _stream(dev)), 'gp'))
smem = run(lambda: _lib.check(lib.lcs_gather_peak_smem(_ptr(buf), lat.size, lon.size, lat.size, lon.size, B, iters, _ptr(sink), _stream(dev)), 'gps'))
bytes_ = B * lat.size * lon.size * iters * 16 * 16
print('direct L1/L2 gather : %.3f ms  %.1f TB/s' % (direct, bytes_ / direct / 1e9))
print('smem-tile gather    : %.3f ms  %.1f TB/s (taps only; + 10.5 KB staged per block and round)' % (smem, bytes_ / smem / 1e9))
