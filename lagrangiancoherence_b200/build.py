"""In-tree build of liblcs_b200.so (sm_100a only) with nvcc.

``python -m lagrangiancoherence_b200.build`` or ``build()``; the shared object lands next to this
file so that it travels with the repository snapshot (it is git-ignored, not gpurun-ignored).
Every translation unit is compiled to an object file in ``csrc/_obj`` (in parallel, only when its
sources changed) and the objects are linked into the library.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liblcs_b200.so')
SOURCES = ['advect.cu', 'advect_inst_es3.cu', 'advect_inst_f64.cu', 'advect_inst_orders.cu', 'advect_inst_f32.cu', 'advect_inst_r32.cu',
           'prefilter.cu', 'epilogue.cu', 'filters.cu', 'spectral.cu', 'seams.cu']
HEADERS = ['lcs_device.cuh', 'lcs_internal.h', 'advect_kernels.cuh', os.path.join('..', '..', 'include', 'lcs_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: liblcs_b200.so cannot be built')
    return exe


def _deps():
    return [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]


def is_stale(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False, defines=(), lib=LIB):
    """Compile every CUDA source into ``lib`` (default liblcs_b200.so); returns the library path.
    ``defines``: extra ``-DNAME=VALUE`` macros (tuning experiments build variant libraries this way)."""
    if not force and not is_stale(lib):
        return lib
    tag = hashlib.sha1(' '.join(sorted(defines)).encode()).hexdigest()[:8] if defines else 'default'
    objdir = os.path.join(CSRC, '_obj', tag)
    os.makedirs(objdir, exist_ok=True)
    headers_t = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    flags = NVCC_FLAGS + [f'-D{d}' for d in defines] + (['-Xptxas', '-v'] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, src[:-3] + '.o')
        path = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj) and
                os.path.getmtime(obj) > max(os.path.getmtime(path), headers_t, os.path.getmtime(os.path.abspath(__file__)))):
            return obj, 0, ''
        res = subprocess.run([_nvcc()] + flags + ['-c', '-o', obj, path], capture_output=True, text=True)
        return obj, res.returncode, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    log = ''.join(r[2] for r in results)
    if any(r[1] for r in results):
        sys.stderr.write(log)
        raise RuntimeError('nvcc failed building liblcs_b200.so')
    res = subprocess.run([_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', lib] + [r[0] for r in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking liblcs_b200.so')
    if verbose:
        sys.stderr.write(log)
    return lib


if __name__ == '__main__':
    defs = [a[2:] for a in sys.argv[1:] if a.startswith('-D')]
    out = next((a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--out=')), LIB)
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, defines=defs, lib=out))
