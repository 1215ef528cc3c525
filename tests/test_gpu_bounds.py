"""Memory safety without compute-sanitizer (closed on the measurement pool): a -DLCS_BOUNDS_CHECK build of the library
asserts, on the device, every index the gathers, the persistent kernel's state / candidate / flag tables and the
prefilter form; scripts/sanitize_case.py drives every kernel family through it on ragged sizes, in a subprocess (a
device-side assert poisons the CUDA context)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_kernel_family_under_device_side_bounds_checks(cuda_device):
    from lagrangiancoherence_b200 import build
    lib = os.path.join(ROOT, 'variants', 'liblcs_b200_bounds.so')
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    build.build(defines=['LCS_BOUNDS_CHECK=1'], lib=lib)
    env = dict(os.environ, LCS_B200_LIB=lib)
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'sanitize_case.py')], env=env, capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert 'sanitize_case: all kernels ran' in res.stdout
