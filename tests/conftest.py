import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('a test marked gpu was collected without a CUDA device')
    from lagrangiancoherence_b200 import build
    build.build()           # no-op when the in-tree .so is current
    return 'cuda:0'


def position_parity(name, ex, ey, tol=1e-10, max_fraction=1e-3, max_outlier=None):
    """Departure-point parity bookkeeping shared by the GPU parity tests.  ``ex, ey``: relative errors.  A few particles
    may sit on a discontinuity of the scheme (mod-180 wrap, order-1 'constant' cut-off of the pole rows, clamp
    thresholds, the exit flags of the outer-product clamp) where a 1-ulp difference selects the other branch.  Their
    number is bounded by ``max_fraction`` and -- when ``max_outlier`` is given -- so is their magnitude; both are
    printed so that the test log records them (run pytest with -s / -rP)."""
    import numpy as np
    out = np.concatenate([np.ravel(ex)[np.ravel(ex) > tol], np.ravel(ey)[np.ravel(ey) > tol]])
    worst = float(out.max()) if out.size else 0.0
    print(f'PARITY {name}: {out.size} of {ex.size + ey.size} coordinates beyond {tol:g} (flipped branches), largest {worst:.3e}; '
          f'median error {float(np.median(ex)):.1e} / {float(np.median(ey)):.1e}')
    assert (ex > tol).mean() <= max_fraction and (ey > tol).mean() <= max_fraction, (name, out.size, worst)
    if max_outlier is not None:
        assert worst <= max_outlier, (name, out.size, worst)
    return out.size, worst
