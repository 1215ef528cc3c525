"""B200-native FTLE engine behind the Python API of gabrielmpp/LagrangianCoherence.

    from lagrangiancoherence_b200.LCS.LCS import LCS, flowmap_gradient
    from lagrangiancoherence_b200.LCS.trajectory import parcel_propagation

Every per-particle / per-grid-point operation runs in hand-written sm_100a CUDA kernels
(liblcs_b200.so, C ABI in include/lcs_b200.h).  There is no CPU fallback.
"""
__version__ = '0.1.0'

from .labelled import DataArray, Dataset  # noqa: F401
