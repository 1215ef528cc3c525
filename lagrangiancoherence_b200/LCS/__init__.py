"""Reference-shaped package layout: ``LCS.LCS`` (class LCS, flowmap_gradient),
``LCS.trajectory`` (parcel_propagation) and ``LCS.tools`` (array-level seams), so that
``import lagrangiancoherence_b200 as LagrangianCoherence`` keeps the reference's import paths
(LCS.py:12-15, examples/ideal_vortex.py:5-8) working."""
from . import LCS, trajectory, tools  # noqa: F401
