"""CPU oracle for the FTLE hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on numpy + scipy + numba, what gabrielmpp/LagrangianCoherence
*executes* for the path  departure points -> flow-map Jacobian -> sigma_max field
(SURVEY.md section 8a).  It exists to check the CUDA engine in
``lagrangiancoherence_b200``; nothing in the product package may import it.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.

Pinning status
--------------
The reference ships no tests and no golden vectors, and it cannot be
imported as-is in this image (xarray, dask, IPython, windspharm, xr_tools, cftime are
absent).  The oracle is pinned two ways instead:

* ``oracle/refshim`` provides a minimal stand-in for those absent imports so that the
  UNMODIFIED reference sources under /root/reference are executed in the build container;
  ``oracle/make_golden.py`` records their outputs under ``tests/golden/`` and
  ``tests/test_oracle_golden.py`` holds the oracle to them.  The labelled-array semantics
  of that stand-in (notably orthogonal indexing for ``da[np.where(..)] = c``) are this
  repository's reading of xarray's documented behaviour, so the pin is
  "reference code + emulated xarray", not "reference code + real xarray".
* analytic known-answer tests that follow from the reference code (zero wind, uniform
  zonal wind, see ``tests/test_oracle_kat.py``).
"""
