"""refshim: the reference imports dask.diagnostics.ProgressBar and never uses it (LCS.py:9)."""
