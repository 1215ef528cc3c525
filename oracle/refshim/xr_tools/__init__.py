"""refshim: xr_tools is the reference author's un-vendored helper package (LCS.py:13)."""
