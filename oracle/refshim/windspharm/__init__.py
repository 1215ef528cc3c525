"""refshim: windspharm is imported at module level (tools.py:8, LCS.py:17) but only used by the global
regrid/truncation branch (LCS.py:115-118), which the goldens do not exercise."""
