"""refshim: IPython.core.debugger.set_trace is imported and unused (LCS.py:16)."""
