def set_trace(*a, **k):     # imported, unused (LCS.py:16)
    raise RuntimeError('refshim: set_trace is not expected to run')
