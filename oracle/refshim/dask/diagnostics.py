class ProgressBar:          # imported, unused (LCS.py:9)
    pass
