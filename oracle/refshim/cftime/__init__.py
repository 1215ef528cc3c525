"""refshim: only the isinstance check of trajectory.py:129 touches cftime."""


class Datetime360Day:
    pass
