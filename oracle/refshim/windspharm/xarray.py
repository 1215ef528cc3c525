class VectorWind:
    def __init__(self, *a, **k):
        raise NotImplementedError('refshim: spectral truncation needs the real windspharm')
