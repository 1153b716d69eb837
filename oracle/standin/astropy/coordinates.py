class SkyCoord:  # source side only; not on the hot path
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class SkyOffsetFrame:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('stand-in: sky frames are out of scope')
