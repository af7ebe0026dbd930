class FSDPStrategy:
    def __init__(self, *a, **k):
        raise RuntimeError("stub")


class XLAStrategy(FSDPStrategy):
    pass
