"""stub (tests/golden only): plotting is out of scope"""
def __getattr__(name):
    def _f(*a, **k):
        raise NotImplementedError("matplotlib stub")
    return _f
