"""stub (tests/golden only)"""
def imread(*a, **k):
    raise NotImplementedError
