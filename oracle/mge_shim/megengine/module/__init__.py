class Module:
    def __init__(self, *a, **k):
        pass


class _Init:
    def __getattr__(self, name):
        return lambda *a, **k: None


init = _Init()
Conv2d = Linear = Module
