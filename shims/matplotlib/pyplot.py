"""matplotlib.pyplot stand-in: every plotting call raises (nothing on the env path plots)."""


def __getattr__(name):
    def _missing(*a, **k):
        raise RuntimeError(f"matplotlib.pyplot.{name}: matplotlib is not installed (shims/matplotlib is an import-time stand-in)")
    return _missing
