"""Empty stand-in: wrappers.py sets pandas' plotting backend to "plotly" at import time (plots are not on the step path)."""


def plot(*a, **k):
    raise NotImplementedError("plotly is not installed; plotting is outside the hot path")
