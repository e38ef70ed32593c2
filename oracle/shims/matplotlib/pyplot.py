"""ORACLE / TEST INFRASTRUCTURE ONLY: matplotlib.pyplot stub (see the package docstring)."""
from . import _Null


class _Axes(list):
    """`plt.subplots(n, 1)` returns an indexable, sized collection of axes."""

    def __getattr__(self, name):
        return _Null()


def subplots(nrows=1, ncols=1, **kwargs):
    n = nrows * ncols
    return _Null(), (_Axes(_Null() for _ in range(n)) if n > 1 else _Null())


def show(*a, **k):
    return None


def __getattr__(name):
    return _Null()
