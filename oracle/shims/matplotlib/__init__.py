"""ORACLE / TEST INFRASTRUCTURE ONLY.  Import-time stub of matplotlib (absent from the image) so that the reference's
core/simulate.py and visualization.py can be executed unmodified; nothing is drawn."""


class _Null:
    """Absorbs any attribute access, call, indexing or iteration the plotting code performs."""

    def __getattr__(self, name):
        return _Null()

    def __call__(self, *a, **k):
        return _Null()

    def __getitem__(self, k):
        return _Null()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):            # `handles, labels = ax.get_legend_handles_labels()`
        return iter((_Null(), _Null()))

    def __len__(self):
        return 2


rcParams = {}
