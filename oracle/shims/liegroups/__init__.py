"""ORACLE / TEST INFRASTRUCTURE ONLY.  `liegroups` import surface (both
`from liegroups import SE3` -- reference `core/simulate.py:5`, `transformations/*.py` --
and `from liegroups.numpy import SE3` -- reference `dynamics/dynamics.py:6` -- are used).
See liegroups/numpy/__init__.py for what is restated and how it is pinned."""
from .numpy import SE3, SO3  # noqa: F401

__version__ = "1.1.0+oracle-shim"
