"""Drop-in replacements for the reference's top-level packages `dynamics` and `transformations`.

Put THIS directory on sys.path ahead of the reference checkout (PYTHONPATH=.../rigid_body_manipulation_b200/dropin) and
`import dynamics as dyn`, `from dynamics import StateSpace, StateSpaceConfig`, `from transformations import Poses,
homogenize` (reference core/simulate.py:13-15, core/core.py:23-25, controllers/lqr.py:10) resolve to the B200-backed
implementations with the same names, signatures, defaults and exception types.  See INTEGRATION.md.
"""
