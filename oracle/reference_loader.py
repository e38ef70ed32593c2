"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Executes the reference's OWN hot-path files, unmodified, from the read-only checkout at
/root/reference (this container only; the directory does not exist on the GPU box):

    dynamics/dynamics.py, transformations/transformations.py, transformations/poses.py,
    planners/joint_position_planner.py

Their third-party imports that are absent from the image (`liegroups`, `mujoco`,
`omegaconf`) are satisfied by the shims in oracle/shims/ (see the headers there).  The
modules are imported under private names and every sys.path / sys.modules change is
undone afterwards, so the product's own `dynamics` / `transformations` drop-in packages
are never shadowed.

Used by oracle/gen_golden.py (to produce tests/golden/*.npz) and by the `not gpu` tests
that validate oracle/dynamics_oracle.py against the reference itself.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("RBM_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_SHADOWED = ("dynamics", "transformations", "planners", "utilities", "liegroups", "mujoco", "omegaconf", "matplotlib", "controllers", "sensors",
             "visualization", "_rbm_ref_simulate")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dynamics", "dynamics.py"))


class ReferenceModules:
    """Namespace holding the reference modules (`.dynamics`, `.transformations`, `.planner`, `.liegroups`)."""


_cache = None


def load() -> ReferenceModules:
    """Import the reference's modules on top of the shims; returns a namespace of module objects."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"reference checkout not found at {REFERENCE_ROOT}")

    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _SHADOWED}
    for k in saved_mods:
        del sys.modules[k]
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        sys.path.insert(0, _SHIMS)
        ns = ReferenceModules()
        ns.liegroups = importlib.import_module("liegroups")
        ns.transformations = importlib.import_module("transformations")
        ns.dynamics = importlib.import_module("dynamics")
        # planners/__init__.py star-imports the planner module
        ns.planner = importlib.import_module("planners.joint_position_planner")
        assert os.path.realpath(ns.dynamics.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _SHADOWED]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
    _cache = ns
    return ns


_sim_cache = None


def load_simulation() -> ReferenceModules:
    """The reference's closed loop, unmodified: `.simulate` (core/simulate.py, loaded by file path because core/__init__.py pulls in
    dm_control), `.controllers` (controllers/lqr.py), `.planner`, `.dynamics` -- on top of the shims, with `mujoco._functions` bound to
    the functional stand-in of oracle/mujoco_standin.py and matplotlib stubbed out."""
    global _sim_cache
    if _sim_cache is not None:
        return _sim_cache
    if not available():
        raise FileNotFoundError(f"reference checkout not found at {REFERENCE_ROOT}")
    import importlib.util

    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _SHADOWED}
    for k in saved_mods:
        del sys.modules[k]
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        sys.path.insert(0, _SHIMS)
        ns = ReferenceModules()
        ns.dynamics = importlib.import_module("dynamics")
        ns.planner = importlib.import_module("planners.joint_position_planner")
        ns.controllers = importlib.import_module("controllers")
        spec = importlib.util.spec_from_file_location("_rbm_ref_simulate", os.path.join(REFERENCE_ROOT, "core", "simulate.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ns.simulate = mod
        assert os.path.realpath(ns.controllers.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _SHADOWED]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
    _sim_cache = ns
    return ns


_dropin_sim_cache = None


def load_simulation_on_dropin(dropin_root: str) -> ReferenceModules:
    """The reference's `core/simulate.py`, unmodified, importing the DROP-IN packages instead of its own: `dropin_root` (the product's
    rigid_body_manipulation_b200/dropin directory: `dynamics`, `transformations`, `controllers`, `planners`) precedes the reference
    checkout on sys.path, so `import dynamics as dyn`, `from transformations import Poses` (core/simulate.py:13-15) and the controller /
    planner modules resolve to the replacement, while `sensors`, `utilities`, `visualization` still come from the reference and
    `liegroups` / `mujoco` / `matplotlib` from the shims.  Returns `.simulate`, `.dynamics`, `.transformations`, `.controllers`,
    `.planner` (module objects), each asserted to come from where it should."""
    global _dropin_sim_cache
    if _dropin_sim_cache is not None:
        return _dropin_sim_cache
    if not available():
        raise FileNotFoundError(f"reference checkout not found at {REFERENCE_ROOT}")
    import importlib.util

    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _SHADOWED}
    for k in saved_mods:
        del sys.modules[k]
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        sys.path.insert(0, _SHIMS)
        sys.path.insert(0, dropin_root)
        ns = ReferenceModules()
        ns.dynamics = importlib.import_module("dynamics")
        ns.transformations = importlib.import_module("transformations")
        ns.planner = importlib.import_module("planners.joint_position_planner")
        ns.controllers = importlib.import_module("controllers")
        spec = importlib.util.spec_from_file_location("_rbm_ref_simulate", os.path.join(REFERENCE_ROOT, "core", "simulate.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ns.simulate = mod
        real = os.path.realpath
        for m_ in (ns.dynamics, ns.transformations, ns.planner, ns.controllers):
            assert real(m_.__file__).startswith(real(dropin_root)), m_.__file__
        assert mod.dyn is ns.dynamics and real(sys.modules["sensors"].__file__).startswith(real(REFERENCE_ROOT))
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in _SHADOWED]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
    _dropin_sim_cache = ns
    return ns
